"""Shared machinery of the two built-in step rules: the arithmetic lives in
libstein_b200.so (stein_clip_adam_step / stein_clip_adagrad_step)."""
import ctypes

import numpy as np

from ..runtime import context, ptr
from .abstract_gradient_descent import AbstractGradientDescent


class FusedGradientDescent(AbstractGradientDescent):
    """Base of the built-in rules.  Two ways of running:

    * bound to a sampler's engine (the normal case): the engine performs
      clip + update + `X +=` in one kernel and this object only mirrors
      `n_iters` / `learning_rate` and exposes the moments;
    * stand-alone `update(phi)` on a host array (reference API): the same kernel
      is launched on a zero "particle" buffer with the clip disabled, so the
      returned array is exactly the step.
    """
    _kind = None            # "adam" | "adagrad"
    _state_names = ()       # attribute names of the moment buffers (m1, m2)

    def __init__(self, learning_rate, decay):
        super().__init__(learning_rate, decay)
        self._engine = None
        self._dev = None    # stand-alone device state: dict of torch tensors

    # -- engine binding ---------------------------------------------------------
    def _hyper(self):
        raise NotImplementedError()

    def _bind(self, engine):
        self._engine = engine
        self._pushed = self._hyper()

    def _push_hyper(self):
        """Forward hyper-parameters changed on this object since the last step (a learning-rate
        schedule, `gd.learning_rate = x`) to the engine, which otherwise keeps its own copy."""
        h = self._hyper()
        if self._engine is not None and h != getattr(self, "_pushed", None):
            self._engine.set_hyper(h["learning_rate"], h["decay"], h["p1"], h["p2"])
        self._pushed = h

    def _after_engine_step(self):
        self.n_iters += 1
        if self._kind == "adam":
            self.learning_rate *= self.decay   # adam_gradient_descent.py:56
        if getattr(self, "_pushed", None) is not None:
            self._pushed = self._hyper()       # the engine applied the same decay to its copy

    def _moment(self, which):
        if self._engine is not None:
            st = self._engine.get_state()
            return st["m1"] if which == 0 else st["m2"]
        if self._dev is None:
            raise AttributeError("optimizer has not been stepped yet")
        n, d = self._dev["shape"]
        return self._dev["m%d" % (which + 1)][:n, :d].double().cpu().numpy()

    # -- stand-alone update -----------------------------------------------------
    def update(self, phi):
        import torch
        ctx = context()
        phi = np.asarray(phi, dtype=np.float64)
        if phi.ndim != 2:
            raise ValueError("phi must be (n_particles x n_params)")
        p = ctx.to_padded(phi)
        if self._dev is None or self._dev["m1"].shape != p.shape:
            self._dev = {"shape": phi.shape, "m1": torch.zeros_like(p), "m2": torch.zeros_like(p),
                         "zero": torch.zeros(1, dtype=torch.float64, device=p.device)}
        step = torch.zeros_like(p)               # X = 0  ->  X += step  ->  step
        self._launch(ctx, step, p)
        self._after_engine_step()
        n, d = phi.shape
        return step[:n, :d].double().cpu().numpy()

    def _launch(self, ctx, X, p):
        raise NotImplementedError()
