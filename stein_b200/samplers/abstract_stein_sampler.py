import ctypes
from abc import abstractmethod

import numpy as np

from ..engine import SvgdEngine
from ..kernels import SquaredExponentialKernel
from ..log_p.base import LogPosterior, Output
from ..log_p.graph_log_p import GraphLogPosterior, is_graph_tensor
from ..optimizers._fused import FusedGradientDescent
from ..runtime import context, ptr
from ..utilities import convert_array_to_dictionary, convert_dictionary_to_array


class AbstractSteinSampler:
    """Stein variational gradient descent (Liu & Wang 2016) on a B200.

    Mirrors stein/samplers/abstract_stein_sampler.py:7-196.  Same public surface:
    `n_particles`, `model_vars`, `log_p`, `theta` (dict variable -> (n, *shape)
    float64 array), `compute_phi`, `update_particles`, `function_posterior`.

    Differences in mechanism, not in results: the particles, scores and optimizer
    moments are device resident (fp32) in a `stein_engine` and `theta` is
    materialised on the host only when read; there is no tf.Session (`sess` is
    None).
    """

    def __init__(self, n_particles, log_p, theta=None):
        if is_graph_tensor(log_p):
            # a `log_p` tensor recorded by the TF1 stand-in of compat/ (the reference's own scripts)
            log_p = GraphLogPosterior(log_p)
        if not isinstance(log_p, LogPosterior):
            raise TypeError(
                "log_p must be a stein_b200.log_p.LogPosterior (LinearRegression, "
                "LogisticRegression, RegressionNeuralNetwork, TorchLogPosterior) or a graph "
                "tensor built with the `tensorflow` stand-in of compat/; real TensorFlow graphs "
                "are not supported")
        self.n_particles = n_particles
        self.sess = None
        self.log_p = log_p
        self.model_vars = list(log_p.model_vars)       # abstract_stein_sampler.py:49-51
        self.grad_log_p = log_p.scores                 # :55 -- batched, all particles at once
        self._access_indices = log_p.column_slices()
        self._theta_cache = None
        if theta is not None:
            self._init_theta = dict(theta)             # :66-67
        else:
            # :69-74 -- same NumPy global-RNG draws, in model_vars (creation) order
            self._init_theta = {
                v: np.random.normal(size=[self.n_particles] + v.get_shape().as_list()) * 0.01
                for v in self.model_vars
            }
        self._engine = None

    # -- engine -------------------------------------------------------------------
    def _make_engine(self, gd):
        hyper = gd._hyper() if isinstance(gd, FusedGradientDescent) else dict(optimizer="adam")
        self._engine = SvgdEngine(self.n_particles, self.log_p.n_params, **hyper)
        if isinstance(gd, FusedGradientDescent):
            gd._bind(self._engine)
        full = convert_dictionary_to_array(self._init_theta)[0]
        if full.shape != (self.n_particles, self.log_p.n_params):
            raise ValueError("theta has shape %r, expected %r"
                             % (full.shape, (self.n_particles, self.log_p.n_params)))
        e = self._engine
        e.set_particles(full[e.row_begin:e.row_begin + e.n_local])
        self._init_theta = None

    @property
    def engine(self):
        return self._engine

    # -- particles as the reference exposes them -------------------------------------
    @property
    def theta(self):
        """{variable: (n_local, *shape) float64}.  With one GPU n_local == n_particles."""
        if self._theta_cache is None:
            arr = self._engine.get_particles(np.float64)
            self._theta_cache = convert_array_to_dictionary(arr, self._access_indices)
        return self._theta_cache

    @theta.setter
    def theta(self, dictionary):
        arr = convert_dictionary_to_array(dictionary)[0]
        self._engine.set_particles(arr)
        self._theta_cache = None

    # -- phi ----------------------------------------------------------------------------
    def compute_phi(self, theta_array, grads_array):
        """(K.S + dK) / n for host arrays (abstract_stein_sampler.py:76-105), through
        the fused GPU path: median -> bandwidth -> phi, K never materialised."""
        import torch
        ctx = context()
        theta_array = np.asarray(theta_array)
        n, d = theta_array.shape
        if not isinstance(self.kernel, SquaredExponentialKernel):
            # any other kernel operator (abstract_kernel.py:45-62): the reference's own formula
            # (K.dot(S) + dK) / n (:105), with the GEMM on the device
            K, dK = self.kernel.kernel_and_grad(theta_array)
            return self._phi_from_kernel(ctx, ctx.to_square(K), ctx.to_padded(dK), ctx.to_padded(grads_array),
                                         n, d)[0][:n, :d].double().cpu().numpy()
        X, S = ctx.to_padded(theta_array), ctx.to_padded(grads_array)
        rows, ld = X.shape
        r = torch.empty(rows, dtype=torch.float32, device=X.device)
        ctx.check(ctx.lib.stein_row_norms(ctx.handle, ptr(X), n, d, ld, ptr(r)))
        bw = self.kernel._bandwidth_dev(ctx, X, r, n, d)
        ws_bytes = int(ctx.lib.stein_phi_workspace_bytes(ctx.handle, n, n, d))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=X.device)
        phi = torch.empty_like(X)
        sumsq = torch.zeros(1, dtype=torch.float64, device=X.device)
        ctx.check(ctx.lib.stein_phi(ctx.handle, ptr(X), ptr(S), ptr(r), n, d, ld, 0, n, float(bw),
                                    ptr(ws), ws_bytes, ptr(phi), ptr(sumsq)))
        return phi[:n, :d].double().cpu().numpy()

    @staticmethod
    def _phi_from_kernel(ctx, K, dK, S, n, d, phi=None, sumsq=None):
        """(K S + dK) / n on the device (stein_phi_from_kernel); returns (phi, sumsq) tensors."""
        import torch
        rows, ld = S.shape
        phi = torch.empty_like(S) if phi is None else phi
        sumsq = torch.zeros(1, dtype=torch.float64, device=S.device) if sumsq is None else sumsq
        ws = torch.empty(rows * ld * 4 + 16384, dtype=torch.uint8, device=S.device)
        ctx.check(ctx.lib.stein_phi_from_kernel(ctx.handle, ptr(K), K.shape[1], ptr(dK), ptr(S), n, d, ld, ptr(ws),
                                                ws.numel(), ptr(phi), ptr(sumsq)))
        return phi, sumsq

    def _plugin_kernel_step(self):
        """One iteration with a kernel operator other than the squared-exponential one: K and dK from
        `self.kernel` (on the device when the operator offers kernel_and_grad_dev, else through its
        host arrays like the reference), phi = (K S + dK) / n, then the engine's clip + optimizer."""
        import torch
        e, ctx = self._engine, self._engine.ctx
        if e.n_local != e.n_particles:
            raise NotImplementedError("kernel operators other than SquaredExponentialKernel run on one GPU")
        n, d = e.n_particles, e.n_params
        X, S = e.particles_dev, e.scores_dev
        if hasattr(self.kernel, "kernel_and_grad_dev"):
            r = torch.empty(e.rows_padded, dtype=torch.float32, device=X.device)
            ctx.check(ctx.lib.stein_row_norms(ctx.handle, ptr(X), n, d, e.ld, ptr(r)))
            K, dK = self.kernel.kernel_and_grad_dev(ctx, X, r, n, d)
        else:
            Kh, dKh = self.kernel.kernel_and_grad(e.get_particles(np.float64))
            K = ctx.to_square(Kh)
            dK = torch.zeros_like(X)
            dK[:n, :d] = torch.from_numpy(np.ascontiguousarray(dKh, dtype=np.float32)).to(X.device)
        self._phi_from_kernel(ctx, K, dK, S, n, d, phi=e.phi_dev, sumsq=e.sumsq_dev)

    def update_particles(self, grads_array):
        """abstract_stein_sampler.py:107-127 for a host score matrix."""
        e = self._engine
        grads_array = np.ascontiguousarray(grads_array)
        fused_kernel = isinstance(self.kernel, SquaredExponentialKernel)
        if isinstance(self.gd, FusedGradientDescent) and fused_kernel:
            self._sync_kernel_bandwidth()
            self.gd._push_hyper()          # the reference reads lr / decay / betas at every update()
            e.update_particles_host(grads_array)
            self.gd._after_engine_step()
        elif isinstance(self.gd, FusedGradientDescent):
            e.set_scores(grads_array)
            self._plugin_kernel_step()
            self.gd._push_hyper()
            e.apply_phi()
            self.gd._after_engine_step()
        else:
            # user-defined step rule: same sequence as the reference, phi from the GPU
            e.set_scores(grads_array)
            self._device_phi_only()
            phi = e.get_phi(np.float64)
            phi *= 10. / max(10., np.linalg.norm(phi))
            e.set_particles(e.get_particles(np.float64) + self.gd.update(phi))
        self._theta_cache = None

    def _sync_kernel_bandwidth(self):
        """The engine follows `self.kernel`: a kernel built with `bandwidth=h` makes it
        skip the median (the kernel object may be replaced after construction)."""
        want = getattr(self.kernel, "fixed_bandwidth", None)
        if want != getattr(self, "_engine_bandwidth", None):
            self._engine.set_bandwidth(want)
            self._engine_bandwidth = want

    def _device_phi_only(self):
        """phi for the scores in the engine, without the optimizer step (stein_engine_phi_only:
        the engine's own workspace, leading dimension and row shard)."""
        if not isinstance(self.kernel, SquaredExponentialKernel):
            return self._plugin_kernel_step()
        self._sync_kernel_bandwidth()
        self._engine.phi_only()

    # -- posterior functionals --------------------------------------------------------------
    def function_posterior(self, func, feed_dict, axis=None):
        """abstract_stein_sampler.py:129-168: `func` evaluated for every particle,
        one row per particle; `axis` averages.  `func` is an Output handle of the
        model (e.g. model.logits, model.pred): all particles in one kernel."""
        if is_graph_tensor(func) and isinstance(self.log_p, GraphLogPosterior):
            dist = self.log_p.evaluate(func, self._engine, feed_dict)     # any tensor of the graph
        elif isinstance(func, Output):
            dist = func.model.evaluate(func, self._engine, feed_dict)
        else:
            raise TypeError("func must be an Output handle of the model (e.g. model.logits)")
        dist = dist.reshape(dist.shape[0], -1)
        if axis is not None:
            return dist.double().mean(dim=axis).cpu().numpy()
        return dist.double().cpu().numpy()

    @abstractmethod
    def train_on_batch(self, batch_feed):
        raise NotImplementedError()
