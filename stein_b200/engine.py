"""Python handle on `stein_engine` (device-resident particles, scores, phi and
optimizer moments; one SVGD update per step()).

Replaces AbstractSteinSampler.update_particles,
stein/samplers/abstract_stein_sampler.py:107-127 (reference keeps all of this in
float64 NumPy on the host and crosses into TensorFlow twice per iteration).
"""
import ctypes

import numpy as np

from . import _lib
from .runtime import context, _torch


class SvgdEngine:
    def __init__(self, n_particles, n_params, optimizer="adam", learning_rate=1e-3, decay=1.0,
                 p1=0.9, p2=0.999, ctx=None, peer_push=True):
        self.ctx = ctx or context()
        self.lib = self.ctx.lib
        self.n_particles, self.n_params = int(n_particles), int(n_params)
        kind = {"adam": _lib.OPT_ADAM, "adagrad": _lib.OPT_ADAGRAD}[optimizer]
        self.handle = ctypes.c_void_p()
        self.ctx.check(self.lib.stein_engine_create(
            ctypes.byref(self.handle), self.ctx.handle, self.n_particles, self.n_params, kind,
            float(learning_rate), float(decay), float(p1), float(p2)))
        rb, nl = ctypes.c_int64(), ctypes.c_int64()
        self.ctx.check(self.lib.stein_engine_local_rows(self.handle, ctypes.byref(rb), ctypes.byref(nl)))
        self.row_begin, self.n_local = rb.value, nl.value
        x, s, p = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        ld, rows = ctypes.c_int64(), ctypes.c_int64()
        self.ctx.check(self.lib.stein_engine_buffers(self.handle, ctypes.byref(x), ctypes.byref(s),
                                                     ctypes.byref(p), ctypes.byref(ld), ctypes.byref(rows)))
        self.ld, self.rows_padded = ld.value, rows.value
        self.x_ptr, self.s_ptr, self.phi_ptr = x.value, s.value, p.value
        # sharded run on the library's NCCL transport: let the optimizer kernel push the updated
        # rows into the peers' buffers (replaces the all-gather of the particles)
        self.peer_push = False
        if peer_push and getattr(self.ctx, "_native_comm_group", None) is not None:
            from .distributed import connect_peers
            self.peer_push = connect_peers(self)

    def close(self, collective=True):
        """Frees the engine.  With peers connected this is COLLECTIVE: the other ranks' kernels
        store into this engine's buffers, so every rank first finishes its queued work and meets
        the others at a barrier."""
        if self.handle:
            if self.peer_push and collective:
                from .distributed import quiesce_peers
                quiesce_peers(self)
            self.lib.stein_engine_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close(collective=False)     # garbage collection is not a collective moment
        except Exception:
            pass

    # ---- device views (torch tensors aliasing the engine's buffers) ---------
    def _view(self, ptr):
        from .distributed import _DevMem
        torch = _torch()
        t = torch.as_tensor(_DevMem(ptr, self.rows_padded * self.ld, "<f4"),
                            device="cuda:%d" % self.ctx.device)
        return t.view(self.rows_padded, self.ld)

    @property
    def particles_dev(self):
        """View of the local particle rows (the score kernels read it).  A caller that WRITES through it calls
        particles_changed() before the next step."""
        return self._view(self.x_ptr)

    def particles_changed(self):
        """The particle buffer was modified through particles_dev: drop what the engine derived ahead of time
        from the old particles (the prefetched median of update_particles_host, the peers' copies of the rows)."""
        self.ctx.check(self.lib.stein_engine_particles_changed(self.handle))

    @property
    def scores_dev(self):
        return self._view(self.s_ptr)

    @property
    def phi_dev(self):
        return self._view(self.phi_ptr)

    @property
    def sumsq_dev(self):
        """1-element float64 device tensor: sum(phi^2) that the clip of the next apply_phi() reads."""
        from .distributed import _DevMem
        p = ctypes.c_void_p()
        self.ctx.check(self.lib.stein_engine_sumsq_dev(self.handle, ctypes.byref(p)))
        return _torch().as_tensor(_DevMem(p.value, 1, "<f8"), device="cuda:%d" % self.ctx.device)

    # ---- host transfers --------------------------------------------------------
    def _host(self, a):
        a = np.asarray(a)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        a = np.ascontiguousarray(a)
        if a.shape != (self.n_local, self.n_params):
            raise ValueError("expected a (%d, %d) array, got %r" % (self.n_local, self.n_params, a.shape))
        return a, int(a.dtype == np.float64)

    def set_particles(self, X):
        a, f64 = self._host(X)
        self.ctx.sync_stream()
        self.ctx.check(self.lib.stein_engine_set_particles(self.handle, a.ctypes.data_as(ctypes.c_void_p), f64))

    def get_particles(self, dtype=np.float64, out=None):
        out = np.empty((self.n_local, self.n_params), dtype) if out is None else out
        self.ctx.sync_stream()
        self.ctx.check(self.lib.stein_engine_get_particles(
            self.handle, out.ctypes.data_as(ctypes.c_void_p), int(out.dtype == np.float64)))
        return out

    def set_scores(self, S):
        a, f64 = self._host(S)
        self.ctx.sync_stream()
        self.ctx.check(self.lib.stein_engine_set_scores(self.handle, a.ctypes.data_as(ctypes.c_void_p), f64))

    def get_phi(self, dtype=np.float64):
        out = np.empty((self.n_local, self.n_params), dtype)
        self.ctx.sync_stream()
        self.ctx.check(self.lib.stein_engine_get_phi(
            self.handle, out.ctypes.data_as(ctypes.c_void_p), int(dtype == np.float64)))
        return out

    # ---- the iteration ------------------------------------------------------------
    def step(self):
        """One update_particles() on the scores currently in the S buffer."""
        self.ctx.sync_stream()
        self.ctx.check(self.lib.stein_engine_step(self.handle))

    def update_particles_host(self, S_host, X_out=None):
        """Host-buffer drop-in for update_particles(grads_array): H2D scores,
        step, (optionally) D2H particles.  Arrays must be C-contiguous float32 or
        float64 of shape (n_local, n_params)."""
        self.ctx.sync_stream()
        S_host, f64 = self._host(S_host)           # shape, dtype, contiguity
        if X_out is not None:
            if not isinstance(X_out, np.ndarray) or X_out.dtype != S_host.dtype:
                raise ValueError("S_host and X_out must share a dtype (float32 or float64)")
            if X_out.shape != (self.n_local, self.n_params) or not X_out.flags.c_contiguous:
                raise ValueError("X_out must be a C-contiguous (%d, %d) array" % (self.n_local, self.n_params))
        self.ctx.check(self.lib.stein_engine_update_particles_host(
            self.handle, S_host.ctypes.data_as(ctypes.c_void_p),
            X_out.ctypes.data_as(ctypes.c_void_p) if X_out is not None else None, f64))
        return X_out

    def set_prefetch(self, on):
        """update_particles_host enqueues the next iteration's median behind the particle download (default on)."""
        self.ctx.check(self.lib.stein_engine_set_prefetch(self.handle, int(bool(on))))

    def set_device_bandwidth(self, on):
        """phi is enqueued ahead of the host's median result, reading the bandwidth from the device (default on)."""
        self.ctx.check(self.lib.stein_engine_set_device_bandwidth(self.handle, int(bool(on))))

    def device_bandwidth_stats(self):
        u, r = ctypes.c_int64(), ctypes.c_int64()
        self.ctx.check(self.lib.stein_engine_device_bandwidth_stats(self.handle, ctypes.byref(u), ctypes.byref(r)))
        return {"used": u.value, "redone": r.value}

    def prefetch_stats(self):
        b, u = ctypes.c_int64(), ctypes.c_int64()
        self.ctx.check(self.lib.stein_engine_prefetch_stats(self.handle, ctypes.byref(b), ctypes.byref(u)))
        return {"begun": b.value, "used": u.value}

    def phi_only(self):
        """compute_phi() on the scores in the S buffer, into the phi buffer; no optimizer step."""
        self.ctx.sync_stream()
        self.ctx.check(self.lib.stein_engine_phi_only(self.handle))

    def apply_phi(self):
        """Clip + optimizer step on the phi / sum(phi^2) currently in the engine's buffers."""
        self.ctx.sync_stream()
        self.ctx.check(self.lib.stein_engine_apply_phi(self.handle))

    def set_hyper(self, learning_rate, decay, p1, p2):
        self.ctx.check(self.lib.stein_engine_set_hyper(self.handle, float(learning_rate), float(decay),
                                                       float(p1), float(p2)))

    def set_bandwidth(self, bandwidth=None):
        """Fixed bandwidth h for the following steps; None / 0 = the reference's
        median heuristic (abstract_kernel.py:40)."""
        self.ctx.check(self.lib.stein_engine_set_bandwidth(self.handle, float(bandwidth or 0.0)))

    def last(self):
        med, bw, nrm, sw = ctypes.c_float(), ctypes.c_float(), ctypes.c_double(), ctypes.c_int32()
        self.ctx.check(self.lib.stein_engine_last(self.handle, ctypes.byref(med), ctypes.byref(bw),
                                                  ctypes.byref(nrm), ctypes.byref(sw)))
        return {"median": med.value, "bandwidth": bw.value, "phi_norm": nrm.value, "sweeps": sw.value}

    def get_state(self, moments=True):
        it, lr = ctypes.c_int64(), ctypes.c_double()
        m1 = np.empty((self.n_local, self.n_params), np.float64) if moments else None
        m2 = np.empty((self.n_local, self.n_params), np.float64) if moments else None
        self.ctx.check(self.lib.stein_engine_get_state(
            self.handle, ctypes.byref(it), ctypes.byref(lr),
            m1.ctypes.data_as(ctypes.c_void_p) if moments else None,
            m2.ctypes.data_as(ctypes.c_void_p) if moments else None, 1))
        return {"n_iters": it.value, "learning_rate": lr.value, "m1": m1, "m2": m2}

    def set_state(self, n_iters, learning_rate, m1=None, m2=None):
        def p(a):
            if a is None:
                return None, None
            a = np.ascontiguousarray(np.asarray(a, np.float64))
            return a, a.ctypes.data_as(ctypes.c_void_p)
        a1, p1 = p(m1)
        a2, p2 = p(m2)
        self.ctx.check(self.lib.stein_engine_set_state(self.handle, int(n_iters), float(learning_rate),
                                                       p1, p2, 1))
