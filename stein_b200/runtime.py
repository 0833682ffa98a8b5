"""Process-wide handle on the C-ABI context (one per CUDA device) and small
helpers for moving NumPy data into the library's padded device layout.

PyTorch is the carrier only: it allocates device tensors, provides streams and
(optionally) torch.distributed; all arithmetic happens in libstein_b200.so.
"""
import ctypes

import numpy as np

from . import _lib

_contexts = {}


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.SteinLibraryError(
            "stein_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


class Context:
    """Owns a `stein_ctx*` bound to torch's current stream on `device`."""

    def __init__(self, device=0):
        torch = _torch()
        self.lib = _lib.load()
        self.device = int(device)
        torch.cuda.set_device(self.device)
        self.handle = ctypes.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.stein_ctx_create(ctypes.byref(self.handle), self.device,
                                             ctypes.c_void_p(stream)))
        self._comm_keepalive = None

    def check(self, rc):
        _lib.check(rc, self.handle)

    def sync_stream(self):
        """Re-bind to torch's current stream (cheap; call before enqueueing)."""
        torch = _torch()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.check(self.lib.stein_ctx_set_stream(self.handle, ctypes.c_void_p(stream)))

    def set_phi_impl(self, impl):
        self.check(self.lib.stein_ctx_set_phi_impl(self.handle, int(impl)))

    def set_median_impl(self, impl):
        self.check(self.lib.stein_ctx_set_median_impl(self.handle, int(impl)))

    def phi_route(self):
        """Route of the last guarded phi call: {"route": "fast" | "precise" | "ffma" | None, "kappa",
        "predicted_fast_error"} (include/stein_b200.h stein_ctx_phi_route)."""
        r, k, e = ctypes.c_int32(), ctypes.c_float(), ctypes.c_float()
        self.check(self.lib.stein_ctx_phi_route(self.handle, ctypes.byref(r), ctypes.byref(k), ctypes.byref(e)))
        return {"route": {0: "fast", 1: "precise", 2: "ffma"}.get(r.value), "kappa": k.value, "predicted_fast_error": e.value}

    def set_phi_guard_tol(self, tol):
        self.check(self.lib.stein_ctx_set_phi_guard_tol(self.handle, float(tol)))

    @property
    def launch_count(self):
        return int(self.lib.stein_ctx_launch_count(self.handle))

    # ---- padded device layout -------------------------------------------
    def ld(self, d):
        return int(self.lib.stein_ld(d))

    def rows_padded(self, n):
        return int(self.lib.stein_rows_padded(n))

    def to_padded(self, array, ld=None):
        """host (n x d) -> zero-padded fp32 device tensor (rows_padded x ld); ld defaults to stein_ld(d)
        (pass 128 / 256 / 512 / 768 / 1024 to make the matrix eligible for the tensor-core kernels)."""
        torch = _torch()
        a = np.ascontiguousarray(np.asarray(array, dtype=np.float32))
        if a.ndim != 2:
            raise ValueError("expected a 2-D (n_particles x n_params) array")
        n, d = a.shape
        out = torch.zeros((self.rows_padded(n), int(ld) if ld else self.ld(d)), dtype=torch.float32,
                          device="cuda:%d" % self.device)
        out[:n, :d] = torch.from_numpy(a).to(out.device)
        return out

    def to_square(self, K):
        """host (n x n) kernel matrix -> zero-padded fp32 device tensor (rows_padded x rows_padded)."""
        torch = _torch()
        a = np.ascontiguousarray(np.asarray(K, dtype=np.float32))
        if a.ndim != 2 or a.shape[0] != a.shape[1]:
            raise ValueError("expected a square (n_particles x n_particles) kernel matrix")
        rows = self.rows_padded(a.shape[0])
        out = torch.zeros((rows, rows), dtype=torch.float32, device="cuda:%d" % self.device)
        out[:a.shape[0], :a.shape[0]] = torch.from_numpy(a).to(out.device)
        return out

    def dense(self, array, dtype=np.float32):
        """host array -> contiguous device tensor, no padding (data matrices)."""
        torch = _torch()
        a = np.ascontiguousarray(np.asarray(array, dtype=dtype))
        return torch.from_numpy(a).to("cuda:%d" % self.device)


def context(device=None):
    torch = _torch()
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    if device not in _contexts:
        _contexts[device] = Context(device)
    ctx = _contexts[device]
    ctx.sync_stream()
    return ctx


def ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr())
