"""Stage 1 -- lifting.  Drop-ins for `net.Encoder(x)` (duffing.py:764,847,884) and
`rbf(x, cx)` (duffing_RBF.py:20-23), batched over a leading scenario axis."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._tensors import like_input, ptr, stream_ptr, to_dev
from .weights import load_encoder_weights

LIFT_RAW, LIFT_OFFSET, LIFT_STACK = 0, 1, 2
RBF_PYTHON, RBF_MATLAB = 0, 1
PREC_FP64, PREC_TC = 0, 1   # include/kmpc.h KMPC_PREC_*: fp64 DMMA path | tcgen05 bf16x3 split path (~1e-7)


class Encoder:
    """theta_E on the GPU.  `enc(x)`: x (S, n) or (n,) -> (S, nz) or (nz,); numpy in -> numpy out,
    CPU tensor in -> CPU tensor out (host<->device copies included), CUDA tensor in -> CUDA
    tensor out with no copies.  mode: LIFT_RAW | LIFT_OFFSET | LIFT_STACK."""

    def __init__(self, Ws, bs, mode=LIFT_RAW):
        L = _lib.lib()
        from ._tensors import require_cuda
        require_cuda()
        self.Ws = [np.ascontiguousarray(W, dtype=np.float64) for W in Ws]
        self.bs = [np.ascontiguousarray(b, dtype=np.float64).reshape(-1) for b in bs]
        self.mode = mode
        nl = len(self.Ws)
        dims = [self.Ws[0].shape[1]] + [W.shape[0] for W in self.Ws]
        for l in range(1, nl):
            if self.Ws[l].shape[1] != dims[l]:
                raise ValueError("layer %d input width mismatch" % l)
        self.dims = dims
        Wp = (ctypes.c_void_p * nl)(*[W.ctypes.data for W in self.Ws])
        bp = (ctypes.c_void_p * nl)(*[b.ctypes.data for b in self.bs])
        dp = (ctypes.c_int * (nl + 1))(*dims)
        h = ctypes.c_void_p()
        _lib.check(L.kmpc_encoder_create(ctypes.byref(h), Wp, bp, dp, nl, stream_ptr()))
        self._h = h
        self.n = dims[0]
        self.nz = dims[-1]

    @classmethod
    def from_file(cls, path, mode=LIFT_RAW):
        Ws, bs = load_encoder_weights(path)
        return cls(Ws, bs, mode)

    @property
    def handle(self):
        return self._h

    def out_dim(self, mode=None):
        return int(_lib.lib().kmpc_encoder_out_dim(self._h, self.mode if mode is None else mode))

    @property
    def has_tc(self):
        """True when the net fits the tcgen05 split-precision kernel (PREC_TC)."""
        return bool(_lib.lib().kmpc_encoder_has_tc(self._h))

    def encode_into(self, x_dev, z_dev, mode=None, precision=PREC_FP64):
        """Zero-copy form: x_dev (S, n), z_dev (S, out_dim) CUDA float64 tensors."""
        S = x_dev.shape[0]
        _lib.check(_lib.lib().kmpc_encode_ex(self._h, ptr(x_dev), ptr(z_dev), S,
                                             self.mode if mode is None else mode, int(precision), stream_ptr()))
        return z_dev

    def __call__(self, x, mode=None, precision=PREC_FP64):
        single = (x.ndim == 1)
        xd = to_dev(x).reshape(-1, self.n)
        z = torch.empty((xd.shape[0], self.out_dim(mode)), dtype=torch.float64, device=xd.device)
        self.encode_into(xd, z, mode, precision)
        if single:
            z = z[0]
        return like_input(z, x)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _lib.lib().kmpc_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rbf(X, cx, variant=RBF_PYTHON):
    """Thin-plate RBF lift.  Reference call shape `rbf(x, cx)` with x (n,) -> (nz, 1)
    (duffing_RBF.py:20-23); batched: X (S, n) -> (S, nz)."""
    single = (X.ndim == 1)
    cxd = to_dev(cx)
    nz, n = cxd.shape
    xd = to_dev(X).reshape(-1, n)
    z = torch.empty((xd.shape[0], nz), dtype=torch.float64, device=xd.device)
    _lib.check(_lib.lib().kmpc_rbf_lift(ptr(xd), ptr(cxd), ptr(z), xd.shape[0], n, nz, variant, stream_ptr()))
    if single:
        z = z.reshape(nz, 1)
    return like_input(z, X)
