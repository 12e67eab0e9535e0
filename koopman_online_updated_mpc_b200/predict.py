"""Open-loop multi-step predictor: how the reference judges an EDMD model (duffing.py:290-343,
vanderpol.py:292-348).  Along `plot_time` consecutive snapshots the lifted state is re-encoded from
the true state every 10 steps and propagated with z+ = A z + B u in between; the read-out C z and
the lifted state are logged before the propagation; RMSE of one read-out row."""
import torch

from . import _lib
from ._tensors import like_input, ptr, stream_ptr, to_dev


def open_loop_predict(psi, X, U, A, B, C, plot_time, reset_every=10, rmse_row=0, n_seq=1, seq_stride=None):
    """Reference shapes: psi (nz, M) = lifted snapshots PHIX, X (n, M), U (1, M), A (nz, nz),
    B (nz, 1), C (n, nz) -> test_Y (n, plot_time), decoder_X (nz, plot_time), RMSE (float) for
    n_seq = 1 (the reference's case); for n_seq > 1 sequences starting every `seq_stride` snapshots:
    test_Y (n_seq, n, plot_time), decoder_X (n_seq, nz, plot_time), RMSE (n_seq,)."""
    psi_d = to_dev(psi).t().contiguous()     # snapshot-major (M, nz)
    x_d = to_dev(X).t().contiguous()
    u_d = to_dev(U).reshape(-1)
    A_d, B_d, C_d = to_dev(A), to_dev(B).reshape(-1), to_dev(C)
    M, nz = psi_d.shape
    n = x_d.shape[1]
    T = int(plot_time)
    stride = T if seq_stride is None else int(seq_stride)
    if (n_seq - 1) * stride + T > M:
        raise ValueError("%d sequences of %d steps at stride %d do not fit %d snapshots" % (n_seq, T, stride, M))
    dec = torch.empty((n_seq, T, nz), dtype=torch.float64, device=psi_d.device)
    ty = torch.empty((n_seq, T, n), dtype=torch.float64, device=psi_d.device)
    rm = torch.empty(n_seq, dtype=torch.float64, device=psi_d.device)
    _lib.check(_lib.lib().kmpc_open_loop_predict(ptr(psi_d), ptr(x_d), ptr(u_d), ptr(A_d), ptr(B_d), ptr(C_d),
                                                 nz, n, n_seq, T, stride, int(reset_every), int(rmse_row),
                                                 ptr(dec), ptr(ty), ptr(rm), stream_ptr()))
    ty, dec = ty.transpose(1, 2), dec.transpose(1, 2)
    if n_seq == 1:
        return like_input(ty[0].contiguous(), psi), like_input(dec[0].contiguous(), psi), float(rm.item())
    return like_input(ty.contiguous(), psi), like_input(dec.contiguous(), psi), like_input(rm, psi)
