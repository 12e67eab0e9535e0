"""Batched evaluation of the auto-encoder training losses on loaded weights.  Drop-in for
duffing.py:179-235 (the same loop body trains the network in
DeepLearning_KoopmanControl_Approach3.py:462-563): for every window start k

    Loss_rec  = || Decoder(psi_k) - x_k ||^2
    Loss_lin  = sum_{p=1..H} || A^p psi_k + sum_{s=1..p} A^(p-s) B u_{k+s-1} - psi_{k+p} ||^2
    Loss_pred = sum_{p=1..H} || x_{k+p} - Decoder(A^p psi_k + ...) ||^2          (H = pred_horizon = 30)

(`criterion = nn.MSELoss(reduction='sum')`, duffing.py:70).  The reference evaluates the O(H^2)
`matrix_power` sums one window at a time; here all windows go through four kernels: theta_E lift of the
snapshot set, one linear rollout per window (kmpc_open_loop_predict, stride-1 windows), ONE Decoder pass
over all W x (H + 1) lifted states, and the per-window sums (kmpc_window_losses).  The reference's own
accumulation over windows -- Loss_lin and Loss_pred are NOT reset between windows and are divided by H
in every iteration (l.221-222), the L1 weight term is added once per window (l.226-231) -- is host
arithmetic on the W x 3 sums and is reproduced literally by `reference_totals`."""
import numpy as np
import torch

from . import _lib
from ._tensors import ptr, stream_ptr, to_dev
from .lift import Encoder


def window_losses(encoder, decoder, X, U, A, B, pred_horizon=30, n_windows=None, k0=0, stride=1, psi=None):
    """X (n, M) / U (1, M) in the reference's layout (or snapshot-major CUDA tensors (M, n) / (M,)),
    A (nz, nz), B (nz, 1).  Returns a CUDA tensor (W, 3): rec, lin, pred of every window."""
    if isinstance(X, np.ndarray):
        X, U = X.T.copy(), np.asarray(U).reshape(-1)
    x_d, u_d = to_dev(X), to_dev(U).reshape(-1)
    M, n = x_d.shape
    H = int(pred_horizon)
    T = H + 1
    if n_windows is None:
        n_windows = (M - k0 - T) // stride + 1
    W = int(n_windows)
    if W < 1 or k0 + (W - 1) * stride + T > M:
        raise ValueError("%d windows of %d steps from %d at stride %d do not fit %d snapshots" % (W, T, k0, stride, M))
    psi_d = encoder(x_d) if psi is None else to_dev(psi)
    nz = psi_d.shape[1]
    A_d, B_d = to_dev(A).reshape(nz, nz), to_dev(B).reshape(nz)
    L = _lib.lib()
    dev = x_d.device
    zpred = torch.empty((W, T, nz), dtype=torch.float64, device=dev)
    ty = torch.empty((W, T, n), dtype=torch.float64, device=dev)      # C z read-out of the predictor: unused here
    Cz = torch.zeros((n, nz), dtype=torch.float64, device=dev)
    _lib.check(L.kmpc_open_loop_predict(ptr(psi_d[k0:]), None, ptr(u_d[k0:]), ptr(A_d), ptr(B_d), ptr(Cz), nz, n, W, T,
                                        int(stride), T, 0, ptr(zpred), ptr(ty), None, stream_ptr()))
    xdec = decoder(zpred.reshape(W * T, nz)).reshape(W, T, n)
    out = torch.empty((W, 3), dtype=torch.float64, device=dev)
    _lib.check(L.kmpc_window_losses(ptr(psi_d), ptr(x_d), ptr(zpred), ptr(xdec.contiguous()), nz, n, W, T, int(k0),
                                    int(stride), ptr(out), stream_ptr()))
    return out


def l1_weight(*nets):
    """sum |param| over the networks' parameters (duffing.py:226-228)."""
    return float(sum(np.abs(W).sum() + np.abs(b).sum() for net in nets for W, b in zip(net.Ws, net.bs)))


def reference_totals(sums, weight, pred_horizon=30, alphas=(1.0, 10.0, 50.0, 1e-6), batch_size=100):
    """The reference's accumulation over windows j = 0, 1, ... (duffing.py:179-235), literally:
    Loss_lin / Loss_pred carry over from window to window and are divided by H each time."""
    sums = np.asarray(sums.cpu() if hasattr(sums, "cpu") else sums, dtype=np.float64)
    a1, a2, a3, a4 = alphas
    Loss = Loss_rec = Loss_lin = Loss_pred = 0.0
    for rec, lin, pred in sums:
        Loss_rec = rec
        Loss_lin = (Loss_lin + lin) / pred_horizon
        Loss_pred = (Loss_pred + pred) / pred_horizon
        Loss = Loss + a1 * Loss_rec + a2 * Loss_lin + a3 * Loss_pred + a4 * weight
    return {"Loss_rec": Loss_rec, "Loss_lin": Loss_lin, "Loss_pred": Loss_pred, "Loss": Loss / batch_size,
            "weight": weight}


def training_losses(encoder, decoder, X, U, A, B, pred_horizon=30, batch=100, batch_size=100,
                    alphas=(1.0, 10.0, 50.0, 1e-6)):
    """duffing.py:179-235 with i = 0: windows k = 0 .. until `batch - k <= pred_horizon` (l.185-186)."""
    n_windows = min(batch_size, max(batch - pred_horizon, 0))
    sums = window_losses(encoder, decoder, X, U, A, B, pred_horizon, n_windows)
    out = reference_totals(sums, l1_weight(encoder, decoder), pred_horizon, alphas, batch_size)
    out["window_sums"] = sums
    return out


def load_decoder(path):
    """Decoder half as an `Encoder`-type handle (MAT file with W1.., b1.. in nn.Linear layout)."""
    return Encoder.from_file(path)
