"""Oracle: training-loss evaluation on loaded weights (test infrastructure).

duffing.py:179-235, restated literally: per window k (i = 0: k = j) the reconstruction loss of the
decoded lift, and for p = 1..pred_horizon the multi-step linear prediction `A^p psi_k + sum_s A^(p-s) B
u_{k+s-1}` (formed with matrix powers, as the reference does) against psi_{k+p} and its decoding against
x_{k+p}; Loss_lin / Loss_pred are NOT reset between windows and are divided by pred_horizon in every
iteration (l.221-222); the L1 weight term is added per window (l.226-231); Loss / batch_size at the end."""
import numpy as np

from .lift import encoder_forward


def training_losses(enc, dec, X, U, A, B, pred_horizon=30, batch=100, batch_size=100, N=100, N_Traj=100,
                    alphas=(1.0, 10.0, 50.0, 1e-6)):
    """enc, dec: (Ws, bs) tuples; X (n, M), U (1, M); returns the reference's printed quantities and
    the per-window (rec, lin, pred) sums."""
    a1, a2, a3, a4 = alphas
    PHIX = encoder_forward(enc[0], enc[1], X.T).T
    B = np.asarray(B).reshape(-1, 1)
    weight = float(sum(np.abs(W).sum() + np.abs(b).sum() for W, b in zip(enc[0] + dec[0], enc[1] + dec[1])))
    Apow = [np.linalg.matrix_power(A, p) for p in range(pred_horizon + 1)]   # the reference recomputes them
    Loss = Loss_rec = Loss_lin = Loss_pred = 0.0
    sums = []
    for j in range(batch_size):
        k = j
        if N * N_Traj - k <= pred_horizon or batch - k <= pred_horizon:
            break
        phix = PHIX[:, k]
        rec = float(np.sum((encoder_forward(dec[0], dec[1], phix) - X[:, k]) ** 2))
        lin = pred = 0.0
        for p in range(1, pred_horizon + 1):
            acc = np.zeros((A.shape[0], 1))
            for s in range(1, p + 1):
                acc = acc + Apow[p - s] @ B * U[0, k + s - 1]
            zp = (Apow[p] @ phix.reshape(-1, 1) + acc).ravel()
            lin += float(np.sum((zp - PHIX[:, k + p]) ** 2))
            pred += float(np.sum((X[:, k + p] - encoder_forward(dec[0], dec[1], zp)) ** 2))
        sums.append((rec, lin, pred))
        Loss_rec = rec
        Loss_lin = (Loss_lin + lin) / pred_horizon
        Loss_pred = (Loss_pred + pred) / pred_horizon
        Loss = Loss + a1 * Loss_rec + a2 * Loss_lin + a3 * Loss_pred + a4 * weight
    return {"Loss_rec": Loss_rec, "Loss_lin": Loss_lin, "Loss_pred": Loss_pred, "Loss": Loss / batch_size,
            "weight": weight, "window_sums": np.array(sums)}
