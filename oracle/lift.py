"""Oracle: lifting functions (test infrastructure, see oracle/__init__.py).

theta_E encoder  : duffing.py:21-29 (nn.Sequential Linear/ReLU chain, no final activation),
                   call sites duffing.py:153,155,764,847,884; MATLAB Encoder_Duffing.m:3-6,
                   Encoder_VDP.m:3-6 (4 layers), Encoder_Tank.m:3-5 (3 layers).
thin-plate RBF   : duffing_RBF.py:20-23 (python variant), rbf.m:24-29 (MATLAB variant).
"""
import numpy as np


def encoder_forward(Ws, bs, x):
    """z = W_L(...relu(W_1 x + b_1)...) + b_L.  x: (S, n) or (n,) -> (S, nz) or (nz,)."""
    h = np.asarray(x, dtype=np.float64)
    for i, (W, b) in enumerate(zip(Ws, bs)):
        h = h @ np.asarray(W).T + np.asarray(b).reshape(-1)
        if i < len(Ws) - 1:
            h = np.maximum(h, 0.0)
    return h


LIFT_RAW, LIFT_OFFSET, LIFT_STACK = 0, 1, 2


def lift_mlp(Ws, bs, x, mode=LIFT_RAW):
    """mode RAW    : theta(x)                               (duffing.py:764, vanderpol.py:760)
       mode OFFSET : theta(x) - theta(0)                    (Koopman_update_Tracking_Lift.m:65)
       mode STACK  : [x; theta(x)] - [0; theta(0)]          (Koopman_update.m:67)"""
    x = np.asarray(x, dtype=np.float64)
    z = encoder_forward(Ws, bs, x)
    if mode == LIFT_RAW:
        return z
    z0 = encoder_forward(Ws, bs, np.zeros(x.shape[-1]))
    z = z - z0
    if mode == LIFT_OFFSET:
        return z
    return np.concatenate([x, z], axis=-1)


RBF_PYTHON, RBF_MATLAB = 0, 1


def rbf_lift(x, cx, variant=RBF_PYTHON):
    """x: (S, n) or (n,), cx: (nz, n) centres -> (S, nz) or (nz,).
    python : d^2 * log(d + 1e-4), d = ||x - c||            (duffing_RBF.py:20-23; sklearn's
             euclidean_distances uses the |x|^2-2xc+|c|^2 expansion, we use the direct form)
    matlab : r2 * log(sqrt(r2)), NaN -> 0                  (rbf.m:24-29)"""
    x = np.asarray(x, dtype=np.float64)
    single = x.ndim == 1
    X = np.atleast_2d(x)
    diff = X[:, None, :] - np.asarray(cx)[None, :, :]
    r2 = np.sum(diff * diff, axis=2)
    if variant == RBF_PYTHON:
        d = np.sqrt(r2)
        out = r2 * np.log(d + 1e-4)
    else:
        with np.errstate(divide="ignore", invalid="ignore"):
            out = r2 * np.log(np.sqrt(r2))
        out[np.isnan(out)] = 0.0
    return out[0] if single else out
