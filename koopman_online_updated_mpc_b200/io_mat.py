"""On-disk formats either side of the hot path, in the reference's own MAT-v5 layouts, so that its
MATLAB encoders and plotting scripts consume GPU results unchanged.

* `save_model_weights`   -- `model_weights.mat` as duffing.py:61-64 / vanderpol.py:59-62 write it:
  W1..WL in nn.Linear layout (out, in), b1..bL as 1 x out rows; read back by Encoder_Duffing.m:2,
  Encoder_VDP.m:2, Encoder_Tank.m:2 and by `weights.load_mat`.
* `save_nn_encoder`      -- `NN_Encoder.mat` (duffing.py:1172, vanderpol.py:1112):
  X_Collection_NO (n, T) = closed loop without update, X_Collection (n, T) = with online update,
  U_Collection (1, T); the file VDP_Revise_2/NN_Encoder.mat of the reference has this layout.
* `save_trajectory`      -- `DuffingPlot_trajectory.mat` / `VDPPlot_trajectory.mat`
  (duffing.py:344, vanderpol.py:350): snapshot set and open-loop predictor outputs.
* `closed_loop_logs`     -- one scenario of a `ClosedLoop` batch as the reference's
  logX / logU arrays ((n, T) and (1, T), duffing.py:786-800)."""
import numpy as np


def _np(a):
    try:
        import torch
        if isinstance(a, torch.Tensor):
            return a.detach().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(a)


def save_model_weights(path, Ws, bs):
    import scipy.io as sio
    d = {}
    for i, (W, b) in enumerate(zip(Ws, bs), start=1):
        W, b = _np(W).astype(np.float64), _np(b).astype(np.float64).reshape(-1)
        if W.ndim != 2 or W.shape[0] != b.shape[0]:
            raise ValueError("layer %d: W %s does not match b %s" % (i, W.shape, b.shape))
        d["W%d" % i] = W
        d["b%d" % i] = b          # scipy stores 1-D arrays as 1 x out rows, like the reference's files
    sio.savemat(path, d)


def closed_loop_logs(loop, scenario=0, T=None):
    """(logX (n, T), logU (1, T)) of one scenario of a ClosedLoop run with log_steps >= T."""
    if loop.log_x is None:
        raise ValueError("the ClosedLoop was built without log_steps")
    T = loop.step_index if T is None else T
    T = min(T, loop.log_x.shape[0])
    lx = _np(loop.log_x[:T, scenario]).T.copy()
    lu = _np(loop.log_u[:T, scenario]).reshape(1, -1).copy()
    return lx, lu


def save_nn_encoder(path, logX_no_update, logX_update, logU):
    import scipy.io as sio
    a, b, u = _np(logX_no_update), _np(logX_update), _np(logU)
    if a.shape != b.shape or u.reshape(1, -1).shape[1] != b.shape[1]:
        raise ValueError("shapes: X_Collection_NO %s, X_Collection %s, U_Collection %s" % (a.shape, b.shape, u.shape))
    sio.savemat(path, {"X_Collection_NO": a, "X_Collection": b, "U_Collection": u.reshape(1, -1)})


def save_trajectory(path, X, Y, U, test_Y, decoder_X, test_X=None, marker_X=None, h=0.05, reset_every=10):
    import scipy.io as sio
    X, Y, U, test_Y, decoder_X = (_np(v) for v in (X, Y, U, test_Y, decoder_X))
    T = test_Y.shape[1]
    tspan = np.arange(T) * h
    markers = np.arange(0, T, reset_every)
    d = {"tspan": tspan, "tspan_pred": tspan, "X": X, "Y": Y, "U": U.reshape(1, -1), "test_Y": test_Y,
         "marker_T": markers * h, "marker_originX": X[:, markers], "decoder_X": decoder_X,
         "marker_X": decoder_X[:, markers] if marker_X is None else _np(marker_X),
         "Uplot": U.reshape(-1)[:T]}
    if test_X is not None:
        d["test_X"] = _np(test_X)
    sio.savemat(path, d)
