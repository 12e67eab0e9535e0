#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING the reference scripts.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py [--only duffing,vanderpol,duffing_rbf,vanderpol_rbf,weights,vdp_mat]

What it does ("4-point shim"):
  * copies the reference script text into a scratch dir, truncated right after the two
    closed loops (the rest is plotting), with `maxStep` shortened;
  * prepends a prologue that stubs `matplotlib` / `lmi_sdp` (not installed here) and makes
    `torch.load` pass `weights_only=False` (the .pkl is a full-module pickle);
  * rewrites `+ u])` -> `+ np.ravel(u)])` in the plant lambdas (NumPy >= 1.24 rejects the
    ragged list the original builds);
  * appends an epilogue that dumps the variables we pin the oracle against;
  * runs it as `__main__` in the scratch dir (so `__main__.AutoEncoder` resolves).
No reference source is copied into this repository: only numeric outputs are kept.

Also written:
  * weights/*.mat  -- the encoder weights re-exported with the reference's own
    `model_weights.mat` recipe (duffing.py:61-64: W1..W4, b1..b4), so tests/bench on the GPU
    box (no /root/reference there) load the same numbers through the same MAT-v5 path.
  * vdp_nn_encoder_head.npz -- the first 600 columns of VDP_Revise_2/NN_Encoder.mat, the only
    golden vector the reference itself ships (written by vanderpol.py:1112).
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

PROLOGUE = r'''
import sys as _sys
from unittest import mock as _mock
for _m in ("matplotlib", "matplotlib.pyplot", "lmi_sdp"):
    _sys.modules[_m] = _mock.MagicMock()
import torch as _torch
_orig_load = _torch.load
def _load(f, *a, **k):
    k.setdefault("weights_only", False)
    return _orig_load(f, *a, **k)
_torch.load = _load
'''

EPILOGUE_ENC = r'''
import numpy as _np
def _n(v):
    try:
        return v.detach().numpy()
    except AttributeError:
        return _np.asarray(v)
_np.savez(_OUT,
    A=_n(A), B=_n(B), C=_n(C),
    X_head=_n(X)[:, :10000][:, :300], PHIX_head=_n(PHIX)[:, :300], PHIY_head=_n(PHIY)[:, :300],
    U_head=_n(inputs_u)[:, :300],
    X_sum=_n(inputs_x).sum(axis=1), U_sum=_n(inputs_u).sum(axis=1), PHIX_sum=_n(PHIX).sum(axis=1),
    logX=logX, logU=logU, logXloc=logXloc, logUloc=logUloc,
    logXlift=logXlift, logXLOClift=logXLOClift,
    Aloc=_n(Aloc_d), Bloc=_n(Bloc_d), Cloc=_n(Cloc_d),
    K_A=_n(K_A), inv_K_G=_n(inv_K_G), bar_X=_n(bar_X), bar_Q=_n(bar_Q),
    A_error=_np.asarray(A_error), B_error=_np.asarray(B_error), C_error=_np.asarray(C_error),
    maxStep=maxStep)
'''

EPILOGUE_RBF = r'''
import numpy as _np
_np.savez(_OUT,
    A=_np.asarray(A), B=_np.asarray(B), C=_np.asarray(C), cx=_np.asarray(cx),
    X_head=_np.asarray(X)[:, :300], PHIX_head=_np.asarray(PHIX)[:, :300],
    PHIY_head=_np.asarray(PHIY)[:, :300],
    PHIX_sum=_np.asarray(PHIX).sum(axis=1),
    logX=logX, logU=logU, logXloc=logXloc, logUloc=logUloc,
    logXlift=logXlift, logXLOClift=logXLOClift,
    Aloc=_np.asarray(Aloc_d), Bloc=_np.asarray(Bloc_d), Cloc=_np.asarray(Cloc_d),
    A_error=_np.asarray(A_error), B_error=_np.asarray(B_error), C_error=_np.asarray(C_error),
    maxStep=maxStep)
'''

EPILOGUE_PRED = r'''
import numpy as _np
def _n(v):
    try:
        return v.detach().numpy()
    except AttributeError:
        return _np.asarray(v)
# test snapshot set of the open-loop check (np.random.seed(33 / 50) + data_generate) and the
# predictor outputs of duffing.py:272-343 / vanderpol.py:272-348
_np.savez(_OUT,
    A=_n(A), B=_n(B), C=_n(C),
    X_head=_n(X)[:, :600], Y_head=_n(Y)[:, :600], U_head=_n(U)[:, :600],
    X_sum=_n(X).sum(axis=1), Y_sum=_n(Y).sum(axis=1), U_sum=_n(U).sum(axis=1),
    X_tail=_n(X)[:, -300:], Y_tail=_n(Y)[:, -300:], U_tail=_n(U)[:, -300:],
    n_snap=_n(X).shape[1],
    test_Y=_n(test_Y), decoder_X=_n(decoder_X), marker_X=_n(marker_X), test_X=_n(test_X),
    RMSE=_np.asarray(RMSE), plotTime=plotTime)
'''

EPILOGUE_LOSS = r'''
import numpy as _np
import scipy.io as _sio
def _n(v):
    try:
        return v.detach().numpy()
    except AttributeError:
        return _np.asarray(v)
# training-loss evaluation of duffing.py:179-235 on the loaded weights (rec / multi-step lin / pred,
# the reference's own accumulation order) + the Decoder half of the auto-encoder it needs
_w = net.state_dict()
_sio.savemat(_DEC, {"W%d" % (i + 1): _w["Decoder.%d.weight" % (2 * i)].numpy() for i in range(4)}
             | {"b%d" % (i + 1): _w["Decoder.%d.bias" % (2 * i)].numpy() for i in range(4)})
_np.savez(_OUT, Loss_rec=_n(Loss_rec), Loss_lin=_n(Loss_lin), Loss_pred=_n(Loss_pred), Loss=_n(Loss),
          weight=_n(weight), A=_n(A), B=_n(B), pred_horizon=pred_horizon, batch=batch, batch_size=batch_size,
          alphas=_np.array([alpha_1, alpha_2, alpha_3, alpha_4]), last_j=j)
'''

JOBS = {
    # name: (script, truncate-after-line (1-based, inclusive), steps, epilogue, needed files)
    "duffing": ("duffing.py", 1013, 300, EPILOGUE_ENC, ["AutoEncoder_20220418_duffing_2.pkl"]),
    "vanderpol": ("vanderpol.py", 952, 400, EPILOGUE_ENC, ["AutoEncoder_20220414_4.pkl"]),
    "duffing_rbf": ("duffing_RBF.py", 527, 120, EPILOGUE_RBF, []),
    "vanderpol_rbf": ("vanderpol_RBF.py", 527, 120, EPILOGUE_RBF, []),
    # open-loop multi-step predictor + its test snapshot set (no closed loop: steps = None)
    "duffing_predict": ("duffing.py", 343, None, EPILOGUE_PRED, ["AutoEncoder_20220418_duffing_2.pkl"]),
    "vanderpol_predict": ("vanderpol.py", 348, None, EPILOGUE_PRED, ["AutoEncoder_20220414_4.pkl"]),
    # training-loss evaluation on the loaded weights (duffing.py:179-235), seed-101 snapshot set
    "duffing_losses": ("duffing.py", 235, None, EPILOGUE_LOSS, ["AutoEncoder_20220418_duffing_2.pkl"]),
}


def run_script(name):
    script, cut, steps, epilogue, files = JOBS[name]
    work = tempfile.mkdtemp(prefix="kmpc_golden_")
    try:
        shutil.copy(os.path.join(REF, "data_generate.py"), work)
        for f in files:
            shutil.copy(os.path.join(REF, f), work)
        with open(os.path.join(REF, script), encoding="utf-8") as fh:
            lines = fh.read().split("\n")[:cut]
        text = "\n".join(lines)
        text = text.replace("+ u])", "+ np.ravel(u)])")
        if steps is not None:
            text, nsub = re.subn(r"^maxStep = 10000$", "maxStep = %d" % steps, text, flags=re.M)
            assert nsub == 1, "maxStep patch point not found"
        out = os.path.join(HERE, "ref_%s.npz" % name)
        dec = os.path.join(HERE, "weights", "duffing_decoder_weights.mat")
        body = PROLOGUE + text + "\n_OUT = %r\n_DEC = %r\n" % (out, dec) + epilogue
        path = os.path.join(work, "run_" + script)
        with open(path, "w", encoding="utf-8") as fh:
            fh.write(body)
        env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
        with open(os.devnull, "w") as devnull:
            subprocess.run([sys.executable, path], cwd=work, check=True, stdout=devnull, env=env)
        print("wrote", out)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def export_weights():
    """Re-export encoder weights exactly as duffing.py:61-64 does (model_weights.mat)."""
    import scipy.io as sio
    import torch
    import torch.nn as nn

    class AutoEncoder(nn.Module):  # shape-only stand-in so the full-module pickle resolves
        def __init__(self):
            super().__init__()

    sys.modules["__main__"].AutoEncoder = AutoEncoder
    wdir = os.path.join(HERE, "weights")
    os.makedirs(wdir, exist_ok=True)
    for pkl, out in (("AutoEncoder_20220418_duffing_2.pkl", "duffing_model_weights.mat"),
                     ("AutoEncoder_20220414_4.pkl", "vdp_model_weights.mat")):
        net = torch.load(os.path.join(REF, pkl), weights_only=False)
        w = net.state_dict()
        sio.savemat(os.path.join(wdir, out), {
            "W1": w["Encoder.0.weight"].numpy(), "W2": w["Encoder.2.weight"].numpy(),
            "W3": w["Encoder.4.weight"].numpy(), "W4": w["Encoder.6.weight"].numpy(),
            "b1": w["Encoder.0.bias"].numpy(), "b2": w["Encoder.2.bias"].numpy(),
            "b3": w["Encoder.4.bias"].numpy(), "b4": w["Encoder.6.bias"].numpy()})
        ref_mat = {"duffing_model_weights.mat": "Revise_2/duffing_weights.mat",
                   "vdp_model_weights.mat": "VDP_Revise_2/Good_VDP.mat"}[out]
        m = sio.loadmat(os.path.join(REF, ref_mat))
        for k in ("W1", "W2", "W3", "W4"):
            assert np.array_equal(m[k], w["Encoder.%d.weight" % (2 * (int(k[1]) - 1))].numpy()), (pkl, k)
        print("wrote", out, "(bit-identical to", ref_mat + ")")
    # Tank: 3-layer encoder, only exists as .mat (Weights/Tank_New.mat, Encoder_Tank.m:2-5)
    m = sio.loadmat(os.path.join(REF, "Weights/Tank_New.mat"))
    sio.savemat(os.path.join(wdir, "tank_model_weights.mat"),
                {k: m[k] for k in ("W1", "W2", "W3", "b1", "b2", "b3")})
    print("wrote tank_model_weights.mat")


def export_vdp_mat():
    import scipy.io as sio
    m = sio.loadmat(os.path.join(REF, "VDP_Revise_2/NN_Encoder.mat"))
    n = 600
    np.savez(os.path.join(HERE, "vdp_nn_encoder_head.npz"),
             X_Collection_NO=m["X_Collection_NO"][:, :n], X_Collection=m["X_Collection"][:, :n],
             U_Collection=m["U_Collection"][:, :n],
             X_Collection_tail=m["X_Collection"][:, -5:], X_Collection_NO_tail=m["X_Collection_NO"][:, -5:])
    print("wrote vdp_nn_encoder_head.npz")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="weights,vdp_mat,duffing,vanderpol,duffing_rbf,vanderpol_rbf,"
                                       "duffing_predict,vanderpol_predict")
    args = ap.parse_args()
    for name in args.only.split(","):
        if name == "weights":
            export_weights()
        elif name == "vdp_mat":
            export_vdp_mat()
        else:
            run_script(name)


if __name__ == "__main__":
    main()
