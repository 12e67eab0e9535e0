"""Weight loaders for the theta_E encoder -- the same files the reference loads.

* `.pkl`  : full-module torch pickle written by `torch.save(net, ...)`
            (DeepLearning_KoopmanControl_Approach3.py:565; read by duffing.py:57,
            vanderpol.py:57).  The pickle references `__main__.AutoEncoder`; we register a
            structural stand-in so it resolves without the reference's script.
* `.mat`  : MAT-v5 with W1..WL (out,in) and b1..bL stored as 1xH rows (duffing.py:61-64
            `model_weights.mat`, Revise_2/duffing_weights.mat, Weights/Tank_New.mat; read by
            Encoder_Duffing.m:2, Encoder_Tank.m:2).
* `.npz`  : W1.., b1.. arrays.
Returns (Ws, bs): lists of float64 numpy arrays in nn.Linear layout."""
import sys

import numpy as np


def _from_dict(m):
    Ws, bs = [], []
    i = 1
    while ("W%d" % i) in m:
        Ws.append(np.ascontiguousarray(np.asarray(m["W%d" % i], dtype=np.float64)))
        bs.append(np.ascontiguousarray(np.asarray(m["b%d" % i], dtype=np.float64).reshape(-1)))
        i += 1
    if not Ws:
        raise ValueError("no W1/b1 entries found")
    for W, b in zip(Ws, bs):
        if W.shape[0] != b.shape[0]:
            raise ValueError("weight/bias shape mismatch: %s vs %s" % (W.shape, b.shape))
    return Ws, bs


def load_mat(path):
    import scipy.io as sio
    return _from_dict(sio.loadmat(path))


def load_npz(path):
    return _from_dict(np.load(path))


def load_pkl(path):
    """Full-module pickle: needs weights_only=False and an importable `__main__.AutoEncoder`."""
    import torch
    import torch.nn as nn

    main = sys.modules["__main__"]
    injected = False
    if not hasattr(main, "AutoEncoder"):
        class AutoEncoder(nn.Module):  # structure comes from the pickle itself
            def __init__(self):
                super().__init__()
        AutoEncoder.__module__ = "__main__"
        main.AutoEncoder = AutoEncoder
        injected = True
    try:
        net = torch.load(path, map_location="cpu", weights_only=False)
    finally:
        if injected:
            del main.AutoEncoder
    sd = net.state_dict() if hasattr(net, "state_dict") else net
    keys = sorted((k for k in sd if k.startswith("Encoder.") and k.endswith(".weight")),
                  key=lambda k: int(k.split(".")[1]))
    if not keys:
        raise ValueError("no Encoder.*.weight entries in %s" % path)
    Ws = [sd[k].detach().to(torch.float64).numpy().copy() for k in keys]
    bs = [sd[k.replace(".weight", ".bias")].detach().to(torch.float64).numpy().copy() for k in keys]
    return Ws, bs


def load_encoder_weights(path):
    p = str(path).lower()
    if p.endswith(".mat"):
        return load_mat(path)
    if p.endswith(".npz"):
        return load_npz(path)
    if p.endswith(".pkl") or p.endswith(".pt") or p.endswith(".pth"):
        return load_pkl(path)
    raise ValueError("unknown weight file type: %s" % path)


def save_model_weights_mat(path, Ws, bs):
    """Write the reference's `model_weights.mat` layout (duffing.py:61-64)."""
    import scipy.io as sio
    d = {}
    for i, (W, b) in enumerate(zip(Ws, bs), start=1):
        d["W%d" % i] = np.asarray(W)
        d["b%d" % i] = np.asarray(b).reshape(-1)
    sio.savemat(path, d)
