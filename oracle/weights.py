"""Oracle-side weight access (test infrastructure): plain scipy.io read of the MAT-v5 layout the
reference exports (duffing.py:61-64: W1..W4 (out,in), b1..b4 stored as 1xH rows)."""
import numpy as np
import scipy.io as sio


def load_mat_encoder(path):
    m = sio.loadmat(path)
    Ws, bs = [], []
    i = 1
    while "W%d" % i in m:
        Ws.append(np.ascontiguousarray(m["W%d" % i], dtype=np.float64))
        bs.append(np.ascontiguousarray(m["b%d" % i], dtype=np.float64).reshape(-1))
        i += 1
    return Ws, bs
