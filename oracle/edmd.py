"""Oracle: EDMD regression of the lifted A, B, C (test infrastructure).

python form : duffing.py:167-177   K = PHIY @ pinv([PHIX; U]) ; C = X @ pinv(PHIX)
gram  form  : Tank_System.m:93-100 M = (W V') pinv(V V'), W = [Ylift; X], V = [Xlift; U]
              -> [A B; C 0]  (C here comes from the JOINT regression: not the python C)
Gram pack layout shared with the CUDA path (include/kmpc.h):
    G  = V V'  (nv x nv),  Aq = PHIY V' (nz x nv),  XV = X V' (n x nv),  count
"""
import numpy as np

C_PYTHON, C_JOINT = 0, 1


def edmd_pinv(PHIX, PHIY, U, X):
    V = np.concatenate([PHIX, U], axis=0)
    K = PHIY @ np.linalg.pinv(V)
    nz = PHIX.shape[0]
    return K[:, :nz], K[:, nz:], X @ np.linalg.pinv(PHIX)


def gram_pack(PHIX, PHIY, U, X):
    V = np.concatenate([PHIX, U], axis=0)
    return V @ V.T, PHIY @ V.T, X @ V.T


def edmd_from_gram(G, Aq, XV, nz, c_variant=C_PYTHON):
    K = Aq @ np.linalg.pinv(G)
    A, B = K[:, :nz], K[:, nz:]
    if c_variant == C_PYTHON:
        C = XV[:, :nz] @ np.linalg.pinv(G[:nz, :nz])
    else:
        C = (XV @ np.linalg.pinv(G))[:, :nz]
    return A, B, C
