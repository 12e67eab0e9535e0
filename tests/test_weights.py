"""Weight loaders: the same .pkl / .mat files the reference reads."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

import helpers as H
from koopman_online_updated_mpc_b200 import weights


def test_mat_loader_on_reference_export():
    Ws, bs = weights.load_encoder_weights(H.weights_path("duffing"))
    assert [W.shape for W in Ws] == [(100, 2), (100, 100), (100, 100), (8, 100)]
    assert [b.shape for b in bs] == [(100,), (100,), (100,), (8,)]
    Wt, bt = weights.load_encoder_weights(H.weights_path("tank"))
    assert [W.shape for W in Wt] == [(100, 2), (100, 100), (10, 100)]   # Encoder_Tank.m: 3 layers
    oW, ob = H.oracle_weights("duffing")
    assert all(np.array_equal(a, b) for a, b in zip(Ws, oW))


class AutoEncoder(nn.Module):  # same attribute layout as duffing.py:17-44
    def __init__(self):
        super().__init__()
        self.Encoder = nn.Sequential(nn.Linear(2, 16), nn.ReLU(), nn.Linear(16, 16), nn.ReLU(), nn.Linear(16, 8))
        self.Decoder = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 2))


def test_full_module_pickle_roundtrip(tmp_path):
    """`torch.save(net, ...)` full-module pickle (DeepLearning_KoopmanControl_Approach3.py:565)."""
    torch.manual_seed(0)
    net = AutoEncoder().double()
    AutoEncoder.__module__ = "__main__"
    sys.modules["__main__"].AutoEncoder = AutoEncoder
    try:
        path = str(tmp_path / "AutoEncoder_test.pkl")
        torch.save(net, path)
    finally:
        del sys.modules["__main__"].AutoEncoder
    Ws, bs = weights.load_encoder_weights(path)   # resolves __main__.AutoEncoder by itself
    sd = net.state_dict()
    assert len(Ws) == 3
    np.testing.assert_array_equal(Ws[2], sd["Encoder.4.weight"].numpy())
    np.testing.assert_array_equal(bs[0], sd["Encoder.0.bias"].numpy())
    out = str(tmp_path / "model_weights.mat")
    weights.save_model_weights_mat(out, Ws, bs)  # duffing.py:61-64 layout
    W2, b2 = weights.load_encoder_weights(out)
    assert all(np.array_equal(a, b) for a, b in zip(Ws, W2)) and all(np.array_equal(a, b) for a, b in zip(bs, b2))


@pytest.mark.skipif(not os.path.exists(H.REF), reason="reference checkout only exists in the build container")
def test_reference_pkl_equals_reference_mat():
    """the .mat files are bit-identical to the Encoder halves of the .pkl files."""
    for pkl, mat in (("AutoEncoder_20220418_duffing_2.pkl", "Revise_2/duffing_weights.mat"),
                     ("AutoEncoder_20220414_4.pkl", "VDP_Revise_2/Good_VDP.mat")):
        Wp, bp = weights.load_encoder_weights(os.path.join(H.REF, pkl))
        Wm, bm = weights.load_encoder_weights(os.path.join(H.REF, mat))
        assert all(np.array_equal(a, b) for a, b in zip(Wp, Wm)) and all(np.array_equal(a, b) for a, b in zip(bp, bm))
    Wg, _ = weights.load_encoder_weights(H.weights_path("duffing"))
    Wp, _ = weights.load_encoder_weights(os.path.join(H.REF, "AutoEncoder_20220418_duffing_2.pkl"))
    assert all(np.array_equal(a, b) for a, b in zip(Wg, Wp))


def test_mat_writers_use_the_reference_layouts(tmp_path):
    """io_mat writes model_weights.mat / NN_Encoder.mat / *_trajectory.mat in the layouts of
    duffing.py:61-64, 1172 and 344 (the reference's own NN_Encoder.mat head is the layout fixture)."""
    import scipy.io as sio
    from koopman_online_updated_mpc_b200 import io_mat, weights as W
    Ws, bs = H.oracle_weights("duffing")
    p = tmp_path / "model_weights.mat"
    io_mat.save_model_weights(p, Ws, bs)
    m = sio.loadmat(p)
    ref = sio.loadmat(H.weights_path("duffing"))          # written with the reference's own recipe
    for k in ("W1", "W2", "W3", "W4", "b1", "b2", "b3", "b4"):
        assert m[k].shape == ref[k].shape and np.array_equal(m[k], ref[k])
    Ws2, bs2 = W.load_mat(p)
    assert all(np.array_equal(a, b) for a, b in zip(Ws, Ws2)) and all(np.array_equal(a, b) for a, b in zip(bs, bs2))
    g = H.golden("vdp_nn_encoder_head.npz")
    q = tmp_path / "NN_Encoder.mat"
    io_mat.save_nn_encoder(q, g["X_Collection_NO"], g["X_Collection"], g["U_Collection"])
    m = sio.loadmat(q)
    for k in ("X_Collection_NO", "X_Collection", "U_Collection"):
        assert m[k].shape == g[k].shape and np.array_equal(m[k], g[k])
    with pytest.raises(ValueError):
        io_mat.save_nn_encoder(q, g["X_Collection_NO"][:, :5], g["X_Collection"], g["U_Collection"])
    gp = H.golden("ref_duffing_predict.npz")
    t = tmp_path / "DuffingPlot_trajectory.mat"
    io_mat.save_trajectory(t, gp["X_head"], gp["Y_head"], gp["U_head"], gp["test_Y"], gp["decoder_X"], test_X=gp["test_X"])
    m = sio.loadmat(t)
    T = int(gp["plotTime"])
    assert m["test_Y"].shape == (2, T) and m["decoder_X"].shape == (8, T) and m["marker_X"].shape == (8, T // 10)
    np.testing.assert_allclose(m["marker_X"], gp["marker_X"], rtol=0, atol=1e-12)
    assert m["marker_T"].shape[1] == T // 10 and m["Uplot"].shape[1] == T
