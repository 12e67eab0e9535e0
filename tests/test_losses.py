"""Training-loss evaluation: the oracle's literal restatement of duffing.py:179-235
against the reference script's own printed values (tests/golden/ref_duffing_losses.npz, produced by
running the script in the build container), and -- on the GPU -- the batched kernels against both."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import losses as olosses
from oracle import plant as oplant
from oracle import weights as ow


def _setup():
    g = H.golden("ref_duffing_losses.npz")
    enc = H.oracle_weights("duffing")
    dec = ow.load_mat_encoder(os.path.join(H.WEIGHTS, "duffing_decoder_weights.mat"))
    X, Y, U = oplant.generate_snapshots(100, 100, oplant.DUFFING_PRE, np.random.RandomState(101))
    return g, enc, dec, X, U


def test_oracle_losses_match_the_reference_script():
    g, enc, dec, X, U = _setup()
    assert [W.shape for W in dec[0]] == [(100, 8), (100, 100), (100, 100), (2, 100)]
    out = olosses.training_losses(enc, dec, X, U, g["A"], g["B"])
    assert len(out["window_sums"]) == int(g["last_j"]) == 70
    for k in ("Loss_rec", "Loss_lin", "Loss_pred", "Loss", "weight"):
        assert abs(out[k] - float(g[k])) <= 1e-10 * abs(float(g[k])), (k, out[k], float(g[k]))


@pytest.mark.gpu
def test_gpu_losses_match_the_reference_script_and_the_oracle():
    import koopman_online_updated_mpc_b200 as K
    g, enc, dec, X, U = _setup()
    e = K.Encoder(*enc)
    d = K.losses.load_decoder(os.path.join(H.WEIGHTS, "duffing_decoder_weights.mat"))
    assert d.n == 8 and d.nz == 2 and not d.has_tc
    out = K.losses.training_losses(e, d, X, U, g["A"], g["B"])
    for k in ("Loss_rec", "Loss_lin", "Loss_pred", "Loss", "weight"):
        assert abs(out[k] - float(g[k])) <= 1e-9 * abs(float(g[k])), (k, out[k], float(g[k]))
    want = olosses.training_losses(enc, dec, X, U, g["A"], g["B"])["window_sums"]
    got = out["window_sums"].cpu().numpy()
    assert got.shape == want.shape == (70, 3)
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    # every window of the snapshot set at once (9 970 windows): the first 70 are the reference's
    allw = K.losses.window_losses(e, d, X, U, g["A"], g["B"]).cpu().numpy()
    assert allw.shape == (10000 - 31 + 1, 3) and np.array_equal(allw[:70], got) and np.isfinite(allw).all()
    # from the package's own EDMD: A, B agree with the reference's A_hat, B_hat, hence the losses
    A, B, C, _ = K.scripts.identify(e, X, oplant.generate_snapshots(100, 100, oplant.DUFFING_PRE,
                                                                   np.random.RandomState(101))[1], U, n_step=100)
    out2 = K.losses.training_losses(e, d, X, U, A, B)
    assert abs(out2["Loss"] - float(g["Loss"])) <= 1e-7 * float(g["Loss"])
