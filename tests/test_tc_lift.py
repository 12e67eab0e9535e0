"""tcgen05 / TMEM / TMA split-precision lift (csrc/tc_lift.cu, KMPC_PREC_TC) against the oracle, the
fp64 tensor path and the reference's own EDMD run.

Tolerances (north_star: 1e-4 relative on lifted states and Koopman matrices; the round-2 bar for
this kernel: lifted states <= 1e-6 relative, A, B, C <= 1e-4 relative):
  lifted states   |z_tc - z_oracle| <= 1e-6 * max|z|          (bf16 x 3 split, fp32 accumulation)
  A, B, C         <= 1e-4 relative vs tests/golden/ref_duffing.npz (the reference script's own run)
                  and vs the fp64 path on a large synthetic snapshot set
"""
import numpy as np
import pytest
import torch

import helpers as H
import koopman_online_updated_mpc_b200 as K
from koopman_online_updated_mpc_b200 import edmd as kedmd
from koopman_online_updated_mpc_b200.lift import LIFT_OFFSET, LIFT_RAW, LIFT_STACK, PREC_FP64, PREC_TC
from oracle import lift as olift
from oracle import plant as oplant

pytestmark = pytest.mark.gpu


def _enc(system):
    Ws, bs = H.oracle_weights(system)
    return K.Encoder(Ws, bs), Ws, bs


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("system", ["duffing", "vdp", "tank"])
@pytest.mark.parametrize("S", [1, 127, 128, 129, 4096 + 77, 148 * 256 * 3 + 5])
def test_tc_lift_matches_oracle(system, S):
    """Ragged sizes around the 128-row tile / 2-tile pair / grid boundaries."""
    enc, Ws, bs = _enc(system)
    assert enc.has_tc
    rs = np.random.default_rng(S)
    x = rs.uniform(-2.5, 2.5, (S, 2))
    want = olift.encoder_forward(Ws, bs, x) if S <= 5000 else None
    xd = torch.from_numpy(x).cuda()
    z64 = enc(xd, precision=PREC_FP64).cpu().numpy()
    ztc = enc(xd, precision=PREC_TC).cpu().numpy()
    assert ztc.shape == z64.shape
    assert np.isfinite(ztc).all()
    assert _rel(ztc, z64) <= 1e-6, _rel(ztc, z64)
    if want is not None:
        assert _rel(ztc, want) <= 1e-6


def test_tc_lift_modes_and_known_answers():
    """OFFSET / STACK lift modes (Koopman_update_Tracking_Lift.m:65, Koopman_update.m:67) and the
    Appendix-A known answers of SURVEY.md (theta_E([1, 0]) of the VDP net)."""
    enc, Ws, bs = _enc("vdp")
    x = np.array([[1.0, 0.0], [0.0, 0.0], [-2.0, -2.0]])
    xd = torch.from_numpy(x).cuda()
    raw = enc(xd, mode=LIFT_RAW, precision=PREC_TC).cpu().numpy()
    kat = np.array([-0.0409514541, 0.2196833613, 0.4919650204, -1.0396789232, -0.9837153132, -0.5268935956,
                    0.7626905720, -0.5200081464])
    assert np.abs(raw[0] - kat).max() < 2e-6
    z0 = olift.encoder_forward(Ws, bs, np.zeros((1, 2)))[0]
    off = enc(xd, mode=LIFT_OFFSET, precision=PREC_TC).cpu().numpy()
    assert np.abs(off - (raw - z0)).max() < 1e-12      # same lift, theta(0) subtracted in fp64
    assert np.abs(off[1]).max() < 2e-6
    st = enc(xd, mode=LIFT_STACK, precision=PREC_TC).cpu().numpy()
    assert st.shape == (3, 10)
    assert np.array_equal(st[:, :2], x)
    assert np.abs(st[:, 2:] - off).max() < 1e-12


def test_tc_edmd_matches_reference_run():
    """A, B, C from the tcgen05 lift + fp64 Gram vs the reference script's own matrices
    (tests/golden/ref_duffing.npz: duffing.py:167-177 on its seed-101 snapshot set, 10 k snapshots)."""
    g = H.golden("ref_duffing.npz")
    enc, Ws, bs = _enc("duffing")
    X, Y, U = oplant.generate_snapshots(100, 100, oplant.DUFFING_PRE, np.random.RandomState(101))
    for fn, kw in ((kedmd.gram_from_snapshots, {}), (kedmd.gram_from_trajectories, {"n_step": 100})):
        pack = fn(enc, X.T.copy(), Y.T.copy(), U.reshape(-1), precision=PREC_TC, **kw)
        A, B, C, status = kedmd.edmd_solve(pack, 8, 2)
        assert int(status.item()) == 0
        for got, want in ((A, g["A"]), (B, g["B"]), (C, g["C"])):
            assert _rel(got.cpu().numpy(), want) <= 1e-4, _rel(got.cpu().numpy(), want)


def test_tc_edmd_matches_fp64_path_at_scale():
    """BASELINE configs[3] size: 10 M synthetic duffing snapshots (100 000 trajectories x 100 steps, generated
    on the GPU like data_generate.py:17-57): the Koopman matrices from the tcgen05 lift agree with the fp64
    lift to the north-star tolerance (1e-4 relative)."""
    enc, _, _ = _enc("duffing")
    n_traj, n_step = 100000, 100
    rs = np.random.default_rng(7)
    X, Y, U = K.data_generate.generate_snapshots(rs.uniform(-2, 2, (n_traj, 2)), rs.uniform(-2, 2, (n_step, n_traj)),
                                                 K.plant.DUFFING_PRE)
    mats = {}
    for name, prec in (("fp64", PREC_FP64), ("tc", PREC_TC)):
        pack = kedmd.gram_from_trajectories(enc, X, Y, U, n_step, precision=prec)
        A, B, C, status = kedmd.edmd_solve(pack, 8, 2)
        assert int(status.item()) == 0
        mats[name] = [m.cpu().numpy() for m in (A, B, C)]
    for got, want in zip(mats["tc"], mats["fp64"]):
        assert _rel(got, want) <= 1e-4, _rel(got, want)


def test_tc_unsupported_net_is_reported_not_emulated():
    """A net outside the kernel's shape (hidden width 120 > 112) has no tensor path: the _ex entry
    points say so instead of silently computing something else."""
    rs = np.random.default_rng(0)
    Ws = [rs.standard_normal((120, 2)), rs.standard_normal((120, 120)) * 0.1, rs.standard_normal((8, 120)) * 0.1]
    bs = [rs.standard_normal(120), rs.standard_normal(120), rs.standard_normal(8)]
    enc = K.Encoder(Ws, bs)
    assert not enc.has_tc
    with pytest.raises(K.KmpcError):
        enc(torch.zeros((4, 2), dtype=torch.float64, device="cuda"), precision=PREC_TC)
