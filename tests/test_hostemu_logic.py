"""Kernel LOGIC on the CPU: the warp-per-scenario device functions (percase.cuh / loopbody.cuh)
compiled for the host with sequential lanes, checked against the oracle.  This is not a product
path (nothing in the package can reach it); the same cases run on the real kernels in
tests/test_gpu_parity.py."""
import numpy as np
import pytest

import cases
import helpers as H
from oracle import closed_loop as ocl
from oracle import edmd as oedmd
from oracle import rls as orls

EMU = H.EmuBackend()


@pytest.mark.parametrize("nz,ny,N,kw", [
    (8, 2, 10, {}), (8, 8, 10, dict(identity=True)), (11, 1, 20, {}), (8, 2, 50, dict(S=6)),
    (8, 2, 10, dict(shared=True)), (8, 2, 10, dict(r_full=True)), (8, 2, 10, dict(terminal=True)),
    (8, 8, 10, dict(identity=True, terminal=True)), (8, 2, 10, dict(wide=True)), (3, 1, 1, {}),
    (16, 4, 64, dict(S=3)), (10, 2, 10, {}),
])
def test_qp_matches_oracle(nz, ny, N, kw):
    act = cases.check_qp(EMU, nz, ny, N, **kw)
    if not kw.get("wide") and 10 <= N <= 50:
        assert act > 1.0  # the random cases exercise the active-set iterations


def test_qp_close_to_reference_lbfgsb():
    cases.check_qp_vs_literal(EMU)


@pytest.mark.parametrize("nz,lam,update_c,skip_first,p0,q0", [
    (8, 1.0, True, False, 1e4, 100.0), (8, 1.0, True, False, 1e5, 1e5), (10, 0.98, True, False, 1e4, 1e4),
    (10, 1.0, False, False, 1e4, 1e4), (10, 1.0, True, True, 1e4, 1e4), (1, 1.0, True, False, 10.0, 10.0),
    (16, 1.0, True, False, 1e3, 1e3),
])
def test_rls_matches_oracle(nz, lam, update_c, skip_first, p0, q0):
    cases.check_rls(EMU, nz, 2, lam, update_c, skip_first, p0, q0)


def test_plant_matches_oracle():
    cases.check_plant(EMU)


def test_rbf_matches_oracle_and_reference():
    cases.check_rbf(EMU)


def _run_emu(case, T, warm=None):
    emu = H.EmuClosedLoop(case["spec"], case["x0"], case["A"], case["B"], case["C"], case["r"], case["Ws"],
                          case["bs"], case["cx"], warm=warm, log_steps=T).run(T)
    return dict(log_x=emu.log_x, log_u=emu.log_u, status=emu.status, A=emu.A, B=emu.B, C=emu.C, emu=emu)


@pytest.mark.parametrize("name,T", [("duffing", 150), ("duffing_frozen", 130), ("vdp", 150), ("vdp_frozen", 110),
                                    ("duffing_rbf_frozen", 110)])
def test_closed_loop_matches_oracle_and_reference(name, T):
    case = cases.loop_case(name)
    run = _run_emu(case, T)
    cases.compare_loop(run, cases.oracle_loops(case, T), "update" if case["update"] else "frozen", T)
    gx, gu = case["gold"]  # scenario 0 is the reference's own run (x0 = [-2,-2])
    assert np.abs(run["log_x"][:, 0].T - gx[:, :T]).max() < 2e-4
    # chunked execution continues seamlessly (step index drives the plant switch at 102)
    emu2 = H.EmuClosedLoop(case["spec"], case["x0"], case["A"], case["B"], case["C"], case["r"], case["Ws"],
                           case["bs"], case["cx"], log_steps=T)
    emu2.run(60).run(T - 60)
    assert np.array_equal(emu2.log_x, run["log_x"]) and np.array_equal(emu2.log_u, run["log_u"])


def test_closed_loop_rbf_warm_rls():
    """duffing_RBF.py update loop ('storage method') == RLS warm-started from the offline Gram."""
    import koopman_online_updated_mpc_b200 as K
    from oracle import lift as olift, plant as oplant
    g = H.golden("ref_duffing_rbf.npz")
    X, Y, U = oplant.generate_snapshots(100, 100, oplant.DUFFING_PRE, np.random.RandomState(101))
    PX, PY = olift.rbf_lift(X.T, g["cx"]).T, olift.rbf_lift(Y.T, g["cx"]).T
    G, Aq, XV = oedmd.gram_pack(PX, PY, U, X)
    x0 = np.array([[-2.0, -2.0], [1.0, 0.5]])
    S, T = 2, 110
    warm = H.new_rls_state(S, 8, 2, 1.0, 1.0)
    warm["KA"][:], warm["P"][:] = Aq, np.linalg.pinv(G)
    warm["barX"][:], warm["barQ"][:] = XV[:, :8], np.linalg.pinv(G[:8, :8])
    emu = H.EmuClosedLoop(K.rbf_spec(), x0, g["A"], g["B"], g["C"], np.array([1.0, 0.0]), cx=g["cx"], warm=warm,
                          log_steps=T).run(T)
    cfg = ocl.rbf_config(g["cx"])
    for s in range(S):
        o = ocl.run_loop(cfg, g["A"], g["B"], g["C"], x0[s], T, update=ocl.UPDATE_RLS, qp="exact",
                         warm=orls.RLSState.warm(G, Aq, XV[:, :8], G[:8, :8]))
        assert np.abs(o["X"] - emu.log_x[:, s]).max() < 1e-7
        np.testing.assert_allclose(emu.A[s], o["A"], atol=1e-7)
    assert np.abs(emu.log_x[:, 0].T - g["logXloc"][:, :T]).max() < 1e-4  # the reference's own run


def test_closed_loop_tank_velocity_form():
    """Tank_System.m:170-291 with the Encoder_Tank lift (BASELINE config 3)."""
    import koopman_online_updated_mpc_b200 as K
    t = cases.tank_setup()
    x0 = np.array([[0.0, 0.0], [0.5, 1.5], [2.0, 0.2]])
    T = 160
    emu = H.EmuClosedLoop(K.tank_spec(), x0, t["A"], t["B"], t["C"], np.array([1.0]), t["Ws"], t["bs"],
                          log_steps=T).run(T)
    for s in range(len(x0)):
        o = ocl.run_loop(t["cfg"], t["A"], t["B"], t["C"], x0[s], T, update=ocl.UPDATE_RLS, qp="exact")
        # cond(H) reaches 1e16 while the restarted model is rank deficient: compare
        # states loosely in the transient, tightly once the model has converged
        assert np.abs(o["X"] - emu.log_x[:, s]).max() < 5e-3, s
        assert np.abs(o["X"][60:100] - emu.log_x[60:100, s]).max() < 1e-5, s
        assert np.all((emu.log_x[:, s] >= 0.0))
    du = np.abs(np.diff(np.concatenate([np.zeros((1, len(x0))), emu.log_u]), axis=0))
    assert du.max() <= 0.5 + 1e-9 and np.abs(emu.log_u).max() <= 8.0 + 1e-9


def test_spd_right_solve():
    rs = np.random.RandomState(0)
    for n, rows in ((9, 10), (11, 12), (1, 1), (8, 2)):
        M = rs.randn(n, n + 3)
        G = M @ M.T + 0.1 * np.eye(n)
        Bm = rs.randn(rows, n)
        want = Bm @ np.linalg.inv(G)
        Gc, Bc = G.copy(), Bm.copy()
        status = H.hostemu().emu_spd_right_solve(H.dp(Gc), n, H.dp(Bc), rows)
        assert status == 0
        np.testing.assert_allclose(Bc, want, rtol=1e-10, atol=1e-12)
    G = np.array([[1.0, 2.0], [2.0, 1.0]])  # indefinite -> pivot flag
    assert H.hostemu().emu_spd_right_solve(H.dp(G), 2, H.dp(np.ones((1, 2))), 1) == 4


@pytest.mark.parametrize("name,x0", [("duffing", [1.91195805, 0.15398348]), ("vdp", [-2.0, -2.0])])
def test_teacher_forced_single_steps_along_the_horizon(name, x0):
    """Every step k = 1..299 of an oracle trajectory (including the post-switch chattering regime
    that makes free-running comparisons chaotic) as one batch of single-step problems."""
    case = cases.loop_case(name)
    batch, want = cases.teacher_forced_batch(case, np.array(x0), 300)
    S = len(batch["x"])
    warm = dict(KA=batch["KA"].copy(), P=batch["P"].copy(), barX=batch["barX"].copy(), barQ=batch["barQ"].copy(),
                A=np.zeros_like(batch["A"]), B=np.zeros((S, 8)), C=np.zeros_like(batch["C"]))
    emu = H.EmuClosedLoop(case["spec"], batch["x"], batch["A"], batch["B"], batch["C"], case["r"], case["Ws"],
                          case["bs"], None, warm=warm, log_steps=1, params_pre=batch["params"],
                          params_post=batch["params"])
    emu.u_prev[:] = batch["u_prev"]
    emu.run(1)
    cases.compare_teacher_forced(dict(u=emu.log_u[0], x=emu.log_x[0], z=emu.z, A=emu.A, B=emu.B, C=emu.C), want)
