"""The oracle against the reference: golden vectors produced by RUNNING the reference scripts
(tests/golden/make_golden.py), the reference's own golden file NN_Encoder.mat, and the
known-answer values of SURVEY.md Appendix A."""
import numpy as np
import pytest

import helpers as H
from oracle import closed_loop as ocl
from oracle import edmd, lift, mpc, plant, rls

KAT = {  # SURVEY.md Appendix A (numpy W x + b / ReLU chain on the reference .mat files)
    "duffing": {(0.0, 0.0): [-0.1307017308, 0.0116525620, 0.1553210699, 0.3012312445, 0.2159267715,
                             0.2903769236, 0.1455947749, -0.3163627459],
                (-2.0, -2.0): [-2.7544418928, 1.1253887376, -2.5303745365, 0.8120528076, 0.9200802247,
                               -1.1628101268, 0.2653595713, -1.4517699873]},
    "vdp": {(1.0, 0.0): [-0.0409514541, 0.2196833613, 0.4919650204, -1.0396789232, -0.9837153132,
                         -0.5268935956, 0.7626905720, -0.5200081464]},
    "tank": {(0.5, 1.5): [-0.1764189935, -0.0845278244, 0.5288223995, 0.2961597918, -0.8305293793,
                          0.2674247656, 0.1717034730, -0.6081194051, 0.4651725101, 0.2575188385]},
}


@pytest.mark.parametrize("system", ["duffing", "vdp", "tank"])
def test_encoder_known_answers(system):
    Ws, bs = H.oracle_weights(system)
    for x, z in KAT[system].items():
        np.testing.assert_allclose(lift.encoder_forward(Ws, bs, np.array(x)), z, rtol=0, atol=2e-10)


def test_lift_modes():
    Ws, bs = H.oracle_weights("duffing")
    x = np.array([[0.3, -0.7], [0.0, 0.0]])
    raw = lift.lift_mlp(Ws, bs, x, lift.LIFT_RAW)
    off = lift.lift_mlp(Ws, bs, x, lift.LIFT_OFFSET)
    stk = lift.lift_mlp(Ws, bs, x, lift.LIFT_STACK)
    np.testing.assert_allclose(off, raw - raw[1], atol=1e-15)
    assert np.all(np.abs(off[1]) < 1e-15) and stk.shape == (2, 10)
    np.testing.assert_allclose(stk[:, :2], x)
    np.testing.assert_allclose(stk[:, 2:], off)


def test_snapshots_and_edmd_match_reference_run():
    g = H.golden("ref_duffing.npz")
    Ws, bs = H.oracle_weights("duffing")
    X, Y, U = plant.generate_snapshots(100, 100, plant.DUFFING_PRE, np.random.RandomState(101))
    np.testing.assert_allclose(X[:, :300], g["X_head"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(U[:, :300], g["U_head"], rtol=0, atol=0)
    np.testing.assert_allclose(X.sum(1), g["X_sum"], rtol=1e-12)
    PHIX, PHIY = lift.encoder_forward(Ws, bs, X.T).T, lift.encoder_forward(Ws, bs, Y.T).T
    np.testing.assert_allclose(PHIX[:, :300], g["PHIX_head"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(PHIY[:, :300], g["PHIY_head"], rtol=0, atol=1e-13)
    A, B, C = edmd.edmd_pinv(PHIX, PHIY, U, X)
    for got, name in ((A, "A"), (B, "B"), (C, "C")):
        np.testing.assert_allclose(got, g[name].reshape(got.shape), rtol=0, atol=1e-12)
    assert abs(A[0, 0] - 1.01890695) < 1e-8 and abs(B[0, 0] - 0.03713868) < 1e-8
    # Gram form (Tank_System.m:93-100) == pinv form when V has full row rank
    A2, B2, C2 = edmd.edmd_from_gram(*edmd.gram_pack(PHIX, PHIY, U, X), 8)
    np.testing.assert_allclose(A2, A, atol=1e-10)
    np.testing.assert_allclose(C2, C, atol=1e-10)


def test_rbf_lift_and_edmd_match_reference_run():
    g = H.golden("ref_duffing_rbf.npz")
    X, Y, U = plant.generate_snapshots(100, 100, plant.DUFFING_PRE, np.random.RandomState(101))
    PHIX, PHIY = lift.rbf_lift(X.T, g["cx"]).T, lift.rbf_lift(Y.T, g["cx"]).T
    np.testing.assert_allclose(PHIX[:, :300], g["PHIX_head"], rtol=1e-10, atol=1e-12)
    A, B, C = edmd.edmd_pinv(PHIX, PHIY, U, X)
    np.testing.assert_allclose(A, g["A"], atol=1e-11)
    np.testing.assert_allclose(C, g["C"], atol=1e-11)


def _teacher_forced(g, cfg, T, p0, q0):
    Xg, Ug = g["logXloc"].T, g["logUloc"][0]
    x, st = np.array([-2.0, -2.0]), rls.RLSState(8, 1, 2, p0, q0)
    A, B, C = g["A"], g["B"], g["C"]
    du_exact = []
    for k in range(T):
        zl = cfg.lift_fn(x)
        ue, _, status = ocl.mpc_move(cfg, A, B, C, zl, 0.0, "exact")
        assert status == 0
        du_exact.append(abs(ue - Ug[k]))
        p = cfg.p_pre if k < cfg.first_post_step else cfg.p_post
        np.testing.assert_allclose(plant.rk4_step(x, Ug[k], np.array(p)), Xg[k], rtol=0, atol=5e-15)
        A, B, C = rls.rls_update(st, zl, Ug[k], cfg.lift_fn(Xg[k]), Xg[k])
        x = Xg[k]
    return np.array(du_exact), st, (A, B, C)


def test_duffing_update_loop_teacher_forced():
    """Feed the reference's own (x_k, u_k): plant and RLS must reproduce it to round-off, the
    exact QP minimiser must sit within the L-BFGS-B noise of the reference's controls."""
    g = H.golden("ref_duffing.npz")
    cfg = ocl.duffing_config(*H.oracle_weights("duffing"))
    du, st, (A, B, C) = _teacher_forced(g, cfg, 300, 1e4, 100.0)
    np.testing.assert_allclose(st.KA, g["K_A"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(st.P, g["inv_K_G"], rtol=0, atol=1e-9 * np.abs(g["inv_K_G"]).max())
    np.testing.assert_allclose(st.barX, g["bar_X"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(st.barQ, g["bar_Q"], rtol=0, atol=1e-9 * np.abs(g["bar_Q"]).max())
    np.testing.assert_allclose(A, g["Aloc"], rtol=0, atol=1e-6 * np.abs(g["Aloc"]).max())
    np.testing.assert_allclose(C, g["Cloc"], rtol=0, atol=1e-8)
    assert np.median(du) < 2e-5 and du[150:].max() < 1e-4 and du.max() < 5e-3


def test_vdp_update_loop_teacher_forced():
    g = H.golden("ref_vanderpol.npz")
    cfg = ocl.vanderpol_config(*H.oracle_weights("vdp"))
    du, st, _ = _teacher_forced(g, cfg, 400, 1e5, 1e5)
    np.testing.assert_allclose(st.P, g["inv_K_G"], rtol=0, atol=1e-8 * np.abs(g["inv_K_G"]).max())
    # the reference's controls are loosely converged L-BFGS-B iterates (ftol 2.2e-9, FD gradient):
    # on the +-6 box they sit within ~1e-4 of the exact minimiser
    assert np.median(du) < 1e-4 and du[200:].max() < 5e-4


@pytest.mark.parametrize("system", ["duffing", "vdp"])
def test_free_running_exact_loop_vs_reference_run(system):
    g = H.golden("ref_duffing.npz" if system == "duffing" else "ref_vanderpol.npz")
    Ws, bs = H.oracle_weights(system)
    cfg = ocl.duffing_config(Ws, bs) if system == "duffing" else ocl.vanderpol_config(Ws, bs)
    T = int(g["maxStep"])
    out = ocl.run_loop(cfg, g["A"], g["B"], g["C"], [-2.0, -2.0], T, update=ocl.UPDATE_RLS, qp="exact")
    assert np.abs(out["X"].T - g["logXloc"]).max() < 2e-4      # reference's own floor: 4e-5
    assert np.all(out["status"] == 0)
    out = ocl.run_loop(cfg, g["A"], g["B"], g["C"], [-2.0, -2.0], T, update=ocl.UPDATE_NONE, qp="exact")
    assert np.abs(out["X"].T - g["logX"]).max() < 2e-4


def test_vdp_update_loop_vs_reference_golden_file():
    """VDP_Revise_2/NN_Encoder.mat (written by vanderpol.py:1112): the update loop restarts the
    model from scratch, so it is reproducible although vanderpol.py's offline data is unseeded."""
    nn, g = H.golden("vdp_nn_encoder_head.npz"), H.golden("ref_vanderpol.npz")
    cfg = ocl.vanderpol_config(*H.oracle_weights("vdp"))
    out = ocl.run_loop(cfg, g["A"], g["B"], g["C"], [-2.0, -2.0], 600, update=ocl.UPDATE_RLS, qp="exact")
    assert np.abs(out["X"].T - nn["X_Collection"][:, :600]).max() < 1e-4
    np.testing.assert_allclose(out["X"][0], [-2.06079263, -0.60566339], atol=1e-8)
    assert out["U"][0] == 6.0


def test_literal_solver_short_run_matches_reference():
    g = H.golden("ref_duffing.npz")
    cfg = ocl.duffing_config(*H.oracle_weights("duffing"))
    out = ocl.run_loop(cfg, g["A"], g["B"], g["C"], [-2.0, -2.0], 20, update=ocl.UPDATE_RLS, qp="literal")
    np.testing.assert_allclose(out["X"].T, g["logXloc"][:, :20], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["U"][:3], [2.0, -2.0, 2.0], atol=1e-9)


def test_rbf_storage_method_is_warm_started_rls():
    """duffing_RBF.py:434-438 re-regresses over all stored data every step; that equals the RLS
    warm-started from the offline Gram (what the CUDA path runs)."""
    g = H.golden("ref_duffing_rbf.npz")
    X, Y, U = plant.generate_snapshots(100, 100, plant.DUFFING_PRE, np.random.RandomState(101))
    PHIX, PHIY = lift.rbf_lift(X.T, g["cx"]).T, lift.rbf_lift(Y.T, g["cx"]).T
    cfg, T = ocl.rbf_config(g["cx"]), 60
    lit = ocl.run_loop(cfg, g["A"], g["B"], g["C"], [-2.0, -2.0], T, update=ocl.UPDATE_STORAGE, qp="exact",
                       storage=dict(PHIX=PHIX, PHIY=PHIY, U=U, X=X))
    assert np.abs(lit["X"].T - g["logXloc"][:, :T]).max() < 1e-4
    G, Aq, XV = edmd.gram_pack(PHIX, PHIY, U, X)
    warm = rls.RLSState.warm(G, Aq, XV[:, :8], G[:8, :8])
    rl = ocl.run_loop(cfg, g["A"], g["B"], g["C"], [-2.0, -2.0], T, update=ocl.UPDATE_RLS, qp="exact", warm=warm)
    np.testing.assert_allclose(rl["X"], lit["X"], atol=1e-9)
    np.testing.assert_allclose(rl["A"], lit["A"], atol=1e-9)
    np.testing.assert_allclose(rl["C"], lit["C"], atol=1e-9)


def test_vanderpol_rbf_update_loop_regresses_C_on_the_duffing_states():
    """vanderpol_RBF.py:127-128 re-seeds (101) and calls `duffing_generate()` again -- the line is
    shared with duffing_RBF.py -- so from there on `X` holds the DUFFING snapshot states (and U the
    same inputs: identical draws), and the storage-method read-out `C = X pinv(X_EX)` (l.438)
    regresses duffing states on the VDP lifts.  With that quirk (warm bar_X = X_duffing PHIX') the
    oracle reproduces the reference run's update loop to round-off; its frozen loop switches to the
    duffing post-plant (l.328)."""
    g = H.golden("ref_vanderpol_rbf.npz")
    X, Y, U = plant.generate_snapshots(100, 100, plant.VDP_PRE, np.random.RandomState(101))
    Xd, _, Ud = plant.generate_snapshots(100, 100, plant.DUFFING_PRE, np.random.RandomState(101))
    assert np.array_equal(U, Ud)
    PHIX, PHIY = lift.rbf_lift(X.T, g["cx"]).T, lift.rbf_lift(Y.T, g["cx"]).T
    np.testing.assert_allclose(PHIX[:, :300], g["PHIX_head"], rtol=1e-10, atol=1e-12)
    A, B, C = edmd.edmd_pinv(PHIX, PHIY, U, X)
    np.testing.assert_allclose(A, g["A"], atol=1e-10)
    np.testing.assert_allclose(C, g["C"], atol=1e-10)
    cfg, T = ocl.rbf_config(g["cx"], "vanderpol"), int(g["maxStep"])
    G, Aq, _ = edmd.gram_pack(PHIX, PHIY, U, X)
    warm = rls.RLSState.warm(G, Aq, Xd @ PHIX.T, G[:8, :8])
    o = ocl.run_loop(cfg, g["A"], g["B"], g["C"], [-2.0, -2.0], T, update=ocl.UPDATE_RLS, qp="exact", warm=warm)
    assert np.abs(o["X"].T - g["logXloc"][:, :T]).max() < 1e-9      # saturated controls: exact agreement
    assert np.abs(o["U"] - g["logUloc"][0, :T]).max() < 1e-9
    np.testing.assert_allclose(np.linalg.norm(o["C"] - g["C"], 2) > 0.5, True)   # the read-out is indeed off
    cfgf = ocl.rbf_config(g["cx"], "vanderpol")
    cfgf.p_post = plant.DUFFING_POST
    f = ocl.run_loop(cfgf, g["A"], g["B"], g["C"], [-2.0, -2.0], T, update=ocl.UPDATE_NONE, qp="exact")
    assert np.abs(f["X"].T - g["logX"][:, :T]).max() < 2e-3


def test_exact_qp_solver_against_bvls_and_lbfgsb():
    rs = np.random.RandomState(0)
    for _ in range(60):
        N = int(rs.choice([1, 3, 10, 20]))
        M = rs.randn(N + 5, N)
        Hm, f = M.T @ M / N + 1e-3 * np.eye(N), rs.randn(N) * rs.choice([0.1, 1, 10])
        lb, ub = -rs.rand() - 0.01, rs.rand() + 0.01
        x, status, _ = mpc.solve_box_qp_exact(Hm, f, lb, ub)
        assert status == 0
        np.testing.assert_allclose(x, mpc.solve_box_qp_bvls(Hm, f, lb, ub), atol=1e-10)
    # KKT conditions hold
    g = 2 * Hm @ x + f
    assert np.all(g[(x > lb) & (x < ub)] < 1e-8) and np.all(g[x <= lb] >= -1e-8) and np.all(g[x >= ub] <= 1e-8)


def test_condensed_cost_equals_rollout_cost():
    """U'HU + f'U + const == costFunction(U) (duffing.py:540-581)."""
    rs = np.random.RandomState(2)
    A, B, C, z0 = rs.randn(8, 8) * 0.3, rs.randn(8, 1), rs.randn(2, 8), rs.randn(8)
    r = np.repeat(np.array([[1.0], [0.0]]), 10, axis=1)
    Hm, f = mpc.condense(A, B, C, z0, r, 100.0, 1e-4, 10)
    AB = np.concatenate([A, B], axis=1)
    c0 = mpc.cost_function_literal(np.zeros(10), r, AB, C, z0)
    for _ in range(5):
        U = rs.randn(10)
        assert abs(U @ Hm @ U + f @ U + c0 - mpc.cost_function_literal(U, r, AB, C, z0)) < 1e-8 * c0


def test_tank_restatement_behaves_like_the_paper():
    """MATLAB path is unpinned; check the qualitative behaviour SURVEY.md App. B records: level
    x2 -> 1 before the plant change, dips at the change (step 100), recovers."""
    import cases
    t = cases.tank_setup()
    out = ocl.run_loop(t["cfg"], t["A"], t["B"], t["C"], [0.0, 0.0], 260, update=ocl.UPDATE_RLS, qp="exact")
    x2 = out["X"][:, 1]
    assert abs(x2[95] - 1.0) < 5e-3 and x2[100:130].min() < 0.9 and abs(x2[255] - 1.0) < 5e-3
    assert np.all(np.abs(np.diff(np.concatenate([[0.0], out["U"]]))) <= 0.5 + 1e-9)  # dU box


# ------------------------------------------------ snapshot generator + open-loop predictor (N1, N2) --
@pytest.mark.parametrize("name,wsys,row,seed", [("duffing", "duffing", 0, 33), ("vanderpol", "vdp", 1, 50)])
def test_generator_and_open_loop_predictor_match_reference_run(name, wsys, row, seed):
    """oracle/plant.generate_snapshots and oracle/predict.py against the reference's own run of
    data_generate.py + duffing.py:262-343 / vanderpol.py:263-348 (ref_*_predict.npz)."""
    from oracle import lift as olift, plant as oplant, predict as opredict
    g = H.golden("ref_%s_predict.npz" % name)
    Ws, bs = H.oracle_weights(wsys)
    p = oplant.DUFFING_PRE if name == "duffing" else oplant.VDP_PRE
    X, Y, U = oplant.generate_snapshots(100, 100, p, np.random.RandomState(seed))
    assert X.shape[1] == int(g["n_snap"])
    np.testing.assert_allclose(X[:, :600], g["X_head"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(Y[:, -300:], g["Y_tail"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(U[:, :600], g["U_head"], rtol=0, atol=0)
    np.testing.assert_allclose(X.sum(axis=1), g["X_sum"], rtol=0, atol=1e-10)
    T = int(g["plotTime"])
    tY, dX, mX = opredict.open_loop_predict(lambda x: olift.lift_mlp(Ws, bs, x), g["A"], g["B"], g["C"], X, U, T)
    np.testing.assert_allclose(tY, g["test_Y"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(dX, g["decoder_X"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(mX, g["marker_X"], rtol=0, atol=1e-12)
    assert abs(opredict.rmse(tY, X, T, row) - float(g["RMSE"])) < 1e-14
