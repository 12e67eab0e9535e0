"""Parity tests proper: the sm_100a kernels, called through the C ABI (ctypes -> libkmpc.so),
against the oracle and the reference goldens.  Run on the B200 box with `-m gpu`."""
from dataclasses import replace

import numpy as np
import pytest
import torch

import cases
import helpers as H
import koopman_online_updated_mpc_b200 as K
from oracle import closed_loop as ocl
from oracle import edmd as oedmd
from oracle import lift as olift
from oracle import plant as oplant
from oracle import rls as orls

pytestmark = pytest.mark.gpu
CUDA = H.CudaBackend()


# ------------------------------------------------------------------------------ stage 1: lift --
@pytest.mark.parametrize("system", ["duffing", "vdp", "tank"])
def test_encoder_matches_oracle(system):
    Ws, bs = H.oracle_weights(system)
    enc = K.Encoder(Ws, bs)
    rs = np.random.RandomState(0)
    for S in (1, 31, 32, 33, 257, 4097):          # ragged tiles (kTileS = 32)
        x = rs.uniform(-2.5, 2.5, (S, 2))
        z = enc(x)
        assert z.shape == (S, enc.nz)
        np.testing.assert_allclose(z, olift.encoder_forward(Ws, bs, x), rtol=0, atol=1e-12)
    from test_oracle_golden import KAT
    for xk, zk in KAT[system].items():            # SURVEY Appendix A known answers
        np.testing.assert_allclose(enc(np.array(xk)), zk, rtol=0, atol=2e-10)
    assert enc(np.zeros((0, 2))).shape == (0, enc.nz)   # empty batch


def test_encoder_lift_modes_and_tensor_kinds():
    Ws, bs = H.oracle_weights("duffing")
    enc = K.Encoder(Ws, bs)
    x = np.random.RandomState(1).uniform(-2, 2, (100, 2))
    for mode in (0, 1, 2):
        np.testing.assert_allclose(enc(x, mode), olift.lift_mlp(Ws, bs, x, mode), rtol=0, atol=1e-12)
    xt = torch.from_numpy(x)
    assert isinstance(enc(xt), torch.Tensor) and not enc(xt).is_cuda
    zc = enc(xt.cuda())
    assert zc.is_cuda
    np.testing.assert_allclose(zc.cpu().numpy(), enc(x), rtol=0, atol=0)


def test_encoder_matches_reference_lifted_snapshots():
    g = H.golden("ref_duffing.npz")
    enc = K.Encoder.from_file(H.weights_path("duffing"))
    np.testing.assert_allclose(enc(g["X_head"].T.copy()).T, g["PHIX_head"], rtol=0, atol=1e-12)


def test_rbf_matches_oracle_and_reference():
    cases.check_rbf(CUDA)


# ------------------------------------------------------------------------------ stage 2: EDMD --
def test_gram_and_edmd_match_oracle_and_reference_run():
    g = H.golden("ref_duffing.npz")
    Ws, bs = H.oracle_weights("duffing")
    X, Y, U = oplant.generate_snapshots(100, 100, oplant.DUFFING_PRE, np.random.RandomState(101))
    PHIX, PHIY = olift.encoder_forward(Ws, bs, X.T).T, olift.encoder_forward(Ws, bs, Y.T).T
    A, B, C = K.edmd.edmd(PHIX, PHIY, U, X)          # reference call shape, numpy in / numpy out
    np.testing.assert_allclose(A, g["A"], rtol=0, atol=1e-9)   # vs the reference script's own A, B, C
    np.testing.assert_allclose(B, g["B"].reshape(8, 1), rtol=0, atol=1e-9)
    np.testing.assert_allclose(C, g["C"], rtol=0, atol=1e-9)
    pack = K.edmd.gram_accumulate(PHIX.T.copy(), PHIY.T.copy(), U.ravel(), X.T.copy()).cpu().numpy()
    G, Aq, XV = oedmd.gram_pack(PHIX, PHIY, U, X)
    np.testing.assert_allclose(pack[:-1], np.concatenate([G.ravel(), Aq.ravel(), XV.ravel()]), rtol=1e-12)
    assert pack[-1] == 10000
    # fused lift + Gram straight from raw snapshots, and the MATLAB joint-C variant
    enc = K.Encoder(Ws, bs)
    pack2 = K.edmd.gram_from_snapshots(enc, X.T.copy(), Y.T.copy(), U.ravel())
    np.testing.assert_allclose(pack2.cpu().numpy(), pack, rtol=1e-11)
    A2, B2, C2, st = K.edmd.edmd_solve(pack2, 8, 2, K.edmd.C_JOINT)
    Ao, Bo, Co = oedmd.edmd_from_gram(G, Aq, XV, 8, oedmd.C_JOINT)
    assert int(st.item()) == 0
    np.testing.assert_allclose(C2.cpu().numpy(), Co, rtol=0, atol=1e-9)
    np.testing.assert_allclose(A2.cpu().numpy(), Ao, rtol=0, atol=1e-9)


@pytest.mark.parametrize("nz", [3, 10, 11])
def test_gram_other_dimensions(nz):
    rs = np.random.RandomState(nz)
    M = 1237                                         # ragged: not a multiple of the block size
    PX, PY, U, X = rs.randn(nz, M), rs.randn(nz, M), rs.randn(1, M), rs.randn(2, M)
    pack = K.edmd.gram_accumulate(PX.T.copy(), PY.T.copy(), U.ravel(), X.T.copy()).cpu().numpy()
    G, Aq, XV = oedmd.gram_pack(PX, PY, U, X)
    np.testing.assert_allclose(pack[:-1], np.concatenate([G.ravel(), Aq.ravel(), XV.ravel()]), rtol=1e-11, atol=1e-11)
    A, B, C = K.edmd.edmd(PX, PY, U, X)
    Ao, Bo, Co = oedmd.edmd_pinv(PX, PY, U, X)
    np.testing.assert_allclose(A, Ao, atol=1e-10)
    np.testing.assert_allclose(C, Co, atol=1e-10)


@pytest.mark.parametrize("M", [1, 3, 4, 5, 1237, 148 * 2 * 8 * 8 * 4 + 3, 700001])
def test_gram_tensor_path_ragged_sizes(M):
    """nz = 8: the Gram runs on the fp64 tensor path (gram_dmma_kernel: groups of 4 snapshots, contiguous
    ranges per warp, eight groups in flight).  Sizes around a group, around one group per warp and with
    idle warps; pack and snapshot count against numpy."""
    rs = np.random.RandomState(M % 1000)
    PX, PY, U, X = rs.randn(8, M), rs.randn(8, M), rs.randn(1, M), rs.randn(2, M)
    pack = K.edmd.gram_accumulate(PX.T.copy(), PY.T.copy(), U.ravel(), X.T.copy()).cpu().numpy()
    G, Aq, XV = oedmd.gram_pack(PX, PY, U, X)
    want = np.concatenate([G.ravel(), Aq.ravel(), XV.ravel()])
    np.testing.assert_allclose(pack[:-1], want, rtol=0, atol=1e-12 * max(1.0, np.abs(want).max()) * max(1.0, M ** 0.5))
    assert pack[-1] == M


@pytest.mark.parametrize("n_step,n_traj", [(1, 7), (2, 5), (3, 9), (5, 13), (11, 37), (100, 211)])
def test_gram_tensor_path_trajectory_layout(n_step, n_traj):
    """Trajectory layout (n_step + 1 lifted states per trajectory, the snapshot's successor is the next row):
    trajectories shorter than a group of 4, groups straddling trajectory boundaries."""
    enc = K.Encoder.from_file(H.weights_path("duffing"))
    rs = np.random.default_rng(n_step * 100 + n_traj)
    X, Y, U = K.data_generate.generate_snapshots(rs.uniform(-2, 2, (n_traj, 2)), rs.uniform(-2, 2, (n_step, n_traj)),
                                                 K.plant.DUFFING_PRE)
    a = K.edmd.gram_from_trajectories(enc, X, Y, U, n_step, verify=True).cpu().numpy()
    b = K.edmd.gram_from_snapshots(enc, X, Y, U).cpu().numpy()
    Ws, bs = H.oracle_weights("duffing")
    Xh, Yh = X.cpu().numpy(), Y.cpu().numpy()
    G, Aq, XV = oedmd.gram_pack(olift.encoder_forward(Ws, bs, Xh).T, olift.encoder_forward(Ws, bs, Yh).T,
                                U.cpu().numpy().reshape(1, -1), Xh.T)
    want = np.concatenate([G.ravel(), Aq.ravel(), XV.ravel()])
    tol = 1e-11 * max(1.0, np.abs(want).max())
    np.testing.assert_allclose(a[:-1], want, rtol=0, atol=tol)
    np.testing.assert_allclose(b[:-1], want, rtol=0, atol=tol)
    assert a[-1] == b[-1] == n_step * n_traj


def test_edmd_rank_deficient_is_flagged():
    rs = np.random.RandomState(0)
    PX = rs.randn(8, 500)
    PX[7] = PX[0]                                    # duplicate observable -> singular Gram
    with pytest.raises(K.KmpcError):
        K.edmd.edmd(PX, rs.randn(8, 500), rs.randn(1, 500), rs.randn(2, 500))


# ------------------------------------------------------------------------------ stage 3: RLS ---
@pytest.mark.parametrize("nz,lam,update_c,skip_first,p0,q0", [
    (8, 1.0, True, False, 1e4, 100.0), (8, 1.0, True, False, 1e5, 1e5), (10, 0.98, True, False, 1e4, 1e4),
    (10, 1.0, False, False, 1e4, 1e4), (10, 1.0, True, True, 1e4, 1e4), (16, 1.0, True, False, 1e3, 1e3),
])
def test_rls_matches_oracle(nz, lam, update_c, skip_first, p0, q0):
    cases.check_rls(CUDA, nz, 2, lam, update_c, skip_first, p0, q0)


def test_rls_duffing_trace_reproduces_reference_state():
    """Teacher-forced on the reference's own (x_k, u_k) trace: after 300 updates the device RLS
    state equals the K_A / inv_K_G / bar_X / bar_Q the reference script ended with."""
    g = H.golden("ref_duffing.npz")
    Ws, bs = H.oracle_weights("duffing")
    enc = K.Encoder(Ws, bs)
    Xg, Ug = g["logXloc"].T, g["logUloc"][0]
    xs = np.concatenate([[[-2.0, -2.0]], Xg[:-1]])
    Z, Yl = enc(xs), enc(Xg)
    st = K.RLSState(1, 8, 2, 1e4, 100.0)
    for k in range(300):
        A, B, C = K.rls_update(st, Z[k:k + 1], Ug[k:k + 1], Yl[k:k + 1], Xg[k:k + 1])
    np.testing.assert_allclose(st.KA[0].cpu().numpy(), g["K_A"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(st.P[0].cpu().numpy(), g["inv_K_G"], rtol=0, atol=1e-8 * np.abs(g["inv_K_G"]).max())
    np.testing.assert_allclose(A[0].cpu().numpy(), g["Aloc"], rtol=0, atol=1e-5 * np.abs(g["Aloc"]).max())
    np.testing.assert_allclose(C[0].cpu().numpy(), g["Cloc"], rtol=0, atol=1e-7)


# ------------------------------------------------------------------------------ stage 4: QP ----
@pytest.mark.parametrize("nz,ny,N,kw", [
    (8, 2, 10, {}), (8, 8, 10, dict(identity=True)), (11, 1, 20, {}), (8, 2, 50, dict(S=6)),
    (8, 2, 10, dict(shared=True)), (8, 2, 10, dict(r_full=True)), (8, 2, 10, dict(terminal=True)),
    (8, 8, 10, dict(identity=True, terminal=True)), (8, 2, 10, dict(wide=True)), (3, 1, 1, {}),
    (16, 4, 64, dict(S=3)), (8, 2, 10, dict(S=1029)),
])
def test_qp_matches_oracle(nz, ny, N, kw):
    cases.check_qp(CUDA, nz, ny, N, **kw)


def test_qp_close_to_reference_lbfgsb():
    cases.check_qp_vs_literal(CUDA)


def test_plant_matches_oracle():
    cases.check_plant(CUDA)


# ------------------------------------------------------------------------------ fused loop -----
def _run_cuda(case, T, x0=None, warm=None, **kw):
    enc = K.Encoder(case["Ws"], case["bs"]) if case["Ws"] is not None else None
    loop = K.ClosedLoop(case["spec"], case["x0"] if x0 is None else x0, case["A"], case["B"], case["C"],
                        case["r"], encoder=enc, cx=case["cx"], rls_state=warm, log_steps=T, **kw)
    loop.run(T)
    torch.cuda.synchronize()
    return dict(log_x=loop.log_x.cpu().numpy(), log_u=loop.log_u.cpu().numpy(), status=loop.status.cpu().numpy(),
                A=loop.A.cpu().numpy(), B=loop.B.cpu().numpy(), C=loop.C.cpu().numpy(), loop=loop)


@pytest.mark.parametrize("name,T", [("duffing", 150), ("duffing_frozen", 130), ("vdp", 150), ("vdp_frozen", 110),
                                    ("duffing_rbf_frozen", 110)])
def test_closed_loop_matches_oracle_and_reference(name, T):
    case = cases.loop_case(name)
    run = _run_cuda(case, T)
    cases.compare_loop(run, cases.oracle_loops(case, T), "update" if case["update"] else "frozen", T)
    gx, gu = case["gold"]           # scenario 0 == the reference script's own run from x0 = [-2,-2]
    assert np.abs(run["log_x"][:, 0].T - gx[:, :T]).max() < 2e-4
    enc = K.Encoder(case["Ws"], case["bs"]) if case["Ws"] is not None else None
    loop2 = K.ClosedLoop(case["spec"], case["x0"], case["A"], case["B"], case["C"], case["r"], encoder=enc,
                         cx=case["cx"], log_steps=T)
    loop2.run(60).run(T - 60)       # chunked execution is bit-identical
    assert np.array_equal(loop2.log_x.cpu().numpy(), run["log_x"])
    assert loop2.step_index == T


def test_vdp_closed_loop_vs_reference_golden_file():
    """600 steps against VDP_Revise_2/NN_Encoder.mat (vanderpol.py:1112), update loop."""
    nn = H.golden("vdp_nn_encoder_head.npz")
    case = cases.loop_case("vdp")
    run = _run_cuda(case, 600, x0=np.array([[-2.0, -2.0]]))
    assert np.abs(run["log_x"][:, 0].T - nn["X_Collection"][:, :600]).max() < 1e-4
    assert run["status"][0] == 0


@pytest.mark.parametrize("name,x0", [("duffing", [1.91195805, 0.15398348]), ("vdp", [-2.0, -2.0])])
def test_teacher_forced_single_steps_along_the_horizon(name, x0):
    case = cases.loop_case(name)
    batch, want = cases.teacher_forced_batch(case, np.array(x0), 300)
    S = len(batch["x"])
    warm = K.RLSState(S, 8, 2)
    for k in ("KA", "P", "barX", "barQ"):
        getattr(warm, k).copy_(torch.from_numpy(batch[k]))
    enc = K.Encoder(case["Ws"], case["bs"])
    loop = K.ClosedLoop(case["spec"], batch["x"], batch["A"], batch["B"], batch["C"], case["r"], encoder=enc,
                        rls_state=warm, log_steps=1, params_pre=batch["params"], params_post=batch["params"],
                        u_prev=batch["u_prev"])
    loop.run(1)
    got = dict(u=loop.log_u[0].cpu().numpy(), x=loop.log_x[0].cpu().numpy(), z=loop.z.cpu().numpy(),
               A=loop.A.cpu().numpy(), B=loop.B.cpu().numpy(), C=loop.C.cpu().numpy())
    cases.compare_teacher_forced(got, want)


def test_closed_loop_rbf_warm_rls():
    g = H.golden("ref_duffing_rbf.npz")
    X, Y, U = oplant.generate_snapshots(100, 100, oplant.DUFFING_PRE, np.random.RandomState(101))
    PX, PY = olift.rbf_lift(X.T, g["cx"]).T, olift.rbf_lift(Y.T, g["cx"]).T
    G, Aq, XV = oedmd.gram_pack(PX, PY, U, X)
    x0 = np.array([[-2.0, -2.0], [1.0, 0.5]])
    T = 110
    warm = K.RLSState.warm(2, G, Aq, XV[:, :8], G[:8, :8])
    loop = K.ClosedLoop(K.rbf_spec(), x0, g["A"], g["B"], g["C"], np.array([1.0, 0.0]), cx=g["cx"],
                        rls_state=warm, log_steps=T).run(T)
    lx = loop.log_x.cpu().numpy()
    cfg = ocl.rbf_config(g["cx"])
    for s in range(2):
        o = ocl.run_loop(cfg, g["A"], g["B"], g["C"], x0[s], T, update=ocl.UPDATE_RLS, qp="exact",
                         warm=orls.RLSState.warm(G, Aq, XV[:, :8], G[:8, :8]))
        assert np.abs(o["X"] - lx[:, s]).max() < 1e-7
    assert np.abs(lx[:, 0].T - g["logXloc"][:, :T]).max() < 1e-4


def test_closed_loop_tank_velocity_form():
    t = cases.tank_setup()
    x0 = np.array([[0.0, 0.0], [0.5, 1.5], [2.0, 0.2]])
    T = 160
    enc = K.Encoder(t["Ws"], t["bs"])
    loop = K.ClosedLoop(K.tank_spec(), x0, t["A"], t["B"], t["C"], np.array([1.0]), encoder=enc, log_steps=T).run(T)
    lx, lu = loop.log_x.cpu().numpy(), loop.log_u.cpu().numpy()
    for s in range(len(x0)):
        o = ocl.run_loop(t["cfg"], t["A"], t["B"], t["C"], x0[s], T, update=ocl.UPDATE_RLS, qp="exact")
        assert np.abs(o["X"] - lx[:, s]).max() < 5e-3, s
        assert np.abs(o["X"][60:100] - lx[60:100, s]).max() < 1e-5, s
    du = np.abs(np.diff(np.concatenate([np.zeros((1, len(x0))), lu]), axis=0))
    assert du.max() <= 0.5 + 1e-9 and np.abs(lu).max() <= 8.0 + 1e-9 and lx.min() >= 0.0


# ------------------------------------------------------------------------------ full size ------
def test_full_size_vdp_properties_and_shard_invariance():
    """BASELINE config 2 shape (4096 VDP scenarios, tracking + online update): size-independent
    properties -- bounds respected, finite, status clean, per-scenario results bit-identical
    whether the batch runs whole or as two shards (scenarios are independent)."""
    Ws, bs = H.oracle_weights("vdp")
    g = H.golden("ref_vanderpol.npz")
    enc = K.Encoder(Ws, bs)
    rs = np.random.default_rng(20240601)
    S, T = 4096, 60
    x0 = rs.uniform(-2, 2, (S, 2))
    x0[0] = [-2.0, -2.0]
    xref = np.stack([rs.uniform(-1, 1, S), np.zeros(S)], axis=1)
    xref[0] = [1.0, 0.0]
    r = enc(xref)
    spec = K.vanderpol_spec()

    def run(sl):
        loop = K.ClosedLoop(spec, x0[sl], g["A"], g["B"], g["C"], r[sl], encoder=enc, log_steps=T).run(T)
        torch.cuda.synchronize()
        return loop.log_x.cpu().numpy(), loop.log_u.cpu().numpy(), loop.status.cpu().numpy(), loop.A.cpu().numpy()

    lx, lu, st, A = run(slice(0, S))
    assert np.all(np.isfinite(lx)) and np.all(np.isfinite(A)) and np.all(st == 0)
    assert lu.min() >= -6.0 and lu.max() <= 6.0
    assert np.abs(lx[:, 0].T - g["logXloc"][:, :T]).max() < 1e-4     # scenario 0 is the reference's own
    lx1, lu1, _, A1 = run(slice(0, S // 2))
    lx2, lu2, _, A2 = run(slice(S // 2, S))
    assert np.array_equal(np.concatenate([lx1, lx2], axis=1), lx)
    assert np.array_equal(np.concatenate([lu1, lu2], axis=1), lu)
    assert np.array_equal(np.concatenate([A1, A2], axis=0), A)
    o = ocl.run_loop(ocl.vanderpol_config(Ws, bs, xref[77]), g["A"], g["B"], g["C"], x0[77], T,
                     update=ocl.UPDATE_RLS, qp="exact")
    assert np.abs(o["X"] - lx[:, 77]).max() < 1e-4                    # a random scenario vs the oracle


def test_launch_counter_counts_kernels():
    before = K.launch_count()
    K.lift.rbf(np.zeros((4, 2)), np.ones((8, 2)))
    assert K.launch_count() == before + 1


# ------------------------------------------------------------------------------ fused kernel ---
@pytest.mark.parametrize("name,T", [("duffing", 150), ("duffing_frozen", 130), ("vdp", 150), ("vdp_frozen", 110),
                                    ("duffing_rbf_frozen", 110)])
def test_fused_kernel_matches_generic_three_kernel_path(name, T):
    """The persistent fused kernel (fused.cu: state in registers, T steps per launch) against the
    generic qp_plant -> lift -> rls kernels (closed_loop.cu) on the same scenarios."""
    case = cases.loop_case(name)
    fused = _run_cuda(case, T)
    assert fused["loop"].fused
    generic = _run_cuda(dict(case, spec=replace(case["spec"], path=K.closed_loop.PATH_GENERIC)), T)
    assert not generic["loop"].fused
    assert np.array_equal(fused["status"], generic["status"])
    # same bar as against the oracle (cases.loop_tolerances): the RLS restart makes the first steps
    # of the update loops amplify rounding differences between the two arithmetic orders
    xa, ua, ul = cases.loop_tolerances("update" if case["update"] else "frozen")
    late = slice(3 * T // 4, T)
    assert np.abs(fused["log_x"] - generic["log_x"]).max() <= xa
    assert np.abs(fused["log_u"] - generic["log_u"]).max() <= ua
    assert np.abs(fused["log_u"][late] - generic["log_u"][late]).max() <= ul * max(1.0, np.abs(generic["log_u"][late]).max())
    for k in ("A", "B", "C"):
        scale = np.abs(generic[k]).max()
        np.testing.assert_allclose(fused[k], generic[k], rtol=0, atol=1e-4 * scale)
    lf, lg = fused["loop"], generic["loop"]
    np.testing.assert_allclose(lf.z.cpu().numpy(), lg.z.cpu().numpy(), rtol=0, atol=xa)
    if lf.rls is not None:
        for k in ("KA", "P", "barX", "barQ"):
            a, b = getattr(lf.rls, k).cpu().numpy(), getattr(lg.rls, k).cpu().numpy()
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-4 * np.abs(b).max())


def test_fused_kernel_ragged_tiles_and_single_steps():
    """S not a multiple of the 32-scenario tile; T = 1 launches equal one T = 40 launch bit for bit."""
    Ws, bs = H.oracle_weights("vdp")
    g = H.golden("ref_vanderpol.npz")
    enc = K.Encoder(Ws, bs)
    rs = np.random.default_rng(5)
    S, T = 77, 40
    x0 = rs.uniform(-1.5, 1.5, (S, 2))
    r = enc(np.stack([rs.uniform(-1, 1, S), np.zeros(S)], axis=1))
    one = K.ClosedLoop(K.vanderpol_spec(), x0, g["A"], g["B"], g["C"], r, encoder=enc, log_steps=T).run(T)
    many = K.ClosedLoop(K.vanderpol_spec(), x0, g["A"], g["B"], g["C"], r, encoder=enc, log_steps=T)
    for _ in range(T):
        many.run(1)
    torch.cuda.synchronize()
    assert np.array_equal(one.log_x.cpu().numpy(), many.log_x.cpu().numpy())
    assert np.array_equal(one.rls.P.cpu().numpy(), many.rls.P.cpu().numpy())
    assert one.step_index == T and int(one.status.max().item()) == 0


def test_fused_timed_phases_sum_below_launch_time():
    case = cases.loop_case("vdp")
    enc = K.Encoder(case["Ws"], case["bs"])
    loop = K.ClosedLoop(case["spec"], case["x0"], case["A"], case["B"], case["C"], case["r"], encoder=enc)
    loop.run(5)
    ms = loop.run_timed(20)
    assert all(v > 0 for v in ms.values()) and loop.step_index == 25


def test_reset_starts_an_identical_episode():
    case = cases.loop_case("vdp")
    enc = K.Encoder(case["Ws"], case["bs"])
    T = 80
    loop = K.ClosedLoop(case["spec"], case["x0"], case["A"], case["B"], case["C"], case["r"], encoder=enc, log_steps=T)
    assert loop.fused
    first = loop.run(T).log_x.clone()
    P1 = loop.rls.P.clone()
    loop.reset().run(T)
    assert loop.step_index == T
    assert torch.equal(first, loop.log_x) and torch.equal(P1, loop.rls.P)
    x1 = np.array([[0.3, -0.2], [1.0, 1.0], [-0.5, 0.7]])
    loop.reset(x1).run(T)
    fresh = K.ClosedLoop(case["spec"], x1, case["A"], case["B"], case["C"], case["r"], encoder=enc, log_steps=T).run(T)
    assert torch.equal(fresh.log_x, loop.log_x)


def test_plant_blow_up_matches_the_reference_arithmetic():
    """x0 near (-2, -1.3) with a far set-point drives |x1| past 2.4, where the reference's RK4 plant
    (h = 0.05) is numerically unstable: the oracle and the kernel must diverge the same way (same
    trajectory until the blow-up, non-finite within a step of each other, flagged, not aborted)."""
    Ws, bs = H.oracle_weights("vdp")
    g = H.golden("ref_vanderpol.npz")
    enc = K.Encoder(Ws, bs)
    x0 = np.array([[-1.9679426952760442, -1.2564560344690783], [0.5, 0.5]])
    xref = np.array([[0.8085404454242318, 0.0], [0.2, 0.0]])
    T = 90
    loop = K.ClosedLoop(K.vanderpol_spec(), x0, g["A"], g["B"], g["C"], enc(xref), encoder=enc, log_steps=T).run(T)
    lx, st = loop.log_x.cpu().numpy(), loop.status.cpu().numpy()
    with np.errstate(all="ignore"):
        o = ocl.run_loop(ocl.vanderpol_config(Ws, bs, xref[0]), g["A"], g["B"], g["C"], x0[0], T, update=ocl.UPDATE_RLS,
                         qp="exact")
    bad_gpu = int(np.argmax(~np.isfinite(lx[:, 0]).all(axis=1)))
    bad_ref = int(np.argmax(~np.isfinite(o["X"]).all(axis=1)))
    assert bad_gpu > 0 and abs(bad_gpu - bad_ref) <= 1
    assert np.abs(lx[:65, 0] - o["X"][:65]).max() < 1e-4
    assert st[0] & 2 and st[1] == 0 and np.isfinite(lx[:, 1]).all()      # the healthy neighbour is untouched


def test_fp64_peak_probe():
    dmma, dfma = K.measure_fp64_peak()
    assert 5.0 < dmma < 100.0 and 5.0 < dfma < 100.0


# ------------------------------------------------------------------------------ full sizes -----
def test_full_size_edmd_snapshots_linearity():
    """BASELINE config 4 shape on one GPU: 10 M duffing-like snapshots through the fused lift + Gram.
    Size-independent properties: Gram(all) == Gram(first part) + Gram(rest) (what the all-reduce
    relies on), the count, and A, B, C equal to the regression over a 10 k subsample within the
    statistical tolerance of the snapshot distribution."""
    Ws, bs = H.oracle_weights("duffing")
    enc = K.Encoder(Ws, bs)
    M = 10_000_000
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand((M, 2), generator=g, device="cuda", dtype=torch.float64) * 4 - 2
    u = torch.rand(M, generator=g, device="cuda", dtype=torch.float64) * 4 - 2
    y = K.plant.f_update(x, u, K.plant.DUFFING_PRE)
    whole = K.edmd.gram_from_snapshots(enc, x, y, u)
    cut = 3_333_337
    parts = K.edmd.gram_from_snapshots(enc, x[:cut], y[:cut], u[:cut])
    parts = K.edmd.gram_from_snapshots(enc, x[cut:], y[cut:], u[cut:], pack=parts)
    assert float(whole[-1].item()) == M and float(parts[-1].item()) == M
    np.testing.assert_allclose(parts.cpu().numpy(), whole.cpu().numpy(), rtol=1e-10)
    A, B, C, st = K.edmd.edmd_solve(whole, 8, 2)
    assert int(st.item()) == 0 and bool(torch.isfinite(A).all())
    sub = slice(0, 20000)
    xs, ys, us = x[sub].cpu().numpy(), y[sub].cpu().numpy(), u[sub].cpu().numpy()
    Ao, Bo, Co = oedmd.edmd_pinv(olift.encoder_forward(Ws, bs, xs).T, olift.encoder_forward(Ws, bs, ys).T,
                                 us.reshape(1, -1), xs.T)
    assert np.abs(A.cpu().numpy() - Ao).max() < 0.05 * np.abs(Ao).max()
    assert np.abs(C.cpu().numpy() - Co).max() < 0.05 * np.abs(Co).max()


def test_full_size_tank_shard_properties():
    """BASELINE config 3 shape: 65 536 tank scenarios (velocity form, N = 20) for a few steps on the
    generic kernels: input and rate bounds respected, levels non-negative, finite, status clean,
    and a shard of the batch reproduces its slice bit for bit."""
    t = cases.tank_setup()
    enc = K.Encoder(t["Ws"], t["bs"])
    rs = np.random.default_rng(3)
    S, T = 65536, 6
    x0 = np.maximum(rs.uniform(0, 2, (S, 2)), 0.0)
    loop = K.ClosedLoop(K.tank_spec(), x0, t["A"], t["B"], t["C"], np.array([1.0]), encoder=enc, log_steps=T).run(T)
    assert not loop.fused
    lx, lu, st = loop.log_x.cpu().numpy(), loop.log_u.cpu().numpy(), loop.status.cpu().numpy()
    assert np.isfinite(lx).all() and np.isfinite(lu).all()
    assert lx.min() >= 0.0
    # no iteration cap, nothing non-finite; the rank-1 model of the very first RLS steps can make
    # the Delta-u Hessian numerically semi-definite (cond ~ 2e16), which the
    # kernel flags as KMPC_STATUS_PIVOT instead of aborting: rare, and never fatal
    assert ((st & 3) == 0).all(), np.unique(st)
    assert (st != 0).mean() < 0.05, (st != 0).mean()
    du = np.abs(np.diff(np.concatenate([np.zeros((1, S)), lu]), axis=0))
    assert du.max() <= 0.5 + 1e-9 and np.abs(lu).max() <= 8.0 + 1e-9
    sl = slice(40000, 40000 + 4099)
    part = K.ClosedLoop(K.tank_spec(), x0[sl], t["A"], t["B"], t["C"], np.array([1.0]), encoder=enc, log_steps=T).run(T)
    assert np.array_equal(part.log_x.cpu().numpy(), lx[:, sl])
    # nearly empty tanks: at step 5 the restarted RLS model makes the Hessian numerically singular;
    # the kernel and the oracle apply the same relative pivot floor and stay together
    xs = np.array([[0.06378558970830861, 0.003204316873218982], [0.026345235171531645, 0.0067565715286719286]])
    T2 = 60
    hard = K.ClosedLoop(K.tank_spec(), xs, t["A"], t["B"], t["C"], np.array([1.0]), encoder=enc, log_steps=T2).run(T2)
    hx, hs = hard.log_x.cpu().numpy(), hard.status.cpu().numpy()
    assert np.isfinite(hx).all() and ((hs & 3) == 0).all()
    for k in range(2):
        o = ocl.run_loop(t["cfg"], t["A"], t["B"], t["C"], xs[k], T2, update=ocl.UPDATE_RLS, qp="exact")
        assert np.abs(o["X"] - hx[:, k]).max() < 5e-3, k
        assert np.abs(o["X"][40:] - hx[40:, k]).max() < 1e-5, k


def test_full_size_rbf_horizon50_properties():
    """BASELINE config 5 shape (per-GPU shard): 125 000 RBF-lifted duffing scenarios, horizon 50,
    warm-started update, a few steps: bounds, finiteness, status, and scenario 0 against the oracle."""
    g = H.golden("ref_duffing_rbf.npz")
    X, Y, U = oplant.generate_snapshots(100, 100, oplant.DUFFING_PRE, np.random.RandomState(101))
    PX, PY = olift.rbf_lift(X.T, g["cx"]).T, olift.rbf_lift(Y.T, g["cx"]).T
    G, Aq, XV = oedmd.gram_pack(PX, PY, U, X)
    rs = np.random.default_rng(9)
    S, T = 125000, 4
    x0 = rs.uniform(-2, 2, (S, 2))
    x0[0] = [-2.0, -2.0]
    warm = K.RLSState.warm(S, G, Aq, XV[:, :8], G[:8, :8])
    spec = K.rbf_spec(N=50)
    loop = K.ClosedLoop(spec, x0, g["A"], g["B"], g["C"], np.array([1.0, 0.0]), cx=g["cx"], rls_state=warm,
                        log_steps=T).run(T)
    lx, lu, st = loop.log_x.cpu().numpy(), loop.log_u.cpu().numpy(), loop.status.cpu().numpy()
    assert np.isfinite(lx).all() and (st == 0).all()
    assert lu.min() >= -2.0 - 1e-12 and lu.max() <= 2.0 + 1e-12
    cfg = ocl.rbf_config(g["cx"])
    cfg.N = 50
    o = ocl.run_loop(cfg, g["A"], g["B"], g["C"], x0[0], T, update=ocl.UPDATE_RLS, qp="exact",
                     warm=orls.RLSState.warm(G, Aq, XV[:, :8], G[:8, :8]))
    assert np.abs(o["X"] - lx[:, 0]).max() < 1e-7


def test_generic_qp_start_modes_reach_the_same_minimiser():
    """The QP has ONE minimiser (strictly convex): the generic kernels' start modes -- warm + sweeps (0),
    cold like the reference (1), warm + primal only (2), warm + damped sweeps (3, tank_spec's default) --
    differ in the path, not in the answer.  Horizon-50 RBF loop (frozen model and warm-started update,
    well conditioned): identical trajectories to 1e-9; and the damped mode against the oracle."""
    g = H.golden("ref_duffing_rbf.npz")
    X, Y, U = oplant.generate_snapshots(100, 100, oplant.DUFFING_PRE, np.random.RandomState(101))
    PX, PY = olift.rbf_lift(X.T, g["cx"]).T, olift.rbf_lift(Y.T, g["cx"]).T
    G, Aq, XV = oedmd.gram_pack(PX, PY, U, X)
    rs = np.random.default_rng(11)
    S, T = 96, 25
    x0 = rs.uniform(-2, 2, (S, 2))
    for update in (False, True):
        logs = {}
        for mode in (0, 1, 2, 3):
            warm = K.RLSState.warm(S, G, Aq, XV[:, :8], G[:8, :8]) if update else None
            loop = K.ClosedLoop(K.rbf_spec(N=50, update=update, qp_cold=mode), x0, g["A"], g["B"], g["C"],
                                np.array([1.0, 0.0]), cx=g["cx"], rls_state=warm, log_steps=T).run(T)
            assert not loop.fused and int(loop.status.max().item()) == 0
            logs[mode] = (loop.log_x.cpu().numpy(), loop.log_u.cpu().numpy())
        for mode in (1, 2, 3):
            assert np.abs(logs[mode][0] - logs[0][0]).max() < 1e-9, (update, mode)
            assert np.abs(logs[mode][1] - logs[0][1]).max() < 1e-8, (update, mode)
    cfg = ocl.rbf_config(g["cx"])
    cfg.N = 50
    for s in range(8):   # the horizon-50 kernel factors with the DMMA panel pre-pass: held to the oracle directly
        o = ocl.run_loop(cfg, g["A"], g["B"], g["C"], x0[s], T, update=ocl.UPDATE_RLS, qp="exact",
                         warm=orls.RLSState.warm(G, Aq, XV[:, :8], G[:8, :8]))
        for mode in (0, 3):
            assert np.abs(o["X"] - logs[mode][0][:, s]).max() < 1e-7, (mode, s)
            assert np.abs(o["U"] - logs[mode][1][:, s]).max() < 1e-6, (mode, s)


# ------------------------------------------- snapshot generator + open-loop predictor (N1, N2) --
@pytest.mark.parametrize("name,seed", [("duffing", 33), ("vanderpol", 50)])
def test_snapshot_generator_matches_oracle_and_reference_run(name, seed):
    """kmpc_generate_snapshots through the reference's call shape (data_generate.generate)."""
    from koopman_online_updated_mpc_b200 import data_generate as DG
    g = H.golden("ref_%s_predict.npz" % name)
    np.random.seed(seed)
    gen = DG.generate(100, 100)
    X, Y, U = gen.duffing_generate() if name == "duffing" else gen.vanderpol_generate()
    assert X.shape == (2, 10000) and Y.shape == (2, 10000) and U.shape == (1, 10000)
    np.testing.assert_allclose(X[:, :600], g["X_head"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(Y[:, :600], g["Y_head"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(X[:, -300:], g["X_tail"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(U[:, -300:], g["U_tail"], rtol=0, atol=0)
    p = oplant.DUFFING_PRE if name == "duffing" else oplant.VDP_PRE
    Xo, Yo, Uo = oplant.generate_snapshots(100, 100, p, np.random.RandomState(seed))
    np.testing.assert_allclose(X, Xo, rtol=0, atol=1e-12)
    np.testing.assert_allclose(Y, Yo, rtol=0, atol=1e-12)
    assert np.array_equal(U, Uo)


def test_snapshot_generator_ragged_shapes_and_variants():
    from koopman_online_updated_mpc_b200 import data_generate as DG
    rs = np.random.RandomState(3)
    for n_traj, n_step in ((1, 1), (127, 7), (129, 8), (300, 19), (5, 100)):   # ragged tiles / chunks
        x0 = rs.uniform(-1.5, 1.5, (n_traj, 2))
        u0 = rs.uniform(-2, 2, (n_step, n_traj))
        for variant in (oplant.RK4_PYTHON, oplant.RK4_MATLAB):
            X, Y, U = DG.generate_snapshots(x0, u0, oplant.VDP_PRE, rk4_variant=variant)
            x, Xs, Ys = x0.copy(), [], []
            for j in range(n_step):
                xn = oplant.rk4_step(x, u0[j], np.asarray(oplant.VDP_PRE), variant=variant)
                Xs.append(x)
                Ys.append(xn)
                x = xn
            np.testing.assert_allclose(X.cpu().numpy(), np.stack(Xs, 1).reshape(-1, 2), rtol=0, atol=1e-12)
            np.testing.assert_allclose(Y.cpu().numpy(), np.stack(Ys, 1).reshape(-1, 2), rtol=0, atol=1e-12)
            assert np.array_equal(U.cpu().numpy(), u0.T.reshape(-1))
    x0 = np.abs(rs.uniform(0, 2, (40, 2)))        # tank map (Tank_System.m:9-10, clamp at 0)
    u0 = rs.uniform(-1, 3, (30, 40))
    X, Y, U = DG.generate_snapshots(x0, u0, oplant.TANK_PRE, kind=oplant.PLANT_TANK)
    x = x0.copy()
    for j in range(30):
        xn = oplant.tank_step(x, u0[j], np.asarray(oplant.TANK_PRE))
        np.testing.assert_allclose(Y.cpu().numpy().reshape(40, 30, 2)[:, j], xn, rtol=0, atol=1e-12)
        x = xn
    Xe, Ye, Ue = DG.generate_snapshots(np.zeros((0, 2)), np.zeros((5, 0)), oplant.VDP_PRE)   # empty set
    assert Xe.shape == (0, 2) and Ue.shape == (0,)


@pytest.mark.parametrize("name,wsys,row", [("duffing", "duffing", 0), ("vanderpol", "vdp", 1)])
def test_open_loop_predictor_matches_reference_run(name, wsys, row):
    """kmpc_open_loop_predict against the reference's own predictor outputs and RMSE
    (duffing.py:290-343 / vanderpol.py:292-348) and the oracle on many sequences."""
    from koopman_online_updated_mpc_b200 import predict as P
    from oracle import predict as opredict
    g = H.golden("ref_%s_predict.npz" % name)
    Ws, bs = H.oracle_weights(wsys)
    enc = K.Encoder(Ws, bs)
    T = int(g["plotTime"])
    X, U = g["X_head"], g["U_head"]
    PHIX = enc(X.T.copy()).T
    tY, dX, rmse = P.open_loop_predict(PHIX, X, U, g["A"], g["B"], g["C"], T, rmse_row=row)
    np.testing.assert_allclose(tY, g["test_Y"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(dX, g["decoder_X"], rtol=0, atol=1e-11)
    assert abs(rmse - float(g["RMSE"])) < 1e-13
    # many sequences, ragged last segment (T not a multiple of reset_every), other reset period
    T2, stride, nseq = 37, 50, 12
    tYs, dXs, rm = P.open_loop_predict(PHIX, X, U, g["A"], g["B"], g["C"], T2, reset_every=7, rmse_row=row,
                                       n_seq=nseq, seq_stride=stride)
    for s in range(nseq):
        Xs, Us = X[:, s * stride:], U[:, s * stride:]
        oY, oD, _ = opredict.open_loop_predict(lambda x: olift.lift_mlp(Ws, bs, x), g["A"], g["B"], g["C"], Xs, Us, T2,
                                               reset_every=7)
        np.testing.assert_allclose(tYs[s], oY, rtol=0, atol=1e-11)
        np.testing.assert_allclose(dXs[s], oD, rtol=0, atol=1e-11)
        assert abs(rm[s] - opredict.rmse(oY, Xs, T2, row)) < 1e-13


def test_gram_from_trajectories_equals_gram_from_snapshots():
    """One encode per state on consecutive trajectory-major snapshots == two encodes per snapshot."""
    from koopman_online_updated_mpc_b200 import data_generate as DG, edmd as E
    rs = np.random.RandomState(7)
    for wsys, n_traj, n_step in (("duffing", 37, 11), ("tank", 64, 20), ("duffing", 3000, 100)):
        enc = K.Encoder.from_file(H.weights_path(wsys))
        x0 = rs.uniform(0.1, 1.5, (n_traj, 2))
        u0 = rs.uniform(-1, 1, (n_step, n_traj))
        X, Y, U = DG.generate_snapshots(x0, u0, oplant.DUFFING_PRE)
        a = E.gram_from_snapshots(enc, X, Y, U).cpu().numpy()
        b = E.gram_from_trajectories(enc, X, Y, U, n_step, verify=True).cpu().numpy()
        assert a[-1] == b[-1] == n_traj * n_step
        np.testing.assert_allclose(b, a, rtol=1e-12, atol=1e-12 * np.abs(a).max())
    with pytest.raises(ValueError):     # shuffled snapshots are not consecutive: refused when verified
        perm = torch.randperm(X.shape[0], device=X.device)
        E.gram_from_trajectories(enc, X[perm], Y[perm], U[perm], n_step, verify=True)
