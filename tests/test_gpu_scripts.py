"""The reference's driver scripts end to end through the package (scripts.run_*), the closed loops of
the two `Koopman_update*.m` scripts, per-step teacher-forced parity of the Tank loop and of the
benchmarked 4096-scenario VDP batch.  Run on the B200 box with `-m gpu`.

Tolerances: A, B, C from the package's own EDMD vs the reference run 1e-9 relative; closed loops from
those matrices vs the reference's logged trajectories at the bars of test_gpu_parity.py (the
reference's L-BFGS-B answers carry ~5e-5 of solver noise); teacher-forced single
steps 1e-7 (controls / states) and 1e-6 relative (Koopman matrices)."""
from dataclasses import replace

import numpy as np
import pytest
import torch

import cases
import helpers as H
import koopman_online_updated_mpc_b200 as K
from koopman_online_updated_mpc_b200 import scripts
from oracle import closed_loop as ocl
from oracle import rls as orls

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(np.asarray(a).reshape(np.shape(b)) - b).max() / np.abs(b).max()


# ------------------------------------------------------------------ python scripts end to end ---
@pytest.mark.parametrize("system", ["duffing", "vanderpol"])
def test_script_end_to_end_matches_the_reference_run(system, tmp_path):
    """duffing.py:57 -> 1015 / vanderpol.py:57 -> 1112 through the package: weights file ->
    data_generate (seed 101 / 50) -> theta_E -> EDMD -> frozen loop -> update loop -> .mat files,
    against tests/golden/ref_<system>.npz (the reference script's own run in the build container)."""
    g = H.golden("ref_%s.npz" % system)
    T = int(g["maxStep"])
    wsys = "duffing" if system == "duffing" else "vdp"
    Ws, bs = H.oracle_weights(wsys)
    if system == "duffing":
        out = scripts.run_duffing(H.weights_path(wsys), max_step=T, save_dir=str(tmp_path))
        for k in ("A", "B", "C"):
            assert _rel(out[k], g[k]) < 1e-9, (k, _rel(out[k], g[k]))
        assert np.abs(out["X"][:, :300] - g["X_head"]).max() < 1e-12      # the reference's snapshot set
    else:
        # vanderpol.py draws its EDMD snapshots unseeded (l.17 commented out): the package's own
        # identification is checked against the oracle's pinv form on the same draws, the closed
        # loops are replayed from the reference run's logged matrices
        own = scripts.run_vanderpol(H.weights_path(wsys), max_step=5, seed=7)
        from oracle import edmd as oedmd, lift as olift
        Ao, Bo, Co = oedmd.edmd_pinv(olift.encoder_forward(Ws, bs, own["X"].T).T, olift.encoder_forward(Ws, bs, own["Y"].T).T,
                                     own["U"], own["X"])
        for k, want in (("A", Ao), ("B", Bo), ("C", Co)):
            assert _rel(own[k], want) < 1e-9, (k, _rel(own[k], want))
        out = scripts.run_vanderpol(H.weights_path(wsys), max_step=T, save_dir=str(tmp_path), model=(g["A"], g["B"], g["C"]))
    # update loop: the reference's own floor is 4e-5
    assert np.abs(out["logXloc"][0] - g["logXloc"][:, :T]).max() < 2e-4
    # frozen loop: L-BFGS-B noise accumulates in the steady-state offset (1.3e-2 between scipy builds)
    assert np.abs(out["logX"][0] - g["logX"][:, :T]).max() < 2e-2
    assert np.abs(out["logX"][0][:, :60] - g["logX"][:, :60]).max() < 2e-4
    assert (out["status_frozen"] == 0).all() and (out["status_update"] == 0).all()
    # RLS end state of the update loop vs the reference's (duffing.py:927-953)
    assert _rel(out["K_A"][0], g["K_A"]) < 1e-3 and _rel(out["Aloc"][0], g["Aloc"]) < 1e-2
    # files in the reference's layouts, readable by its loaders
    import scipy.io as sio
    nn = sio.loadmat(str(tmp_path / "NN_Encoder.mat"))
    assert nn["X_Collection"].shape == (2, T) and nn["X_Collection_NO"].shape == (2, T) and nn["U_Collection"].shape == (1, T)
    W2, _ = K.weights.load_encoder_weights(str(tmp_path / "model_weights.mat"))
    assert all(np.array_equal(a, b) for a, b in zip(W2, Ws))
    # the same chain against the oracle from the matrices the loops used (exact QP): tight
    cfg = ocl.duffing_config(Ws, bs) if system == "duffing" else ocl.vanderpol_config(Ws, bs)
    Al, Bl, Cl = (out["A"], out["B"], out["C"]) if system == "duffing" else (g["A"], g["B"], g["C"])
    o = ocl.run_loop(cfg, Al, Bl, Cl, np.array([-2.0, -2.0]), 120, update=ocl.UPDATE_NONE, qp="exact")
    assert np.abs(o["X"].T - out["logX"][0][:, :120]).max() < 1e-8


def test_script_with_the_tensor_core_lift():
    """Same duffing.py chain with the EDMD-side lift on tcgen05 (KMPC_PREC_TC): Koopman matrices
    within 1e-4 relative of the reference run, the closed loops (fp64 lift) follow."""
    g = H.golden("ref_duffing.npz")
    out = scripts.run_duffing(H.weights_path("duffing"), max_step=120, precision=K.lift.PREC_TC)
    for k in ("A", "B", "C"):
        assert _rel(out[k], g[k]) < 1e-4, (k, _rel(out[k], g[k]))
    assert np.abs(out["logX"][0][:, :60] - g["logX"][:, :60]).max() < 1e-3


@pytest.mark.parametrize("system", ["duffing", "vanderpol"])
def test_rbf_script_matches_the_reference_run(system):
    """duffing_RBF.py / vanderpol_RBF.py (tests/golden/ref_*_rbf.npz; centres = the reference run's
    KMeans centres): EDMD matrices, frozen loop (vanderpol_RBF.py:328 switches to the DUFFING
    post-plant) and the 'storage method' update loop (l.434-438, VDP post-plant l.506)."""
    g = H.golden("ref_%s_rbf.npz" % system)
    T = int(g["maxStep"])
    out = scripts.run_rbf(g["cx"], system=system, max_step=T)
    for k in ("A", "B", "C"):
        assert _rel(out[k], g[k]) < 1e-8, (k, _rel(out[k], g[k]))
    assert np.abs(out["logX"][0] - g["logX"][:, :T]).max() < 2e-3
    assert np.abs(out["logX"][0][:, :60] - g["logX"][:, :60]).max() < 2e-4
    assert np.abs(out["logXloc"][0] - g["logXloc"][:, :T]).max() < 2e-4
    assert (out["status_update"] == 0).all()


# ------------------------------------------------------------------ MATLAB scripts -------------
def _gpu_loop(spec, m, x0, T, enc, warm=None):
    loop = K.ClosedLoop(spec, x0, m["A"], m["B"], m["C"], m["cfg"].r, encoder=enc, rls_state=warm, log_steps=T).run(T)
    torch.cuda.synchronize()
    return loop


def test_koopman_update_m_closed_loop():
    """Revise_2/Koopman_update.m:67-70 (stacked lift, nz = 10), :130-142,185-188 (Q = 10, R = 0.01,
    +-2), :258-278 (warm start from the offline Gram, lambda, C frozen): GPU loop vs the oracle, and
    the whole script (scripts.run_koopman_update: GPU data generation, lift, Gram, solve) vs the
    oracle's identification on the same draws."""
    m = cases.matlab_offline("koopman_update")
    enc = K.Encoder(m["Ws"], m["bs"])
    nz, T = m["nz"], 100
    x0 = np.array([[-1.0, 1.0], [0.5, -0.5], [1.5, 1.0], [-2.0, -2.0]])
    spec = K.LoopSpec(nz=nz, out_mode=K.closed_loop.OUT_C, lift_mode=K.lift.LIFT_STACK, rk4_variant=K.plant.RK4_MATLAB,
                      first_post_step=1 << 30, q=10.0, rw=0.01, lb=-2.0, ub=2.0, lam=1.0, update_c=False)
    for lam in (1.0, 0.98):
        warm = K.RLSState.warm(len(x0), m["G"], m["Aq"], m["XV"][:, :nz], m["G"][:nz, :nz])
        loop = _gpu_loop(replace(spec, lam=lam), m, x0, T, enc, warm)
        assert not loop.fused                                   # nz = 10: generic kernels
        lx, lu = loop.log_x.cpu().numpy(), loop.log_u.cpu().numpy()
        cfg = replace(m["cfg"], lam=lam)
        # lambda = 1 (the value in the file): the free-running loops agree to 1e-7 over all 100 steps.
        # lambda < 1 winds the covariance up (P ~ lambda^-k in the unexcited directions): a 1e-16
        # perturbation of P grows to 8e-8 by step 98 in the oracle itself, so the free-running bar is
        # 1e-7 over the first 40 steps and 1e-3 overall, and every step is checked teacher-forced below
        n_tight = T if lam == 1.0 else 40
        for s in range(len(x0)):
            w = orls.RLSState.warm(m["G"], m["Aq"], m["XV"][:, :nz], m["G"][:nz, :nz])
            o = ocl.run_loop(cfg, m["A"], m["B"], m["C"], x0[s], T, update=ocl.UPDATE_RLS, qp="exact", warm=w)
            assert np.abs(o["X"] - lx[:, s])[:n_tight].max() < 1e-7, (lam, s)
            assert np.abs(o["U"] - lu[:, s])[:n_tight].max() < 1e-6, (lam, s)
            assert np.abs(o["X"] - lx[:, s]).max() < 1e-3, (lam, s)
            if lam == 1.0:
                assert _rel(loop.A[s].cpu().numpy(), o["A"]) < 1e-6
        assert int(loop.status.max().item()) == 0
        w = orls.RLSState.warm(m["G"], m["Aq"], m["XV"][:, :nz], m["G"][:nz, :nz])
        got, want, _ = _teacher_forced(cfg, m["A"], m["B"], m["C"], x0[0], T, replace(spec, lam=lam), enc, cfg.r, warm=w)
        assert np.abs(got["u"] - want["u"]).max() < 1e-7 and np.abs(got["x"] - want["x"]).max() < 1e-9
        assert np.all(np.abs(got["A"] - want["A"]) <= 1e-6 * np.abs(want["A"]).reshape(len(want["A"]), -1).max(axis=1).reshape(-1, 1, 1))
        assert torch.equal(loop.C.cpu(), torch.from_numpy(np.broadcast_to(m["C"], (len(x0), 2, nz)).copy()))
    out = scripts.run_koopman_update(H.weights_path("duffing"), max_step=T, seed=m["seed"])
    for k in ("A", "B", "C"):
        assert _rel(out[k], m[k]) < 1e-8, (k, _rel(out[k], m[k]))
    w = orls.RLSState.warm(m["G"], m["Aq"], m["XV"][:, :nz], m["G"][:nz, :nz])
    o = ocl.run_loop(m["cfg"], m["A"], m["B"], m["C"], np.array([-1.0, 1.0]), T, update=ocl.UPDATE_RLS, qp="exact", warm=w)
    assert np.abs(o["X"].T - out["logXloc"][0]).max() < 1e-6


def test_tracking_lift_m_closed_loop():
    """VDP_Revise_2/Koopman_update_Tracking_Lift.m:65 (offset lift), :99 (C = I), :108-113 (Q = 100 I_8,
    R = 1e-4, N = 10), :151 (+-6), :181-195 (P0 = 1e5 I restart): the fused kernel serves this loop."""
    m = cases.matlab_offline("tracking_lift")
    enc = K.Encoder(m["Ws"], m["bs"])
    T = 160
    x0 = np.array([[1.0, 1.0], [0.3, -0.4], [-1.2, 0.5]])
    spec = K.vanderpol_spec(lift_mode=K.lift.LIFT_OFFSET, rk4_variant=K.plant.RK4_MATLAB, first_post_step=100,
                            update_c=False)
    runs = {}
    for name, path in (("fused", K.closed_loop.PATH_AUTO), ("generic", K.closed_loop.PATH_GENERIC)):
        loop = _gpu_loop(replace(spec, path=path), m, x0, T, enc)
        assert loop.fused == (name == "fused")
        runs[name] = (loop.log_x.cpu().numpy(), loop.log_u.cpu().numpy(), loop.status.cpu().numpy())
    xa, ua, ul = cases.loop_tolerances("update")
    late = slice(3 * T // 4, T)
    for name, (lx, lu, st) in runs.items():
        assert (st == 0).all()
        for s in range(len(x0)):
            o = ocl.run_loop(m["cfg"], m["A"], m["B"], m["C"], x0[s], T, update=ocl.UPDATE_RLS, qp="exact")
            assert np.abs(o["X"] - lx[:, s]).max() <= xa, (name, s)
            assert np.abs(o["U"] - lu[:, s]).max() <= ua, (name, s)
            # the plant switch at step 100 starts a second transient (the input oscillates while the RLS
            # re-identifies): the late window of this loop gets the transient bar, single steps 1e-7 below
            assert np.abs(o["U"][late] - lu[late, s]).max() <= ua, (name, s)
    for x in x0[:2]:
        got, want, _ = _teacher_forced(m["cfg"], m["A"], m["B"], m["C"], x, T, spec, enc, m["cfg"].r)
        assert np.abs(got["u"] - want["u"]).max() < 1e-7 and np.abs(got["x"] - want["x"]).max() < 1e-9
        assert np.abs(got["z"] - want["z_next"]).max() < 1e-9
        assert np.all(np.abs(got["A"] - want["A"]) <= 1e-6 * np.maximum(np.abs(want["A"]).reshape(len(want["A"]), -1).max(axis=1), 1e-3).reshape(-1, 1, 1))
    out = scripts.run_tracking_lift(H.weights_path("vdp"), max_step=T, seed=m["seed"])
    for k in ("A", "B"):
        assert _rel(out[k], m[k]) < 1e-8, (k, _rel(out[k], m[k]))
    assert np.abs(out["logXloc"][0].T - runs["fused"][0][:, 0]).max() < 1e-6


# ------------------------------------------------------------------ teacher-forced single steps --
def _teacher_forced(cfg, A, B, C, x0, T, spec, enc, r, warm=None):
    """Oracle trajectory with every pre-step state recorded -> ONE GPU batch of T - 1 single-step
    problems (steps 1 .. T - 1, the RLS already running) -> per-step comparison."""
    o = ocl.run_loop(cfg, A, B, C, x0, T, update=ocl.UPDATE_RLS, qp="exact", record_states=True, record_models=True,
                     warm=warm)
    ps = o["pre_states"][1:]
    S = len(ps)
    nz = cfg.nz
    warm = K.RLSState(S, nz, 2)
    for k in ("KA", "P", "barX", "barQ"):
        getattr(warm, k).copy_(torch.from_numpy(np.array([getattr(p["rls"], k) for p in ps])))
    params = np.array([cfg.p_pre if p["k"] < cfg.first_post_step else cfg.p_post for p in ps])
    loop = K.ClosedLoop(spec, np.array([p["x"] for p in ps]), np.array([p["A"] for p in ps]),
                        np.array([p["B"] for p in ps]), np.array([p["C"] for p in ps]), r, encoder=enc,
                        rls_state=warm, log_steps=1, params_pre=params, params_post=params,
                        u_prev=np.array([p["u_prev"] for p in ps]))
    loop.run(1)
    torch.cuda.synchronize()
    got = dict(u=loop.log_u[0].cpu().numpy(), x=loop.log_x[0].cpu().numpy(), z=loop.z.cpu().numpy(),
               A=loop.A.cpu().numpy(), B=loop.B.cpu().numpy(), C=loop.C.cpu().numpy(), status=loop.status.cpu().numpy())
    want = dict(u=o["U"][1:], x=o["X"][1:], A=np.array([mm[0] for mm in o["models"][1:]]),
                B=np.array([mm[1] for mm in o["models"][1:]]), C=np.array([mm[2] for mm in o["models"][1:]]),
                z_next=np.concatenate([o["Z"][2:], cfg.lift_fn(o["X"][-1])[None]]))
    return got, want, o


def test_tank_teacher_forced_single_steps():
    """Tank_System.m:170-291, every step k = 1..219 of the oracle's run from identical states: the
    free-running comparison of the Tank loop is only good to 5e-3 while cond(H) ~ 1e16 in the RLS
    transient; step by step the kernels hold 1e-7 on the applied input and 1e-9 on the state."""
    t = cases.tank_setup()
    enc = K.Encoder(t["Ws"], t["bs"])
    for x0 in (np.array([0.0, 0.0]), np.array([0.5, 1.5])):
        got, want, o = _teacher_forced(t["cfg"], t["A"], t["B"], t["C"], x0, 220, K.tank_spec(), enc, np.array([1.0]))
        # steps whose Hessian is numerically singular (rank-deficient restarted model: pivot floor
        # applied by both sides) are flagged by the kernel; everywhere else the bar is 1e-7
        clean = (got["status"] & K.mpc.STATUS_PIVOT) == 0
        assert clean.mean() > 0.9 and ((got["status"] & 3) == 0).all()
        assert np.abs(got["u"] - want["u"])[clean].max() < 1e-7
        assert np.abs(got["x"] - want["x"])[clean].max() < 1e-9
        assert np.abs(got["u"] - want["u"]).max() < 5e-3        # flagged steps: the loose bar
        assert np.abs(got["z"] - want["z_next"]).max() < 1e-9
        for k in ("A", "B", "C"):
            w = want[k].reshape(len(want[k]), -1)
            gk = got[k].reshape(w.shape)
            scale = np.maximum(np.abs(w).max(axis=1, keepdims=True), 1e-3)
            assert np.all(np.abs(gk - w)[clean] <= 1e-6 * scale[clean]), k


def test_bench_batch_teacher_forced_over_the_full_horizon():
    """The BENCHMARKED regime (bench.py: S = 4096, T = 400, seed 20240601, x0 ~ U[-2,2]^2, set-points
    U[-1,1]): 64 scenarios -- the ones that chatter between the input bounds hardest, every scenario
    the GPU flags, and a random rest -- are followed by the oracle over all 400 steps; each of their
    steps is re-run on the GPU from the oracle's state (teacher forcing) and compared at 1e-7, and
    the scenarios whose RK4 plant blows up must blow up at the same step on both sides."""
    Ws, bs = H.oracle_weights("vdp")
    g = H.golden("ref_vanderpol.npz")
    enc = K.Encoder(Ws, bs)
    rs = np.random.default_rng(20240601)
    S, T = 4096, 400
    x0 = rs.uniform(-2, 2, (S, 2))
    xref = np.stack([rs.uniform(-1, 1, S), np.zeros(S)], axis=1)
    r = enc(xref)
    spec = K.vanderpol_spec()
    loop = K.ClosedLoop(spec, x0, g["A"], g["B"], g["C"], r, encoder=enc, log_steps=T).run(T)
    lx, lu, st = loop.log_x.cpu().numpy(), loop.log_u.cpu().numpy(), loop.status.cpu().numpy()
    flagged = np.nonzero(st != 0)[0]
    with np.errstate(all="ignore"):
        chatter = np.nan_to_num(np.abs(np.diff(lu[150:], axis=0)).sum(axis=0), nan=0.0, posinf=0.0)
    chatter[flagged] = 0.0
    pick = list(flagged[:16]) + list(np.argsort(-chatter)[:24])
    rest = [int(s) for s in np.random.default_rng(1).permutation(S) if s not in set(pick)]
    pick = [int(s) for s in pick] + rest[:64 - len(pick)]
    assert len(pick) == 64
    r_np = np.asarray(r)
    n_blow = n_steps = n_loose = 0
    for s in pick:
        cfg = ocl.vanderpol_config(Ws, bs, xref[s])
        with np.errstate(all="ignore"):
            o = ocl.run_loop(cfg, g["A"], g["B"], g["C"], x0[s], T, update=ocl.UPDATE_RLS, qp="exact",
                             record_states=True, record_models=True)
        fin_o = np.isfinite(o["X"]).all(axis=1)
        fin_g = np.isfinite(lx[:, s]).all(axis=1)
        if not fin_o.all() or not fin_g.all():          # RK4 left the reals: same step on both sides
            n_blow += 1
            bo, bg = int(np.argmin(fin_o)), int(np.argmin(fin_g))
            assert not fin_o.all() and not fin_g.all(), s
            assert abs(bo - bg) <= 1, (s, bo, bg)
            assert st[s] & K.mpc.STATUS_NONFINITE
            n_ok = max(min(bo, bg) - 2, 2)
            n_model = max(n_ok - 10, 2)       # the last steps before the blow-up: |x| >> 1, the model is garbage
        else:
            assert st[s] == 0, (s, st[s])
            n_ok = n_model = T
        ps = o["pre_states"][1:n_ok]
        B_ = len(ps)
        warm = K.RLSState(B_, 8, 2)
        for k in ("KA", "P", "barX", "barQ"):
            getattr(warm, k).copy_(torch.from_numpy(np.array([getattr(p["rls"], k) for p in ps])))
        params = np.array([cfg.p_pre if p["k"] < cfg.first_post_step else cfg.p_post for p in ps])
        one = K.ClosedLoop(spec, np.array([p["x"] for p in ps]), np.array([p["A"] for p in ps]),
                           np.array([p["B"] for p in ps]), np.array([p["C"] for p in ps]),
                           np.broadcast_to(r_np[s], (B_, 8)).copy(), encoder=enc, rls_state=warm, log_steps=1,
                           params_pre=params, params_post=params, u_prev=np.array([p["u_prev"] for p in ps]))
        one.run(1)
        gu, gx = one.log_u[0].cpu().numpy(), one.log_x[0].cpu().numpy()
        # controls: 1e-7 on (almost) every step; where the input leaves a bound the free block of the QP
        # is nearly singular and the two exact solvers differ by up to ~1e-5 (measured 1.1e-5 on one of
        # ~25 000 steps): those steps are held to the north-star bound (1e-4) and counted
        du = np.abs(gu - o["U"][1:n_ok])
        assert du.max() < 1e-4, (s, float(du.max()))
        n_steps += len(du)
        n_loose += int((du > 1e-7).sum())
        dx = np.abs(gx - o["X"][1:n_ok]).max(axis=1)
        assert np.all(dx <= 1e-9 * max(1.0, np.abs(o["X"][1:n_ok]).max()) + 0.1 * du), s   # h = 0.05: dx ~ h du
        Aw = np.array([mm[0] for mm in o["models"][1:n_model]])
        scale = np.maximum(np.abs(Aw).reshape(len(Aw), -1).max(axis=1), 1e-3).reshape(-1, 1, 1)
        dA = np.abs(one.A.cpu().numpy()[:len(Aw)] - Aw) / scale
        # healthy scenarios: 1e-6 relative on the Koopman matrix (north_star: 1e-4); scenarios on their way to
        # the RK4 blow-up carry |x1| > 2 lifts through the P0 = 1e5 restart: 1e-5 (measured 1.1e-6)
        assert dA.max() <= (1e-6 if n_ok == T else 1e-5), (s, float(dA.max()), int(dA.reshape(len(Aw), -1).max(axis=1).argmax()), n_ok)
        one.close()
    assert n_blow == min(len(flagged), 16)
    assert n_steps > 20000 and n_loose <= 1e-3 * n_steps, (n_steps, n_loose)


def test_examples_run_end_to_end(tmp_path):
    """examples/*.py: the reference's scripts through the package, as a user would start them."""
    import os
    import subprocess
    import sys
    for name, args in (("duffing", ["--steps", "40", "--scenarios", "3"]), ("vanderpol", ["--steps", "40", "--tc"]),
                       ("tank", ["--steps", "30", "--scenarios", "2"])):
        r = subprocess.run([sys.executable, os.path.join(H.ROOT, "examples", name + ".py"), "--out", str(tmp_path)] + args,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        # Tank: the rank-deficient restarted model makes the du-form Hessian numerically singular for a
        # few steps (flag 4 = KMPC_STATUS_PIVOT: pivot floor applied, never fatal)
        ok = ("status: frozen [0], update [0]",) if name != "tank" else ("status: frozen [0], update [0]",
                                                                         "status: frozen [0], update [0 4]")
        assert any(t in r.stdout for t in ok), r.stdout
