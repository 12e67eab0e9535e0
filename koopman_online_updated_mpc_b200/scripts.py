"""The reference's driver scripts with the five call sites swapped for the B200 kernels.

Each function is one reference script from the line that loads the weights to the line that saves
the logs, batched over scenarios (scenario 0 with the defaults is the reference's own run):

    run_duffing        duffing.py:57 -> 1015      theta_E lift, C-output cost, bounds +-2
    run_vanderpol      vanderpol.py:57 -> 1112    theta_E lift, lifted-reference cost, bounds +-6
    run_rbf            duffing_RBF.py / vanderpol_RBF.py:44 -> 526   thin-plate RBF lift, "storage method" update
    run_tank           Tank_System.m:4 -> 291     Encoder_Tank lift (l.70-71), du form, N = 20
    run_koopman_update Revise_2/Koopman_update.m:10 -> 278            stacked lift, lambda, warm start, C frozen
    run_tracking_lift  VDP_Revise_2/Koopman_update_Tracking_Lift.m:9 -> 195   offset lift, C = I

Pipeline of every script: snapshots (data_generate.generate / kmpc_generate_snapshots) -> lift
(Encoder / rbf) -> EDMD (edmd.gram_* + edmd_solve) -> closed loop with the frozen model
(ClosedLoop(update=False)) -> closed loop with the online update (ClosedLoop(update=True)) ->
logs in the reference's array names, optionally written with io_mat in the reference's layouts.
Everything numerical runs in libkmpc.so; numpy is used for the random draws (the reference's own
streams) and to hand arrays back.
"""
import numpy as np
import torch

from dataclasses import replace as _replace

from . import closed_loop as _cl
from . import data_generate as _dg
from . import edmd as _edmd
from . import io_mat as _io
from . import lift as _lift
from . import plant as _plant
from .rls import RLSState


def _logs(loop, T):
    return loop.log_x[:T].permute(1, 2, 0).cpu().numpy(), loop.log_u[:T].t().cpu().numpy()


def identify(encoder, X, Y, U, n_step=None, mode=None, precision=_lift.PREC_FP64, c_variant=_edmd.C_PYTHON):
    """duffing.py:152-177: lift the snapshot set and regress A, B, C.  X, Y (n, M) / U (1, M) in the
    reference's layout (numpy) or snapshot-major CUDA tensors (M, n) / (M,).  n_step: the snapshots
    are consecutive n_step-step trajectories (data_generate.py:63-74) -> one encode per state."""
    if isinstance(X, np.ndarray):
        X, Y, U = X.T.copy(), Y.T.copy(), U.reshape(-1)
    if n_step:
        pack = _edmd.gram_from_trajectories(encoder, X, Y, U, n_step, mode=mode, precision=precision)
    else:
        pack = _edmd.gram_from_snapshots(encoder, X, Y, U, mode=mode, precision=precision)
    nz = encoder.out_dim(mode)
    A, B, C, status = _edmd.edmd_solve(pack, nz, 2, c_variant)
    if int(status.item()) != 0:
        raise RuntimeError("EDMD Gram matrix is not positive definite (status %d)" % int(status.item()))
    return A, B, C, pack


def _two_loops(spec, x0, A, B, C, r, max_step, **kw):
    """duffing.py:738-805 (frozen model) then 823-1012 (online update) from the same x0."""
    out = {}
    frozen = _cl.ClosedLoop(_replace(spec, update=False), x0, A, B, C, r, log_steps=max_step, **kw).run(max_step)
    out["logX"], out["logU"] = _logs(frozen, max_step)
    out["status_frozen"] = frozen.status.cpu().numpy()
    upd = _cl.ClosedLoop(_replace(spec, update=True), x0, A, B, C, r, log_steps=max_step, **kw).run(max_step)
    out["logXloc"], out["logUloc"] = _logs(upd, max_step)
    out["status_update"] = upd.status.cpu().numpy()
    out["Aloc"], out["Bloc"], out["Cloc"] = (t.cpu().numpy() for t in (upd.A, upd.B, upd.C))
    if upd.rls is not None:
        out["K_A"], out["inv_K_G"] = upd.rls.KA.cpu().numpy(), upd.rls.P.cpu().numpy()
        out["bar_X"], out["bar_Q"] = upd.rls.barX.cpu().numpy(), upd.rls.barQ.cpu().numpy()
    frozen.close()
    upd.close()
    return out


def _script(system, weights, max_step, x0, seed, save_dir, precision, model=None):
    enc = weights if isinstance(weights, _lift.Encoder) else _lift.Encoder.from_file(weights)
    if seed is not None:
        np.random.seed(seed)                               # duffing.py:47 (vanderpol.py:17 is commented out)
    gen = _dg.generate(100, 100)                           # duffing.py:72-76
    X, Y, U = gen.duffing_generate() if system == "duffing" else gen.vanderpol_generate()
    A, B, C, _ = identify(enc, X, Y, U, n_step=100, precision=precision)
    if model is not None:                                  # replay a run from logged matrices
        A, B, C = (torch.as_tensor(np.asarray(M, dtype=np.float64), device="cuda") for M in model)
    x0 = np.array([[-2.0, -2.0]]) if x0 is None else np.asarray(x0, dtype=np.float64).reshape(-1, 2)
    if system == "duffing":
        spec, r = _cl.duffing_spec(), np.array([1.0, 0.0])            # duffing.py:748-759
    else:
        spec = _cl.vanderpol_spec()
        r = enc(np.array([1.0, 0.0]))                                 # vanderpol.py:657-675: lifted set-point
    out = {"A": A.cpu().numpy(), "B": B.cpu().numpy(), "C": C.cpu().numpy(), "X": X, "Y": Y, "U": U}
    out.update(_two_loops(spec, x0, A, B, C, r, max_step, encoder=enc))
    if save_dir is not None:
        import os
        _io.save_model_weights(os.path.join(save_dir, "model_weights.mat"), enc.Ws, enc.bs)       # duffing.py:61-64
        _io.save_nn_encoder(os.path.join(save_dir, "NN_Encoder.mat"), out["logX"][0], out["logXloc"][0],
                            out["logU"][0])                                                       # duffing.py:1172
    return out


def run_duffing(weights, max_step=300, x0=None, seed=101, save_dir=None, precision=_lift.PREC_FP64, model=None):
    """duffing.py end to end.  Returns the reference's arrays: A, B, C, logX / logU (frozen model),
    logXloc / logUloc (online update) as (S, n, T) / (S, T), Aloc, Bloc, Cloc, K_A, inv_K_G, bar_X,
    bar_Q after the last step."""
    return _script("duffing", weights, max_step, x0, seed, save_dir, precision, model)


def run_vanderpol(weights, max_step=400, x0=None, seed=None, save_dir=None, precision=_lift.PREC_FP64, model=None):
    """vanderpol.py end to end.  The reference draws its EDMD snapshot set UNSEEDED (`np.random.seed(50)`
    at vanderpol.py:17 is commented out; the first live seed, l.263, is for the open-loop test set), so
    its A, B, C are not reproducible from a seed: `model=(A, B, C)` replays the closed loops from logged
    matrices (tests/golden/ref_vanderpol.npz holds the reference run's)."""
    return _script("vanderpol", weights, max_step, x0, seed, save_dir, precision, model)


def run_rbf(cx, system="duffing", max_step=120, x0=None, seed=101, N=10):
    """duffing_RBF.py / vanderpol_RBF.py with the centres `cx` (nz, 2) as an input (the reference
    draws them with an unseeded KMeans, duffing_RBF.py:44-46).  The online update is the "storage
    method" (l.434-438): the RLS warm-started from the offline Gram."""
    np.random.seed(seed)
    gen = _dg.generate(100, 100)
    X, Y, U = gen.duffing_generate() if system == "duffing" else gen.vanderpol_generate()
    cx_d = torch.as_tensor(np.asarray(cx, dtype=np.float64), device="cuda")
    PX = _lift.rbf(torch.from_numpy(X.T.copy()).cuda(), cx_d)
    PY = _lift.rbf(torch.from_numpy(Y.T.copy()).cuda(), cx_d)
    nz = cx_d.shape[0]
    pack = _edmd.gram_accumulate(PX, PY, U.reshape(-1), X.T.copy())
    A, B, C, status = _edmd.edmd_solve(pack, nz, 2)
    if int(status.item()) != 0:
        raise RuntimeError("EDMD Gram matrix is not positive definite")
    nv = nz + 1
    p = pack.cpu().numpy()
    G, Aq, XV = p[:nv * nv].reshape(nv, nv), p[nv * nv:nv * nv + nz * nv].reshape(nz, nv), \
        p[nv * nv + nz * nv:nv * nv + (nz + 2) * nv].reshape(2, nv)
    if system != "duffing":
        # vanderpol_RBF.py:127-128 (a line shared with duffing_RBF.py) re-seeds and calls
        # `duffing_generate()` again: from there on `X` is the DUFFING snapshot set (same inputs U: same
        # draws), and the storage-method read-out C = X pinv(X_EX) (l.438) regresses those states on the
        # VDP lifts.  Reproduced as is: the warm bar_X is X_duffing PHIX'.
        np.random.seed(seed)
        Xd, _, _ = _dg.generate(100, 100).duffing_generate()
        pd = _edmd.gram_accumulate(PX, PY, U.reshape(-1), Xd.T.copy()).cpu().numpy()
        XV = pd[nv * nv + nz * nv:nv * nv + (nz + 2) * nv].reshape(2, nv)
    x0 = np.array([[-2.0, -2.0]]) if x0 is None else np.asarray(x0, dtype=np.float64).reshape(-1, 2)
    S = x0.shape[0]
    pre, post = (_plant.DUFFING_PRE, _plant.DUFFING_POST) if system == "duffing" else (_plant.VDP_PRE, _plant.VDP_POST)
    spec = _cl.rbf_spec(nz=nz, N=N, params_pre=pre, params_post=post)
    r = np.array([1.0, 0.0])
    out = {"A": A.cpu().numpy(), "B": B.cpu().numpy(), "C": C.cpu().numpy(), "cx": np.asarray(cx)}
    # vanderpol_RBF.py:328 -- the FROZEN loop of the VDP script still switches to the duffing
    # post-plant (left over from duffing_RBF.py); its update loop (l.506) uses the VDP one
    frozen_post = _plant.DUFFING_POST
    frozen = _cl.ClosedLoop(_replace(spec, update=False, params_post=frozen_post), x0, A, B, C, r, cx=cx_d,
                            log_steps=max_step).run(max_step)
    out["logX"], out["logU"] = _logs(frozen, max_step)
    warm = RLSState.warm(S, G, Aq, XV[:, :nz], G[:nz, :nz])
    upd = _cl.ClosedLoop(spec, x0, A, B, C, r, cx=cx_d, rls_state=warm, log_steps=max_step).run(max_step)
    out["logXloc"], out["logUloc"] = _logs(upd, max_step)
    out["Aloc"], out["Bloc"], out["Cloc"] = (t.cpu().numpy() for t in (upd.A, upd.B, upd.C))
    out["status_update"] = upd.status.cpu().numpy()
    frozen.close()
    upd.close()
    return out


def tank_identify(encoder, n_traj=60, n_step=60, seed=55, x0=None, u0=None):
    """Tank_System.m:29-113 with the Encoder_Tank lift: random inputs u ~ U[-5, 5], x0 ~ U[-2, 2]^2
    clamped at 0 (l.36-45), joint Gram regression [A B; C 0] = W V' pinv(V V') (l.93-100).  The
    numpy RandomState(seed) stream stands in for MATLAB's rng(55) (unverified)."""
    if u0 is None or x0 is None:
        rs = np.random.RandomState(seed)
        u0 = (10 * rs.rand(n_step * n_traj) - 5).reshape((n_step, n_traj), order="F")
        x0 = (4 * rs.rand(2 * n_traj) - 2).reshape((2, n_traj), order="F").T
        x0 = np.maximum(x0, 0.0)
    X, Y, U = _dg.generate_snapshots(x0, u0, _plant.TANK_PRE, kind=_plant.PLANT_TANK)
    return identify(encoder, X, Y, U, c_variant=_edmd.C_JOINT)


def run_tank(weights, max_step=300, x0=None, seed=55, yr=1.0):
    """Tank_System.m end to end: identification, then the du-form closed loop with online update
    (l.170-291) from x0 (default [0, 0], l.163)."""
    enc = weights if isinstance(weights, _lift.Encoder) else _lift.Encoder.from_file(weights)
    A, B, C, _ = tank_identify(enc, seed=seed)
    x0 = np.zeros((1, 2)) if x0 is None else np.asarray(x0, dtype=np.float64).reshape(-1, 2)
    out = {"A": A.cpu().numpy(), "B": B.cpu().numpy(), "C": C.cpu().numpy()}
    out.update(_two_loops(_cl.tank_spec(), x0, A, B, C, np.array([yr]), max_step, encoder=enc))
    return out


def _offline_stats(encoder, mode, X, Y, U):
    """Gram statistics of the offline data in the lifted space: G = V V', Aq = PHIY V', XV = X V'."""
    pack = _edmd.gram_from_snapshots(encoder, X, Y, U, mode=mode)
    nz = encoder.out_dim(mode)
    nv = nz + 1
    p = pack.cpu().numpy()
    return (pack, p[:nv * nv].reshape(nv, nv), p[nv * nv:nv * nv + nz * nv].reshape(nz, nv),
            p[nv * nv + nz * nv:nv * nv + (nz + 2) * nv].reshape(2, nv))


def run_koopman_update(weights, max_step=100, x0=None, seed=2141444, lam=1.0, n_traj=100, n_step=100):
    """Revise_2/Koopman_update.m: lift [x; theta(x)] - [0; theta(0)] (nz = 10, l.67-70), Gram EDMD
    (l.94-101), N = 10, Q = 10 I, R = 0.01, bounds +-2 (l.130-142, 185-188), RLS with forgetting
    factor (`lambda = 1.0` in the file, l.257) warm-started from the offline Gram, C not updated
    (l.258-278), MATLAB RK4 (l.21-25), x0 = [-1; 1] (l.136).
    `Steps = 100` (l.154): the plant switch (l.230-239) never fires.  The SDP terminal weight
    (l.314-381) is out of scope."""
    enc = weights if isinstance(weights, _lift.Encoder) else _lift.Encoder.from_file(weights)
    rs = np.random.RandomState(seed)
    u0 = 4 * rs.rand(n_step, n_traj) - 2
    xi = 4 * rs.rand(n_traj, 2) - 2
    X, Y, U = _dg.generate_snapshots(xi, u0, _plant.DUFFING_PRE, rk4_variant=_plant.RK4_MATLAB)
    mode = _lift.LIFT_STACK
    pack, G, Aq, XV = _offline_stats(enc, mode, X, Y, U)
    nz = enc.out_dim(mode)
    A, B, C, status = _edmd.edmd_solve(pack, nz, 2, _edmd.C_JOINT)
    if int(status.item()) != 0:
        raise RuntimeError("EDMD Gram matrix is not positive definite")
    x0 = np.array([[-1.0, 1.0]]) if x0 is None else np.asarray(x0, dtype=np.float64).reshape(-1, 2)
    S = x0.shape[0]
    spec = _cl.LoopSpec(nz=nz, out_mode=_cl.OUT_C, lift_mode=mode, rk4_variant=_plant.RK4_MATLAB,
                        first_post_step=1 << 30, q=10.0, rw=0.01, lb=-2.0, ub=2.0, lam=lam, update_c=False)
    r = np.array([1.0, 0.0])
    out = {"A": A.cpu().numpy(), "B": B.cpu().numpy(), "C": C.cpu().numpy()}
    frozen = _cl.ClosedLoop(_replace(spec, update=False), x0, A, B, C, r, encoder=enc, log_steps=max_step).run(max_step)
    out["logX"], out["logU"] = _logs(frozen, max_step)
    warm = RLSState.warm(S, G, Aq, XV[:, :nz], G[:nz, :nz])
    upd = _cl.ClosedLoop(spec, x0, A, B, C, r, encoder=enc, rls_state=warm, log_steps=max_step).run(max_step)
    out["logXloc"], out["logUloc"] = _logs(upd, max_step)
    out["Aloc"], out["Bloc"], out["Cloc"] = (t.cpu().numpy() for t in (upd.A, upd.B, upd.C))
    out["status_update"] = upd.status.cpu().numpy()
    frozen.close()
    upd.close()
    return out


def run_tracking_lift(weights, max_step=300, x0=None, seed=2141444, n_traj=100, n_step=100, xref=(-1.0, 0.0)):
    """VDP_Revise_2/Koopman_update_Tracking_Lift.m: lift theta(x) - theta(0) (l.65), C = I (l.99: the
    cost tracks the lifted reference), Q = 100 I_8, R = 1e-4, N = 10, bounds +-6 (l.108-113, 151),
    RLS restart P0 = pinv(1e-5 I) (l.181-195), MATLAB RK4, plant switch tested before the plant
    call (l.157-171), x0 = [1; 1] (l.118), set-point liftFun([-1; 0]) (l.109)."""
    enc = weights if isinstance(weights, _lift.Encoder) else _lift.Encoder.from_file(weights)
    rs = np.random.RandomState(seed)
    u0 = 4 * rs.rand(n_step, n_traj) - 2
    xi = 4 * rs.rand(n_traj, 2) - 2
    X, Y, U = _dg.generate_snapshots(xi, u0, _plant.VDP_PRE, rk4_variant=_plant.RK4_MATLAB)
    mode = _lift.LIFT_OFFSET
    A, B, C, _ = identify(enc, X, Y, U, mode=mode, c_variant=_edmd.C_JOINT)
    x0 = np.array([[1.0, 1.0]]) if x0 is None else np.asarray(x0, dtype=np.float64).reshape(-1, 2)
    spec = _cl.vanderpol_spec(lift_mode=mode, rk4_variant=_plant.RK4_MATLAB, first_post_step=100, update_c=False)
    r = enc(np.asarray(xref, dtype=np.float64), mode=mode)
    out = {"A": A.cpu().numpy(), "B": B.cpu().numpy(), "C": C.cpu().numpy()}
    out.update(_two_loops(spec, x0, A, B, C, r, max_step, encoder=enc))
    return out
