"""Oracle: closed-loop Koopman MPC, one scenario at a time (test infrastructure).

One unified step covers every reference loop (the CUDA context implements the same semantics):

    zl = lift(x)                                            duffing.py:847  Tank_System.m:290
    zq = [zl; u_prev] and (A,B,C) -> ([A B;0 1],[B;1],C[I 0])   if du_aug   Tank_System.m:110-113,265-268
    move = argmin box-QP ; u = move (or u_prev + move)      duffing.py:856-861  Tank_System.m:188,192
    p = pre-switch params if k < first_post_step else post  duffing.py:991-992  Tank_System.m:193-195
    x+ = plant(x, u, p)                                     duffing.py:871      Tank_System.m:210-211
    if update: RLS with (zl, u) -> lift(x+), C paired with x+ (python) or x (tank)
                                                            duffing.py:927-984  Tank_System.m:234-263

Switch timing: the python scripts test `i > 100` at the END of the loop body, so the first step
integrated with the new plant is 0-based k = 102 (first_post_step = 102); the MATLAB scripts test
it BEFORE the plant call with 1-based i, so first_post_step = 100.
"""
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np

from . import edmd, lift, mpc, plant, rls

OUT_C, OUT_IDENTITY, OUT_C_ROW = 0, 1, 2
UPDATE_NONE, UPDATE_RLS, UPDATE_STORAGE = 0, 1, 2


@dataclass
class LoopConfig:
    name: str
    lift_fn: Callable            # (n,) -> (nz,)
    nz: int
    plant_kind: int = plant.PLANT_POLY2
    p_pre: tuple = plant.DUFFING_PRE
    p_post: tuple = plant.DUFFING_POST
    first_post_step: int = 102
    rk4_variant: int = plant.RK4_PYTHON
    h: float = 0.05
    N: int = 10
    q: float = 100.0
    rw: float = 1e-4
    lb: float = -2.0
    ub: float = 2.0
    out_mode: int = OUT_C        # y = C z (ny=2) | y = z (ny=nz) | y = (C z)[out_row] (ny=1)
    out_row: int = 1
    r: Optional[np.ndarray] = None   # (ny,) constant reference over the horizon
    du_aug: bool = False         # velocity form (Tank_System.m:110-113)
    u_lb: float = -8.0           # absolute input bounds on the first move when du_aug
    u_ub: float = 8.0
    p0: float = 1e4              # RLS P0 = p0*I   (duffing.py:929-930 -> 1e4 ; vanderpol.py:874 -> 1e5)
    q0: float = 100.0            # RLS bar_Q0      (duffing.py:946 -> 100 ; vanderpol.py:888 -> 1e5)
    lam: float = 1.0
    update_c: bool = True
    c_pairs_next: bool = True    # python: bar_X += x+ z'  ; tank: bar_X += x z'
    skip_first_barx: bool = False  # Tank_System.m:252-254


def duffing_config(Ws, bs):
    return LoopConfig("duffing", lambda x: lift.encoder_forward(Ws, bs, x), 8,
                      p_pre=plant.DUFFING_PRE, p_post=plant.DUFFING_POST,
                      r=np.array([1.0, 0.0]), lb=-2.0, ub=2.0, p0=1e4, q0=100.0)


def vanderpol_config(Ws, bs, xref=(1.0, 0.0)):
    enc = lambda x: lift.encoder_forward(Ws, bs, x)
    return LoopConfig("vanderpol", enc, 8, p_pre=plant.VDP_PRE, p_post=plant.VDP_POST,
                      out_mode=OUT_IDENTITY, r=enc(np.asarray(xref, dtype=np.float64)),
                      lb=-6.0, ub=6.0, p0=1e5, q0=1e5)


def rbf_config(cx, system="duffing"):
    # vanderpol_RBF.py:118,342 run the VDP plant; its update loop switches to the VDP post-plant
    # (l.506) but its frozen loop still switches to the *duffing* post-plant (l.328, left over
    # from duffing_RBF.py) -- callers reproduce that by overriding p_post.
    pre, post = ((plant.DUFFING_PRE, plant.DUFFING_POST) if system == "duffing"
                 else (plant.VDP_PRE, plant.VDP_POST))
    return LoopConfig(system + "_rbf", lambda x: lift.rbf_lift(x, cx, lift.RBF_PYTHON), cx.shape[0],
                      p_pre=pre, p_post=post, r=np.array([1.0, 0.0]), lb=-2.0, ub=2.0)


def tank_config(lift_fn, nz):
    return LoopConfig("tank", lift_fn, nz, plant_kind=plant.PLANT_TANK, p_pre=plant.TANK_PRE,
                      p_post=plant.TANK_POST, first_post_step=100, N=20, q=10.0, rw=1e-3,
                      lb=-0.5, ub=0.5, out_mode=OUT_C_ROW, out_row=1, r=np.array([1.0]),
                      du_aug=True, u_lb=-8.0, u_ub=8.0, p0=1e4, q0=1e4, c_pairs_next=False,
                      skip_first_barx=True)


def koopman_update_config(Ws, bs, lam=1.0):
    """Revise_2/Koopman_update.m: lift [x; theta(x)] - [0; theta(0)] (nz = 10, l.67-70), Cy = I_2 on the
    jointly regressed C (l.94-101, 108), N = 10, Q = 10 I_2, R = 0.01 (l.110-113), bounds +-2 (l.213),
    MATLAB RK4 with k4 = f(x + k1 dt) (l.21-25), RLS with forgetting factor `lambda` (= 1.0 in the
    file, l.257) warm-started from the offline Gram (l.264-265), C never updated.  Steps = 100
    (l.154) so the `i > 100` plant switch (l.230-239) never fires: first_post_step is out of reach.
    The SDP terminal weight (l.314-381) is out of scope."""
    return LoopConfig("koopman_update", lambda x: lift.lift_mlp(Ws, bs, x, lift.LIFT_STACK), 10,
                      p_pre=plant.DUFFING_PRE, p_post=plant.DUFFING_POST, first_post_step=1 << 30,
                      rk4_variant=plant.RK4_MATLAB, N=10, q=10.0, rw=0.01, lb=-2.0, ub=2.0,
                      r=np.array([1.0, 0.0]), lam=lam, update_c=False)


def tracking_lift_config(Ws, bs, xref=(-1.0, 0.0)):
    """VDP_Revise_2/Koopman_update_Tracking_Lift.m: lift theta(x) - theta(0) (l.65), C = I (l.99: the
    cost tracks the lifted reference Yr = liftFun([-1; 0]), l.109), N = 10, Q = 100 I_8, R = 1e-4
    (l.106-108), bounds +-6 (l.151), RLS restart P0 = pinv(1e-5 I) (l.183-185), C never updated,
    MATLAB RK4 (l.21-25), plant switch tested BEFORE the plant call with 1-based i (l.157-165):
    first new-plant step is 0-based k = 100."""
    f = lambda x: lift.lift_mlp(Ws, bs, x, lift.LIFT_OFFSET)
    return LoopConfig("tracking_lift", f, 8, p_pre=plant.VDP_PRE, p_post=plant.VDP_POST, first_post_step=100,
                      rk4_variant=plant.RK4_MATLAB, N=10, q=100.0, rw=1e-4, lb=-6.0, ub=6.0,
                      out_mode=OUT_IDENTITY, r=f(np.asarray(xref, dtype=np.float64)), p0=1e5, q0=1e5,
                      update_c=False)


def qp_model(cfg, A, B, C):
    """Matrices the QP sees: optional du-augmentation and the output selection."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64).reshape(-1, 1)
    nz = A.shape[0]
    if cfg.out_mode == OUT_IDENTITY:
        Cy = np.eye(nz)
    elif cfg.out_mode == OUT_C:
        Cy = np.asarray(C, dtype=np.float64)
    else:
        Cy = np.asarray(C, dtype=np.float64)[cfg.out_row:cfg.out_row + 1]
    if cfg.du_aug:
        Aa = np.block([[A, B], [np.zeros((1, nz)), np.eye(1)]])
        Ba = np.concatenate([B, np.eye(1)], axis=0)
        Cya = np.concatenate([Cy, np.zeros((Cy.shape[0], 1))], axis=1)
        return Aa, Ba, Cya
    return A, B, Cy


def mpc_move(cfg, A, B, C, zl, u_prev, qp="exact"):
    """Returns (u applied, full move sequence, status)."""
    Aq, Bq, Cy = qp_model(cfg, A, B, C)
    zq = np.concatenate([zl, [u_prev]]) if cfg.du_aug else zl
    lb = np.full(cfg.N, cfg.lb)
    ub = np.full(cfg.N, cfg.ub)
    if cfg.du_aug:  # Tank_System.m:182-188: umin <= U0 + dU_1 <= umax folded into the box
        lb[0] = max(lb[0], cfg.u_lb - u_prev)
        ub[0] = min(ub[0], cfg.u_ub - u_prev)
    if qp == "literal":
        assert not cfg.du_aug
        r = np.repeat(np.asarray(cfg.r).reshape(-1, 1), cfg.N, axis=1)
        U = mpc.solve_literal(r, np.concatenate([Aq, Bq], axis=1),
                              None if cfg.out_mode == OUT_IDENTITY else Cy, zq, cfg.lb, cfg.ub, cfg.N)
        status = 0
    else:
        H, f = mpc.condense(Aq, Bq, Cy, zq, cfg.r, cfg.q, cfg.rw, cfg.N)
        U, status, _ = mpc.solve_box_qp_exact(H, f, lb, ub)
    u = u_prev + U[0] if cfg.du_aug else U[0]
    return float(u), U, status


def run_loop(cfg, A, B, C, x0, T, update=UPDATE_RLS, qp="exact", warm=None, storage=None,
             u_prev=0.0, record_models=False, start_step=0, record_states=False):
    """Run T closed-loop steps for one scenario.  `warm` = rls.RLSState to continue from
    (Koopman_update.m:264-265); `storage` = dict(PHIX, PHIY, U, X) for the literal storage-method
    update of duffing_RBF.py:406-438.  Returns a dict of logs and the final model/state."""
    x = np.asarray(x0, dtype=np.float64).copy()
    A, B, C = (np.array(M, dtype=np.float64) for M in (A, B, C))
    B = B.reshape(-1, 1)
    st = warm
    logX, logU, logZ, logStatus, models, pre_states = [], [], [], [], [], []
    if update == UPDATE_STORAGE:
        X_EX, Y_EX = storage["PHIX"].copy(), storage["PHIY"].copy()
        U_EX, Xs = storage["U"].copy(), storage["X"].copy()
    zl = cfg.lift_fn(x)
    for k in range(start_step, start_step + T):
        if record_states:  # everything step k starts from (teacher-forcing fixtures)
            pre_states.append(dict(k=k, x=x.copy(), z=zl.copy(), u_prev=u_prev, A=A.copy(), B=B.copy(),
                                   C=C.copy(), rls=None if st is None else st.copy()))
        u, _, status = mpc_move(cfg, A, B, C, zl, u_prev, qp)
        p = cfg.p_pre if k < cfg.first_post_step else cfg.p_post
        xn = plant.plant_step(cfg.plant_kind, x, u, np.asarray(p), cfg.h, cfg.rk4_variant)
        yl = cfg.lift_fn(xn)
        logZ.append(zl)
        logX.append(xn)
        logU.append(u)
        logStatus.append(status)
        if update == UPDATE_RLS:
            if st is None:
                st = rls.RLSState(cfg.nz, 1, x.shape[0], cfg.p0, cfg.q0)
                first = True
            else:
                first = False
            xc = xn if cfg.c_pairs_next else x
            An, Bn, Cn = rls.rls_update(st, zl, u, yl, xc, cfg.lam, cfg.update_c,
                                        accumulate_barx=not (cfg.skip_first_barx and first and warm is None))
            A, B = An, Bn
            if Cn is not None:
                C = Cn
        elif update == UPDATE_STORAGE:
            X_EX = np.concatenate([X_EX, zl.reshape(-1, 1)], axis=1)
            Y_EX = np.concatenate([Y_EX, yl.reshape(-1, 1)], axis=1)
            U_EX = np.concatenate([U_EX, np.array([[u]])], axis=1)
            Xs = np.concatenate([Xs, xn.reshape(-1, 1)], axis=1)
            XU = np.concatenate([X_EX, U_EX], axis=0)
            K = (Y_EX @ XU.T) @ np.linalg.pinv(XU @ XU.T)
            A, B = K[:, :cfg.nz], K[:, cfg.nz:]
            C = Xs @ np.linalg.pinv(X_EX)
        if record_models:
            models.append((A.copy(), B.copy(), C.copy()))
        x, zl, u_prev = xn, yl, u
    return dict(X=np.array(logX), U=np.array(logU), Z=np.array(logZ), status=np.array(logStatus),
                A=A, B=B, C=C, rls=st, x=x, u_prev=u_prev, models=models, pre_states=pre_states)
