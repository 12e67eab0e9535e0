"""The fused closed loop: one object per scenario batch, `run(T)` advances every scenario T steps.

A scenario-step is the body of duffing.py:823-992 (lift -> QP -> plant -> lift -> RLS), of
duffing.py:738-805 when `update=False`, or of Tank_System.m:170-291 with `du_aug=True`.  The
presets below carry the constants the reference scripts hard-code."""
import ctypes
from dataclasses import dataclass, replace

import torch

from . import _lib, lift as _lift, plant as _plant
from ._tensors import ptr, stream_ptr, to_dev

OUT_C, OUT_IDENTITY, OUT_C_ROW = 0, 1, 2
LIFTKIND_MLP, LIFTKIND_RBF = 0, 1
PATH_AUTO, PATH_GENERIC = 0, 1


@dataclass
class LoopSpec:
    nz: int
    n: int = 2
    N: int = 10                    # MPCHorizon = ControlHorizon = 10 (duffing.py:632-633)
    out_mode: int = OUT_C
    out_row: int = 1
    du_aug: bool = False
    update: bool = True
    update_c: bool = True
    c_pairs_next: bool = True
    skip_first_barx: bool = False
    lift_kind: int = LIFTKIND_MLP
    lift_mode: int = _lift.LIFT_RAW
    plant_kind: int = _plant.PLANT_POLY2
    rk4_variant: int = _plant.RK4_PYTHON
    first_post_step: int = 102     # python: `if i > 100` at the END of the body (duffing.py:991-992)
    max_iter: int = 0
    h: float = 0.05                # duffing.py:250
    q: float = 100.0               # cost weights inside costFunction (duffing.py:580)
    rw: float = 1e-4
    lb: float = -2.0               # duffing.py:636
    ub: float = 2.0
    u_lb: float = -8.0             # Tank_System.m:144-145
    u_ub: float = 8.0
    lam: float = 1.0
    p0: float = 1e4                # duffing.py:929-930 (pinv(1e-4 I))
    q0: float = 100.0              # duffing.py:946
    tol: float = 0.0
    path: int = PATH_AUTO          # PATH_GENERIC: force the per-step generic kernels (cross-check of the fused one)
    qp_cold: int = 0               # generic kernels: 0 warm start + sweeps, 1 cold start (duffing.py:634), 2 warm primal only, 3 warm + damped sweeps (Tank)
    params_pre: tuple = _plant.DUFFING_PRE
    params_post: tuple = _plant.DUFFING_POST

    @property
    def nzq(self):
        return self.nz + (1 if self.du_aug else 0)

    @property
    def ny(self):
        if self.out_mode == OUT_IDENTITY:
            return self.nzq
        return self.n if self.out_mode == OUT_C else 1


def duffing_spec(**kw):
    """duffing.py: r = (1, 0), bounds +-2, P0 = 1e4 I, bar_Q0 = 100 I."""
    return replace(LoopSpec(nz=8), **kw)


def vanderpol_spec(**kw):
    """vanderpol.py: tracks the LIFTED reference (y = z, l.456-459), bounds +-6 (l.542-544),
    P0 = bar_Q0 = 1e5 I (l.874, 888)."""
    return replace(LoopSpec(nz=8, out_mode=OUT_IDENTITY, lb=-6.0, ub=6.0, p0=1e5, q0=1e5,
                            params_pre=_plant.VDP_PRE, params_post=_plant.VDP_POST), **kw)


def rbf_spec(nz=8, **kw):
    """duffing_RBF.py: thin-plate RBF lift, same cost; its online update ('storage method',
    l.434-438) is the RLS warm-started from the offline Gram: pass rls_state=RLSState.warm(...)."""
    return replace(LoopSpec(nz=nz, lift_kind=LIFTKIND_RBF, lift_mode=_lift.RBF_PYTHON), **kw)


def tank_spec(nz=10, **kw):
    """Tank_System.m: velocity form (l.110-113), N = 20, Q = 10, R = 1e-3 (l.116-118),
    dU in +-0.5, u in +-8 (l.144-159), P0 = bar_Q0 = 1e4 I, switch tested before the plant call.
    qp_cold = 3: damped primal-dual sweeps (the plain sweeps cycle in more than half of this loop's heavy steps)."""
    return replace(LoopSpec(nz=nz, N=20, out_mode=OUT_C_ROW, out_row=1, du_aug=True, c_pairs_next=False,
                            skip_first_barx=True, plant_kind=_plant.PLANT_TANK, first_post_step=100,
                            q=10.0, rw=1e-3, lb=-0.5, ub=0.5, p0=1e4, q0=1e4, qp_cold=3,
                            params_pre=_plant.TANK_PRE, params_post=_plant.TANK_POST), **kw)


def make_config(spec, S, shared_model, log_steps=0):
    """LoopSpec -> the C struct kmpc_loop_config (include/kmpc.h)."""
    return _lib.LoopConfigC(
        S=S, nz=spec.nz, n=spec.n, N=spec.N, out_mode=spec.out_mode, out_row=spec.out_row,
        du_aug=int(spec.du_aug), update=int(spec.update), rls_flags=(1 if spec.update_c else 0),
        c_pairs_next=int(spec.c_pairs_next), skip_first_barx=int(spec.skip_first_barx),
        shared_model=int(shared_model), lift_kind=spec.lift_kind, lift_mode=spec.lift_mode,
        plant_kind=spec.plant_kind, rk4_variant=spec.rk4_variant,
        first_post_step=spec.first_post_step, max_iter=spec.max_iter, h=spec.h, q=spec.q,
        rw=spec.rw, lb=spec.lb, ub=spec.ub, u_lb=spec.u_lb, u_ub=spec.u_ub, lam=spec.lam,
        p0=spec.p0, q0=spec.q0, tol=spec.tol, path=int(spec.path), qp_cold=int(spec.qp_cold))


def _as_model(M, tail, S, name):
    """(tail) or (1|S, tail) -> CUDA (1|S, *tail); B may come without its trailing 1."""
    t = to_dev(M)
    numel = tail[0] * tail[1]
    if t.numel() == numel:
        return t.reshape(1, *tail)
    if t.numel() == S * numel and t.shape[0] == S:
        return t.reshape(S, *tail)
    raise ValueError("%s has shape %s: expected %s or (%d, %d, %d)" % (name, tuple(t.shape), tail, S, *tail))


def _check_rls_state(st, S, nz, n):
    want = {"KA": (S, nz, nz + 1), "P": (S, nz + 1, nz + 1), "barX": (S, n, nz), "barQ": (S, nz, nz)}
    for k, shp in want.items():
        t = getattr(st, k)
        if tuple(t.shape) != shp or t.dtype != torch.float64 or not t.is_cuda or not t.is_contiguous():
            raise ValueError("RLSState.%s must be a contiguous CUDA float64 tensor of shape %s, got %s"
                             % (k, shp, tuple(t.shape)))


class ClosedLoop:
    """Device-resident batch of S closed-loop scenarios.

    x0 (S,2); A (nz,nz) / B (nz,1) / C (n,nz) shared or per-scenario (S,...); r (ny,) or (S,ny);
    encoder: lift.Encoder (MLP) or cx: (nz,n) centres (RBF); rls_state: optional warm RLSState;
    log_steps: capacity of the (T,S,.) logs."""

    def __init__(self, spec, x0, A, B, C, r, encoder=None, cx=None, rls_state=None, log_steps=0,
                 params_pre=None, params_post=None, u_prev=None):
        from .rls import RLSState
        L = _lib.lib()
        self.spec = spec
        f64 = dict(dtype=torch.float64, device="cuda")
        self.x = to_dev(x0).reshape(-1, spec.n).clone()
        S = self.S = self.x.shape[0]
        nz, n = spec.nz, spec.n
        # A, B, C: one sharing mode for all three.  The kernels index every matrix with the same
        # scenario offset, so a shared tensor is expanded as soon as any of them is per-scenario
        # (or the loop updates the model online).
        A_d = _as_model(A, (nz, nz), S, "A")
        B_d = _as_model(B, (nz, 1), S, "B")
        C_d = _as_model(C, (n, nz), S, "C")
        shared = (not spec.update) and all(t.shape[0] == 1 for t in (A_d, B_d, C_d))
        self.shared_model = shared
        if not shared:
            A_d, B_d, C_d = (t.expand(S, *t.shape[1:]) for t in (A_d, B_d, C_d))
        self.A = A_d.contiguous().clone()
        self.B = B_d.contiguous().clone()
        self.C = C_d.contiguous().clone()
        r_d = to_dev(r)
        if r_d.ndim == 1:
            r_d = r_d.reshape(1, -1).expand(S, -1)
        self.r = r_d.contiguous().clone()
        assert self.r.shape == (S, spec.ny), (self.r.shape, spec.ny)

        def params(p, default):
            p_d = to_dev(default if p is None else p)
            if p_d.ndim == 1:
                p_d = p_d.reshape(1, 5).expand(S, 5)
            return p_d.contiguous().clone()
        self.params_pre = params(params_pre, spec.params_pre)
        self.params_post = params(params_post, spec.params_post)
        self.u_prev = torch.zeros(S, **f64) if u_prev is None else to_dev(u_prev).reshape(S).clone()
        self.encoder, self.cx = encoder, (None if cx is None else to_dev(cx).clone())
        self.z = torch.empty((S, nz), **f64)
        if spec.lift_kind == LIFTKIND_MLP:
            encoder.encode_into(self.x, self.z, spec.lift_mode)
        else:
            _lib.check(L.kmpc_rbf_lift(ptr(self.x), ptr(self.cx), ptr(self.z), S, n, nz, spec.lift_mode,
                                       stream_ptr()))
        self.rls = None
        self._rls0 = None            # warm start of construction, restored by reset()
        rls_started = 0
        if spec.update:
            if rls_state is not None:
                _check_rls_state(rls_state, S, nz, n)
                # the loop updates its OWN copy in place: the caller's warm state stays intact and
                # reset() can start every episode from it
                self.rls, rls_started = rls_state.clone(), 1
                self._rls0 = rls_state.clone()
            else:
                self.rls = RLSState(S, nz, n, spec.p0, spec.q0)
        # initial values kept for reset()
        self._x0, self._A0, self._B0, self._C0 = self.x.clone(), self.A.clone(), self.B.clone(), self.C.clone()
        self._u0 = self.u_prev.clone()
        self.log_x = torch.zeros((log_steps, S, n), **f64) if log_steps else None
        self.log_u = torch.zeros((log_steps, S), **f64) if log_steps else None
        self.status = torch.zeros(S, dtype=torch.int32, device="cuda")
        cfg = make_config(spec, S, shared, log_steps)
        rs = self.rls
        buf = _lib.LoopBuffersC(
            x=ptr(self.x), z=ptr(self.z), u_prev=ptr(self.u_prev), A=ptr(self.A), B=ptr(self.B),
            C=ptr(self.C), KA=ptr(rs.KA) if rs else None, P=ptr(rs.P) if rs else None,
            barX=ptr(rs.barX) if rs else None, barQ=ptr(rs.barQ) if rs else None, r=ptr(self.r),
            params_pre=ptr(self.params_pre), params_post=ptr(self.params_post), cx=ptr(self.cx),
            log_x=ptr(self.log_x), log_u=ptr(self.log_u), status=ptr(self.status),
            log_capacity=log_steps)
        h = ctypes.c_void_p()
        _lib.check(L.kmpc_ctx_create(ctypes.byref(h), ctypes.byref(cfg), ctypes.byref(buf),
                                     encoder.handle if spec.lift_kind == LIFTKIND_MLP else None,
                                     rls_started, stream_ptr()))
        self._h = h

    @property
    def fused(self):
        """True when run(T) is ONE persistent fused kernel launch (nz = 8, N = 10 loops)."""
        return bool(_lib.lib().kmpc_ctx_is_fused(self._h))

    def reset(self, x0=None, rls_state=None, cold=False):
        """Start a new episode on the same device buffers (the reference's `for i in range(maxStep)`
        begins again): x <- x0 (default: the x0 of construction), z <- lift(x), u_prev <- 0,
        A, B, C <- the initial model, step index <- 0.  RLS: a loop built with a warm `rls_state`
        restarts from that same warm state (or from the `rls_state` given here); a loop built cold,
        or `cold=True`, restarts from P = p0 I, bar_Q = q0 I (duffing.py:927-930)."""
        L = _lib.lib()
        spec = self.spec
        if x0 is not None:
            self._x0.copy_(to_dev(x0).reshape(self.S, spec.n), non_blocking=True)
        self.x.copy_(self._x0)
        self.A.copy_(self._A0)
        self.B.copy_(self._B0)
        self.C.copy_(self._C0)
        self.u_prev.copy_(self._u0)
        self.status.zero_()
        if spec.lift_kind == LIFTKIND_MLP:
            self.encoder.encode_into(self.x, self.z, spec.lift_mode)
        else:
            _lib.check(L.kmpc_rbf_lift(ptr(self.x), ptr(self.cx), ptr(self.z), self.S, spec.n, spec.nz,
                                       spec.lift_mode, stream_ptr()))
        started = 0
        if spec.update:
            src = rls_state if rls_state is not None else (None if cold else self._rls0)
            if src is not None:
                _check_rls_state(src, self.S, spec.nz, spec.n)
                for k in ("KA", "P", "barX", "barQ"):
                    getattr(self.rls, k).copy_(getattr(src, k))
                started = 1
        _lib.check(L.kmpc_ctx_reset(self._h, started, stream_ptr()))
        return self

    @property
    def step_index(self):
        return int(_lib.lib().kmpc_ctx_step_index(self._h))

    def run(self, T):
        """Advance all scenarios by T closed-loop steps (asynchronous on the current stream)."""
        _lib.check(_lib.lib().kmpc_closed_loop_steps(self._h, int(T), stream_ptr()))
        return self

    def run_timed(self, T):
        """run(T) with per-kernel CUDA-event timing; returns summed device milliseconds
        {"qp_plant": .., "lift": .., "rls": ..} (synchronises)."""
        ms = (ctypes.c_float * 3)()
        _lib.check(_lib.lib().kmpc_closed_loop_steps_timed(self._h, int(T), stream_ptr(), ms))
        return {"qp_plant": ms[0], "lift": ms[1], "rls": ms[2]}

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().kmpc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
