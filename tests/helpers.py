"""Shared test plumbing: golden paths, the host lane-emulator, backend shims."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
WEIGHTS = os.path.join(GOLDEN, "weights")
REF = "/root/reference"


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def weights_path(system):
    return os.path.join(WEIGHTS, {"duffing": "duffing_model_weights.mat", "vdp": "vdp_model_weights.mat",
                                  "tank": "tank_model_weights.mat"}[system])


def oracle_weights(system):
    from oracle import weights
    return weights.load_mat_encoder(weights_path(system))


_EMU = None


def hostemu():
    """Build (g++) and load the TEST-ONLY host lane-emulator of the warp kernels."""
    global _EMU
    if _EMU is None:
        src = os.path.join(HERE, "hostemu", "hostemu.cpp")
        out_dir = os.path.join(HERE, "hostemu", "_build")
        os.makedirs(out_dir, exist_ok=True)
        lib = os.path.join(out_dir, "libkmpc_hostemu.so")
        deps = [src] + [os.path.join(ROOT, "koopman_online_updated_mpc_b200", "csrc", f)
                        for f in ("percase.cuh", "loopbody.cuh")] + [os.path.join(ROOT, "include", "kmpc.h")]
        if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
            subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", lib, src], check=True)
        _EMU = ctypes.CDLL(lib)
    return _EMU


def dp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class EmuBackend:
    """Stage entry points on the host emulator, numpy in / numpy out (mirrors the package API)."""
    name = "hostemu"

    def rls_update(self, st, z, u, y, xc, lam=1.0, update_c=True, skip_barx=False):
        S, nz = z.shape
        n = xc.shape[1]
        flags = (1 if update_c else 0) | (2 if skip_barx else 0)
        hostemu().emu_rls_update(dp(st["KA"]), dp(st["P"]), dp(st["barX"]), dp(st["barQ"]), dp(f64(z)),
                                 dp(f64(u)), dp(f64(y)), dp(f64(xc)), dp(st["A"]), dp(st["B"]), dp(st["C"]),
                                 ctypes.c_int64(S), nz, n, ctypes.c_double(lam), flags)
        return st["A"], st["B"], st["C"]

    def qp(self, A, B, C, z0, r, lb, ub, N, q, rw, PN=None):
        S, nz = z0.shape
        flags = 0
        if A.ndim == 2:
            flags |= 1
        if C is None:
            flags |= 4
            ny = nz
        else:
            ny = C.shape[-2]
        if r.ndim == 3:
            flags |= 2
        u0 = np.zeros(S)
        U = np.zeros((S, N))
        st = np.zeros(S, dtype=np.int32)
        hostemu().emu_qp_first_move(dp(f64(A)), dp(f64(B)), dp(None if C is None else f64(C)), dp(f64(z0)),
                                    dp(f64(r)), dp(f64(lb)), dp(f64(ub)), dp(None if PN is None else f64(PN)),
                                    ctypes.c_double(q), ctypes.c_double(rw), N, ny, nz, ctypes.c_int64(S),
                                    flags, dp(u0), dp(U), dp(st), 0, ctypes.c_double(0.0), None, None)
        return u0, U, st

    def plant(self, x, u, params, kind, variant, h=0.05):
        S = x.shape[0]
        out = np.zeros((S, 2))
        hostemu().emu_plant_step(dp(f64(x)), dp(f64(u)), dp(f64(params)), dp(out), ctypes.c_int64(S), kind,
                                 variant, ctypes.c_double(h))
        return out

    def rbf(self, x, cx, variant):
        S, n = x.shape
        nz = cx.shape[0]
        z = np.zeros((S, nz))
        hostemu().emu_rbf_lift(dp(f64(x)), dp(f64(cx)), dp(z), ctypes.c_int64(S), n, nz, variant)
        return z


class CudaBackend:
    """The product path: package functions -> libkmpc.so -> sm_100a kernels."""
    name = "cuda"

    def rls_update(self, st, z, u, y, xc, lam=1.0, update_c=True, skip_barx=False):
        import torch
        from koopman_online_updated_mpc_b200 import rls
        S, nz = z.shape
        n = xc.shape[1]
        dev = rls.RLSState(S, nz, n)
        for k in ("KA", "P", "barX", "barQ"):
            getattr(dev, k).copy_(torch.from_numpy(st[k]))
        A, B, C = rls.rls_update(dev, z, u, y, xc, lam, update_c, skip_barx)
        for k in ("KA", "P", "barX", "barQ", "A", "C"):
            st[k][...] = getattr(dev, k).cpu().numpy()
        st["B"][...] = dev.B.cpu().numpy().reshape(st["B"].shape)
        return st["A"], st["B"], st["C"]

    def qp(self, A, B, C, z0, r, lb, ub, N, q, rw, PN=None):
        from koopman_online_updated_mpc_b200 import mpc
        u0, U, st = mpc.mpc_first_move(A, B, C, z0, r, lb, ub, N=N, q=q, rw=rw, PN=PN, return_sequence=True)
        return u0, U, st

    def plant(self, x, u, params, kind, variant, h=0.05):
        from koopman_online_updated_mpc_b200 import plant
        return plant.f_update(x, u, params, kind, variant, h)

    def rbf(self, x, cx, variant):
        from koopman_online_updated_mpc_b200 import lift
        return lift.rbf(x, cx, variant)


def new_rls_state(S, nz, n, p0, q0):
    nv = nz + 1
    return dict(KA=np.zeros((S, nz, nv)), P=np.tile(p0 * np.eye(nv), (S, 1, 1)),
                barX=np.zeros((S, n, nz)), barQ=np.tile(q0 * np.eye(nz), (S, 1, 1)),
                A=np.zeros((S, nz, nz)), B=np.zeros((S, nz)), C=np.zeros((S, n, nz)))


class EmuClosedLoop:
    """Host-emulated twin of koopman_online_updated_mpc_b200.ClosedLoop (numpy buffers)."""

    def __init__(self, spec, x0, A, B, C, r, Ws=None, bs=None, cx=None, warm=None, log_steps=0,
                 params_pre=None, params_post=None):
        from koopman_online_updated_mpc_b200 import _lib
        from koopman_online_updated_mpc_b200.closed_loop import make_config
        from oracle import lift as olift
        self.spec = spec
        self.x = f64(x0).reshape(-1, 2).copy()
        S = self.S = self.x.shape[0]
        nz, n = spec.nz, spec.n
        A, B, C = f64(A), f64(B), f64(C)
        shared = A.ndim == 2
        if spec.update and shared:
            A, B, C = (np.tile(M.reshape((1,) + s), (S, 1, 1)) for M, s in
                       ((A, (nz, nz)), (B, (nz, 1)), (C, (n, nz))))
            shared = False
        self.A, self.B, self.C = f64(A).copy(), f64(B).reshape(-1, nz, 1).copy(), f64(C).copy()
        r = f64(r)
        self.r = np.tile(r.reshape(1, -1), (S, 1)) if r.ndim == 1 else r.copy()
        pp = lambda p, d: np.tile(f64(d if p is None else p).reshape(1, 5), (S, 1)) if f64(d if p is None else p).ndim == 1 else f64(p).copy()
        self.params_pre, self.params_post = pp(params_pre, spec.params_pre), pp(params_post, spec.params_post)
        self.u_prev = np.zeros(S)
        self.Ws, self.bs, self.cx = Ws, bs, (None if cx is None else f64(cx))
        if spec.lift_kind == 0:
            self.z = f64(olift.lift_mlp(Ws, bs, self.x, spec.lift_mode))
        else:
            self.z = f64(olift.rbf_lift(self.x, self.cx, spec.lift_mode))
        self.rls_started = 0
        self.st = None
        if spec.update:
            if warm is not None:
                self.st, self.rls_started = warm, 1
            else:
                self.st = new_rls_state(S, nz, n, spec.p0, spec.q0)
        self.log_x = np.zeros((log_steps, S, n)) if log_steps else None
        self.log_u = np.zeros((log_steps, S)) if log_steps else None
        self.status = np.zeros(S, dtype=np.int32)
        self.cfg = make_config(spec, S, shared, log_steps)
        st = self.st
        self.buf = _lib.LoopBuffersC(
            x=dp(self.x), z=dp(self.z), u_prev=dp(self.u_prev), A=dp(self.A), B=dp(self.B), C=dp(self.C),
            KA=dp(st["KA"]) if st else None, P=dp(st["P"]) if st else None,
            barX=dp(st["barX"]) if st else None, barQ=dp(st["barQ"]) if st else None, r=dp(self.r),
            params_pre=dp(self.params_pre), params_post=dp(self.params_post), cx=dp(self.cx),
            log_x=dp(self.log_x), log_u=dp(self.log_u), status=dp(self.status), log_capacity=log_steps)
        self.step = 0
        self.qp_x = np.full((S, spec.N), np.nan)   # QP warm start carried from step to step (kmpc_ctx does the same)

    def run(self, T):
        nl = len(self.Ws) if self.Ws else 0
        if nl:
            dims = [self.Ws[0].shape[1]] + [W.shape[0] for W in self.Ws]
            Wc = [f64(W) for W in self.Ws]
            bc = [f64(b).reshape(-1) for b in self.bs]
            Wp = (ctypes.c_void_p * nl)(*[W.ctypes.data for W in Wc])
            bp = (ctypes.c_void_p * nl)(*[b.ctypes.data for b in bc])
            dm = (ctypes.c_int * (nl + 1))(*dims)
        else:
            Wp = bp = dm = None
        hostemu().emu_closed_loop(ctypes.byref(self.cfg), ctypes.byref(self.buf), int(T), self.rls_started,
                                  ctypes.c_int64(self.step), nl, dm, Wp, bp, dp(self.qp_x))
        self.step += T
        if self.spec.update:
            self.rls_started = 1
        return self
