"""CPU study behind DESIGN.md 3.2 (damped primal-dual sweeps, qp_cold = 3): factorisations per step of the Tank
loop (Tank_System.m shape, the bench's draw) under the generic kernel's solver semantics.

    POLICY=cold|shift [PI=8] [REL1=99] python tests/studies/qp_sweeps_tank.py [n_scenarios]

PI = sweeps before the monotone primal fallback, REL1 = from that sweep on only the most negative multiplier is
released (PI=40 REL1=1 is what qp_cold = 3 does).  Test infrastructure: lives under tests/ because it imports the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ[k] = "1"
from oracle import closed_loop as ocl, mpc, plant, rls
POLICY = os.environ.get("POLICY", "shift")

PI = int(os.environ.get("PI", "8")); REL1 = int(os.environ.get("REL1", "99"))
def pdas(H, f, lb, ub, x, W, tol=1e-10, pdas_iters=None, max_iter=220):
    pdas_iters = PI
    """kernel qp_solve_warp from a given (x, W); returns x, W, iterations (= factorisations)."""
    N = f.size; H2 = 2 * H
    mtol = tol * max(1.0, np.abs(f).max())
    g = H2 @ x + f
    it = 0; done = False
    while not done and it < max_iter:
        pd = it < pdas_iters
        free = W == 0
        p, ok = mpc._masked_chol_solve(H2, -g, free)
        it += 1
        alpha, block, side = 1.0, -1, 0
        if not pd:
            for i in range(N):
                if not free[i]: continue
                if p[i] > 0 and x[i] + p[i] > ub[i]:
                    a = (ub[i] - x[i]) / p[i]
                    if a < alpha: alpha, block, side = a, i, 1
                elif p[i] < 0 and x[i] + p[i] < lb[i]:
                    a = (lb[i] - x[i]) / p[i]
                    if a < alpha: alpha, block, side = a, i, -1
        x = x + alpha * p
        if block >= 0:
            x[block] = ub[block] if side > 0 else lb[block]; W[block] = side
        g = H2 @ x + f
        if pd:
            lam = np.where(W < 0, g, np.where(W > 0, -g, np.inf))
            rel = lam < -mtol
            if it > REL1 and rel.sum() > 1:
                w = int(np.argmin(lam)); rel = np.zeros_like(rel); rel[w] = True
            lo = free & (x < lb); hi = free & (x > ub)
            W[rel] = 0; W[lo] = -1; W[hi] = 1
            x[lo] = lb[lo]; x[hi] = ub[hi]
            if not (rel.any() or lo.any() or hi.any()): done = True
            elif lo.any() or hi.any(): g = H2 @ x + f
        elif block < 0:
            lam = np.where(W < 0, g, np.where(W > 0, -g, np.inf))
            w = int(np.argmin(lam))
            if lam[w] >= -mtol: done = True
            else: W[w] = 0
    return x, W, it

def run(idx, T=300):
    import cases
    case = cases.tank_setup()
    cfg = case["cfg"]
    x0 = np.maximum(np.random.default_rng(20240801).uniform(0, 2, (65536, 2)), 0.0)[idx]
    A, B, C = case["A"].copy(), case["B"].reshape(-1, 1).copy(), case["C"].copy()
    x = x0.copy(); zl = cfg.lift_fn(x); st = None; u_prev = 0.0
    N = cfg.N
    prevU = None; prevW = None
    iters = np.zeros(T, int); nact = np.zeros(T, int); nchg = np.zeros(T, int); cond = np.zeros(T)
    with np.errstate(all="ignore"):
        for k in range(T):
            Aq, Bq, Cy = ocl.qp_model(cfg, A, B, C)
            zq = np.concatenate([zl, [u_prev]])
            lb = np.full(N, cfg.lb); ub = np.full(N, cfg.ub)
            lb[0] = max(lb[0], cfg.u_lb - u_prev); ub[0] = min(ub[0], cfg.u_ub - u_prev)
            H, f = mpc.condense(Aq, Bq, Cy, zq, cfg.r, cfg.q, cfg.rw, N)
            if prevU is None or POLICY == "cold":
                U, status, it = mpc.solve_box_qp_exact(H, f, lb, ub)
                it += 1
                W = np.where(U <= lb, -1, np.where(U >= ub, 1, 0))
            else:
                if POLICY == "shift":
                    xs = np.r_[prevU[1:], 0.0]
                elif POLICY == "keep":
                    xs = prevU.copy()
                xs = np.minimum(np.maximum(xs, lb), ub)
                W0 = np.where(xs <= lb, -1, np.where(xs >= ub, 1, 0))
                if POLICY == "shiftW":   # working set carried (shifted), free variables start from the shifted moves
                    xs = np.r_[prevU[1:], 0.0]; xs = np.minimum(np.maximum(xs, lb), ub)
                    W0 = np.r_[prevW[1:], 0]
                    xs = np.where(W0 < 0, lb, np.where(W0 > 0, ub, xs))
                U, W, it = pdas(H, f, lb, ub, xs.copy(), W0.copy())
            if prevW is not None:
                nchg[k] = (W != np.r_[prevW[1:], 0]).sum()
            iters[k] = it; nact[k] = (W != 0).sum()
            ev = np.linalg.eigvalsh(H); cond[k] = ev[-1] / max(ev[0], 1e-300)
            prevU, prevW = U.copy(), W.copy()
            u = u_prev + U[0]
            p = cfg.p_pre if k < cfg.first_post_step else cfg.p_post
            xn = plant.plant_step(cfg.plant_kind, x, u, np.asarray(p), cfg.h, cfg.rk4_variant)
            yl = cfg.lift_fn(xn)
            first = st is None
            if first: st = rls.RLSState(cfg.nz, 1, 2, cfg.p0, cfg.q0)
            A, B, Cn = rls.rls_update(st, zl, u, yl, x, cfg.lam, cfg.update_c, accumulate_barx=not (cfg.skip_first_barx and first))
            if Cn is not None: C = Cn
            x, zl, u_prev = xn, yl, u
    return iters, nact, nchg, cond

if __name__ == "__main__":
    import multiprocessing as mp
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    with mp.get_context("spawn").Pool(8) as pool:
        res = pool.map(run, range(n))
    it = np.array([r[0] for r in res]); na = np.array([r[1] for r in res]); nc = np.array([r[2] for r in res]); cd = np.array([r[3] for r in res])
    print(POLICY, "mean iters/step", it.mean().round(2), "by 25-step bins", it.reshape(n, -1, 25).mean((0, 2)).round(1))
    print("   nact by bins", na.reshape(n, -1, 25).mean((0, 2)).round(1))
    print("   set changes vs shifted prev by bins", nc.reshape(n, -1, 25).mean((0, 2)).round(1))
    print("   median log10 cond(H) by bins", np.median(np.log10(cd).reshape(n, -1, 25), axis=(0, 2)).round(1))
    print("   share of iterations in steps with it>PI:", (it[it>PI].sum()/it.sum()).round(3), "max", it.max())
