"""CPU study behind DESIGN.md 3.1 (adaptive warm start of the fused kernel's QP): active-set sweeps per step of
the BENCH workload (vanderpol.py tracking MPC + online update, the bench's own draw) under the kernel's
primal-dual active-set method, for three warm-start policies.  Oracle arithmetic (numpy), kernel semantics
(sweep counts per scenario, per warp of 4 and per quarter of 8 scenarios in lock step).

    POLICY=shift|noshift|adapt|adapt3 [PI=8] [REL1=99] python tests/studies/qp_sweeps_bench_workload.py [n_scenarios]

shift = round 1 (last optimum shifted by one move), noshift = kept, adapt = whichever matched last step (what
the kernel does), adapt3 = plus the optimum of two steps ago.  PI / REL1: sweeps before the primal fallback /
single-release damping from that sweep on.  Test infrastructure: lives under tests/ because it imports the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ[k] = "1"
from oracle import closed_loop as ocl, mpc, plant, rls, weights as ow, lift

GOLD = os.path.join(ROOT, "tests/golden/ref_vanderpol.npz")
WEIGHTS = os.path.join(ROOT, "tests/golden/weights/vdp_model_weights.mat")

def chol_solve(H2, rhs, free):
    p, ok = mpc._masked_chol_solve(H2, rhs, free)
    return p

PI = int(os.environ.get('PI', '8')); REL1 = int(os.environ.get('REL1', '99'))
def warm_pdas(H, f, lb, ub, wlo, whi, tol=1e-10, pdas_iters=8, max_iter=120):
    pdas_iters = PI
    N = f.size
    H2 = 2 * H
    mtol = tol * max(1.0, np.max(np.abs(f)))
    x0 = min(max(0.0, lb), ub)
    x = np.where(wlo, lb, np.where(whi, ub, x0))
    g = H2 @ x + f
    done = False
    it = 0
    hist = []
    while not done and it < max_iter:
        pdas = it < pdas_iters
        masked = wlo | whi
        hist.append((wlo.copy(), whi.copy()))
        p = chol_solve(H2, -g, ~masked)
        it += 1
        if not pdas:
            alpha, block, side = 1.0, -1, 0
            for i in range(N):
                if masked[i]: continue
                if p[i] > 0 and x[i] + p[i] > ub:
                    a = (ub - x[i]) / p[i]
                    if a < alpha: alpha, block, side = a, i, 1
                elif p[i] < 0 and x[i] + p[i] < lb:
                    a = (lb - x[i]) / p[i]
                    if a < alpha: alpha, block, side = a, i, -1
            x = x + alpha * p
            if block >= 0:
                x[block] = ub if side > 0 else lb
                if side > 0: whi[block] = True
                else: wlo[block] = True
            g = H2 @ x + f
            if block < 0:
                lam = np.where(wlo, g, np.where(whi, -g, np.inf))
                w = int(np.argmin(lam))
                if lam[w] >= -mtol: done = True
                else: wlo[w] = False; whi[w] = False
            continue
        xn = x + p
        x = xn.copy()
        g = H2 @ x + f
        lam = np.where(wlo, g, np.where(whi, -g, np.inf))
        rel = lam < -mtol
        if it > REL1 and rel.sum() > 1:
            w = int(np.argmin(lam)); rel = np.zeros_like(rel); rel[w] = True
        lo = (~masked) & (x < lb)
        hi = (~masked) & (x > ub)
        wlo = (wlo & ~rel) | lo
        whi = (whi & ~rel) | hi
        x[lo] = lb
        x[hi] = ub
        if not (rel.any() or lo.any() or hi.any()):
            done = True
        elif lo.any() or hi.any():
            g = H2 @ x + f
    return x, wlo, whi, it, hist

POLICY = os.environ.get('POLICY', 'shift')
def run(idx, seed=20240601 + int(os.environ.get("RANK", "0")), T=400, S=4096):
    noshift = False; prev = None; pol = 0; p2lo = np.zeros(10, bool); p2hi = np.zeros(10, bool)
    gold = np.load(GOLD)
    Ws, bs = ow.load_mat_encoder(WEIGHTS)
    rs = np.random.default_rng(seed)
    x0 = rs.uniform(-2, 2, (S, 2))[idx]
    xref = np.array([rs.uniform(-1, 1, S)[idx], 0.0])
    cfg = ocl.vanderpol_config(Ws, bs, xref)
    A, B, C = gold["A"].copy(), gold["B"].reshape(-1, 1).copy(), gold["C"].copy()
    x = x0.copy()
    zl = cfg.lift_fn(x)
    st = None
    wlo = np.zeros(10, bool); whi = np.zeros(10, bool)
    iters = np.zeros(T, int); nact = np.zeros(T, int); U = np.zeros(T)
    sets = []
    with np.errstate(all="ignore"):
        for k in range(T):
            H, f = mpc.condense(A, B, np.eye(8), zl, cfg.r, cfg.q, cfg.rw, cfg.N)
            sl = np.r_[wlo[1:], wlo[-1]]; sh = np.r_[whi[1:], whi[-1]]
            if POLICY == 'shift' or (POLICY == 'adapt' and not noshift): wl, wh = sl, sh
            else: wl, wh = wlo.copy(), whi.copy()
            pl, ph = wlo.copy(), whi.copy()
            if POLICY == 'adapt3':
                cands = [(sl, sh), (pl, ph), (p2lo, p2hi)]
                wl, wh = cands[pol][0].copy(), cands[pol][1].copy()
            try:
                xs, wlo, whi, it, hist = warm_pdas(H, f, cfg.lb, cfg.ub, wl, wh)
            except AssertionError:
                it = 99
                xs, _, _ = mpc.solve_box_qp_exact(H, f, cfg.lb, cfg.ub)
                wlo = xs <= cfg.lb; whi = xs >= cfg.ub
            if not np.all(np.isfinite(xs)):
                iters[k:] = 1
                break
            if POLICY == 'adapt3':
                match = [(wlo == c[0]).all() and (whi == c[1]).all() for c in cands]
                if not match[pol]:
                    for j in range(3):
                        if match[j]: pol = j; break
                p2lo, p2hi = pl, ph
            if POLICY == 'adapt':
                same_un = (wlo == pl).all() and (whi == ph).all()
                same_sh = (wlo == sl).all() and (whi == sh).all()
                if same_un and not same_sh: noshift = True
                elif same_sh and not same_un: noshift = False
            iters[k] = it; nact[k] = (wlo | whi).sum(); U[k] = xs[0]
            sets.append((wl, wh, wlo.copy(), whi.copy()))
            p = cfg.p_pre if k < cfg.first_post_step else cfg.p_post
            xn = plant.plant_step(cfg.plant_kind, x, xs[0], np.asarray(p), cfg.h, cfg.rk4_variant)
            if not np.all(np.isfinite(xn)):
                iters[k:] = 1
                break
            yl = cfg.lift_fn(xn)
            if st is None:
                st = rls.RLSState(cfg.nz, 1, 2, cfg.p0, cfg.q0)
            A, B, Cn = rls.rls_update(st, zl, xs[0], yl, xn, cfg.lam, cfg.update_c)
            x, zl = xn, yl
    return idx, iters, nact, U, sets

if __name__ == "__main__":
    import multiprocessing as mp
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    t0 = time.time()
    with mp.get_context("spawn").Pool(16) as pool:
        res = pool.map(run, range(n))
    print("time", time.time() - t0)
    it = np.array([r[1] for r in res])
    print("mean iters/step", it.mean(), "per-scenario total: median", np.median(it.sum(1)), "max", it.sum(1).max())
    tot = it.sum(1)
    print("quantiles of per-scenario total", np.quantile(tot, [0.5, 0.75, 0.9, 0.95, 0.99]))
    # warps of 4, quarters of 8
    w4 = it.reshape(-1, 4, it.shape[1]).max(1).sum(1)
    q8 = it.reshape(-1, 8, it.shape[1]).max(1).sum(1)
    print("warp totals: median", np.median(w4), "max", w4.max(), "| quarter totals: median", np.median(q8), "max", q8.max())
    print("by phase: steps 0-100", it[:, :100].mean(), "100-200", it[:, 100:200].mean(), "200-400", it[:, 200:].mean())
