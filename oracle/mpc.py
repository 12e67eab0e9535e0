"""Oracle: the per-step MPC problem (test infrastructure).

literal path : `costFunction` duffing.py:540-581 (vanderpol.py:445-487 tracks the lifted state,
               y = z) minimised by `optimize.minimize(cost, zeros(N), bounds=bounds)` =
               L-BFGS-B with 2-point finite-difference gradient (duffing.py:776-778, 859-861).
condensed    : Tank_System.m:128-159,182-188 / Koopman_update.m:130-142,185-188:
               F = [Cy A; ...; Cy A^N], G = lower block-Toeplitz of Cy A^j B,
               H = G' Qbar G + Rbar, H = (H+H')/2, f = 2 (F z0 - Yr)' Qbar G,
               quadprog(2H, f, ..., lb, ub)  ==  argmin U'HU + f'U  s.t. lb <= U <= ub.
Both describe the same strictly convex box QP (q = 100, rw = 1e-4 in the python scripts).
`solve_box_qp_exact` is the primal active-set method the CUDA kernel implements;
`solve_box_qp_bvls` is an independent cross-check (scipy BVLS on the Cholesky factor).
"""
import numpy as np
from scipy import optimize


def cost_function_literal(U, r, AB, C, z0):
    """duffing.py:540-581 with Np == Nc (tail loop empty) and d = 0.  C=None -> y = z
    (vanderpol.py:456-459).  r: (ny, N)."""
    U = np.asarray(U, dtype=np.float64)
    z = np.asarray(z0, dtype=np.float64).reshape(-1, 1)
    acc = 0.0
    for i, u in enumerate(U):
        zu = np.concatenate([z, np.reshape(u, (1, 1))])
        z = AB @ zu
        y = z if C is None else C @ z
        y = y - np.asarray(r[:, i]).reshape(-1, 1)
        acc += np.sum(np.square(y))
    return 100.0 * acc + 0.0001 * np.sum(np.square(U))


def solve_literal(r, AB, C, z0, lb, ub, N):
    """The reference's own solver call (cold start from zeros every step, duffing.py:634)."""
    res = optimize.minimize(lambda U: cost_function_literal(U, r, AB, C, z0), np.zeros(N),
                            bounds=[(lb, ub)] * N)
    return res.x


def condense(A, B, Cy, z0, r, q, rw, N, PN=None):
    """Return H (N,N), f (N,) of  U'HU + f'U  (constant dropped).  B: (nz,) or (nz,1); Cy (ny,nz);
    r (ny,N) or (ny,) constant; PN optional terminal weight (ny,ny) replacing the last q*I."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64).reshape(-1, 1)
    Cy = np.asarray(Cy, dtype=np.float64)
    ny = Cy.shape[0]
    r = np.asarray(r, dtype=np.float64)
    if r.ndim == 1:
        r = np.repeat(r.reshape(ny, 1), N, axis=1)
    F = np.zeros((N * ny, A.shape[0]))
    G = np.zeros((N * ny, N))
    Ak = np.eye(A.shape[0])
    markov = []
    for k in range(N):
        markov.append(Cy @ Ak @ B)  # Cy A^k B
        Ak = Ak @ A
        F[k * ny:(k + 1) * ny] = Cy @ Ak
    for k in range(N):
        for j in range(k + 1):
            G[k * ny:(k + 1) * ny, j:j + 1] = markov[k - j]
    Qbar = q * np.eye(N * ny)
    if PN is not None:
        Qbar[-ny:, -ny:] = PN
    H = G.T @ Qbar @ G + rw * np.eye(N)
    H = (H + H.T) / 2
    Yr = r.T.reshape(-1, 1)
    f = 2.0 * (F @ np.asarray(z0, dtype=np.float64).reshape(-1, 1) - Yr).T @ Qbar @ G
    return H, f.reshape(-1)


# relative pivot floor of the Cholesky factorisations (same constant in csrc/percase.cuh and
# csrc/fused.cu): a pivot below PIVOT_FLOOR * (its original diagonal entry) is replaced by that
# floor and flagged -- the Delta-u Hessian of Tank_System.m reaches cond ~ 2e16 while the restarted
# RLS model is still rank deficient, where a plain Cholesky breaks down
PIVOT_FLOOR = 1e-13


def _masked_chol_solve(H2, rhs, free):
    """Solve H2[free,free] p = rhs[free] by Cholesky; returns p (zeros elsewhere), ok flag."""
    idx = np.flatnonzero(free)
    p = np.zeros_like(rhs)
    ok = True
    if idx.size:
        M = H2[np.ix_(idx, idx)]
        n = idx.size
        L = np.zeros((n, n))
        for j in range(n):
            d = M[j, j] - L[j, :j] @ L[j, :j]
            floor = PIVOT_FLOOR * M[j, j]
            if not d > floor:       # numerically semi-definite (cond >~ 1e13): regularise, flag
                ok = False
                d = floor
            L[j, j] = np.sqrt(d)
            for i in range(j + 1, n):
                L[i, j] = (M[i, j] - L[i, :j] @ L[j, :j]) / L[j, j]
        yv = np.zeros(n)
        for i in range(n):
            yv[i] = (rhs[idx[i]] - L[i, :i] @ yv[:i]) / L[i, i]
        xv = np.zeros(n)
        for i in range(n - 1, -1, -1):
            xv[i] = (yv[i] - L[i + 1:, i] @ xv[i + 1:]) / L[i, i]
        p[idx] = xv
    return p, ok


STATUS_MAXITER, STATUS_NONFINITE, STATUS_PIVOT = 1, 2, 4


PDAS_ITERS = 8


def solve_box_qp_exact(H, f, lb, ub, max_iter=None, tol=1e-10, pdas_iters=PDAS_ITERS):
    """Exact solve of  min U'HU + f'U,  lb <= U <= ub  (H SPD); the CUDA kernel runs the same steps.
    Start: clipped unconstrained minimiser, clipped variables in the working set W.
    Iterations 0..pdas_iters-1 (primal-dual active-set sweeps): full Newton step on the free face,
    then clip EVERY violated bound into W and release EVERY bound with a negative multiplier; stop
    when nothing changes (KKT holds exactly).  Later iterations (only if the sweeps cycle): the
    monotone primal active-set method -- ratio test, add the blocking bound or drop the most negative
    multiplier -- which always terminates.  Returns U, status, iterations."""
    H = np.asarray(H, dtype=np.float64)
    f = np.asarray(f, dtype=np.float64).reshape(-1)
    N = f.size
    lb = np.broadcast_to(np.asarray(lb, dtype=np.float64), (N,)).copy()
    ub = np.broadcast_to(np.asarray(ub, dtype=np.float64), (N,)).copy()
    if max_iter is None:
        max_iter = 10 * N + 20
    H2 = 2.0 * H
    status = 0
    x, ok = _masked_chol_solve(H2, -f, np.ones(N, dtype=bool))
    if not ok:
        status |= STATUS_PIVOT
    W = np.zeros(N, dtype=np.int64)
    W[x < lb] = -1
    W[x > ub] = 1
    x = np.minimum(np.maximum(x, lb), ub)
    iters = 0
    mtol = tol * max(1.0, np.max(np.abs(f)))
    done = not W.any()
    g = H2 @ x + f
    while not done and iters < max_iter:
        pdas = iters < pdas_iters
        iters += 1
        free = W == 0
        p, ok = _masked_chol_solve(H2, -g, free)
        if not ok:
            status |= STATUS_PIVOT
        alpha, block, side = 1.0, -1, 0
        if not pdas:
            for i in range(N):
                if not free[i]:
                    continue
                if p[i] > 0.0 and x[i] + p[i] > ub[i]:
                    a = (ub[i] - x[i]) / p[i]
                    if a < alpha:
                        alpha, block, side = a, i, 1
                elif p[i] < 0.0 and x[i] + p[i] < lb[i]:
                    a = (lb[i] - x[i]) / p[i]
                    if a < alpha:
                        alpha, block, side = a, i, -1
        x = x + alpha * p
        if block >= 0:
            x[block] = ub[block] if side > 0 else lb[block]
            W[block] = side
        g = H2 @ x + f
        if pdas:
            lam = np.where(W < 0, g, np.where(W > 0, -g, np.inf))
            release = lam < -mtol                      # multipliers at the face minimiser
            lo, hi = free & (x < lb), free & (x > ub)  # violated bounds of the free variables
            W[release] = 0
            W[lo], W[hi] = -1, 1
            x[lo], x[hi] = lb[lo], ub[hi]
            if not (release.any() or lo.any() or hi.any()):
                done = True
            elif lo.any() or hi.any():
                g = H2 @ x + f
        elif block < 0:
            lam = np.where(W < 0, g, np.where(W > 0, -g, np.inf))
            worst = int(np.argmin(lam))
            if lam[worst] >= -mtol:
                done = True
            else:
                W[worst] = 0
    if not done:
        status |= STATUS_MAXITER
    if not np.all(np.isfinite(x)):
        status |= STATUS_NONFINITE
    return x, status, iters


def solve_box_qp_bvls(H, f, lb, ub):
    """Independent check: U'HU + f'U = |R U + 0.5 R^-T f|^2 - const with H = R'R."""
    N = len(f)
    R = np.linalg.cholesky(np.asarray(H)).T
    b = -0.5 * np.linalg.solve(R.T, np.asarray(f, dtype=np.float64))
    res = optimize.lsq_linear(R, b, bounds=(np.broadcast_to(lb, (N,)), np.broadcast_to(ub, (N,))),
                              method="bvls", tol=1e-14, max_iter=1000)
    return res.x


def mpc_first_move_exact(A, B, Cy, z0, r, lb, ub, q, rw, N, PN=None):
    H, f = condense(A, B, Cy, z0, r, q, rw, N, PN)
    U, status, iters = solve_box_qp_exact(H, f, lb, ub)
    return U, status, iters
