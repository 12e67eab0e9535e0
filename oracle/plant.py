"""Oracle: plants (test infrastructure).

RK4 (python): duffing.py:250-261 (k4 uses k3), data_generate.py:24-30, h = 0.05.
RK4 (matlab): Koopman_update.m:21-25 (k4 = f(x + k1*dt) -- uses k1, not k3: reference quirk H7).
Right-hand sides, written as one 2-state polynomial family  p = (p0..p4):
    x1' = p0*x2
    x2' = p1*x2 + p2*x1 + p3*x1^3 + p4*x1^2*x2 + u
  duffing   (duffing.py:255)        : (1, -0.5,  1.0, -1.0,   0)  -> after switch (duffing.py:803)
                                      (1, -5.0,  2.0, -0.5,   0)
  vanderpol (vanderpol.py:252)      : (2,  2.0, -0.8,  0.0, -10)  -> after switch (vanderpol.py:714)
                                      (1, -3.0, -3.0,  0.0, -10)
Tank (Tank_System.m:9-10, switch l.194-195, clamp l.40,45,211): discrete map
    x1+ = x1 - a*sqrt(x1) + b*u ; x2+ = x2 + c*sqrt(x1) - d*sqrt(x2) ; negatives -> 0
  (a,b,c,d) = (0.5,0.4,0.2,0.3) -> (0.53,0.3,0.1,0.35)
"""
import numpy as np

DUFFING_PRE = (1.0, -0.5, 1.0, -1.0, 0.0)
DUFFING_POST = (1.0, -5.0, 2.0, -0.5, 0.0)
VDP_PRE = (2.0, 2.0, -0.8, 0.0, -10.0)
VDP_POST = (1.0, -3.0, -3.0, 0.0, -10.0)
TANK_PRE = (0.5, 0.4, 0.2, 0.3, 0.0)
TANK_POST = (0.53, 0.3, 0.1, 0.35, 0.0)

RK4_PYTHON, RK4_MATLAB = 0, 1


def poly2_rhs(x, u, p):
    """x: (..., 2), u: (...,), p: (..., 5) broadcastable."""
    x1, x2 = x[..., 0], x[..., 1]
    p = np.asarray(p, dtype=np.float64)
    d1 = p[..., 0] * x2
    d2 = p[..., 1] * x2 + p[..., 2] * x1 + p[..., 3] * (x1 * x1 * x1) + p[..., 4] * (x1 * x1 * x2) + u
    return np.stack([d1, d2], axis=-1)


def rk4_step(x, u, p, h=0.05, variant=RK4_PYTHON):
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    k1 = poly2_rhs(x, u, p)
    k2 = poly2_rhs(x + 0.5 * h * k1, u, p)
    k3 = poly2_rhs(x + 0.5 * h * k2, u, p)
    if variant == RK4_PYTHON:
        k4 = poly2_rhs(x + h * k3, u, p)
    else:
        k4 = poly2_rhs(x + h * k1, u, p)
    return x + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def tank_step(x, u, p):
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    s1, s2 = np.sqrt(x[..., 0]), np.sqrt(x[..., 1])
    n1 = x[..., 0] - p[..., 0] * s1 + p[..., 1] * u
    n2 = x[..., 1] + p[..., 2] * s1 - p[..., 3] * s2
    out = np.stack([n1, n2], axis=-1)
    out[out < 0] = 0.0
    return out


PLANT_POLY2, PLANT_TANK = 0, 1


def plant_step(kind, x, u, p, h=0.05, variant=RK4_PYTHON):
    if kind == PLANT_POLY2:
        return rk4_step(x, u, p, h, variant)
    return tank_step(x, u, p)


def generate_snapshots(n_step, n_traj, p, rs, h=0.05):
    """data_generate.py:17-74 (duffing) / 82-152 (vanderpol): u0 = 4*rand(N,N_Traj)-2, then
    x0 = 4*rand(n,N_Traj)-2 from the numpy legacy global stream `rs`; N vectorised RK4 steps;
    output re-ordered trajectory-major (l.63-74).  Returns X, Y (2, M), U (1, M)."""
    u0 = 4.0 * rs.rand(n_step, n_traj) - 2.0
    x = (4.0 * rs.rand(2, n_traj) - 2.0).T  # (n_traj, 2)
    Xs, Ys = [], []
    for i in range(n_step):
        xn = rk4_step(x, u0[i], np.asarray(p))
        Xs.append(x)
        Ys.append(xn)
        x = xn
    X = np.stack(Xs, axis=1).reshape(n_traj * n_step, 2).T  # trajectory-major
    Y = np.stack(Ys, axis=1).reshape(n_traj * n_step, 2).T
    U = u0.T.reshape(1, n_traj * n_step)
    return X, Y, U
