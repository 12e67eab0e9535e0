"""Snapshot generator on the GPU.  Drop-in for the reference's `data_generate.generate`
(data_generate.py:12-152): `generate(N, N_Traj).duffing_generate()` / `.vanderpol_generate()` return
X, Y (n, M) and U (m, M), M = N * N_Traj, trajectory-major (l.63-74), float64.

The random draws are the reference's own -- `u0 = 4 rand(N, N_Traj) - 2` then
`x0 = 4 rand(n, N_Traj) - 2` from numpy's global legacy stream (data_generate.py:33, 41), so
`np.random.seed(101)` before the call reproduces the reference's snapshot set -- and the N_Traj x N
RK4 steps and the re-ordering run in one CUDA kernel (`kmpc_generate_snapshots`).  For snapshot
sets that should never touch the host use `generate_snapshots` with device-resident x0 / u0."""
import numpy as np
import torch

from . import _lib, plant as _plant
from ._tensors import ptr, stream_ptr, to_dev


def generate_snapshots(x0, u0, params, kind=_plant.PLANT_POLY2, rk4_variant=_plant.RK4_PYTHON, h=0.05):
    """x0 (n_traj, 2), u0 (n_step, n_traj) [reference layout], params (5,) ->
    X, Y (M, 2), U (M,) CUDA tensors, snapshot traj * n_step + j = step j of trajectory traj."""
    x0_d = to_dev(x0).reshape(-1, 2)
    u0_d = to_dev(u0)
    n_traj = x0_d.shape[0]
    if u0_d.ndim != 2 or u0_d.shape[1] != n_traj:
        raise ValueError("u0 must be (n_step, n_traj) = (*, %d), got %s" % (n_traj, tuple(u0_d.shape)))
    n_step = u0_d.shape[0]
    p_d = to_dev(params).reshape(5)
    M = n_traj * n_step
    X = torch.empty((M, 2), dtype=torch.float64, device=x0_d.device)
    Y = torch.empty_like(X)
    U = torch.empty(M, dtype=torch.float64, device=x0_d.device)
    _lib.check(_lib.lib().kmpc_generate_snapshots(ptr(x0_d), ptr(u0_d), ptr(p_d), kind, rk4_variant, float(h),
                                                  n_traj, n_step, ptr(X), ptr(Y), ptr(U), stream_ptr()))
    return X, Y, U


class generate:
    """Same constructor and method names as the reference class (data_generate.py:12-15)."""

    def __init__(self, n_step, n_traj):
        self.N = n_step
        self.N_Traj = n_traj

    def _run(self, params):
        u0 = 4.0 * np.random.rand(self.N, self.N_Traj) - 2.0     # data_generate.py:33
        x0 = 4.0 * np.random.rand(2, self.N_Traj) - 2.0          # data_generate.py:41
        X, Y, U = generate_snapshots(x0.T, u0, params)
        return X.t().cpu().numpy(), Y.t().cpu().numpy(), U.reshape(1, -1).cpu().numpy()

    def duffing_generate(self):
        return self._run(_plant.DUFFING_PRE)

    def vanderpol_generate(self):
        return self._run(_plant.VDP_PRE)
