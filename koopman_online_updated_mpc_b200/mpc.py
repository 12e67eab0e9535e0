"""Stage 4 -- the per-step condensed MPC QP.  Drop-in for
`optimize.minimize(costFunction, zeros(N), bounds=bounds).x[0]` (duffing.py:540-581, 776-778) and
`quadprog(2H, f, ..., lb, ub)` (Tank_System.m:128-159, 188): exact box-QP minimiser of
    sum_j q |Cy z_j - r_j|^2 + rw u_{j-1}^2,   z_j = A z_{j-1} + B u_{j-1},  lb <= u <= ub."""
import torch

from . import _lib
from ._tensors import like_input, ptr, stream_ptr, to_dev

QP_SHARED_MODEL, QP_R_FULL, QP_CY_IDENTITY = 1, 2, 4
STATUS_MAXITER, STATUS_NONFINITE, STATUS_PIVOT = 1, 2, 4


def mpc_first_move(A, B, C, z0, r, lb, ub, N=10, q=100.0, rw=1e-4, PN=None, return_sequence=False,
                   max_iter=0, tol=0.0):
    """A (S,nz,nz) or (nz,nz) shared; B (S,nz[,1]) or (nz[,1]); C (S,ny,nz) / (ny,nz) or None for
    y = z (vanderpol.py:456-459); z0 (S,nz); r (S,ny) constant over the horizon or (S,N,ny);
    lb, ub scalars or (S,N).  Returns u0 (S,) [, U (S,N)], status (S,) int32."""
    z0_d = to_dev(z0)
    if z0_d.ndim == 1:
        z0_d = z0_d.reshape(1, -1)
    S, nz = z0_d.shape
    flags = 0
    if C is None:
        flags |= QP_CY_IDENTITY
        ny = nz
    else:
        C_t = to_dev(C)
        if C_t.ndim < 2 or C_t.shape[-1] != nz:
            raise ValueError("C has shape %s: expected (ny, %d) or (S, ny, %d)" % (tuple(C_t.shape), nz, nz))
        ny = C_t.shape[-2]

    def model(M, tail, name):
        """(tail) shared or (S, tail) per scenario -> (1|S, *tail)"""
        t = to_dev(M)
        numel = 1
        for d in tail:
            numel *= d
        if t.numel() == numel:
            return t.reshape(1, *tail)
        if t.numel() == S * numel and t.shape[0] == S:
            return t.reshape(S, *tail)
        raise ValueError("%s has shape %s: expected %s or (%d, ...)" % (name, tuple(t.shape), tail, S))
    # A, B, C, PN: ONE sharing mode.  The kernel indexes all of them with the same scenario offset,
    # so a shared matrix is expanded as soon as any other is per-scenario.
    mats = {"A": model(A, (nz, nz), "A"), "B": model(B, (nz,), "B")}
    if C is not None:
        mats["C"] = model(C, (ny, nz), "C")
    if PN is not None:
        mats["PN"] = model(PN, (ny, ny), "PN")
    if all(t.shape[0] == 1 for t in mats.values()):
        flags |= QP_SHARED_MODEL
    else:
        mats = {k: t.expand(S, *t.shape[1:]).contiguous() for k, t in mats.items()}
    A_d, B_d = mats["A"].contiguous(), mats["B"].contiguous()
    C_d, PN_d = mats.get("C"), mats.get("PN")
    if C_d is not None:
        C_d = C_d.contiguous()
    if PN_d is not None:
        PN_d = PN_d.contiguous()
    r_d = to_dev(r)
    if r_d.ndim == 1:
        r_d = r_d.reshape(1, -1).expand(S, -1).contiguous()
    if r_d.ndim == 3:
        flags |= QP_R_FULL
        assert r_d.shape == (S, N, ny)
    else:
        assert r_d.shape == (S, ny)
    dev = z0_d.device

    def bound(b):
        if isinstance(b, (int, float)):
            return torch.full((S, N), float(b), dtype=torch.float64, device=dev)
        t = to_dev(b)
        if t.ndim <= 1:
            if t.numel() not in (1, N):
                raise ValueError("bound has %d entries: expected a scalar, (N,) or (S, N)" % t.numel())
            return t.reshape(1, -1).expand(S, N).contiguous()
        if tuple(t.shape) != (S, N):
            raise ValueError("bound has shape %s: expected (%d, %d)" % (tuple(t.shape), S, N))
        return t
    lb_d, ub_d = bound(lb), bound(ub)
    u0 = torch.empty(S, dtype=torch.float64, device=dev)
    U = torch.empty((S, N), dtype=torch.float64, device=dev) if return_sequence else None
    status = torch.zeros(S, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().kmpc_qp_first_move(
        ptr(A_d), ptr(B_d), ptr(C_d), ptr(z0_d), ptr(r_d), ptr(lb_d), ptr(ub_d), ptr(PN_d),
        float(q), float(rw), N, ny, nz, S, flags, ptr(u0), ptr(U), ptr(status), int(max_iter),
        float(tol), stream_ptr()))
    if return_sequence:
        return like_input(u0, z0), like_input(U, z0), like_input(status, z0)
    return like_input(u0, z0), like_input(status, z0)
