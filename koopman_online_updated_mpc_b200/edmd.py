"""Stage 2 -- EDMD regression of A, B, C.  Drop-in for duffing.py:167-177
(`K_hat = PHIY @ pinverse([PHIX; U])`, `C = X @ pinv(PHIX)`) in Gram form
(Tank_System.m:93-100), so that a sharded snapshot set only has to all-reduce the Gram pack."""
import torch

from . import _lib
from ._tensors import like_input, ptr, stream_ptr, to_dev

C_PYTHON, C_JOINT = 0, 1


def gram_pack_len(nz, n=2):
    return int(_lib.lib().kmpc_gram_pack_len(nz, n))


def gram_accumulate(psi, psi_next, u, x, pack=None):
    """Snapshot-major inputs: psi, psi_next (M, nz), u (M,) or (M,1), x (M, n).  Adds into `pack`
    (CUDA float64, zero-initialised when None) and returns it."""
    psi_d, psin_d, x_d = to_dev(psi), to_dev(psi_next), to_dev(x)
    u_d = to_dev(u).reshape(-1)
    M, nz = psi_d.shape
    n = x_d.shape[1]
    if pack is None:
        pack = torch.zeros(gram_pack_len(nz, n), dtype=torch.float64, device=psi_d.device)
    _lib.check(_lib.lib().kmpc_gram_accumulate(ptr(psi_d), ptr(psin_d), ptr(u_d), ptr(x_d), M, nz, n,
                                               ptr(pack), stream_ptr()))
    return pack


def gram_from_snapshots(encoder, x, y, u, pack=None, mode=None, precision=0):
    """Fused lift + Gram over raw snapshots x, y (M, n), u (M,).  precision: lift.PREC_FP64 |
    lift.PREC_TC (tcgen05 split-precision lift; the pack is accumulated in fp64 either way)."""
    x_d, y_d = to_dev(x), to_dev(y)
    u_d = to_dev(u).reshape(-1)
    mode = encoder.mode if mode is None else mode
    nz = encoder.out_dim(mode)
    if pack is None:
        pack = torch.zeros(gram_pack_len(nz, x_d.shape[1]), dtype=torch.float64, device=x_d.device)
    _lib.check(_lib.lib().kmpc_gram_from_snapshots_ex(encoder.handle, mode, int(precision), ptr(x_d), ptr(y_d),
                                                      ptr(u_d), x_d.shape[0], ptr(pack), stream_ptr()))
    return pack


def gram_from_trajectories(encoder, x, y, u, n_step, pack=None, mode=None, verify=False, precision=0):
    """Fused lift + Gram over CONSECUTIVE trajectory-major snapshots (what data_generate.py:63-74
    returns: y of snapshot j is x of snapshot j + 1 of the same trajectory): every state is lifted
    once, n_step + 1 encodes per trajectory instead of 2 n_step.  `verify=True` checks the
    precondition on the device (synchronises)."""
    x_d, y_d = to_dev(x), to_dev(y)
    u_d = to_dev(u).reshape(-1)
    M, n = x_d.shape
    if M % n_step:
        raise ValueError("%d snapshots are not a whole number of %d-step trajectories" % (M, n_step))
    if verify:
        xv, yv = x_d.view(-1, n_step, n), y_d.view(-1, n_step, n)
        if not torch.equal(yv[:, :-1], xv[:, 1:]):
            raise ValueError("snapshots are not consecutive within trajectories: use gram_from_snapshots")
    mode = encoder.mode if mode is None else mode
    nz = encoder.out_dim(mode)
    if pack is None:
        pack = torch.zeros(gram_pack_len(nz, n), dtype=torch.float64, device=x_d.device)
    _lib.check(_lib.lib().kmpc_gram_from_trajectories_ex(encoder.handle, mode, int(precision), ptr(x_d), ptr(y_d),
                                                         ptr(u_d), M // n_step, int(n_step), ptr(pack),
                                                         stream_ptr()))
    return pack


def edmd_solve(pack, nz, n=2, c_variant=C_PYTHON):
    """pack -> A (nz,nz), B (nz,1), C (n,nz), status (int tensor, KMPC_STATUS_PIVOT on failure)."""
    dev = pack.device
    A = torch.empty((nz, nz), dtype=torch.float64, device=dev)
    B = torch.empty((nz, 1), dtype=torch.float64, device=dev)
    C = torch.empty((n, nz), dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().kmpc_edmd_solve(ptr(pack), nz, n, c_variant, ptr(A), ptr(B), ptr(C),
                                          ptr(status), stream_ptr()))
    return A, B, C, status


def edmd(PHIX, PHIY, U, X, c_variant=C_PYTHON):
    """Reference call shape (duffing.py:167-177): PHIX, PHIY (nz, M), U (1, M), X (n, M) ->
    A (nz, nz), B (nz, 1), C (n, nz)."""
    psi = to_dev(PHIX).t().contiguous()
    psin = to_dev(PHIY).t().contiguous()
    u = to_dev(U).reshape(-1)
    x = to_dev(X).t().contiguous()
    pack = gram_accumulate(psi, psin, u, x)
    A, B, C, status = edmd_solve(pack, psi.shape[1], x.shape[1], c_variant)
    if int(status.item()) != 0:
        raise _lib.KmpcError("EDMD Gram matrix is not positive definite (status %d): the Cholesky "
                             "path needs a full-rank snapshot set" % int(status.item()))
    return like_input(A, PHIX), like_input(B, PHIX), like_input(C, PHIX)
