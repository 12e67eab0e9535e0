"""Stage 3 -- online recursive (rank-1 / Sherman-Morrison) Koopman update.
Drop-in for duffing.py:900,927-953,965-984 / Koopman_update.m:258-278 / Tank_System.m:234-263,
batched over scenarios.  fp64 only (fp32 diverges: P0 = 1e4..1e5)."""
import torch

from . import _lib
from ._tensors import ptr, stream_ptr, to_dev

RLS_UPDATE_C, RLS_SKIP_BARX = 1, 2


class RLSState:
    """Per-scenario RLS state on the device: KA (S,nz,nz+1), P (S,nz+1,nz+1), barX (S,n,nz),
    barQ (S,nz,nz) plus the current model A (S,nz,nz), B (S,nz,1), C (S,n,nz)."""

    def __init__(self, S, nz, n=2, p0=1e4, q0=100.0, device="cuda"):
        nv = nz + 1
        f64 = dict(dtype=torch.float64, device=device)
        self.S, self.nz, self.n = S, nz, n
        self.KA = torch.zeros((S, nz, nv), **f64)
        self.P = (p0 * torch.eye(nv, **f64)).repeat(S, 1, 1).contiguous()
        self.barX = torch.zeros((S, n, nz), **f64)
        self.barQ = (q0 * torch.eye(nz, **f64)).repeat(S, 1, 1).contiguous()
        self.A = torch.zeros((S, nz, nz), **f64)
        self.B = torch.zeros((S, nz, 1), **f64)
        self.C = torch.zeros((S, n, nz), **f64)

    @classmethod
    def warm(cls, S, G, Aq, XPsi, PsiPsi, device="cuda"):
        """Warm start from an offline Gram (Koopman_update.m:264-265; the 'storage method' of
        duffing_RBF.py:434-438 is this plus rank-1 updates).  The two small inverses are taken
        once on the host in float64."""
        import numpy as np
        G, Aq, XPsi, PsiPsi = (np.asarray(M.cpu() if hasattr(M, "cpu") else M, dtype=np.float64)
                               for M in (G, Aq, XPsi, PsiPsi))
        nz, n = Aq.shape[0], XPsi.shape[0]
        st = cls(S, nz, n, device=device)
        f64 = dict(dtype=torch.float64, device=device)
        st.KA = torch.tensor(Aq, **f64).repeat(S, 1, 1).contiguous()
        st.P = torch.tensor(np.linalg.pinv(G), **f64).repeat(S, 1, 1).contiguous()
        st.barX = torch.tensor(XPsi, **f64).repeat(S, 1, 1).contiguous()
        st.barQ = torch.tensor(np.linalg.pinv(PsiPsi), **f64).repeat(S, 1, 1).contiguous()
        return st

    def clone(self):
        """Deep copy (same device)."""
        c = object.__new__(RLSState)
        c.S, c.nz, c.n = self.S, self.nz, self.n
        for k in ("KA", "P", "barX", "barQ", "A", "B", "C"):
            setattr(c, k, getattr(self, k).clone())
        return c

    def state_dict(self):
        return {k: getattr(self, k) for k in ("KA", "P", "barX", "barQ", "A", "B", "C")}

    def load_state_dict(self, d):
        for k, v in d.items():
            getattr(self, k).copy_(v)


def rls_update(state, z, u, y, xc, lam=1.0, update_c=True, skip_barx=False):
    """One update for all scenarios with samples (z, u) -> y; xc (S, n) is the plant state paired
    with z in the C regression (x_{k+1} in duffing.py:945-950, x_k in Tank_System.m:260).
    Updates `state` in place and returns (A, B, C) views of it."""
    z_d, y_d, xc_d = to_dev(z), to_dev(y), to_dev(xc)
    u_d = to_dev(u).reshape(-1)
    flags = (RLS_UPDATE_C if update_c else 0) | (RLS_SKIP_BARX if skip_barx else 0)
    _lib.check(_lib.lib().kmpc_rls_update(
        ptr(state.KA), ptr(state.P), ptr(state.barX), ptr(state.barQ), ptr(z_d), ptr(u_d), ptr(y_d),
        ptr(xc_d), ptr(state.A), ptr(state.B), ptr(state.C), state.S, state.nz, state.n, float(lam),
        flags, stream_ptr()))
    return state.A, state.B, state.C
