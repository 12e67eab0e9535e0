"""Oracle: rank-1 recursive (Sherman-Morrison) Koopman update (test infrastructure).

python : duffing.py:900,927-953 ; vanderpol.py:821,872-895
matlab : Koopman_update.m:258-278 (forgetting factor lambda, warm start from the offline Gram),
         Koopman_update_Tracking_Lift.m:181-195, Tank_System.m:234-263
Formula order follows the reference (no symmetrisation of P):
    K_A += y v' ;  P = P/lam - (P v v' P)/lam/(lam + v' P v) ;  [A B] = K_A P
    bar_X += xc z' ; bar_Q -= (bar_Q z z' bar_Q)/(1 + z' bar_Q z) ;  C = bar_X bar_Q
"""
import numpy as np


class RLSState:
    def __init__(self, nz, m=1, n=2, p0=1e4, q0=100.0):
        nv = nz + m
        self.KA = np.zeros((nz, nv))
        self.P = p0 * np.eye(nv)
        self.barX = np.zeros((n, nz))
        self.barQ = q0 * np.eye(nz)

    @classmethod
    def warm(cls, G, Aq, XPsi, PsiPsi):
        """Koopman_update.m:264-265 / duffing_RBF.py:434-438 ('storage method' == RLS warm-started
        from the offline Gram)."""
        st = cls.__new__(cls)
        st.KA = np.array(Aq, dtype=np.float64)
        st.P = np.linalg.pinv(G)
        st.barX = np.array(XPsi, dtype=np.float64)
        st.barQ = np.linalg.pinv(PsiPsi)
        return st

    def copy(self):
        st = RLSState.__new__(RLSState)
        st.KA, st.P, st.barX, st.barQ = self.KA.copy(), self.P.copy(), self.barX.copy(), self.barQ.copy()
        return st


def rls_update(st, z, u, y, xc, lam=1.0, update_c=True, accumulate_barx=True):
    """One update with sample (z, u) -> y; xc is the state paired with z in the C regression
    (x_{k+1} in the python scripts, duffing.py:945-950; x_k in Tank_System.m:260).
    accumulate_barx=False reproduces Tank_System.m:252-254 (first bar_X sample skipped).
    Returns A, B, C (C is None when update_c is False, Koopman_update.m never updates C)."""
    z = np.asarray(z, dtype=np.float64).reshape(-1, 1)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    v = np.concatenate([z, np.asarray(u, dtype=np.float64).reshape(-1, 1)], axis=0)
    nz = z.shape[0]
    st.KA = st.KA + y @ v.T
    P = st.P
    st.P = P / lam - (P @ v @ v.T @ P) / lam / (lam + v.T @ P @ v)
    K = st.KA @ st.P
    C = None
    if update_c:
        if accumulate_barx:
            st.barX = st.barX + np.asarray(xc, dtype=np.float64).reshape(-1, 1) @ z.T
        Q = st.barQ
        st.barQ = Q - (Q @ z @ z.T @ Q) / (1.0 + z.T @ Q @ z)
        C = st.barX @ st.barQ
    return K[:, :nz], K[:, nz:], C
