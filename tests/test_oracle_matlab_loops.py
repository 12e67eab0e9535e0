"""Oracle restatements of the two `Koopman_update*.m` closed loops (CPU suite).  PARITY UNPINNED
against MATLAB (no MATLAB / Octave in the image, the reference ships no outputs of these scripts):
what is checked here is internal consistency -- the recursions equal the batch regressions they
replace, the QP respects its box, the loops regulate to their set-points like the paper's figures."""
import numpy as np

import cases
from oracle import closed_loop as ocl
from oracle import edmd as oedmd
from oracle import lift as olift
from oracle import rls as orls


def test_koopman_update_m_loop_is_the_batch_regression_with_the_new_samples():
    """Revise_2/Koopman_update.m:258-278: the warm-started RLS (lambda = 1) after T closed-loop steps
    equals the Gram regression over the offline data plus the T closed-loop samples."""
    m = cases.matlab_offline("koopman_update")
    cfg, T = m["cfg"], 100
    warm = orls.RLSState.warm(m["G"], m["Aq"], m["XV"][:, :m["nz"]], m["G"][:m["nz"], :m["nz"]])
    o = ocl.run_loop(cfg, m["A"], m["B"], m["C"], np.array([-1.0, 1.0]), T, update=ocl.UPDATE_RLS, qp="exact", warm=warm)
    assert o["status"].max() == 0 and np.abs(o["U"]).max() <= 2.0 + 1e-12
    Z = o["Z"]                                               # lift(x_k), k = 0..T-1
    Zn = np.concatenate([Z[1:], cfg.lift_fn(o["X"][-1])[None]])
    V = np.concatenate([Z.T, o["U"].reshape(1, -1)], axis=0)
    G = m["G"] + V @ V.T
    Aq = m["Aq"] + Zn.T @ V.T
    A, B, _ = oedmd.edmd_from_gram(G, Aq, m["XV"], m["nz"], oedmd.C_JOINT)
    assert np.abs(o["A"] - A).max() < 1e-8 * max(1.0, np.abs(A).max())
    assert np.abs(o["B"] - B).max() < 1e-8 * max(1.0, np.abs(B).max())
    assert np.array_equal(o["C"], m["C"])                   # C is never updated (l.296-309 commented out)
    # the stacked lift carries the state itself: rows 0..1 of lift(x) are x (l.67)
    assert np.allclose(Z[:, :2], np.concatenate([[[-1.0, 1.0]], o["X"][:-1]]))
    # Q = 10 I_2 also penalises x2, so the approach to Yr = [1; 0] is slow: x1 = 0.77 after the script's
    # 100 steps (0.993 after 300), monotonically
    assert np.all(np.diff(o["X"][:, 0]) > 0) and 0.7 < o["X"][-1, 0] < 1.0


def test_tracking_lift_m_loop_tracks_the_lifted_set_point():
    """VDP_Revise_2/Koopman_update_Tracking_Lift.m: restart P0 = 1e5 I, C = I, set-point
    liftFun([-1; 0]); the loop survives the plant switch at step 100 and returns to x1 = -1."""
    m = cases.matlab_offline("tracking_lift")
    cfg, T = m["cfg"], 300
    o = ocl.run_loop(cfg, m["A"], m["B"], m["C"], np.array([1.0, 1.0]), T, update=ocl.UPDATE_RLS, qp="exact")
    assert o["status"].max() == 0 and np.abs(o["U"]).max() <= 6.0 + 1e-12
    assert np.allclose(cfg.lift_fn(np.zeros(2)), 0.0)        # theta(x) - theta(0) vanishes at the origin (l.65)
    assert abs(o["X"][95, 0] + 1.0) < 0.05 and abs(o["X"][-1, 0] + 1.0) < 0.05
    # RLS from P0 = 1e5 I == ridge regression with weight 1e-5 over the closed-loop samples
    Z = o["Z"]
    Zn = np.concatenate([Z[1:], cfg.lift_fn(o["X"][-1])[None]])
    V = np.concatenate([Z.T, o["U"].reshape(1, -1)], axis=0)
    K = (Zn.T @ V.T) @ np.linalg.inv(V @ V.T + 1e-5 * np.eye(9))
    assert np.abs(np.concatenate([o["A"], o["B"]], axis=1) - K).max() < 1e-6 * np.abs(K).max()
