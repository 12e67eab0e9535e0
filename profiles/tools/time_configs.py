#!/usr/bin/env python
"""Profiling aid: closed-loop throughput of the other BASELINE configs on one GPU --
configs[2] Tank (Tank_New.mat encoder, nz 10 + du augmentation, N 20), configs[4] duffing RBF with
horizon 50 and warm-started update, plus the RBF N = 10 and duffing theta_E loops for comparison.

    python profiles/tools/time_configs.py [--tank-scenarios 65536] [--rbf-scenarios 125000]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(loop, T, warm):
    import torch
    if warm:
        loop.run(warm)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run(T)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tank-scenarios", type=int, default=65536)
    ap.add_argument("--rbf-scenarios", type=int, default=125000)
    ap.add_argument("--steps", type=int, default=40)
    a = ap.parse_args()
    import numpy as np
    import torch
    import cases
    import koopman_online_updated_mpc_b200 as K

    import helpers as H
    from oracle import edmd as oedmd, lift as olift, plant as oplant

    T = a.steps
    # configs[2]: Tank (Tank_System.m with the Encoder_Tank lift), steps 120 .. 120 + T
    t = cases.tank_setup()
    enc = K.Encoder(t["Ws"], t["bs"])
    rs = np.random.default_rng(3)
    S = a.tank_scenarios
    x0 = np.maximum(rs.uniform(0, 2, (S, 2)), 0.0)
    loop = K.ClosedLoop(K.tank_spec(), x0, t["A"], t["B"], t["C"], np.array([1.0]), encoder=enc)
    for label, warm in (("restart transient, steps 0..%d" % T, 0), ("steps 120..%d" % (120 + T), 120 - T)):
        ms = timed(loop, T, warm) if warm else timed(loop, T, 0)
        print("cfg3 tank S=%d fused=%s %s: %.1f us/step -> %.1f M scenario-steps/s" % (S, loop.fused, label, ms * 1e3 / T, S * T / ms / 1e3))
    print("   phases of %d steps: %s; scenarios with status %d" % (T, {k: round(v, 2) for k, v in loop.run_timed(T).items()},
                                                                   int((loop.status != 0).sum().item())))
    del loop
    # configs[4]: duffing RBF, warm-started update, N = 50 (and N = 10 on the fused kernel)
    g = H.golden("ref_duffing_rbf.npz")
    X, Y, U = oplant.generate_snapshots(100, 100, oplant.DUFFING_PRE, np.random.RandomState(101))
    PX, PY = olift.rbf_lift(X.T, g["cx"]).T, olift.rbf_lift(Y.T, g["cx"]).T
    G, Aq, XV = oedmd.gram_pack(PX, PY, U, X)
    S = a.rbf_scenarios
    x0 = np.random.default_rng(9).uniform(-2, 2, (S, 2))
    for N in (50, 10):
        warm = K.RLSState.warm(S, G, Aq, XV[:, :8], G[:8, :8])
        loop = K.ClosedLoop(K.rbf_spec(N=N), x0, g["A"], g["B"], g["C"], np.array([1.0, 0.0]), cx=g["cx"], rls_state=warm)
        ms0 = timed(loop, T, 0)
        ms1 = timed(loop, T, 0)
        print("cfg5 rbf N=%d S=%d fused=%s: steps 0..%d %.1f us/step -> %.1f M scenario-steps/s; steps %d..%d %.1f us/step -> %.1f M/s"
              % (N, S, loop.fused, T, ms0 * 1e3 / T, S * T / ms0 / 1e3, T, 2 * T, ms1 * 1e3 / T, S * T / ms1 / 1e3))
        print("   phases of %d steps: %s; scenarios with status %d" % (T, {k: round(v, 2) for k, v in loop.run_timed(T).items()},
                                                                       int((loop.status != 0).sum().item())))
        del loop


def timed0(loop, T):
    return timed(loop, T, 0)


if __name__ == "__main__":
    main()
