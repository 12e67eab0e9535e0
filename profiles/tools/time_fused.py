#!/usr/bin/env python
"""Profiling aid: time the fused closed-loop kernel on the bench workload (vanderpol.py tracking MPC
with online update) for one scenario count and episode length, optionally with phases skipped
(KMPC_FUSED_SKIP bit 0 QP, bit 1 lift, bit 2 RLS -- results are then meaningless, timing only).

    python profiles/tools/time_fused.py [--scenarios S] [--episode T] [--reps R] [--steady]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=4096)
    ap.add_argument("--episode", type=int, default=400)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--steady", action="store_true", help="time steps T..2T instead of an episode from x0")
    ap.add_argument("--spec", default="vdp", choices=["vdp", "duffing"])
    a = ap.parse_args()
    import numpy as np
    import torch
    import koopman_online_updated_mpc_b200 as K

    name = {"vdp": "vanderpol", "duffing": "duffing"}[a.spec]
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name))
    enc = K.Encoder.from_file(os.path.join(ROOT, "tests", "golden", "weights", "%s_model_weights.mat" % a.spec))
    S, T = a.scenarios, a.episode
    rs = np.random.default_rng(20240601)
    x0 = torch.from_numpy(rs.uniform(-2, 2, (S, 2))).cuda()
    if a.spec == "vdp":
        xref = np.stack([rs.uniform(-1, 1, S), np.zeros(S)], axis=1)
        r = enc(torch.from_numpy(xref).cuda())
        spec = K.vanderpol_spec()
    else:
        r = np.array([1.0, 0.0])
        spec = K.duffing_spec()
    loop = K.ClosedLoop(spec, x0, gold["A"], gold["B"], gold["C"], r, encoder=enc, log_steps=T)
    best = 1e30
    for rep in range(a.reps + 2):
        loop.reset()
        if a.steady:
            loop.run(T)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop.run(T)
        e1.record()
        torch.cuda.synchronize()
        if rep >= 2:
            best = min(best, e0.elapsed_time(e1))
    kt = loop.reset().run_timed(T)
    print("S=%d T=%d spec=%s steady=%d skip=%s fused=%s: %.3f ms/launch, %.2f us/step, %.1f M scenario-steps/s; phases %s"
          % (S, T, a.spec, a.steady, os.environ.get("KMPC_FUSED_SKIP", "0"), loop.fused, best, best * 1e3 / T,
             S * T / best / 1e3, {k: round(v, 3) for k, v in kt.items()}))


if __name__ == "__main__":
    main()
