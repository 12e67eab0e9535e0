#!/usr/bin/env python
"""Profiling aid: where along the episode does the fused kernel spend its time?  Runs the bench
workload (vanderpol.py tracking MPC with online update, S = 4096) as consecutive launches of
`chunk` closed-loop steps and prints the CUDA-event time of every launch (us per step)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import koopman_online_updated_mpc_b200 as K  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 10
T = 400
gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_vanderpol.npz"))
enc = K.Encoder.from_file(os.path.join(ROOT, "tests", "golden", "weights", "vdp_model_weights.mat"))
rs = np.random.default_rng(20240601)
x0 = torch.from_numpy(rs.uniform(-2, 2, (S, 2))).cuda()
xref = np.stack([rs.uniform(-1, 1, S), np.zeros(S)], axis=1)
r = enc(torch.from_numpy(xref).cuda())
loop = K.ClosedLoop(K.vanderpol_spec(), x0, gold["A"], gold["B"], gold["C"], r, encoder=enc, log_steps=T)
n = T // chunk
best = np.full(n, 1e30)
for rep in range(4):
    loop.reset()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for k in range(n):
        loop.run(chunk)
        ev[k + 1].record()
    torch.cuda.synchronize()
    if rep:
        best = np.minimum(best, [ev[k].elapsed_time(ev[k + 1]) for k in range(n)])
print("us per step by chunk of %d steps (S=%d):" % (chunk, S), " ".join("%.1f" % (b * 1e3 / chunk) for b in best))
print("sum %.3f ms" % best.sum())
