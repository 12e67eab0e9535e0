"""A/B timing of the generic-path QP start modes on the Tank loop (BASELINE configs[2] shape) and the
RBF horizon-50 loop: qp_cold = 3 (warm + damped primal-dual sweeps: tank_spec default), 0 (warm + plain sweeps), 2 (warm, primal only: round 1), 1 (cold)."""
import json
import os
import sys
from dataclasses import replace

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import koopman_online_updated_mpc_b200 as K  # noqa: E402
from koopman_online_updated_mpc_b200 import scripts as SC  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
enc = K.Encoder.from_file(os.path.join(ROOT, "tests", "golden", "weights", "tank_model_weights.mat"))
A, B, C, _ = SC.tank_identify(enc)
x0 = np.maximum(np.random.default_rng(20240801).uniform(0, 2, (S, 2)), 0.0)
out = {"S": S, "T": T}
ref = None
for mode in (3, 0, 2, 1):
    loop = K.ClosedLoop(replace(K.tank_spec(), qp_cold=mode), x0, A, B, C, np.array([1.0]), encoder=enc, log_steps=T)
    loop.run(T)
    torch.cuda.synchronize()
    loop.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run(T)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    lx = loop.log_x.cpu().numpy()
    if ref is None:
        ref = lx
    ph = loop.reset().run_timed(min(T, 300))
    out["mode%d" % mode] = {"ms": ms, "scenario_steps_per_s": S * T / ms * 1e3, "status_nonzero": int((loop.status != 0).sum().item()),
                            "max_abs_dx_vs_first_mode": float(np.abs(lx - ref).max()), "phase_ms": ph}
    loop.close()
print(json.dumps(out))
