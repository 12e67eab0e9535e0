"""Times theta_E on both lift paths (fp64 DMMA vs tcgen05 split precision) and the BASELINE configs[3]
EDMD regression (10 M duffing snapshots) with each.  CUDA events, best of 5 after warm-up.
usage: python profiles/tools/time_tc_lift.py [rows] [n_traj]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import koopman_online_updated_mpc_b200 as K  # noqa: E402
from koopman_online_updated_mpc_b200 import edmd as kedmd  # noqa: E402

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
n_traj = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
enc = K.Encoder.from_file(os.path.join(ROOT, "tests", "golden", "weights", "duffing_model_weights.mat"))
g = torch.Generator(device="cuda").manual_seed(1)
x = (torch.rand((rows, 2), dtype=torch.float64, device="cuda", generator=g) * 4 - 2)
z = torch.empty((rows, 8), dtype=torch.float64, device="cuda")


def best(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return min(t)


out = {"rows": rows, "n_traj": n_traj}
for name, prec in (("fp64", 0), ("tc", 1)):
    ms = best(lambda: enc.encode_into(x, z, precision=prec))
    out["encode_%s_ms" % name] = ms
    out["encode_%s_rows_per_s" % name] = rows / ms * 1e3
    out["encode_%s_tflops" % name] = rows * 42308.0 / ms * 1e3 / 1e12
z64 = enc(x[:100000], precision=0)
ztc = enc(x[:100000], precision=1)
out["max_rel_err_tc_vs_fp64"] = float((ztc - z64).abs().max() / z64.abs().max())
n_step = 100
rs = np.random.default_rng(3)
X, Y, U = K.data_generate.generate_snapshots(torch.from_numpy(rs.uniform(-2, 2, (n_traj, 2))).cuda(),
                                             torch.from_numpy(rs.uniform(-2, 2, (n_step, n_traj))).cuda(),
                                             K.plant.DUFFING_PRE)
mats = {}
for name, prec in (("fp64", 0), ("tc", 1)):
    def run():
        pack = kedmd.gram_from_trajectories(enc, X, Y, U, n_step, precision=prec)
        return kedmd.edmd_solve(pack, 8, 2)
    ms = best(run)
    out["edmd_%s_ms" % name] = ms
    out["edmd_%s_snapshots_per_s" % name] = n_traj * n_step / ms * 1e3
    mats[name] = [m.cpu().numpy() for m in run()[:3]]
out["edmd_ABC_rel_err_tc_vs_fp64"] = [float(np.abs(a - b).max() / np.abs(b).max()) for a, b in zip(mats["tc"], mats["fp64"])]
print(json.dumps(out))
