"""Workload for ncu: the BASELINE configs[4] loop (thin-plate RBF lift, horizon 50, warm-started update), 125 000
scenarios, 60 steps.  ncu -k regex:loop_qp_plant --launch-skip 45 -c 1 captures a steady-state QP launch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import koopman_online_updated_mpc_b200 as K  # noqa: E402
from koopman_online_updated_mpc_b200 import edmd as E  # noqa: E402

dev = torch.device("cuda")
gr = np.load(os.path.join(ROOT, "tests", "golden", "ref_duffing_rbf.npz"))
cx = gr["cx"]
np.random.seed(101)
Xs, Ys, Us = K.data_generate.generate(100, 100).duffing_generate()
cx_d = torch.from_numpy(cx).to(dev)
PX = K.lift.rbf(torch.from_numpy(Xs.T.copy()).to(dev), cx_d)
PY = K.lift.rbf(torch.from_numpy(Ys.T.copy()).to(dev), cx_d)
pack = E.gram_accumulate(PX, PY, Us.reshape(-1), Xs.T.copy())
Ar, Br, Cr, _ = E.edmd_solve(pack, 8)
pk = pack.cpu().numpy()
G, Aq, XV = pk[:81].reshape(9, 9), pk[81:153].reshape(8, 9), pk[153:171].reshape(2, 9)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
x0 = np.random.default_rng(20240901).uniform(-2, 2, (S, 2))
warm = K.RLSState.warm(S, G, Aq, XV[:, :8], G[:8, :8])
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
loop = K.ClosedLoop(K.rbf_spec(N=50, qp_cold=mode), x0, Ar, Br, Cr, np.array([1.0, 0.0]), cx=cx_d, rls_state=warm, log_steps=0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
loop.run(T)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("rbf horizon 50, S=%d, T=%d, qp_cold=%d: %.1f ms, %.2f M scenario-steps/s, scenarios with status %d"
      % (S, T, mode, ms, S * T / ms / 1e3, int((loop.status != 0).sum().item())))
