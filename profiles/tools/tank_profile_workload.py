import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import koopman_online_updated_mpc_b200 as K
from koopman_online_updated_mpc_b200 import scripts as SC
ROOT="/root/repo"
enc = K.Encoder.from_file(os.path.join(ROOT, "tests", "golden", "weights", "tank_model_weights.mat"))
A, B, C, _ = SC.tank_identify(enc)
S = 65536
x0 = np.maximum(np.random.default_rng(20240801).uniform(0, 2, (S, 2)), 0.0)
loop = K.ClosedLoop(K.tank_spec(), x0, A, B, C, np.array([1.0]), encoder=enc, log_steps=0)
loop.run(140)
torch.cuda.synchronize()
