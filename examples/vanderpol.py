#!/usr/bin/env python
"""vanderpol.py:57 -> 1112 end to end on the B200 kernels (see examples/README.md)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import koopman_online_updated_mpc_b200 as K  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights", default=os.path.join(ROOT, "tests", "golden", "weights", "vdp_model_weights.mat"))
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--scenarios", type=int, default=1)
    ap.add_argument("--out", default=".", help="directory for the .mat files")
    ap.add_argument("--tc", action="store_true", help="EDMD-side lift on the tcgen05 tensor path")
    a = ap.parse_args()
    x0 = np.array([[-2.0, -2.0]])
    if a.scenarios > 1:
        lo, hi = (-2.0, 2.0)
        x0 = np.concatenate([x0, np.random.default_rng(0).uniform(lo, hi, (a.scenarios - 1, 2))])
    prec = K.lift.PREC_TC if a.tc else K.lift.PREC_FP64
    out = K.scripts.run_vanderpol(a.weights, max_step=a.steps, x0=x0, save_dir=a.out, precision=prec)
    print("A[0,0] = %.8f  B[0] = %.8f  C[0,0] = %.8f" % (out["A"][0, 0], out["B"][0, 0], out["C"][0, 0]))
    for k in range(min(a.scenarios, 4)):
        print("scenario %d: frozen model x(T) = %s, online update x(T) = %s, u(T) = %.6f"
              % (k, out["logX"][k][:, -1], out["logXloc"][k][:, -1], out["logUloc"][k][-1]))
    print("status: frozen %s, update %s" % (np.unique(out["status_frozen"]), np.unique(out["status_update"])))


if __name__ == "__main__":
    main()
