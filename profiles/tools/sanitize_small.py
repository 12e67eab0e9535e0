#!/usr/bin/env python
"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck /
synccheck are ~100x slower than native: sizes are tiny on purpose).

    compute-sanitizer --tool racecheck python profiles/tools/sanitize_small.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch
    import cases
    import helpers as H
    import koopman_online_updated_mpc_b200 as K
    from koopman_online_updated_mpc_b200 import data_generate as DG, edmd as E, predict as P

    rs = np.random.default_rng(0)
    g = H.golden("ref_vanderpol.npz")
    enc = K.Encoder.from_file(H.weights_path("vdp"))
    z = enc(rs.uniform(-2, 2, (43, 2)))                                   # encoder_units_kernel, ragged
    tank = cases.tank_setup()
    enct = K.Encoder(tank["Ws"], tank["bs"])
    zt = enct(rs.uniform(0, 2, (19, 2)))                                  # 10 outputs: two n-tiles in the last layer
    r = enc(np.array([[1.0, 0.0]]))[0]
    loop = K.ClosedLoop(K.vanderpol_spec(), rs.uniform(-1.5, 1.5, (40, 2)), g["A"], g["B"], g["C"], r, encoder=enc,
                        log_steps=6).run(6)                               # fused_loop_kernel, ragged tile
    assert loop.fused
    gd = H.golden("ref_duffing.npz")
    encd = K.Encoder.from_file(H.weights_path("duffing"))
    K.ClosedLoop(K.duffing_spec(update=False), rs.uniform(-1.5, 1.5, (9, 2)), gd["A"], gd["B"], gd["C"],
                 np.array([1.0, 0.0]), encoder=encd, log_steps=4).run(4)  # fused, frozen model, y = C z
    tl = K.ClosedLoop(K.tank_spec(), np.maximum(rs.uniform(0, 2, (6, 2)), 0), tank["A"], tank["B"], tank["C"],
                      np.array([1.0]), encoder=enct, log_steps=5).run(5)  # generic kernels, warm-started QP
    assert not tl.fused
    gr = H.golden("ref_duffing_rbf.npz")
    K.ClosedLoop(K.rbf_spec(N=50, update=False), rs.uniform(-2, 2, (9, 2)), gr["A"], gr["B"], gr["C"],
                 np.array([1.0, 0.0]), cx=gr["cx"], log_steps=3).run(3)   # N = 50: two row slots per lane
    K.ClosedLoop(K.duffing_spec(), rs.uniform(-1.5, 1.5, (12, 2)), gd["A"], gd["B"], gd["C"], np.array([1.0, 0.0]),
                 encoder=encd, log_steps=5).run(3).run(2)                  # fused, update, y = C z, chunked (warm-start state)
    Xs, Ys, Us = DG.generate(20, 20).duffing_generate()
    cxd = torch.from_numpy(gr["cx"]).cuda()
    pkr = E.gram_accumulate(K.lift.rbf(torch.from_numpy(Xs.T.copy()).cuda(), cxd), K.lift.rbf(torch.from_numpy(Ys.T.copy()).cuda(), cxd),
                            Us.reshape(-1), Xs.T.copy()).cpu().numpy()
    Gm, Aq, XV = pkr[:81].reshape(9, 9), pkr[81:153].reshape(8, 9), pkr[153:171].reshape(2, 9)
    warm = K.RLSState.warm(5, Gm, Aq, XV[:, :8], Gm[:8, :8])
    K.ClosedLoop(K.rbf_spec(N=50), rs.uniform(-2, 2, (5, 2)), gr["A"], gr["B"], gr["C"], np.array([1.0, 0.0]),
                 cx=gr["cx"], rls_state=warm, log_steps=3).run(3)          # N = 50 with the warm-started update (cp.async RLS loads)
    if os.environ.get("KMPC_SANITIZE_TC", "1") == "1" and encd.has_tc:
        ztc = encd(rs.uniform(-2, 2, (300, 2)), precision=K.lift.PREC_TC)  # tcgen05 / TMEM / TMA lift, three tiles, ragged
        assert bool(np.isfinite(np.asarray(ztc.cpu() if hasattr(ztc, "cpu") else ztc)).all())
    X, Y, U = DG.generate_snapshots(rs.uniform(-1, 1, (37, 2)), rs.uniform(-2, 2, (11, 37)), K.plant.DUFFING_PRE)
    pk = E.gram_from_trajectories(encd, X, Y, U, 11)
    A, B, C, st = E.edmd_solve(E.gram_from_snapshots(encd, X, Y, U), 8)
    PHIX = encd(X)
    P.open_loop_predict(PHIX.t(), X.t(), U.reshape(1, -1), A, B, C, 11, reset_every=4, n_seq=3, seq_stride=11)
    torch.cuda.synchronize()
    print("sanitize_small ok", float(z.sum()), float(zt.sum()), int(loop.status.max().item()), int(st.item()))


if __name__ == "__main__":
    main()
