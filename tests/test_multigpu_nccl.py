"""N > 1 on real GPUs (NCCL): sharded EDMD with ONE Gram all-reduce and sharded closed loops.
Skipped below two CUDA devices (the CPU suite covers the same host logic with gloo, world size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, H.ROOT)
    sys.path.insert(0, H.HERE)
    import koopman_online_updated_mpc_b200 as K
    from koopman_online_updated_mpc_b200 import data_generate as DG, distributed as D, edmd as E
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    enc = K.Encoder.from_file(H.weights_path("duffing"))
    # the same global draw on every rank, sharded by trajectory
    rs = np.random.RandomState(101)
    n_step, n_traj = 50, 400
    u0 = 4.0 * rs.rand(n_step, n_traj) - 2.0
    x0 = (4.0 * rs.rand(2, n_traj) - 2.0).T
    lo, hi = D.shard_bounds(n_traj, rank, world)
    X, Y, U = DG.generate_snapshots(x0[lo:hi], np.ascontiguousarray(u0[:, lo:hi]), K.plant.DUFFING_PRE)
    A, B, C, st = D.edmd_sharded(lambda: E.gram_from_snapshots(enc, X, Y, U), lambda pack: E.edmd_solve(pack, 8))
    out = {"A": A.cpu().numpy(), "B": B.cpu().numpy(), "C": C.cpu().numpy(), "status": st.cpu().numpy()}
    if rank == 0:   # the whole set on one GPU
        Xa, Ya, Ua = DG.generate_snapshots(x0, u0, K.plant.DUFFING_PRE)
        A1, B1, C1, _ = E.edmd_solve(E.gram_from_snapshots(enc, Xa, Ya, Ua), 8)
        out.update(A1=A1.cpu().numpy(), B1=B1.cpu().numpy(), C1=C1.cpu().numpy())
    # closed loop: contiguous scenario shards, no collective; results gathered for the comparison
    g = H.golden("ref_vanderpol.npz")
    encv = K.Encoder.from_file(H.weights_path("vdp"))
    S, T = 96, 40
    xs = np.random.default_rng(5).uniform(-1.5, 1.5, (S, 2))
    r = encv(np.array([[1.0, 0.0]]))[0]
    slo, shi = D.shard_bounds(S, rank, world)
    loop = K.ClosedLoop(K.vanderpol_spec(), xs[slo:shi], g["A"], g["B"], g["C"], r, encoder=encv, log_steps=T).run(T)
    out.update(lx=loop.log_x.cpu().numpy(), lu=loop.log_u.cpu().numpy(), lo=slo, hi=shi)
    if rank == 0:
        full = K.ClosedLoop(K.vanderpol_spec(), xs, g["A"], g["B"], g["C"], r, encoder=encv, log_steps=T).run(T)
        out.update(lx_full=full.log_x.cpu().numpy(), lu_full=full.log_u.cpu().numpy())
    assert D.max_over_ranks(rank, torch.device("cuda", rank)) == world - 1
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_sharded_edmd_and_closed_loop_world2_nccl(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("A", "B", "C"):
        assert np.array_equal(r0[k], r1[k])                    # replicated solve: bitwise identical
        np.testing.assert_allclose(r0[k], r0[k + "1"], rtol=1e-9, atol=1e-10)   # == unsharded regression
    assert int(r0["status"][0]) == 0
    for r in (r0, r1):                                         # shard invariance of the closed loop
        lo, hi = int(r["lo"]), int(r["hi"])
        assert np.array_equal(r["lx"], r0["lx_full"][:, lo:hi], equal_nan=True)
        assert np.array_equal(r["lu"], r0["lu_full"][:, lo:hi], equal_nan=True)
