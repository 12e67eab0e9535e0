"""N > 1 host logic on CPU: world_size-2 gloo.  Scenario shards need no collective; the EDMD Gram
pack is all-reduced once and every rank solves redundantly (bitwise-identical A, B, C)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


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
    from koopman_online_updated_mpc_b200 import distributed as D
    from oracle import edmd as oedmd, lift as olift, plant as oplant
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Ws, bs = H.oracle_weights("duffing")
    X, Y, U = oplant.generate_snapshots(40, 50, oplant.DUFFING_PRE, np.random.RandomState(101))
    lo, hi = D.shard_bounds(X.shape[1], rank, world)

    def local_gram():  # stand-in for the CUDA Gram kernel: the test injects the oracle
        PX, PY = olift.encoder_forward(Ws, bs, X[:, lo:hi].T).T, olift.encoder_forward(Ws, bs, Y[:, lo:hi].T).T
        G, Aq, XV = oedmd.gram_pack(PX, PY, U[:, lo:hi], X[:, lo:hi])
        return torch.from_numpy(np.concatenate([G.ravel(), Aq.ravel(), XV.ravel(), [hi - lo]]))

    def solve(pack):
        p = pack.numpy()
        G, Aq, XV = p[:81].reshape(9, 9), p[81:153].reshape(8, 9), p[153:171].reshape(2, 9)
        return oedmd.edmd_from_gram(G, Aq, XV, 8) + (p[171],)

    A, B, C, count = D.edmd_sharded(local_gram, solve)
    assert D.max_over_ranks(rank, "cpu") == world - 1 and D.sum_over_ranks(1.0, "cpu") == world
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), A=A, B=B, C=C, count=count)
    dist.destroy_process_group()


def test_gram_allreduce_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("A", "B", "C"):
        assert np.array_equal(r0[k], r1[k])           # replicated solve is bitwise identical
    assert r0["count"] == 2000
    from oracle import edmd as oedmd, lift as olift, plant as oplant
    Ws, bs = H.oracle_weights("duffing")
    X, Y, U = oplant.generate_snapshots(40, 50, oplant.DUFFING_PRE, np.random.RandomState(101))
    A, B, C = oedmd.edmd_pinv(olift.encoder_forward(Ws, bs, X.T).T, olift.encoder_forward(Ws, bs, Y.T).T, U, X)
    np.testing.assert_allclose(r0["A"], A, atol=1e-9)  # sharded Gram == single-process regression
    np.testing.assert_allclose(r0["C"], C, atol=1e-9)


def test_shard_bounds_partition():
    from koopman_online_updated_mpc_b200.distributed import shard_bounds
    for total in (0, 1, 7, 4096, 65536, 10**7 + 3):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
