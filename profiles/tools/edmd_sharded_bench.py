#!/usr/bin/env python
"""BASELINE configs[3]: EDMD regression over 10 M synthetic duffing snapshots (100 000 trajectories
x 100 steps generated like data_generate.py:17-57: u ~ U[-2,2], x0 ~ U[-2,2]^2, RK4 h = 0.05),
sharded by trajectory over the ranks, ONE Gram all-reduce (NCCL), replicated solve.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/tools/edmd_sharded_bench.py
Timed on the device (CUDA events, max over ranks): snapshot generation + theta_E lift + Gram
accumulation + all-reduce + solve, with x0 / u0 resident in HBM.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trajectories", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--two-encodes", action="store_true",
                    help="lift x and y of every snapshot like the reference (default: one encode per state)")
    a = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import koopman_online_updated_mpc_b200 as K
    from koopman_online_updated_mpc_b200 import data_generate as DG, distributed as D, edmd as E

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    enc = K.Encoder.from_file(os.path.join(ROOT, "tests", "golden", "weights", "duffing_model_weights.mat"))
    lo, hi = D.shard_bounds(a.trajectories, rank, world)
    rs = np.random.default_rng(20240601 + rank)
    u0 = torch.from_numpy(rs.uniform(-2, 2, (a.steps, hi - lo))).to(dev)
    x0 = torch.from_numpy(rs.uniform(-2, 2, (hi - lo, 2))).to(dev)
    times, phases = [], None
    for rep in range(a.reps + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev[0].record()
        X, Y, U = DG.generate_snapshots(x0, u0, K.plant.DUFFING_PRE)
        ev[1].record()
        pack = (E.gram_from_snapshots(enc, X, Y, U) if a.two_encodes
                else E.gram_from_trajectories(enc, X, Y, U, a.steps))
        ev[2].record()
        D.allreduce_pack(pack)
        A, B, C, st = E.edmd_solve(pack, 8)
        ev[3].record()
        torch.cuda.synchronize()
        if rep >= 2:
            times.append(D.max_over_ranks(ev[0].elapsed_time(ev[3]), dev))
            phases = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    gathered = [torch.empty_like(A) for _ in range(world)]
    if world > 1:
        dist.all_gather(gathered, A)
    same = all(torch.equal(g, A) for g in gathered) if world > 1 else True
    if rank == 0:
        ms = min(times)
        M = a.trajectories * a.steps
        print(json.dumps({"workload": "EDMD over %d synthetic duffing snapshots (BASELINE configs[3])" % M,
                          "n_gpus": world, "ms": ms, "snapshots_per_s": M / ms * 1e3,
                          "phase_ms_rank0": {"generate": phases[0], "lift_gram": phases[1], "allreduce_solve": phases[2]},
                          "encodes_per_snapshot": 2 if a.two_encodes else (a.steps + 1) / a.steps,
                          "theta_E_tflops": (2 if a.two_encodes else (a.steps + 1) / a.steps) * M * 42308 / ms / 1e9, "status": int(st.item()),
                          "A_bitwise_identical_on_all_ranks": bool(same), "A00": float(A[0, 0].item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
