#!/usr/bin/env python
"""bench.py -- closed-loop Koopman-MPC scenario-steps/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1], SURVEY.md 8d "cfg 2"): vanderpol.py tracking MPC with online
Koopman update, 4096 synthetic scenarios per GPU -- x0 ~ U[-2,2]^2, per-scenario set-point
x1* ~ U[-1,1] lifted through theta_E, bounds +-6, N = 10, nz = ny = 8, RLS restart P0 = 1e5 I,
plant switch at step 102, numpy default_rng(20240601 + rank).  A bench "step" is ONE closed-loop
step of every scenario of the batch (lift -> condensed box-QP -> plant -> lift -> RLS update), so
value = S * K / time [scenario-steps/s], whole job over all ranks (weak scaling: S per GPU fixed,
scenarios are independent, no data-path collective).

Timing: W warm-up steps, then K steps each bracketed by its own pair of CUDA events on the
launching stream with an L2 flush (256 MiB write) between steps -- the per-scenario state (18 MB
at S = 4096) would otherwise sit in the 126 MB L2; barrier + synchronize on both sides; max over
ranks.  `e2e` is the same loop driven from HOST buffers: every step copies the measured states
(pinned host memory) to the device, runs one step through the public API, reads the applied
controls and next states back, and synchronises.

`--impl reference` times the reference's own algorithm on the host cores: the oracle's literal
path (float64 numpy + the same scipy L-BFGS-B call on the rollout cost, duffing.py:776-778), one
process per core on disjoint scenarios.  The reference is Python scripts + MATLAB and cannot be
pip-installed or shipped to the GPU box, so this arm is `cpu_baseline.kind = "port"`.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "closed_loop_koopman_mpc_scenario_steps_per_sec"
UNIT = "scenario-steps/s"
BYTES_PER_SCENARIO_STEP = 4536  # SURVEY.md 8d / BASELINE.md 4, nz = 8, fp64, update variant
# algorithmic split of those bytes over the three kernels of a step (DESIGN.md "Kernels")
KERNEL_BYTES = {"rls": 3728 + 704, "qp_plant": 16 + 16 + 8, "lift": 64}
GOLD = os.path.join(ROOT, "tests", "golden", "ref_vanderpol.npz")
WEIGHTS = os.path.join(ROOT, "tests", "golden", "weights", "vdp_model_weights.mat")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenarios", type=int, default=4096, help="scenarios per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=200)
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(job):
    """One process = one host core = one scenario of the same synthetic workload, literal path."""
    idx, seed, warmup, steps = job
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    import numpy as np
    from oracle import closed_loop as ocl
    from oracle import weights as ow
    gold = np.load(GOLD)
    Ws, bs = ow.load_mat_encoder(WEIGHTS)
    rs = np.random.default_rng(seed)
    x0 = rs.uniform(-2, 2, (4096, 2))[idx % 4096]
    xref = np.array([rs.uniform(-1, 1, 4096)[idx % 4096], 0.0])
    cfg = ocl.vanderpol_config(Ws, bs, xref)
    o = ocl.run_loop(cfg, gold["A"], gold["B"], gold["C"], x0, warmup, update=ocl.UPDATE_RLS, qp="literal")
    t0 = time.perf_counter()
    ocl.run_loop(cfg, o["A"], o["B"], o["C"], o["x"], steps, update=ocl.UPDATE_RLS, qp="literal",
                 warm=o["rls"], u_prev=o["u_prev"], start_step=warmup)
    return time.perf_counter() - t0


def cpu_reference(warmup, steps, seed=20240601):
    import multiprocessing as mp
    cores = min(len(os.sched_getaffinity(0)), 64)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        elapsed = pool.map(_cpu_worker, [(i, seed, warmup, steps) for i in range(cores)])
        wall = time.perf_counter() - t0
    value = cores * steps / max(elapsed)
    sample = ("%d scenarios (one per core) x %d closed-loop steps after %d warm-up steps, literal path "
              "(numpy float64 + scipy L-BFGS-B on the rollout cost), wall %.1f s" % (cores, steps, warmup, wall))
    return value, cores, sample, max(elapsed) / steps * 1e3


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, cores, sample, ms = cpu_reference(args.warmup, args.steps)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, cores),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, s_per_step):
    return {"workload": "vanderpol.py tracking MPC with online Koopman (RLS) update, BASELINE configs[1]",
            "scenarios_per_gpu": args.scenarios, "scenarios_per_step": s_per_step, "horizon": 10, "nz": 8,
            "ny": 8, "bounds": 6.0, "encoder": "AutoEncoder_20220414_4 (2-100-100-100-8, fp64)",
            "plant_switch_step": 102, "l2": "flushed between timed steps (256 MiB write)",
            "parallelism": "independent scenario shards, no collective"}


# ----------------------------------------------------------------------------- clocks sampling
def clocks_start(index):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
        return None


def clocks_stop(proc):
    if proc is None:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    proc.terminate()
    try:
        out, _ = proc.communicate(timeout=5)
    except subprocess.TimeoutExpired:
        proc.kill()
        out, _ = proc.communicate()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in out.strip().splitlines():
        f = [t.strip() for t in ln.split(",")]
        if len(f) < 9:
            continue
        try:
            sm.append(float(f[1]))
            mx.append(float(f[2]))
        except ValueError:
            continue
        for name, v in zip(names, f[5:9]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- B200 arm
def main_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import koopman_online_updated_mpc_b200 as K
    from koopman_online_updated_mpc_b200 import distributed as D

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference(10, args.cpu_steps)   # before any GPU timing
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    gold = np.load(GOLD)
    enc = K.Encoder.from_file(WEIGHTS)
    S, Kst, W = args.scenarios, args.steps, args.warmup
    rs = np.random.default_rng(20240601 + rank)
    x0 = rs.uniform(-2, 2, (S, 2))
    xref = np.stack([rs.uniform(-1, 1, S), np.zeros(S)], axis=1)
    r = enc(torch.from_numpy(xref).to(dev))
    loop = K.ClosedLoop(K.vanderpol_spec(), torch.from_numpy(x0).to(dev), gold["A"], gold["B"], gold["C"], r,
                        encoder=enc)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    loop.run(W)
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(Kst)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(Kst)]
    mon = clocks_start(local) if rank == 0 else None
    launches0 = K.launch_count()
    for k in range(Kst):
        flush.zero_()
        starts[k].record()
        loop.run(1)
        stops[k].record()
    launches = K.launch_count() - launches0
    barrier()
    clocks = clocks_stop(mon) if rank == 0 else None
    ms_total = D.max_over_ranks(sum(s.elapsed_time(e) for s, e in zip(starts, stops)), dev)
    value = world * S * Kst / (ms_total * 1e-3)

    # same loop, back to back without the flush (state L2-resident): reported beside the headline
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run(Kst)
    e1.record()
    barrier()
    ms_resident = D.max_over_ranks(e0.elapsed_time(e1), dev)

    # per-kernel device time (CUDA events recorded inside the library on the launching stream)
    kt = loop.run_timed(min(Kst, 200))
    nk = min(Kst, 200)

    # end to end from host buffers through the public API
    x_host = torch.empty((S, 2), dtype=torch.float64).pin_memory()
    xn_host = torch.empty((S, 2), dtype=torch.float64).pin_memory()
    u_host = torch.empty(S, dtype=torch.float64).pin_memory()
    x_host.copy_(loop.x)
    barrier()
    t0 = time.perf_counter()
    for k in range(Kst):
        loop.x.copy_(x_host, non_blocking=True)       # measured states arrive from the host
        loop.run(1)
        u_host.copy_(loop.u_prev, non_blocking=True)  # applied controls go back
        xn_host.copy_(loop.x, non_blocking=True)
        torch.cuda.synchronize()
        x_host, xn_host = xn_host, x_host
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    status_bad = int((loop.status != 0).sum().item())
    finite = bool(torch.isfinite(loop.x).all().item())

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        ms_step = ms_total / Kst
        achieved = S * BYTES_PER_SCENARIO_STEP / (ms_step * 1e-3) / 1e9
        dominant = max(kt, key=kt.get)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_step")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": Kst, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, S * world),
            "roofline": {
                "bound": "hbm", "kernel": "closed-loop step = qp_plant + lift + rls kernels",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "bytes_per_scenario_step": BYTES_PER_SCENARIO_STEP,
                "dominant_kernel": dominant,
                "kernel_ms_per_step": {k: v / nk for k, v in kt.items()},
                "kernel_hbm_gbs": {k: S * KERNEL_BYTES[k] / (v / nk * 1e-3) / 1e9 for k, v in kt.items() if v > 0},
                "note": "the step is fp64-latency/issue bound, not HBM bound: see DESIGN.md",
            },
            "value_l2_resident": world * S * Kst / (ms_resident * 1e-3),
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e": {"value": world * S * Kst / e2e_s, "unit": UNIT, "h2d_bytes_per_step": S * 2 * 8,
                    "d2h_bytes_per_step": S * 3 * 8},
            "gpu_launches": launches,
            "health": {"scenarios_with_status": status_bad, "finite": finite},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_b200(a)
