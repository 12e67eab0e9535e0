#!/usr/bin/env python
"""bench.py -- closed-loop Koopman-MPC scenario-steps/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1], SURVEY.md 8d "cfg 2"): vanderpol.py tracking MPC with online
Koopman update, 4096 synthetic scenarios per GPU -- x0 ~ U[-2,2]^2, per-scenario set-point
x1* ~ U[-1,1] lifted through theta_E, bounds +-6, N = 10, nz = ny = 8, RLS restart P0 = 1e5 I,
plant switch at step 102, T = 400 closed-loop steps, numpy default_rng(20240601 + rank).

A bench "step" is one pass of the hot path over the batch = one EPISODE: all T = 400 closed-loop
steps (lift -> condensed box-QP -> plant -> lift -> RLS update) of every scenario, from x0, through
ONE `kmpc_closed_loop_steps` call (one persistent fused kernel launch).  value = S * T * K / time
[scenario-steps/s], whole job over all ranks (weak scaling: S per GPU fixed, scenarios are
independent, no data-path collective).

Timing: W warm-up episodes, then K episodes, each bracketed by its own pair of CUDA events on the
launching stream; before every timed episode the inputs are re-staged in HBM (x0, initial model)
and the L2 is flushed with a 256 MiB write; barrier + synchronize on both sides; max over ranks.
`e2e` is the same episode driven from HOST buffers through the public API: initial states copied
from pinned host memory, trajectories (x and u of every step of every scenario) copied back,
synchronised, wall-clock.

`--impl reference` times the reference's own algorithm on the host cores: the oracle's literal
path (float64 numpy + the same scipy L-BFGS-B call on the rollout cost, duffing.py:776-778), one
process per core on disjoint scenarios; each of its K steps is a bounded sample of the episode
(--cpu-chunk closed-loop steps of one scenario per core).  The reference is Python scripts +
MATLAB and cannot be pip-installed or shipped to the GPU box, so this arm is
`cpu_baseline.kind = "port"`.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "closed_loop_koopman_mpc_scenario_steps_per_sec"
UNIT = "scenario-steps/s"
BYTES_PER_SCENARIO_STEP = 4536  # SURVEY.md 8d / BASELINE.md 4, nz = 8, fp64, update variant
GOLD = os.path.join(ROOT, "tests", "golden", "ref_vanderpol.npz")
WEIGHTS = os.path.join(ROOT, "tests", "golden", "weights", "vdp_model_weights.mat")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20, help="timed bench steps (episodes)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--episode", type=int, default=400, help="closed-loop steps per bench step (SURVEY cfg 2: T = 400)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenarios", type=int, default=4096, help="scenarios per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=200, help="closed-loop steps per core of the cpu_baseline leg")
    ap.add_argument("--cpu-chunk", type=int, default=10, help="--impl reference: closed-loop steps per bench step")
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
    chunk = args.cpu_chunk
    value, cores, sample, ms = cpu_reference(args.warmup * chunk, args.steps * chunk)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * chunk, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, cores), closed_loop_steps_per_bench_step=chunk,
                       bench_step="bounded sample of the episode: %d closed-loop steps of one scenario per core" % chunk),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, s_per_step):
    return {"workload": "vanderpol.py tracking MPC with online Koopman (RLS) update, BASELINE configs[1]",
            "scenarios_per_gpu": args.scenarios, "scenarios_per_step": s_per_step, "horizon": 10, "nz": 8,
            "ny": 8, "bounds": 6.0, "encoder": "AutoEncoder_20220414_4 (2-100-100-100-8, fp64)",
            "plant_switch_step": 102, "l2": "flushed between timed bench steps (256 MiB write)",
            "parallelism": "independent scenario shards, no collective"}


# ----------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 10 ms through NVML while the timed
    region runs (the recipe's `nvidia-smi --query-gpu=clocks.sm,...` line, in-process)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        import threading
        self.index, self.sm, self.reasons, self.max_mhz, self.err = index, [], set(), None, None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop.is_set():
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
                self._stop.wait(0.01)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def start(self):
        self._t.start()
        return self

    def stop(self):
        self._stop.set()
        self._t.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples: %s" % self.err]}
        v = sorted(self.sm)
        return {"sm_mhz": v[len(v) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(v)}


# ----------------------------------------------------------------------------- B200 arm
# algorithmic work of ONE scenario-step of this workload (DESIGN.md "Measurement"): theta_E encode
# 2(2*100 + 100*100 + 100*100 + 100*8) + 308 bias adds = 42 308, Krylov chains 19 x 128, H and f
# 110 x 16, Cholesky + two triangular solves 10^3/3 + 200, RLS (P, bar_Q, K_A P, bar_X bar_Q) 2 600,
# RK4 plant 100: 50 kflop with a direct QP solve (SURVEY.md 8d quotes 100 kflop for an iterative one)
FLOPS_PER_SCENARIO_STEP = 50.0e3


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
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the JSON line only
        dist.init_process_group("nccl", device_id=dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_reference(10, args.cpu_steps)   # before any GPU timing
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    gold = np.load(GOLD)
    enc = K.Encoder.from_file(WEIGHTS)
    S, Kst, W, T = args.scenarios, args.steps, args.warmup, args.episode
    rs = np.random.default_rng(20240601 + rank)
    x0 = rs.uniform(-2, 2, (S, 2))
    xref = np.stack([rs.uniform(-1, 1, S), np.zeros(S)], axis=1)
    r = enc(torch.from_numpy(xref).to(dev))
    loop = K.ClosedLoop(K.vanderpol_spec(), torch.from_numpy(x0).to(dev), gold["A"], gold["B"], gold["C"], r,
                        encoder=enc, log_steps=T)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one bench step = one EPISODE: the whole closed loop of the workload (T steps from x0: RLS
    # restart at step 0, plant switch at step 102) for all S scenarios, one library call
    for _ in range(W):
        loop.reset().run(T)
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(Kst)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(Kst)]
    mon = ClockSampler(local).start() if rank == 0 else None
    launches = 0
    for k in range(Kst):
        loop.reset()               # inputs of the step: staged in HBM before the timed region
        flush.zero_()              # evict them (and everything else) from L2
        n0 = K.launch_count()
        starts[k].record()
        loop.run(T)
        stops[k].record()
        launches += K.launch_count() - n0
    barrier()
    clocks = mon.stop() if rank == 0 else None
    ms_total = D.max_over_ranks(sum(s.elapsed_time(e) for s, e in zip(starts, stops)), dev)
    value = world * S * T * Kst / (ms_total * 1e-3)

    # the steady-state tail of the same loop (steps T .. 2T: no restart transient, no switch)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loop.run(T)
    e1.record()
    barrier()
    ms_steady = D.max_over_ranks(e0.elapsed_time(e1), dev)

    # phase split of one episode, measured live (clock64 per phase inside the kernel, scaled to
    # the CUDA-event time of the launch on the launching stream)
    loop.reset()
    kt = loop.run_timed(T)
    dmma_peak, dfma_peak = K.measure_fp64_peak() if rank == 0 else (None, None)

    # end to end from HOST buffers through the public API: initial states in (pinned) host memory
    # -> device, one episode, trajectories (x and u of every step and scenario) back to the host.
    # The device->host copy of an episode's logs runs on a copy stream from a device-side snapshot
    # while the next episode computes (every episode's results still reach the host inside the
    # timed region; the last copy is waited for before the clock stops).
    x_host = torch.from_numpy(x0).pin_memory()
    lx_host = torch.empty((T, S, 2), dtype=torch.float64).pin_memory()
    lu_host = torch.empty((T, S), dtype=torch.float64).pin_memory()
    snap_x, snap_u = torch.empty_like(loop.log_x), torch.empty_like(loop.log_u)
    copy_stream = torch.cuda.Stream(device=dev)
    snap_ready, snap_free = torch.cuda.Event(), torch.cuda.Event()
    n_e2e = max(3, min(Kst, 10))
    barrier()
    main_stream = torch.cuda.current_stream()
    snap_free.record(main_stream)
    t0 = time.perf_counter()
    for k in range(n_e2e):
        loop.reset(x_host)                         # H2D of this episode's inputs
        loop.run(T)
        main_stream.wait_event(snap_free)          # previous snapshot fully copied out
        snap_x.copy_(loop.log_x)
        snap_u.copy_(loop.log_u)
        snap_ready.record(main_stream)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(snap_ready)
            lx_host.copy_(snap_x, non_blocking=True)   # D2H of this episode's results
            lu_host.copy_(snap_u, non_blocking=True)
            snap_free.record(copy_stream)
    torch.cuda.synchronize()
    e2e_s = D.max_over_ranks(time.perf_counter() - t0, dev)
    e2e_ok = bool(torch.equal(torch.nan_to_num(lx_host), torch.nan_to_num(loop.log_x.cpu())))   # the host really holds the last episode
    status_bad = int((loop.status != 0).sum().item())
    finite_scen = int(torch.isfinite(loop.x).all(dim=1).sum().item())

    if rank == 0:
        ms_step = ms_total / Kst
        achieved = S * T * FLOPS_PER_SCENARIO_STEP / (ms_step * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        hbm_peak = None
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak = float(json.load(open(peaks_path))["hbm_gbs"])
        cfg = workload_config(args, S * world)
        cfg.update({"closed_loop_steps_per_bench_step": T,
                    "bench_step": "one episode = %d closed-loop steps of every scenario from x0 (RLS restart, "
                                  "plant switch at 102) in one kmpc_closed_loop_steps call" % T})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": Kst, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": {
                "bound": "tensor", "kernel": "fused_loop_kernel (persistent: QP + plant + theta_E + RLS, fp64)",
                "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
                "traffic": traffic,
                "peak_source": "fp64 tensor path (mma.sync.m8n8k4.f64) measured live by kmpc_measure_fp64_peak on "
                               "this GPU; MEASURED_PEAKS.json has no fp64 entry (its bf16 number does not bound an "
                               "fp64 kernel)",
                "dfma_peak_tflops": dfma_peak,
                "flops_per_scenario_step": FLOPS_PER_SCENARIO_STEP,
                "launch_ms": ms_step, "launches_per_step": launches / Kst,
                "phase_ms_per_launch": kt,
                "hbm_view": {"algorithmic_bytes_per_scenario_step": BYTES_PER_SCENARIO_STEP,
                             "equivalent_gbs_if_state_round_tripped_hbm": value / world * BYTES_PER_SCENARIO_STEP / 1e9,
                             "hbm_peak_gbs": hbm_peak,
                             "note": "the per-scenario state stays in registers for the whole launch, so HBM "
                                     "carries only x0/model in, logs and final state out"},
            },
            "value_steady_state": world * S * T / (ms_steady * 1e-3),
            "us_per_closed_loop_step": ms_step * 1e3 / T,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e": {"value": world * S * T * n_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": S * 2 * 8,
                    "d2h_bytes_per_step": T * S * 3 * 8, "episodes": n_e2e, "results_on_host_verified": e2e_ok,
                    "note": "D2H of episode k overlaps the compute of episode k+1 (copy stream)"},
            "gpu_launches": launches,
            "health": {"scenarios_with_status": status_bad, "finite_scenarios": finite_scen, "scenarios": S,
                       "note": "the reference's RK4 plant itself diverges for |x1| > 2.4 (h*lambda < -2.78); the "
                               "oracle blows up on the same scenarios at the same step (tests)"},
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
