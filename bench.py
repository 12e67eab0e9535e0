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

Besides the headline workload the same invocation measures, bounded to a few seconds each, the other
BASELINE configs (`secondary` list on the JSON line, each with its own roofline and a one-scenario
oracle spot check on rank 0): the duffing.py-shape loop, configs[2] Tank (65 536 scenarios in total,
sharded over the ranks), configs[3] EDMD over 10 M snapshots sharded over the ranks with the NCCL
Gram all-reduce INSIDE the timed region (fp64 lift and tcgen05 lift), configs[4] RBF horizon 50.
`per_rank_ms` lists every rank's own device time of the headline workload.

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
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs")
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


# ----------------------------------------------------------------------------- oracle spot checks
def _spot_worker(job):
    """Checker leg (rank 0 only): the oracle follows ONE scenario of a secondary workload from the
    model the GPU run used; returns its trajectory for comparison with the GPU's."""
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[k] = "1"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from oracle import closed_loop as ocl
    from oracle import lift as olift
    from oracle import rls as orls
    from oracle import weights as ow
    kind, A, B, C, x0, T, extra = job
    wdir = os.path.join(ROOT, "tests", "golden", "weights")
    with np.errstate(all="ignore"):
        if kind == "duffing":
            Ws, bs = ow.load_mat_encoder(os.path.join(wdir, "duffing_model_weights.mat"))
            o = ocl.run_loop(ocl.duffing_config(Ws, bs), A, B, C, x0, T, update=ocl.UPDATE_RLS, qp="exact")
        elif kind == "tank":
            Ws, bs = ow.load_mat_encoder(os.path.join(wdir, "tank_model_weights.mat"))
            cfg = ocl.tank_config(lambda v: olift.encoder_forward(Ws, bs, v), 10)
            o = ocl.run_loop(cfg, A, B, C, x0, T, update=ocl.UPDATE_RLS, qp="exact")
        else:  # rbf, horizon 50, warm-started update
            cfg = ocl.rbf_config(extra["cx"])
            cfg.N = 50
            warm = orls.RLSState.warm(extra["G"], extra["Aq"], extra["XV"][:, :8], extra["G"][:8, :8])
            o = ocl.run_loop(cfg, A, B, C, x0, T, update=ocl.UPDATE_RLS, qp="exact", warm=warm)
    return o["X"], o["U"]


def _spot_compare(jobs, gpu_logs):
    """jobs: {name: job}; gpu_logs: {name: (X (T,2), U (T,))} -> {name: max abs deviation in x}."""
    import multiprocessing as mp
    import numpy as np
    names = list(jobs)
    with mp.get_context("spawn").Pool(len(names)) as pool:
        res = pool.map(_spot_worker, [jobs[k] for k in names])
    out = {}
    for k, (X, U) in zip(names, res):
        gx, gu = gpu_logs[k]
        out[k] = {"scenario": 0, "steps": int(len(U)), "max_abs_dx_vs_oracle": float(np.abs(X - gx[:len(U)]).max()),
                  "max_abs_du_vs_oracle": float(np.abs(U - gu[:len(U)]).max())}
    return out


# ----------------------------------------------------------------------------- secondary workloads
def secondary_workloads(K, D, dev, rank, world, dmma_peak, hbm_peak, bf16_peak):
    """The other BASELINE configs, each bounded to a few seconds.  Returns (list for the JSON line,
    spot-check jobs, GPU logs of scenario 0)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from koopman_online_updated_mpc_b200 import edmd as E
    from koopman_online_updated_mpc_b200 import scripts as SC
    wdir = os.path.join(ROOT, "tests", "golden", "weights")
    out, jobs, logs = [], {}, {}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_loop(loop, T, reps, warm_state=None):
        """reps timed episodes of T steps from x0 (after one warm-up episode); max over ranks of the sum."""
        loop.reset().run(T)
        barrier()
        ms = 0.0
        for _ in range(reps):
            loop.reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loop.run(T)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        barrier()
        return D.max_over_ranks(ms, dev) / reps

    def loop_entry(name, workload, S_rank, S_total, T, ms, flops, bytes_per, fused, status_bad, scaling):
        value = S_total * T / (ms * 1e-3)
        per_gpu = value / world
        ach = per_gpu * flops / 1e12
        return {"name": name, "workload": workload, "value": value, "unit": UNIT, "ms_per_episode": ms,
                "scenarios_per_gpu": S_rank, "scenarios_total": S_total, "closed_loop_steps": T, "scaling": scaling,
                "kernel_path": "fused persistent kernel" if fused else "generic qp_plant -> lift -> rls kernels, 3 launches per step",
                "scenarios_with_status": status_bad,
                "roofline": {"bound": "tensor", "achieved": ach, "peak": dmma_peak, "unit": "TFLOP/s",
                             "frac": (ach / dmma_peak) if dmma_peak else None, "flops_per_scenario_step": flops,
                             "peak_source": "fp64 tensor path measured live (kmpc_measure_fp64_peak)",
                             "hbm_view": {"algorithmic_bytes_per_scenario_step": bytes_per,
                                          "equivalent_gbs": per_gpu * bytes_per / 1e9, "hbm_peak_gbs": hbm_peak,
                                          "frac": (per_gpu * bytes_per / 1e9 / hbm_peak) if hbm_peak else None}}}

    # ---- (a) duffing.py shape: C-output cost, bounds +-2, model from the package's own EDMD ----
    enc_d = K.Encoder.from_file(os.path.join(wdir, "duffing_model_weights.mat"))
    np.random.seed(101)
    Xs, Ys, Us = K.data_generate.generate(100, 100).duffing_generate()
    A, B, C, _ = SC.identify(enc_d, Xs, Ys, Us, n_step=100)
    S, T = 4096, 400
    rs = np.random.default_rng(20240701 + rank)
    x0 = rs.uniform(-2, 2, (S, 2))
    loop = K.ClosedLoop(K.duffing_spec(), x0, A, B, C, np.array([1.0, 0.0]), encoder=enc_d, log_steps=T)
    ms = time_loop(loop, T, 3)
    out.append(loop_entry("duffing_closed_loop", "duffing.py closed loop with online update (l.823-1012): theta_E lift, "
                          "y = C z cost, bounds +-2, S = 4096 per GPU, T = 400, model from the package's EDMD of the "
                          "seed-101 snapshot set", S, S * world, T, ms, 50.0e3, 4536, loop.fused,
                          int((loop.status != 0).sum().item()), "weak"))
    if rank == 0:
        jobs["duffing_closed_loop"] = ("duffing", A.cpu().numpy(), B.cpu().numpy(), C.cpu().numpy(), x0[0], 60, None)
        logs["duffing_closed_loop"] = (loop.log_x[:, 0].cpu().numpy(), loop.log_u[:, 0].cpu().numpy())
    loop.close()
    del loop

    # ---- (b) configs[2]: Tank, 65 536 scenarios in total sharded over the ranks ----
    enc_t = K.Encoder.from_file(os.path.join(wdir, "tank_model_weights.mat"))
    At, Bt, Ct, _ = SC.tank_identify(enc_t)
    S_total, T = 65536, 300
    lo, hi = D.shard_bounds(S_total, rank, world)
    x0 = np.maximum(np.random.default_rng(20240801).uniform(0, 2, (S_total, 2)), 0.0)[lo:hi]
    loop = K.ClosedLoop(K.tank_spec(), x0, At, Bt, Ct, np.array([1.0]), encoder=enc_t, log_steps=T)
    ms = time_loop(loop, T, 1)
    # algorithmic flops per scenario-step: theta_E 2-100-100-10 22.6 k, Krylov chains 20 x 2 x 11^2
    # 4.8 k, H and f 0.9 k, ONE 20 x 20 Cholesky + two triangular solves 3.5 k, RLS (11 x 11 P, 10 x 10 bar_Q, K_A P,
    # bar_X bar_Q) 4.2 k, plant 0.1 k: 36 k
    out.append(loop_entry("tank_closed_loop", "BASELINE configs[2]: Tank_System.m l.170-291 with the Encoder_Tank lift "
                          "(nz = 10 + du augmentation, N = 20), 65 536 scenarios sharded over the ranks, T = 300, model "
                          "from the package's joint Gram regression (l.93-100)", hi - lo, S_total, T, ms, 36.0e3, 6776,
                          loop.fused, int((loop.status != 0).sum().item()), "strong"))
    if rank == 0:
        jobs["tank_closed_loop"] = ("tank", At.cpu().numpy(), Bt.cpu().numpy(), Ct.cpu().numpy(), x0[0], 100, None)
        logs["tank_closed_loop"] = (loop.log_x[:, 0].cpu().numpy(), loop.log_u[:, 0].cpu().numpy())
    loop.close()
    del loop

    # ---- (c) configs[3]: EDMD over 10 M duffing snapshots sharded by trajectory, Gram all-reduce ----
    n_traj, n_step = 100000, 100
    lo, hi = D.shard_bounds(n_traj, rank, world)
    rs = np.random.default_rng(20240601 + rank)
    u0 = torch.from_numpy(rs.uniform(-2, 2, (n_step, hi - lo))).to(dev)
    xi = torch.from_numpy(rs.uniform(-2, 2, (hi - lo, 2))).to(dev)
    M = n_traj * n_step
    for prec_name, prec in (("fp64", K.lift.PREC_FP64), ("tcgen05_bf16x3", K.lift.PREC_TC)):
        if prec == K.lift.PREC_TC and not enc_d.has_tc:
            continue
        times, t_ar = [], []
        for rep in range(7):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            barrier()
            ev[0].record()
            X, Y, U = K.data_generate.generate_snapshots(xi, u0, K.plant.DUFFING_PRE)
            pack = E.gram_from_trajectories(enc_d, X, Y, U, n_step, precision=prec)
            ev[1].record()
            D.allreduce_pack(pack)                       # NCCL all-reduce (sum, fp64) of the Gram pack
            ev[2].record()
            Ae, Be, Ce, st = E.edmd_solve(pack, 8)
            ev[3].record()
            torch.cuda.synchronize()
            if rep >= 2:
                times.append(D.max_over_ranks(ev[0].elapsed_time(ev[3]), dev))
                t_ar.append(ev[1].elapsed_time(ev[2]))
        same = True
        if world > 1:
            gathered = [torch.empty_like(Ae) for _ in range(world)]
            dist.all_gather(gathered, Ae)
            same = all(torch.equal(t, Ae) for t in gathered)
        ms = min(times)
        rows_rank = (hi - lo) * (n_step + 1)
        tfl = rows_rank * 42308.0 / (ms * 1e-3) / 1e12
        if prec == K.lift.PREC_TC:
            # tensor work actually issued: 6 bf16 piece products over the padded 112-wide layers
            peak, peak_src = bf16_peak, "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)"
            tfl_issued = rows_rank * 6 * 2.0 * (2 * 112 * 112 + 112 * 16) / (ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": tfl_issued, "peak": peak, "unit": "TFLOP/s",
                    "frac": (tfl_issued / peak) if peak else None, "peak_source": peak_src,
                    "algorithmic_fp64_equivalent_tflops": tfl,
                    "note": "achieved counts the bf16 piece products issued on tcgen05 (6 per fp32-accurate product, "
                            "100 -> 112 padding); whole regression timed (generation, lift, fp64 Gram, all-reduce, solve)"}
        else:
            roof = {"bound": "tensor", "achieved": tfl, "peak": dmma_peak, "unit": "TFLOP/s",
                    "frac": (tfl / dmma_peak) if dmma_peak else None,
                    "peak_source": "fp64 tensor path measured live (kmpc_measure_fp64_peak)"}
        out.append({"name": "edmd_10M_" + prec_name, "workload": "BASELINE configs[3]: EDMD regression over 10 M synthetic "
                    "duffing snapshots (100 000 trajectories x 100 steps, data_generate.py:17-57 on the GPU), sharded by "
                    "trajectory over the ranks, ONE Gram all-reduce (NCCL, sum, fp64) inside the timed region, replicated solve",
                    "value": M / (ms * 1e-3), "unit": "snapshots/s", "ms": ms, "lift_precision": prec_name,
                    "allreduce_ms_rank0": min(t_ar), "collective": "nccl all_reduce" if world > 1 else "none (1 rank)",
                    "A_bitwise_identical_on_all_ranks": bool(same), "status": int(st.item()), "scaling": "strong",
                    "A00": float(Ae[0, 0].item()), "roofline": roof})

    # ---- (d) configs[4]: RBF lift, horizon 50, warm-started ("storage method") update ----
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
    S, T = 125000, 100
    x0 = np.random.default_rng(20240901 + rank).uniform(-2, 2, (S, 2))
    warm = K.RLSState.warm(S, G, Aq, XV[:, :8], G[:8, :8])
    loop = K.ClosedLoop(K.rbf_spec(N=50), x0, Ar, Br, Cr, np.array([1.0, 0.0]), cx=cx_d, rls_state=warm, log_steps=T)
    ms = time_loop(loop, T, 1)
    out.append(loop_entry("rbf_horizon50_closed_loop", "BASELINE configs[4]: duffing_RBF.py (thin-plate RBF lift, 8 centres "
                          "of the reference run), horizon 50, warm-started update (l.434-438), 125 000 scenarios per GPU "
                          "(1 M over 8), T = 100; flops: Krylov 50 x 2 x 64 + 50 x 36, H (Toeplitz sums, ny = 2) 5.1 k, ONE "
                          "50 x 50 Cholesky + solves 47 k, 8 RBFs 0.3 k, RLS 2.6 k = 64 k with a direct solve (SURVEY.md 8d "
                          "quotes 2.8 M for an iterative one)", S, S * world, T, ms, 64.0e3, 4536, loop.fused,
                          int((loop.status != 0).sum().item()), "weak"))
    if rank == 0:
        jobs["rbf_horizon50_closed_loop"] = ("rbf", Ar.cpu().numpy(), Br.cpu().numpy(), Cr.cpu().numpy(), x0[0], 30,
                                             {"cx": cx, "G": G, "Aq": Aq, "XV": XV})
        logs["rbf_horizon50_closed_loop"] = (loop.log_x[:, 0].cpu().numpy(), loop.log_u[:, 0].cpu().numpy())
    loop.close()
    return out, jobs, logs


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
        # ... including NCCL's own "NCCL version" banner, which it writes to fd 1 when the first communicator
        # comes up: point fd 1 at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

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
    ms_mine = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    ms_total = D.max_over_ranks(ms_mine, dev)
    value = world * S * T * Kst / (ms_total * 1e-3)
    per_rank = torch.tensor([ms_mine / Kst], dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.empty_like(per_rank) for _ in range(world)]
        dist.all_gather(allr, per_rank)
        per_rank_ms = [float(t.item()) for t in allr]
    else:
        per_rank_ms = [float(per_rank.item())]

    # N > 1: the same episodes once more with EVERY rank on rank 0's draw.  The ranks' own draws differ
    # (seed + rank) and a launch lasts as long as its slowest quarter, so the max over ranks mixes "what
    # does a second process cost" with "how heavy is the worst draw"; this leg separates the two.
    same_draw = None
    if world > 1:
        rs0 = np.random.default_rng(20240601)
        x0_0 = rs0.uniform(-2, 2, (S, 2))
        xref_0 = np.stack([rs0.uniform(-1, 1, S), np.zeros(S)], axis=1)
        loop0 = K.ClosedLoop(K.vanderpol_spec(), torch.from_numpy(x0_0).to(dev), gold["A"], gold["B"], gold["C"],
                             enc(torch.from_numpy(xref_0).to(dev)), encoder=enc, log_steps=T)
        loop0.reset().run(T)
        barrier()
        s0 = [torch.cuda.Event(enable_timing=True) for _ in range(Kst)]
        s1 = [torch.cuda.Event(enable_timing=True) for _ in range(Kst)]
        for k in range(Kst):
            loop0.reset()
            flush.zero_()
            s0[k].record()
            loop0.run(T)
            s1[k].record()
        barrier()
        ms0 = sum(a.elapsed_time(b) for a, b in zip(s0, s1))
        mine0 = torch.tensor([ms0 / Kst], dtype=torch.float64, device=dev)
        all0 = [torch.empty_like(mine0) for _ in range(world)]
        dist.all_gather(all0, mine0)
        ms0_max = D.max_over_ranks(ms0, dev)
        same_draw = {"value": world * S * T * Kst / (ms0_max * 1e-3), "unit": UNIT,
                     "per_rank_ms": [float(t.item()) for t in all0],
                     "note": "every rank runs rank 0's draw (seed 20240601): isolates the cost of N processes from "
                             "the spread between the ranks' own draws"}
        loop0.close()
        del loop0

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
    dmma_peak, dfma_peak = K.measure_fp64_peak()

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
    loop.close()
    del loop, snap_x, snap_u, flush

    hbm_peak = bf16_peak = None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        hbm_peak, bf16_peak = float(pk["hbm_gbs"]), float(pk["bf16_tflops"])
    secondary, spot = [], None
    if not args.no_secondary:
        secondary, jobs, logs = secondary_workloads(K, D, dev, rank, world, dmma_peak, hbm_peak, bf16_peak)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            spot = _spot_compare(jobs, logs)      # the oracle as CHECKER (never timed, never on the product path)
            for e in secondary:
                if e["name"] in spot:
                    e["oracle_spot_check"] = spot[e["name"]]

    if rank == 0:
        ms_step = ms_total / Kst
        achieved = S * T * FLOPS_PER_SCENARIO_STEP / (ms_step * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
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
            "per_rank_ms": per_rank_ms,
            "same_draw": same_draw,
            "value_steady_state": world * S * T / (ms_steady * 1e-3),
            "us_per_closed_loop_step": ms_step * 1e3 / T,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "e2e": {"value": world * S * T * n_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": S * 2 * 8,
                    "d2h_bytes_per_step": T * S * 3 * 8, "episodes": n_e2e, "results_on_host_verified": e2e_ok,
                    "note": "per episode: x0 of every scenario host -> device, trajectories device -> host; the initial "
                            "model A, B, C, the set-points and the encoder weights are staged on the device once, "
                            "outside the timed loop (they do not change between episodes); D2H of episode k overlaps "
                            "the compute of episode k+1 (copy stream)"},
            "gpu_launches": launches,
            "secondary": secondary,
            "build": "release: the phase-skip / path-override profiling knobs are compiled out (no KMPC_PROFILING)",
            "health": {"scenarios_with_status": status_bad, "finite_scenarios": finite_scen, "scenarios": S,
                       "value_counting_finite_scenarios_only": value * finite_scen / S,
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
