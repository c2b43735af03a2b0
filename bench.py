#!/usr/bin/env python
"""bench.py — CMA-ES generations/s (and samples/s) on the configuration BASELINE.json quotes its metric on:
N = 1000, lambda = 65536, mu = lambda/2, ill-conditioned ellipsoid (SURVEY.md 8d "config 3"), population sharded
over --gpus N B200s of one box (strong scaling: the population is fixed, each rank owns lambda/N samples).

A "step" is one full generation of the hot path: eigendecomposition, Philox sampling, sampling GEMM, batched
objective, ranking, mean/path updates, rank-mu covariance update (+ all-gather(F) and all-reduce(C partial) for N > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
Under torchrun (N > 1) every rank runs this file; rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337)
WORKLOAD_NAME = "config3: CMA-ES 1000-D ill-conditioned ellipsoid, lambda=65536, mu=32768 (Logarithmic weights), seed 1337"
METRIC, UNIT = "cmaes_generations_per_sec_N1000_lambda65536", "generations/s"


def flops_per_generation(n, lam, mu):
    """Algorithmic flops: sampling 2 N^2 lambda, rank-mu N (N+1) mu."""
    return 2.0 * n * n * lam, float(n) * (n + 1) * mu


def jacobi_blocks(n, num_sms=148):
    """(real 4-row blocks, blocks the tournament runs over, ring order?) — the dispatch rule of launch_jacobi_persistent (eigen.cu)."""
    nb = ((n + 3) // 4 + 1) & ~1
    nb_ring = 2
    while nb_ring < nb:
        nb_ring *= 2
    ring = os.environ.get("KCMA_JACOBI_ORDER") != "rr" and nb_ring // 2 <= num_sms and (nb_ring - nb) * 8 <= nb
    return nb, (nb_ring if ring else nb), ring


def load_peaks():
    out = {}
    for name in ("MEASURED_PEAKS.json", "MEASURED_FP64.json"):
        p = os.path.join(ROOT, name)
        if os.path.exists(p):
            out.update(json.load(open(p)))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, "/tmp/kcma_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate(); self.proc.wait(); self.fh.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_lambda=1024):
    """The reference algorithm on the host: the C restatement in oracle/ (the reference itself needs GSL/Eigen/meson and
    cannot be built here — DESIGN.md). Single thread: CMAES.cpp.base has no OpenMP and its conduits only parallelise
    the user model. One timed step = one generation on a BOUNDED SAMPLE of the population
    (sample_lambda of the 65536 samples, all N = 1000 dimensions); the N x N eigendecomposition, whose cost does
    not depend on lambda, is timed once. generations/s at the full population is extrapolated linearly in lambda."""
    from oracle import oracle as O
    from korali_b200._abi import INJ_BD
    n, lam = WORKLOAD["n"], WORKLOAD["population_size"]
    kw = dict(WORKLOAD); kw["population_size"] = sample_lambda
    o = O.Oracle(**kw)
    o.set_scalar("Oracle/RNG Kind", 1)
    ident = np.concatenate([np.eye(n).ravel(), np.ones(n)])
    times = []
    for it in range(warmup + steps):
        o.inject(INJ_BD, ident)              # eigen timed separately below
        t0 = time.perf_counter()
        o.run_generation()
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
    c = o.get("Covariance Matrix").reshape(n, n)
    t0 = time.perf_counter(); O.eigen(c); t_eig = time.perf_counter() - t0
    t_pop = float(np.mean(times))
    t_full = t_eig + t_pop * (lam / sample_lambda)
    return {"gens_per_sec": 1.0 / t_full, "t_eigen_s": t_eig, "t_population_sample_s": t_pop, "sample_lambda": sample_lambda,
            "ms_per_step": 1e3 * t_full}


def reference_arm(args, rank):
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    lam = WORKLOAD["population_size"]
    sample = ("oracle/okcma.c (C restatement of CMAES.cpp.base, gcc -O2, 1 thread); per step %d of %d samples x all 1000 dims "
              "(%.3f s) extrapolated linearly in lambda + one full 1000x1000 eigendecomposition (%.2f s)"
              % (r["sample_lambda"], lam, r["t_population_sample_s"], r["t_eigen_s"]))
    line = {"impl": "reference", "metric": METRIC, "value": r["gens_per_sec"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "samples_per_sec": r["gens_per_sec"] * lam,
            "config": {"workload": WORKLOAD_NAME},
            "cpu_baseline": {"value": r["gens_per_sec"], "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": r["gens_per_sec"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def hbm_block(phases, steps, n, lam_local, mu_local, hbm_peak):
    """Achieved GB/s of the bandwidth-side kernels (K1 rng, K3 objective, K4 sort, K5 gather-mean)."""
    work = {"rng": ("philox_normal_kernel: Z written (FP64 Box-Muller bound)", 8.0 * n * lam_local),
            "objective": ("objective_kernel: Y read", 8.0 * n * lam_local),
            "sort": ("radix sort, 8 passes x 12 B x lambda (launch-latency bound: 25 launches)", 8 * 12.0 * lam_local),
            "gather_mean": ("gather_mean + mean_reduce: selected rows read + S written", 2 * 8.0 * n * mu_local)}
    out = {}
    for k, (what, nbytes) in work.items():
        ms = phases[k][0] / steps
        gbps = nbytes / (ms * 1e-3) * 1e-9 if ms > 0 else 0.0
        out[k] = {"what": what, "bytes": nbytes, "ms": ms, "GBps": gbps, "frac_of_hbm_peak": (gbps / hbm_peak) if hbm_peak else None}
    return out


# ---------------------------------------------------------------------------------------------------------
def ours_arm(args, rank, world):
    import torch
    import torch.distributed as dist
    from korali_b200 import _lib
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, lam = WORKLOAD["n"], WORKLOAD["population_size"]
    mu = lam // 2
    s = _lib.Solver(device=local_rank, rank=rank, nranks=world, **WORKLOAD)
    s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(_lib.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        s.comm_init(bytes(uid.cpu().tolist()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        s.run_generation()
    # ---- device-timed region: K generations, state resident in HBM --------------------------------------
    s.timing_enable(True); s.timing_reset()
    clocks = ClockSampler(local_rank)
    barrier()
    l0 = s.launch_count()
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        s.run_generation()
    e1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = s.launch_count() - l0
    phases = {p: s.timing(p) for p in ["eigen", "rng", "sample_gemm", "objective", "sort", "gather_mean", "rank_mu", "paths", "collectives", "generation"]}
    sweeps = s.timing("eigen_sweeps")[1] / args.steps   # (counted on the device: read before the e2e leg adds its own)
    s.timing_enable(False)
    # ---- end-to-end through the reference-facing call: kcma_run (Experiment::run loop incl. termination chain) ----
    barrier()
    t0 = time.perf_counter()
    done = s.run(args.steps)
    best = s.scalar("Best Ever Value")       # device -> host read of the step's result
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([ms, t_e2e * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, t_e2e = float(t[0]), float(t[1]) * 1e-3
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    ms_step = ms / args.steps
    gens = 1e3 / ms_step
    f_sample, f_rank = flops_per_generation(n, lam // world, mu // world)
    gemm_ms, gemm_calls = phases["sample_gemm"]
    gemm_avg = gemm_ms / max(gemm_calls, 1)
    achieved = f_sample / (gemm_avg * 1e-3) * 1e-12 if gemm_avg > 0 else 0.0
    peak = peaks.get("fp64_dgemm_tflops_sustained")
    rk_ms, rk_calls = phases["rank_mu"]
    rk_avg = rk_ms / max(rk_calls, 1)
    cpu = cpu_reference_run(8, 1) if world == 1 else None   # ~10 s of host work: 9 x 1024-sample generations + one eigendecomposition
    line = {
        "metric": METRIC, "value": gens, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "samples_per_sec": gens * lam,
        "config": {"workload": WORKLOAD_NAME, "n": n, "lambda": lam, "mu": mu, "parallelism": "population sharded x%d" % world,
                   "l2_hygiene": "inputs larger than L2: Z and Y are %.0f MB each per rank, re-streamed every generation" % (8.0 * n * lam / world / 1e6),
                   "best_ever_value_after_run": best},
        "e2e": {"value": done / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 288,   # sizeof(DevScalars): the termination chain reads it once per generation
                "note": "kcma_run(): Experiment::run loop with the termination chain evaluated on the host every generation (device scalars "
                        "copied back each step); the generation loop takes no per-step host input (samples are drawn on the device from "
                        "Philox(seed, generation) counters). Runs without phase timers (one CUDA-graph replay per generation), "
                        "whereas `value` is timed with eager launches and the per-phase CUDA events enabled. The e2e leg continues from the "
                        "state the timed leg left (generations K+W+1 .. 2K+W): the spectrum has spread a little more by then and the "
                        "warm-started eigensolver needs about one sweep less per generation"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"kernel": "gemm_tn_tma_kernel (sampling GEMM Y = Z (B D)^T, TMA + mbarrier + DMMA.8x8x4)", "bound": "tensor",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if peak else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at N=1 (profiles/r01_ncu_full_summary.md):
                     # 537.5 MB + 492.0 MB = the algorithmic Z read + Y write, i.e. no re-reads
                     "traffic": 1029548288 if world == 1 else None, "traffic_unit": "bytes per launch",
                     "peak_source": "MEASURED_FP64.json: cuBLAS DGEMM 8192^3 sustained on this pool's B200 (FP64 DMMA issue peak 37.1)",
                     "algorithmic_flops_per_launch": f_sample, "avg_launch_ms": gemm_avg},
        "phases_ms_per_generation": {k: (v[0] / args.steps) for k, v in phases.items()},
        "rank_mu": {"kernel": "syrk_tt_kernel", "avg_ms": rk_avg, "achieved_tflops": (f_rank / (rk_avg * 1e-3) * 1e-12) if rk_avg > 0 else 0.0,
                    "algorithmic_flops_per_launch": f_rank},
        # the HBM-side kernels of a generation: algorithmic bytes / phase time against the measured copy bandwidth
        "hbm_kernels": hbm_block(phases, args.steps, n, lam // world, mu // world, peaks.get("hbm_gbs", 6547.2)),
        "eigen": {"kernel": "jacobi_pipe_kernel (persistent cooperative one-sided Jacobi, Gram-update steps on DMMA.8x8x4; replicated on every rank)",
                  "avg_ms": phases["eigen"][0] / args.steps, "sweeps_per_generation": sweeps,
                  "bound": "latency: one dependent tournament step per pair of 4-row blocks and sweep; a step = flag handshake + 64 KB row fetch + Gram + 4 rounds + apply",
                  # executed flops: per step and block pair 3 x (2 * 8 * 8 * N) for Gram, apply G, apply V; (N/4 - 1) * N/8 pair-steps per sweep
                  "executed_flops_per_sweep": 12.0 * n ** 3,
                  "achieved_tflops": (12.0 * n ** 3 * sweeps / (phases["eigen"][0] / args.steps * 1e-3) * 1e-12) if phases["eigen"][0] > 0 else 0.0,
                  "tournament": "ring order, %d steps per sweep (%d real + %d phantom 4-row blocks)" % (jacobi_blocks(n)[1] - 1, jacobi_blocks(n)[0], jacobi_blocks(n)[1] - jacobi_blocks(n)[0])
                                if jacobi_blocks(n)[2] else "round-robin order, %d steps per sweep" % (jacobi_blocks(n)[0] - 1),
                  "us_per_step_of_8_rows": (phases["eigen"][0] / args.steps * 1e3 / max(sweeps * (jacobi_blocks(n)[1] - 1.0), 1e-9))},
        "eigen_sweeps_per_generation": sweeps,
        "gens_per_sec_excluding_eigen": 1e3 / max(ms_step - phases["eigen"][0] / args.steps, 1e-9),
    }
    if cpu:
        line["cpu_baseline"] = {"value": cpu["gens_per_sec"], "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "oracle/okcma.c, 1 thread: %d of %d samples per step (%.3f s) extrapolated linearly in lambda + one "
                                          "1000x1000 eigendecomposition (%.2f s)" % (cpu["sample_lambda"], lam, cpu["t_population_sample_s"], cpu["t_eigen_s"])}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    ours_arm(args, rank, world)


if __name__ == "__main__":
    main()
