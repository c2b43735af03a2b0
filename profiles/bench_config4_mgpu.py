"""config 4 (BASELINE.json): CMA-ES on the 4096-D sphere, lambda = 2^20 with Mirrored Sampling, mu = 2^19, population sharded over the
GPUs of one box. Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/bench_config4_mgpu.py
Not the contract bench (that is /bench.py on config 3): one JSON line with ms/generation and the phase split of rank 0."""
import json
import os
import sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from korali_b200 import _lib  # noqa: E402

rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
n, lam = 4096, 1 << 20
s = _lib.Solver(device=local_rank, rank=rank, nranks=world, n=n, population_size=lam, mu_value=lam // 2, objective="NegSphere", mirrored_sampling=1,
                initial_value=1.0, initial_stddev=1.0, seed=1337)
s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
if world > 1:
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(_lib.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    s.comm_init(bytes(uid.cpu().tolist()))
gens = int(sys.argv[1]) if len(sys.argv) > 1 else 3
s.run_generation()
s.timing_enable(True); s.timing_reset()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(gens):
    s.run_generation()
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ph = {p: round(s.timing(p)[0] / gens, 3) for p in ["eigen", "eigen_sytrd", "eigen_dc", "eigen_back", "rng", "sample_gemm", "objective", "sort", "gather_mean",
                                                        "rank_mu", "paths", "collectives"]}
    print(json.dumps({"config": "config4: N=4096, lambda=2^20 mirrored, mu=2^19, sphere", "n_gpus": world, "generations": gens,
                      "ms_per_generation": float(ms[0]) / gens, "samples_per_sec": lam / (float(ms[0]) / gens * 1e-3),
                      "phases_ms_rank0": ph, "best_ever_value": s.scalar("Best Ever Value"),
                      "sample_gemm_tflops_per_rank": 2.0 * n * n * (lam / 2 / world) / (ph["sample_gemm"] * 1e-3) * 1e-12 if ph["sample_gemm"] else None}), flush=True)
# optional: run on to the optimum (BASELINE.json: convergence within 1e-8 on every config) — `... bench_config4_mgpu.py <gens> converge`
if len(sys.argv) > 2 and sys.argv[2] == "converge":
    import time
    s.timing_enable(False)
    s.set_scalar("Termination Criteria/Max Value", -1e-9)
    s.set_scalar("Termination Criteria/Max Generations", 600)
    t0 = time.perf_counter()
    done = s.run(600)
    best = s.scalar("Best Ever Value")
    dt = time.perf_counter() - t0
    fin, why = s.check_termination()
    if rank == 0:
        print(json.dumps({"config4_convergence": {"generations_total": int(s.scalar("Current Generation")), "best_ever_value": best, "finished": fin,
                                                   "criterion": why, "seconds_for_the_last_%d_generations" % done: dt,
                                                   "max_abs_best_ever_variable": float(abs(s.get("Best Ever Variables")).max())}}), flush=True)
s.close()
if world > 1:
    dist.destroy_process_group()
