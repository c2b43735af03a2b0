"""ms per generation of the Differential Evolution path (include/kdea.h) on one GPU, device objective, and the HBM traffic it implies:
per generation the mutation reads 4 rows (sample, a, b, parent) and writes the candidate, the objective reads the candidate, the
accept step reads candidate + writes sample, the mean reads the population: ~9 x 8 x N x lambda bytes.

    python profiles/microbench/dea_bench.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from korali_b200 import _dea

for n, lam, cr in [(10, 200, 0.9), (100, 4096, 0.5), (100, 65536, 0.5), (1000, 65536, 0.01)]:
    s = _dea.Solver(n=n, population_size=lam, objective="NegSphere", lower_bound=-5.0, upper_bound=5.0, seed=3, crossover_rate=cr)
    for _ in range(5):
        s.run_generation()
    torch.cuda.synchronize()
    l0 = s.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gens = 50
    e0.record()
    for _ in range(gens):
        s.run_generation()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / gens
    print("N=%5d lambda=%6d crossover %.2f: %8.3f ms/generation  %6.1f launches/generation  ~%7.1f GB/s  best %.3e  infeasible draws %d" % (
        n, lam, cr, ms, (s.launch_count() - l0) / gens, 9 * 8.0 * n * lam / (ms * 1e-3) * 1e-9, s.scalar("Best Ever Value"), s.scalar("Infeasible Sample Count")), flush=True)
    s.close()
