"""Times one generation of every BASELINE.json config that fits one GPU (ms/generation, phase split).
Not the contract bench (that is /bench.py on config 3); used for the tables in DESIGN.md."""
import json
import os
import sys
import time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from korali_b200 import _lib  # noqa: E402


def constrained(n):
    iv = np.zeros(n); iv[0] = iv[1] = 4.0; iv[2] = iv[3] = -2.0
    return dict(objective="NegSphereSin2", constraint_family="HalfSpace", n_constraints=4, constraint_shift=np.array([1.0, 1.0, -1.0, -1.0]),
                lower_bound=-10.0, upper_bound=10.0, initial_value=iv, initial_stddev=1.0, is_sigma_bounded=1)


CONFIGS = {
    "config1 N=10 lambda=32 Rosenbrock": (dict(n=10, population_size=32, objective="NegRosenbrock", initial_value=0.0, initial_stddev=0.5), 200),
    "config2 N=100 lambda=4096 Ackley": (dict(n=100, population_size=4096, objective="NegAckley", initial_value=1.0, initial_stddev=3.0), 200),
    "config3 N=1000 lambda=65536 ellipsoid": (dict(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0), 10),
    "config4/8 N=4096 lambda_local=2^17 mirrored sphere (one rank's shard of 2^20)": (
        dict(n=4096, population_size=1 << 17, mu_value=1 << 16, objective="NegSphere", mirrored_sampling=1, initial_value=1.0, initial_stddev=1.0), 2),
    "config5 N=100 lambda=8192 4 constraints": (dict(n=100, population_size=8192, viability_population_size=8192, **constrained(100)), 100),
}

if __name__ == "__main__":
    only = sys.argv[1:] or None
    out = {}
    for name, (kw, gens) in CONFIGS.items():
        if only and not any(o in name for o in only):
            continue
        s = _lib.Solver(seed=1337, **kw)
        s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
        for _ in range(3):
            s.run_generation()
        s.timing_enable(True); s.timing_reset()
        torch.cuda.synchronize()
        l0 = s.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(gens):
            s.run_generation()
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        ph = {p: round(s.timing(p)[0] / gens, 4) for p in ["eigen", "rng", "sample_gemm", "objective", "sort", "gather_mean", "rank_mu", "paths", "constraints", "feasibility"]}
        out[name] = {"ms_per_generation_device": e0.elapsed_time(e1) / gens, "ms_per_generation_wall": 1e3 * (t1 - t0) / gens,
                     "launches_per_generation": (s.launch_count() - l0) / gens, "phases_ms": ph, "best": s.scalar("Best Ever Value")}
        print(name, json.dumps(out[name]), flush=True)
        s.close()
