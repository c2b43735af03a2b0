import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from korali_b200 import _lib
cfgs = {"config1": dict(n=10, population_size=32, objective="NegRosenbrock", initial_value=0.0, initial_stddev=0.5),
        "config2": dict(n=100, population_size=4096, objective="NegAckley", initial_value=1.0, initial_stddev=3.0),
        "config3": dict(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0)}
for name, kw in cfgs.items():
    gens = 20 if name == "config3" else 500
    s = _lib.Solver(seed=1337, **kw)
    s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
    for _ in range(5): s.run_generation()
    torch.cuda.synchronize(); l0 = s.launch_count(); t0 = time.perf_counter()
    for _ in range(gens): s.run_generation()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(name, "KCMA_GRAPH=%s" % os.environ.get("KCMA_GRAPH", "1"), "ms/gen %.4f" % (1e3 * (t1 - t0) / gens), "launches/gen", (s.launch_count() - l0) / gens, "best", s.scalar("Best Ever Value"))
    s.close()
