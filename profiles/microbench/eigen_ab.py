"""A/B of the eigensolvers inside the generation loop (KCMA_EIGEN=tridiag|jacobi, read per call) in ONE process, no torch:
ms per generation, eigen ms and its stages for config 2 (N=100), config 3 (N=1000, lambda=65536) and an N=4096 case.

    python profiles/microbench/eigen_ab.py [generations] [cases, comma separated: c2,c3,n4096]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from korali_b200 import _lib

gens = int(sys.argv[1]) if len(sys.argv) > 1 else 20
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c2", "c3", "n4096"]
CASES = {
    "c2": dict(n=100, population_size=4096, objective="NegAckley", initial_value=1.0, initial_stddev=3.0, seed=1337),
    "c3": dict(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337),
    "n2000": dict(n=2000, population_size=16384, objective="NegSphere", initial_value=1.0, initial_stddev=1.0, seed=1337),
    "n4096": dict(n=4096, population_size=16384, mirrored_sampling=1, objective="NegSphere", initial_value=1.0, initial_stddev=1.0, seed=1337),
}
for name in which:
    case = CASES[name]
    for solver in ("tridiag", "jacobi", "tridiag"):
        os.environ["KCMA_EIGEN"] = solver
        g = gens if case["n"] < 4000 else max(3, gens // 5)
        s = _lib.Solver(**case)
        s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
        for _ in range(3):
            s.run_generation()
        s.timing_enable(True); s.timing_reset()
        s.scalar("Sigma")
        t0 = time.perf_counter()
        for _ in range(g):
            s.run_generation()
        sig = s.scalar("Sigma")
        t1 = time.perf_counter()
        ph = {k: s.timing(k)[0] / g for k in ("eigen", "eigen_sytrd", "eigen_dc", "eigen_back", "sample_gemm", "rank_mu")}
        nn = case["n"]
        cm = s.get("Covariance Matrix").reshape(nn, nn)
        s.ask()
        b = s.get("Covariance Eigenvector Matrix").reshape(nn, nn); d = s.get("Axis Lengths")
        print("%-6s %-7s %8.3f ms/gen (wall)  eigen %8.3f = sytrd %7.3f + dc %7.3f + back %7.3f | gemm %.3f rank_mu %.3f | sigma %.12g best %.10g "
              "|BD2Bt-C|/|C| %.1e |BtB-I| %.1e" % (name, solver, 1e3 * (t1 - t0) / g, ph["eigen"], ph["eigen_sytrd"], ph["eigen_dc"], ph["eigen_back"],
              ph["sample_gemm"], ph["rank_mu"], sig, s.scalar("Best Ever Value"),
              np.abs((b * d**2) @ b.T - cm).max() / np.abs(cm).max(), np.abs(b.T @ b - np.eye(nn)).max()), flush=True)
        s.close()
