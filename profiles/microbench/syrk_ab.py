"""A/B of the rank-mu product inside the generation loop (KCMA_SYRK=splitk | default stream-K, read per call): config 3 and a
one-rank shard of config 4, ms per generation and the rank_mu phase.

    python profiles/microbench/syrk_ab.py [generations]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from korali_b200 import _lib

gens = int(sys.argv[1]) if len(sys.argv) > 1 else 20
CASES = {
    "c3": dict(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337),
    "c3/8": dict(n=1000, population_size=8192, mu_value=4096, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337),
    "c2": dict(n=100, population_size=4096, objective="NegAckley", initial_value=1.0, initial_stddev=3.0, seed=1337),
    "n2000": dict(n=2000, population_size=16384, objective="NegSphere", initial_value=1.0, initial_stddev=1.0, seed=1337),
}
for name, case in CASES.items():
    for mode in ("splitk", "streamk", "splitk", "streamk"):
        if mode == "splitk":
            os.environ["KCMA_SYRK"] = "splitk"
        else:
            os.environ.pop("KCMA_SYRK", None)
        s = _lib.Solver(**case)
        s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
        for _ in range(3):
            s.run_generation()
        s.timing_enable(True); s.timing_reset()
        for _ in range(gens):
            s.run_generation()
        sig = s.scalar("Sigma")
        ph = {k: s.timing(k)[0] / gens for k in ("rank_mu", "paths", "generation")}
        mu = case.get("mu_value", case["population_size"] // 2)
        tf = case["n"] * (case["n"] + 1) * mu / (ph["rank_mu"] * 1e-3) / 1e12
        print("%-6s %-8s rank_mu %.4f ms (%.2f TFLOP/s credited) paths %.4f generation %.3f sigma %.12g best %.10g" % (
            name, mode, ph["rank_mu"], tf, ph["paths"], ph["generation"], sig, s.scalar("Best Ever Value")), flush=True)
        s.close()
