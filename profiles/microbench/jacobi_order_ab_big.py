"""A/B of the tournament order on the LARGE-N eigensolver path (jacobi_big_step_kernel, one launch per step, 16-row blocks):
KCMA_JACOBI_ORDER=rr|ring (read per decomposition). One process, no torch. For each N a CMA-ES-like covariance
C = (1 - c) I + c Z^T Z / m (clustered spectrum) is decomposed from the identity basis with both orders: ms, sweeps,
residual and orthonormality of the result.

    python profiles/microbench/jacobi_order_ab_big.py [N ...]        (default 2048 4096)
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from korali_b200 import _lib

sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096]
for n in sizes:
    rng = np.random.default_rng(n)
    z = rng.standard_normal((2 * n, n))
    c = 0.965 * np.eye(n) + 0.035 * (z.T @ z) / (2 * n)
    c = 0.5 * (c + c.T)
    cmax = np.abs(c).max()
    for order in ("rr", "ring"):
        os.environ["KCMA_JACOBI_ORDER"] = order
        s = _lib.Solver(n=n, population_size=8, objective="NegSphere", initial_value=1.0, initial_stddev=1.0, seed=1)
        s.set("Covariance Matrix", c)
        s.timing_enable(True); s.timing_reset()
        t0 = time.perf_counter()
        s.ask()
        s.scalar("Sigma")
        t1 = time.perf_counter()
        eig = s.timing("eigen")[0]
        sweeps = s.timing("eigen_sweeps")[1]
        b = s.get("Covariance Eigenvector Matrix").reshape(n, n); d = s.get("Axis Lengths")
        res = np.abs((b * d**2) @ b.T - c).max() / cmax
        orth = np.abs(b.T @ b - np.eye(n)).max()
        print("N=%d order=%-4s  eigen %.1f ms (wall of ask %.1f ms)  sweeps %d  |B D^2 B^T - C|/|C| %.1e  |B^T B - I| %.1e  ascending %s"
              % (n, order, eig, 1e3 * (t1 - t0), sweeps, res, orth, bool(np.all(np.diff(d) >= 0))), flush=True)
        s.close()
