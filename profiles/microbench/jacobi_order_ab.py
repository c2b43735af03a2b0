"""A/B of the tournament order of jacobi_pipe_kernel (KCMA_JACOBI_ORDER=rr|ring, read per launch) in ONE process, no torch:
(1) the decomposition itself under the ring order at sizes on both sides of its dispatch rule (cold start);
(2) config 3 (N = 1000, lambda = 65536): ms per generation, eigen ms and sweeps per decomposition for both orders.

    python profiles/microbench/jacobi_order_ab.py [generations] [orders, comma separated]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from korali_b200 import _lib

gens = int(sys.argv[1]) if len(sys.argv) > 1 else 20
orders = sys.argv[2].split(",") if len(sys.argv) > 2 else ["rr", "ring", "rr", "ring"]   # e.g. "ring,anchor,ring,anchor"
os.environ["KCMA_JACOBI_ORDER"] = "ring"
for n in (25, 31, 64, 120, 300, 1000, 1001):
    rng = np.random.default_rng(n)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.sort(10.0 ** rng.uniform(0, 3, n))
    c = (q * lam) @ q.T
    c = 0.5 * (c + c.T)
    w, v = _lib.k_eigen(c)
    print("ring N=%4d  residual %.2e  orthonormality %.2e  eigenvalues %.2e  ascending %s" % (
        n, np.abs(v @ np.diag(w) @ v.T - c).max() / np.abs(c).max(), np.abs(v.T @ v - np.eye(n)).max(),
        np.abs(w - lam).max() / lam.max(), bool(np.all(np.diff(w) >= 0))), flush=True)
n = 64
e = np.random.default_rng(n).standard_normal((n, n))
c = np.eye(n) + 1e-6 * 0.5 * (e + e.T)
w, v = _lib.k_eigen(c)
print("ring clustered N=64  residual %.2e  orthonormality %.2e  eigenvalues %.2e" % (
    np.abs(v @ np.diag(w) @ v.T - c).max(), np.abs(v.T @ v - np.eye(n)).max(), np.abs(w - np.linalg.eigvalsh(c)).max()), flush=True)

case = dict(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337)
for order in orders:
    os.environ["KCMA_JACOBI_ORDER"] = order
    s = _lib.Solver(**case)
    s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
    for _ in range(3):
        s.run_generation()
    s.timing_enable(True); s.timing_reset()
    s.scalar("Sigma")
    t0 = time.perf_counter()
    for _ in range(gens):
        s.run_generation()
    sig = s.scalar("Sigma")
    t1 = time.perf_counter()
    eig = s.timing("eigen")[0] / gens
    sweeps = s.timing("eigen_sweeps")[1] / gens
    nn = case["n"]
    cm = s.get("Covariance Matrix").reshape(nn, nn)
    s.ask()
    b = s.get("Covariance Eigenvector Matrix").reshape(nn, nn); d = s.get("Axis Lengths")
    print("config3 order=%-4s  %.3f ms/generation (wall)  eigen %.3f ms  sweeps/decomposition %.2f  sigma %.12g  best %.10g  "
          "|B D^2 B^T - C|/|C| %.1e  |B^T B - I| %.1e" % (order, 1e3 * (t1 - t0) / gens, eig, sweeps, sig, s.scalar("Best Ever Value"),
          np.abs((b * d**2) @ b.T - cm).max() / np.abs(cm).max(), np.abs(b.T @ b - np.eye(nn)).max()), flush=True)
    s.close()
