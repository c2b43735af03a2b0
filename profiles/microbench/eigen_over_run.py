"""Eigendecomposition time over a whole run to convergence (config 3, one GPU): the divide & conquer stage deflates differently as
the spectrum of C spreads from cond 1 to cond 1e6, and the bench only times 20 early generations. Prints one line per window of
generations: generation, cond(C), sigma, best, ms per generation, eigen ms (= sytrd + D&C + join).

    python profiles/microbench/eigen_over_run.py [window] [max_generations]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from korali_b200 import _lib

window = int(sys.argv[1]) if len(sys.argv) > 1 else 50
max_gens = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
s = _lib.Solver(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337)
s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
s.set_scalar("Termination Criteria/Max Value", -1e-9)
s.timing_enable(True)
g = 0
print("generation  cond(C)      sigma        best ever     ms/gen   eigen = sytrd + dc + join")
while g < max_gens:
    s.timing_reset()
    done = 0
    for _ in range(window):
        fin, why = s.check_termination()
        if fin:
            break
        s.run_generation(); done += 1
    if done == 0:
        break
    g += done
    ph = {k: s.timing(k)[0] / done for k in ("generation", "eigen", "eigen_sytrd", "eigen_dc", "eigen_back")}
    print("%9d  %10.3e  %10.3e  %12.4e  %7.3f  %6.3f = %.3f + %.3f + %.3f" % (
        g, s.scalar("Maximum Covariance Eigenvalue") / s.scalar("Minimum Covariance Eigenvalue"), s.scalar("Sigma"), s.scalar("Best Ever Value"),
        ph["generation"], ph["eigen"], ph["eigen_sytrd"], ph["eigen_dc"], ph["eigen_back"]), flush=True)
    if done < window:
        break
fin, why = s.check_termination()
print("finished:", fin, why, "after", g, "generations")
