"""A few generations of the DEA path for `ncu --metrics gpu__time_duration.sum` launch lists.   python profiles/microbench/dea_once.py [n] [lambda]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from korali_b200 import _dea
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lam = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
s = _dea.Solver(n=n, population_size=lam, objective="NegSphere", lower_bound=-5.0, upper_bound=5.0, seed=3, crossover_rate=0.5)
for _ in range(4):
    s.run_generation()
print("best", s.scalar("Best Ever Value"), "launches", s.launch_count())
