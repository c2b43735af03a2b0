"""Phase timestamps (clock64) of one CTA of jacobi_gram_kernel over 32 consecutive steps: where does a step's time go?
Run with KCMA_JACOBI_DEBUG=1."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ["KCMA_JACOBI_DEBUG"] = "1"
from korali_b200 import _lib
s = _lib.Solver(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337)
for _ in range(5):
    s.run_generation()
buf = (C.c_longlong * 256)()
rc = _lib.lib().kcma_debug_jacobi_timestamps(buf)
t = np.array(buf[:], dtype=np.int64).reshape(32, 8)
names = ["flag wait", "sync", "gram loads+dmma+prefetch", "sync+sum+sync", "inner rotations+sync", "apply+stores", "sync+fence+flag"]
d = np.diff(t, axis=1)
print("rc", rc, "clock cycles per phase (median over 32 steps of CTA 1):")
for i, nm in enumerate(names):
    print("  %-28s %8.0f" % (nm, np.median(d[:, i])))
print("  step period (t0 -> next t0)   %8.0f" % np.median(np.diff(t[:, 0])))
