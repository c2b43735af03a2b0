"""Phase timestamps (clock64) of one CTA of jacobi_pipe_kernel over 32 consecutive steps: where does a step's time go?
Run with KCMA_JACOBI_DEBUG=1 (set below). Slots: see KCMA_TS / KCMA_TSW in korali_b200/csrc/eigen.cu."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("KCMA_JACOBI_DEBUG", "1")   # the sweep to trace
from korali_b200 import _lib
s = _lib.Solver(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337)
for _ in range(5):
    s.run_generation()
buf = (C.c_longlong * 3584)()
rc = _lib.lib().kcma_debug_jacobi_timestamps(buf)
t = np.array(buf[:512], dtype=np.int64).reshape(32, 16)
gs = np.array(buf[512:512 + 249], dtype=np.int64); vs = np.array(buf[1536:1536 + 249], dtype=np.int64)
iv = [("[G] flag wait + bar", 0, 1), ("[G] load + stage + Gram DMMA", 1, 2), ("[G] bar + sum + bar", 2, 3), ("[G] Gamma -> registers", 3, 14),
      ("[G] rotations (registers)", 14, 8), ("[G] ring slot + R + bar", 8, 4), ("[G] apply + stores", 4, 5), ("[G] bar + fence + flags", 5, 6),
      ("[V] wait for R", 9, 10), ("[V] V flags + cp.async issue", 10, 11), ("[V] rows landed (wait + bar)", 11, 12),
      ("[V] apply + stores", 12, 13), ("[V] bar + fence + flags", 13, 15)]
print("rc", rc, "sweep", os.environ["KCMA_JACOBI_DEBUG"], "clock cycles per phase (median over 32 steps of CTA 1):")
for nm, a, b in iv:
    print("  %-38s %8.0f" % (nm, np.median(t[:, b] - t[:, a])))
print("  %-38s %8.0f" % ("[G] step period", np.median(np.diff(t[:, 0]))))
print("  %-38s %8.0f" % ("[V] step period", np.median(np.diff(t[:, 9]))))
print("  %-38s %8.0f" % ("V lag behind G (V step start - G step start)", np.median(t[:, 9] - t[:, 0])))
dg = np.diff(gs)
print("  G step period over the whole sweep: mean %.0f median %.0f; by 25-step bins:" % (dg.mean(), np.median(dg)), [int(dg[i:i + 25].mean()) for i in range(0, 248, 25)])
print("  V lag behind G by 25-step bins:", [int((vs - gs)[i:i + 25].mean()) for i in range(0, 249, 25)])
print("  sweep duration (cycles):", int(gs[-1] - gs[0]))
