# A/B of the sytrd exchange variants inside the generation loop (config 3 and config 2) + clock64 phase stamps
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tridiag.py -x -q > gpurun_out/r02_pf_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pf_tests.log
KCMA_SYTRD_OPT=4 timeout 300 python -m pytest tests/test_gpu_tridiag.py -x -q -k sytrd >> gpurun_out/r02_pf_tests.log 2>&1; echo "pytest opt4 rc=$?" >> gpurun_out/r02_pf_tests.log
KCMA_SYTRD_OPT=4 KCMA_SYTRD_PREFETCH=0 timeout 300 python -m pytest tests/test_gpu_tridiag.py -x -q -k sytrd >> gpurun_out/r02_pf_tests.log 2>&1; echo "pytest opt4 pf0 rc=$?" >> gpurun_out/r02_pf_tests.log
tail -8 gpurun_out/r02_pf_tests.log
for cfg in ${ABCFG:-"0 1 2" "1 1 2" "0 4 2" "1 4 2" "0 4 1" "1 4 1"}; do set -- $cfg
  echo "KCMA_SYTRD_PREFETCH=$1 KCMA_SYTRD_OPT=$2 KCMA_SYTRD_COPIES=$3"
  KCMA_SYTRD_PREFETCH=$1 KCMA_SYTRD_OPT=$2 KCMA_SYTRD_COPIES=$3 timeout 120 python - <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from korali_b200 import _lib
for name, case in (("c3", dict(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337)),
                   ("c2", dict(n=100, population_size=4096, objective="NegAckley", initial_value=1.0, initial_stddev=3.0, seed=1337))):
    s = _lib.Solver(**case)
    s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
    for _ in range(3): s.run_generation()
    s.timing_enable(True); s.timing_reset()
    g = 20
    for _ in range(g): s.run_generation()
    sig = s.scalar("Sigma")
    ph = {k: s.timing(k)[0] / g for k in ("eigen", "eigen_sytrd", "eigen_dc", "eigen_back")}
    print(name, "eigen %.3f sytrd %.3f dc %.3f back %.3f sigma %.12g best %.10g" % (ph["eigen"], ph["eigen_sytrd"], ph["eigen_dc"], ph["eigen_back"], sig, s.scalar("Best Ever Value")), flush=True)
    s.close()
PY
done > gpurun_out/r02_pf_ab.log 2>&1
cat gpurun_out/r02_pf_ab.log
for cfg in ${STCFG:-"0 1" "1 4" "0 4"}; do set -- $cfg
echo "PREFETCH=$1 OPT=$2"
KCMA_SYTRD_PREFETCH=$1 KCMA_SYTRD_OPT=$2 KCMA_SYTRD_PROF=100,1 timeout 120 python profiles/microbench/eigen_once.py 1000 2 | head -20; KCMA_SYTRD_PREFETCH=$1 KCMA_SYTRD_OPT=$2 KCMA_SYTRD_PROF=500,77 timeout 120 python profiles/microbench/eigen_once.py 1000 2 | head -20; done > gpurun_out/r02_pf_stamps.log 2>&1
