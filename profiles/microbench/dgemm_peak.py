"""cuBLAS DGEMM peak on this B200 (roofline denominator for the FP64 GEMM kernels).
Same method as MEASURED_PEAKS.json: torch.matmul f64, best of 10 (burst) and a 4 s sustained loop."""
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3): torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best * 1e-9
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.time(); k = 0; e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5): torch.matmul(a, b, out=c)
    k += 5; torch.cuda.synchronize()
e1.record(); e1.synchronize()
sust = 2 * n**3 * k / e0.elapsed_time(e1) * 1e-9
# also the two shapes of the CMA-ES path
def tm(f, reps=5):
    f(); torch.cuda.synchronize(); bb = 1e9
    for _ in range(reps):
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); e.synchronize(); bb = min(bb, s.elapsed_time(e))
    return bb
N, lam, mu = 1000, 65536, 32768
Z = torch.randn(lam, N, dtype=torch.float64, device="cuda"); A = torch.randn(N, N, dtype=torch.float64, device="cuda")
Y = torch.empty(lam, N, dtype=torch.float64, device="cuda")
t_s = tm(lambda: torch.matmul(Z, A.t(), out=Y))
S = Z[:mu]
P = torch.empty(N, N, dtype=torch.float64, device="cuda")
t_r = tm(lambda: torch.matmul(S.t(), S, out=P))
out = {"fp64_dgemm_tflops": burst, "fp64_dgemm_tflops_sustained": sust, "n": n,
       "cublas_sampling_gemm_ms": t_s, "cublas_sampling_gemm_tflops": 2 * N * N * lam / t_s * 1e-9,
       "cublas_rankmu_full_gemm_ms": t_r, "cublas_rankmu_full_gemm_tflops": 2 * N * N * mu / t_r * 1e-9,
       "gpu": torch.cuda.get_device_name(0), "how": "torch.matmul float64 8192^3 best of 10 / 4 s loop; CMA-ES shapes N=1000 lam=65536 mu=32768"}
print(json.dumps(out))
