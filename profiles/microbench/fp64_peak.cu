// Microbenchmark: FP64 DMMA.8x8x4 vs DFMA issue-rate peak on B200 (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void __launch_bounds__(1024) dmma_loop(double* out, int iters) {
  double c[NACC][2];
  double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-10;
#pragma unroll
  for (int j = 0; j < NACC; j++) c[j][0] = c[j][1] = 0.0;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < NACC; j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < NACC; j++) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void __launch_bounds__(1024) dfma_loop(double* out, int iters) {
  double c[NACC];
  double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 1e-10;
#pragma unroll
  for (int j = 0; j < NACC; j++) c[j] = j;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < NACC; j++) c[j] = fma(c[j], a, b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < NACC; j++) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}
int main() {
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * nsm * 4 * 1024);
  const int iters = 20000;
  int tpbs[] = {128, 256, 512, 1024};
  for (int tpb : tpbs) {
    for (int bps = 1; bps <= 2; bps++) {
      if (tpb * bps > 2048) continue;
      int grid = nsm * bps;
      {
        float ms = timeit([&] { dmma_loop<8><<<grid, tpb>>>(out, iters); });
        double flop = 2.0 * 8 * 8 * 4 * 8.0 * iters * (tpb / 32) * grid;
        printf("DMMA  acc=8 tpb=%4d blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", tpb, bps, ms, flop / ms * 1e-9);
      }
      {
        float ms = timeit([&] { dmma_loop<2><<<grid, tpb>>>(out, iters); });
        double flop = 2.0 * 8 * 8 * 4 * 2.0 * iters * (tpb / 32) * grid;
        printf("DMMA  acc=2 tpb=%4d blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", tpb, bps, ms, flop / ms * 1e-9);
      }
      {
        float ms = timeit([&] { dfma_loop<8><<<grid, tpb>>>(out, iters); });
        double flop = 2.0 * 8.0 * iters * tpb * grid;
        printf("DFMA  acc=8 tpb=%4d blocks/SM=%d : %.3f ms  %.2f TFLOP/s\n", tpb, bps, ms, flop / ms * 1e-9);
      }
    }
  }
  // single-warp DMMA latency (dependent chain)
  {
    float ms = timeit([&] { dmma_loop<1><<<1, 32>>>(out, iters); });
    printf("DMMA dependent chain, 1 warp: %.1f ns per DMMA\n", ms * 1e6 / iters);
    float ms2 = timeit([&] { dfma_loop<1><<<1, 32>>>(out, iters); });
    printf("DFMA dependent chain, 1 warp: %.1f ns per DFMA\n", ms2 * 1e6 / iters);
  }
  return 0;
}
