// Floor of one all-to-all LL exchange per step among G co-resident CTAs (no arithmetic): every CTA publishes NPUB 16-byte slots
// (value + step tag) and polls all G * NPUB slots with 256 threads, 1000 dependent steps. Prints cycles per step.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ll_floor profiles/microbench/ll_exchange_floor.cu && /tmp/ll_floor
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct __align__(16) LL { double v; unsigned long long tag; };
template <int SCOPE>
__device__ __forceinline__ void ll_store(LL* p, double v, unsigned long long tag) {
  if (SCOPE == 0) asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((unsigned long long)__double_as_longlong(v)), "l"(tag) : "memory");
  else asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((unsigned long long)__double_as_longlong(v)), "l"(tag) : "memory");
}
template <int SCOPE>
__device__ __forceinline__ void ll_load2(const LL* p, unsigned long long (&q)[4]) {
  if (SCOPE == 0) asm volatile("ld.volatile.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(p) : "memory");
  else asm volatile("ld.relaxed.gpu.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(p) : "memory");
}
// slots: 2 parities x nslots; CTA b publishes slots b, b + G, ... (npub of them); thread t polls the pairs t, t + 256, ...
template <int SCOPE>
__global__ void __launch_bounds__(256, 1) exch(LL* x, int nslots, int npub, int steps, int work, long long* cyc, double* sink) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  double acc = 0.0;
  long long t0 = clock64();
  for (int i = 0; i < steps; i++) {
    const unsigned long long tag = i + 1ull;
    LL* buf = x + (size_t)(i & 1) * nslots;
    if (tid < npub) ll_store<SCOPE>(buf + b + tid * G, (double)i, tag);
    for (int pr = tid; pr < nslots / 2; pr += 256) {
      unsigned long long q[4];
      do { ll_load2<SCOPE>(buf + 2 * pr, q); } while (q[1] != tag || q[3] != tag);
      acc += __longlong_as_double((long long)q[0]) + __longlong_as_double((long long)q[2]);
    }
    __syncthreads();
    for (int k = 0; k < work; k++) acc = acc * 1.0000001 + 1e-9;   // local phase (dependent FP64 chain, ~8 cycles each)
    __syncthreads();
  }
  if (tid == 0) cyc[b] = clock64() - t0;
  if (acc == 12345.678) sink[0] = acc;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  LL* x; long long* cyc; double* sink;
  cudaMalloc(&x, sizeof(LL) * 2 * 4096); cudaMalloc(&cyc, 8 * 256); cudaMalloc(&sink, 8);
  const int steps = 1000;
  for (int scope = 0; scope < 2; scope++)
    for (int G : {16, 37, 74, 148})
      for (int npub : {1, 7})
        for (int work : {0, 400}) {
          if (G > sms) continue;
          int nslots = G * npub; nslots += nslots & 1;
          if (nslots & 1) nslots++;
          cudaMemset(x, 0, sizeof(LL) * 2 * 4096);
          int st = steps, np = npub, wk = work;
          // an odd nslots would leave an unpublished pad slot: use even G * npub only
          if ((G * npub) & 1) continue;
          void* args[] = {&x, &nslots, &np, &st, &wk, &cyc, &sink};
          const void* fn = scope == 0 ? (const void*)exch<0> : (const void*)exch<1>;
          cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(G), dim3(256), args, 0, 0);
          if (e == cudaSuccess) e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[256]; cudaMemcpy(h, cyc, 8 * G, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int k = 0; k < G; k++) mx = h[k] > mx ? h[k] : mx;
          printf("scope %s G %3d slots/CTA %d polled bytes/CTA/step %6d local work %4d x ~8 cyc: %7.0f cycles/step\n", scope ? "gpu" : "sys(volatile)", G, npub,
                 nslots * 16, work, (double)mx / steps);
        }
  return 0;
}
