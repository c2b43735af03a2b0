// Floor of one all-to-all LL exchange per step among G co-resident CTAs (no matrix arithmetic): every CTA publishes NPUB 16-byte
// slots (value + step tag) into `copies` replicas and polls all G * NPUB slots of replica (b % copies) with 256 threads,
// 1000 dependent steps, optional local phase between the exchanges. copies = G is the push model (a private inbox per consumer).
// Prints cycles per step (max over CTAs).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ll_floor profiles/microbench/ll_exchange_floor.cu && /tmp/ll_floor
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct __align__(16) LL { double v; unsigned long long tag; };
__device__ __forceinline__ void ll_store(LL* p, double v, unsigned long long tag) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((unsigned long long)__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ void ll_load2(const LL* p, unsigned long long (&q)[4]) {
  asm volatile("ld.volatile.global.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(q[0]), "=l"(q[1]), "=l"(q[2]), "=l"(q[3]) : "l"(p) : "memory");
}
// x: 2 parities x copies x nslots. mode bit 0: __nanosleep(64) after a failed poll; bit 1: a single warp polls (lane-strided)
__global__ void __launch_bounds__(256, 1) exch(LL* x, int nslots, int npub, int copies, int steps, int work, int mode, long long* cyc, double* sink) {
  const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  double acc = 0.0;
  const int npoll = (mode & 2) ? 32 : 256;
  long long t0 = clock64();
  for (int i = 0; i < steps; i++) {
    const unsigned long long tag = i + 1ull;
    LL* buf = x + (size_t)(i & 1) * copies * nslots;
    for (int k = tid; k < npub * copies; k += 256) ll_store(buf + (size_t)(k / npub) * nslots + b + (k % npub) * G, (double)i, tag);
    const LL* in = buf + (size_t)(b % copies) * nslots;
    if (tid < npoll)
      for (int pr = tid; pr < nslots / 2; pr += npoll) {
        unsigned long long q[4];
        for (;;) {
          ll_load2(in + 2 * pr, q);
          if (q[1] == tag && q[3] == tag) break;
          if (mode & 1) __nanosleep(64);
        }
        acc += __longlong_as_double((long long)q[0]) + __longlong_as_double((long long)q[2]);
      }
    __syncthreads();
    for (int k = 0; k < work; k++) acc = acc * 1.0000001 + 1e-9;   // local phase (dependent FP64 chain)
    __syncthreads();
  }
  if (tid == 0) cyc[b] = clock64() - t0;
  if (acc == 12345.678) sink[0] = acc;
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  LL* x; long long* cyc; double* sink;
  const size_t cap = (size_t)2 * 148 * 2100;
  cudaMalloc(&x, sizeof(LL) * cap); cudaMalloc(&cyc, 8 * 256); cudaMalloc(&sink, 8);
  int steps = 1000;
  for (int G : {16, 74, 148})
    for (int npub : {1, 7, 14})
      for (int copies : {1, 2, 4, 8, 16, 37, G})
        for (int mode : {0, 1, 2})
          for (int work : {0}) {
            if (G > sms || ((G * npub) & 1) || copies > G || (copies == G && (G == 16 && copies == 16) && false)) continue;
            if ((npub == 14 && G == 148) || (npub == 7 && G == 74 && false)) continue;
            if (mode && !(copies == 1 || copies == 2 || copies == G)) continue;
            int nslots = G * npub;
            if ((size_t)2 * copies * nslots > cap) continue;
            cudaMemset(x, 0, sizeof(LL) * cap);
            void* args[] = {&x, &nslots, &npub, &copies, &steps, &work, &mode, &cyc, &sink};
            cudaError_t e = cudaLaunchCooperativeKernel((const void*)exch, dim3(G), dim3(256), args, 0, 0);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[256]; cudaMemcpy(h, cyc, 8 * G, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int k = 0; k < G; k++) mx = h[k] > mx ? h[k] : mx;
            printf("G %3d values/CTA %2d copies %3d mode %d (polled %6d B, stored %6d B per CTA and step): %7.0f cycles/step\n", G, npub, copies, mode,
                   nslots * 16, npub * copies * 16, (double)mx / steps);
            fflush(stdout);
          }
  return 0;
}
