// Dependent-issue latencies (cycles) of the scalar FP64 / conversion / MUFU / shuffle / shared-memory operations that
// sit on the critical path of a Jacobi rotation round. One warp, clock64 around an unrolled dependent chain.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

#define CHAIN 256
#define MEASURE(name, init, body)                                               \
  {                                                                             \
    init;                                                                       \
    asm volatile("" : "+d"(x));                                                 \
    long long t0 = clock64();                                                   \
    _Pragma("unroll") for (int i = 0; i < CHAIN; i++) { body; }                 \
    asm volatile("" : "+d"(x));                                                 \
    long long t1 = clock64();                                                   \
    if (threadIdx.x == 0) out[slot] = (double)(t1 - t0) / CHAIN;                \
    sink += x;                                                                  \
    slot++;                                                                     \
  }

__device__ __forceinline__ double rsqrt_approx(double a) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); return r; }
__device__ __forceinline__ double rcp_approx(double a) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a)); return r; }
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void lat_kernel(double* out, double seed, double* sinkp) {
  __shared__ double sm[64];
  int slot = 0;
  double sink = 0.0;
  double x;
  sm[threadIdx.x & 63] = seed;
  __syncwarp();
  MEASURE("dfma", x = seed, x = fma(x, 1.0000001, 1e-9))
  MEASURE("dmul", x = seed, x = x * 1.0000001)
  MEASURE("dadd", x = seed, x = x + 1e-9)
  MEASURE("rsqrt64h", x = seed + 1.0, x = rsqrt_approx(x) + 1.0)           // includes one DADD
  MEASURE("rcp64h", x = seed + 1.0, x = rcp_approx(x) + 1.0)               // includes one DADD
  MEASURE("f2f_64_32_64", x = seed, x = (double)((float)x))                // two conversions
  { float y = (float)seed; asm volatile("" : "+f"(y)); long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < CHAIN; i++) y = rsqrtf(y) + 1.0f;
    asm volatile("" : "+f"(y)); long long t1 = clock64(); if (threadIdx.x == 0) out[slot] = (double)(t1 - t0) / CHAIN; sink += y; slot++; }   // MUFU.RSQ + FADD
  { float y = (float)seed; asm volatile("" : "+f"(y)); long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < CHAIN; i++) y = fmaf(y, 1.0001f, 1e-6f);
    asm volatile("" : "+f"(y)); long long t1 = clock64(); if (threadIdx.x == 0) out[slot] = (double)(t1 - t0) / CHAIN; sink += y; slot++; }   // FFMA
  MEASURE("shfl64", x = seed, x = __shfl_xor_sync(0xffffffffu, x, 1))
  MEASURE("lds_sts_syncwarp", x = seed, sm[threadIdx.x] = x; __syncwarp(); x = sm[threadIdx.x ^ 1]; __syncwarp())
  MEASURE("dsqrt", x = seed + 2.0, x = sqrt(x) + 2.0)
  MEASURE("ddiv", x = seed + 2.0, x = 3.0 / x + 1.0)
  { double c0 = seed, c1 = seed; long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < CHAIN; i++) dmma884(c0, c1, 1e-3, 1e-3);
    long long t1 = clock64(); if (threadIdx.x == 0) out[slot] = (double)(t1 - t0) / CHAIN; sink += c0 + c1; slot++; }   // DMMA accumulate chain
  { double c0 = seed, c1 = seed; long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < CHAIN; i++) { dmma884(c0, c1, 1e-3, 1e-3); c0 = c1 * 0.5; }   // DMMA -> A operand dependency (through a DMUL)
    long long t1 = clock64(); if (threadIdx.x == 0) out[slot] = (double)(t1 - t0) / CHAIN; sink += c0 + c1; slot++; }
  if (sink == 123.456) *sinkp = sink;
}

// latency of L2: pointer chase with ld.global.cg over a 32 MB ring; membar.gl after a store; flag ping-pong between 2 CTAs
__global__ void l2_kernel(unsigned long long* ring, int n, double* out, unsigned* flags) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long p = 0;
    long long t0 = clock64();
    for (int i = 0; i < 512; i++) p = __ldcg(ring + p);
    long long t1 = clock64();
    out[0] = (double)(t1 - t0) / 512 + (p == 0xffffffffffffull);
    t0 = clock64();
    for (int i = 0; i < 256; i++) { __stcg(ring + (size_t)i * 64, 0ull + ring[(size_t)i * 64]); __threadfence(); }
    t1 = clock64();
    out[1] = (double)(t1 - t0) / 256;   // load + store + fence
  }
  // ping-pong: CTA 0 and CTA 1 alternate epochs through two flags
  if (threadIdx.x == 0 && blockIdx.x < 2) {
    volatile unsigned* f = flags;
    const int me = blockIdx.x;
    __syncwarp();
    long long t0 = clock64();
    for (unsigned e = 1; e <= 512; e++) {
      if (me == 0) { f[0] = e; __threadfence(); while (f[1] < e) { } }
      else { while (f[0] < e) { } f[1] = e; __threadfence(); }
    }
    long long t1 = clock64();
    if (me == 0) out[2] = (double)(t1 - t0) / 512;   // full round trip (two one-way flag hops)
  }
}

int main() {
  const char* names[] = {"DFMA", "DMUL", "DADD", "RSQ64H+DADD", "RCP64H+DADD", "F2F 64->32->64", "MUFU.RSQ(f32)+FADD", "FFMA", "SHFL 64-bit", "STS+syncwarp+LDS+syncwarp",
                         "sqrt(double)+DADD", "div(double)+DADD", "DMMA.884 accumulate chain", "DMMA.884 + DMUL operand chain"};
  double *out, *sink;
  cudaMalloc(&out, 64 * sizeof(double)); cudaMalloc(&sink, 8);
  for (int rep = 0; rep < 2; rep++) lat_kernel<<<1, 32>>>(out, 1.0, sink);
  double h[64];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  for (int i = 0; i < 14; i++) printf("%-32s %7.1f cycles\n", names[i], h[i]);
  const int n = 4 << 20;   // 32 MB ring, stride 4099 * 8 bytes
  unsigned long long* ring; unsigned* flags;
  cudaMalloc(&ring, (size_t)n * 8); cudaMalloc(&flags, 64); cudaMemset(flags, 0, 64);
  unsigned long long* hr = new unsigned long long[n];
  for (int i = 0; i < n; i++) hr[i] = ((unsigned long long)i + 4099 * 33) % n;
  cudaMemcpy(ring, hr, (size_t)n * 8, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; rep++) { cudaMemset(flags, 0, 64); l2_kernel<<<2, 32>>>(ring, n, out, flags); }
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-32s %7.1f cycles\n%-32s %7.1f cycles\n%-32s %7.1f cycles\n", "ld.global.cg pointer chase (L2)", h[0], "ld + st.cg + __threadfence", h[1], "2-CTA flag ping-pong round trip", h[2]);
  printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
