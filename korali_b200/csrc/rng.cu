// rng.cu — K1: counter-based Philox4x32-10 -> N(0,1) draws (replaces the MT19937/gsl_ran_gaussian loop of
// prepareGeneration, CMAES.cpp.base:449-450 and :467-473; draw order there is sample-major, dimension-minor).
//
// Counter layout (restated on the CPU in oracle/okcma.c philox_normal_pair):
//   ctr = { column pair p = d/2, z-row index (sample, or pair when mirrored), resampling attempt, generation }
//   key = 64-bit seed ("Random Seed" of the Normal Generator)
// Each Philox block gives two 52-bit uniforms in (0,1) -> Box-Muller -> (z[2p], z[2p+1]).
// Results are independent of the launch geometry and of how the population is sharded across GPUs.
#include "common.cuh"
#include "kernels.h"

namespace kc {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double unit_open52(uint32_t lo, uint32_t hi) {
  const unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (double)(v >> 12) * 0x1.0p-52 + 0x1.0p-53;  // exact, in [2^-53, 1-2^-53]
}

// One thread per (row, column pair). Z is row-major with leading dimension ldz (even), pads stay zero.
// row_list (nullable): regenerate only these local rows (resampling); attempt (nullable): per-row attempt counters.
__device__ __forceinline__ void philox_normal_pair(double* __restrict__ Z, int ldz, int n, long long row, int p, unsigned long long grow,
                                                   unsigned att, unsigned generation, unsigned long long seed) {
  uint32_t r[4];
  philox4x32_10((uint32_t)p, (uint32_t)grow, att, generation, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  const double u1 = unit_open52(r[0], r[1]);
  const double u2 = unit_open52(r[2], r[3]);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  double* dst = Z + (size_t)row * ldz + 2 * p;
  if (2 * p + 1 < n) {
    *reinterpret_cast<double2*>(dst) = make_double2(rad * c, rad * s);
  } else {
    dst[0] = rad * c;
  }
}

// WIDE = true (npairs >= 128): a CTA walks rows, its threads the pairs of a row - no 64-bit division per draw
template <bool WIDE>
__global__ void __launch_bounds__(256)
philox_normal_kernel(double* __restrict__ Z, int ldz, long long rows, int n, unsigned long long seed, unsigned generation,
                     unsigned long long row_begin, const unsigned* __restrict__ attempt,
                     const int* __restrict__ row_list, const DevScalars* __restrict__ gen_src) {
  if (generation == kGenFromDevice) generation = (unsigned)gen_src->gen;   // CUDA-graph replay
  const int npairs = (n + 1) >> 1;
  if (WIDE) {
    for (long long li = blockIdx.x; li < rows; li += gridDim.x) {
      const long long row = row_list ? row_list[li] : li;
      const unsigned long long grow = row_begin + (unsigned long long)row;
      const unsigned att = attempt ? attempt[row] : 0u;
      for (int p = threadIdx.x; p < npairs; p += blockDim.x) philox_normal_pair(Z, ldz, n, row, p, grow, att, generation, seed);
    }
    return;
  }
  const long long total = rows * npairs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long li = idx / npairs;
    const int p = (int)(idx - li * npairs);
    const long long row = row_list ? row_list[li] : li;
    const unsigned long long grow = row_begin + (unsigned long long)row;
    const unsigned att = attempt ? attempt[row] : 0u;
    philox_normal_pair(Z, ldz, n, row, p, grow, att, generation, seed);
  }
}

// Discrete mutations of the first _numberOfDiscreteMutations samples, then discretize() of every sample
// (CMAES.cpp.base:515-544 and :453 / :478-481, :862-867; after Hansen 2011, "A CMA-ES for Mixed-Integer Nonlinear
// Optimization"). The reference draws from its _uniformGenerator; here: a second Philox stream,
// key = { seed_lo, seed_hi ^ "DISC" }, ctr = { block, GLOBAL sample index, resampling attempt, generation }, draw k = half
// (k & 1) of block k >> 1 (restated in oracle/okcma.c philox_uniform). One thread per local sample; X is authoritative.
__device__ __forceinline__ double discrete_uniform(unsigned long long seed, unsigned generation, unsigned attempt, unsigned long long sample,
                                                   unsigned k) {
  uint32_t r[4];
  philox4x32_10(k >> 1, (uint32_t)sample, attempt, generation, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x44495343u, r);
  return (k & 1u) ? unit_open52(r[2], r[3]) : unit_open52(r[0], r[1]);
}

__global__ void __launch_bounds__(128)
discrete_mutation_kernel(double* __restrict__ X, int ldx, long long samples, int n, unsigned long long sample_begin, int mirrored,
                         const DevScalars* __restrict__ sc, const double* __restrict__ mask, const double* __restrict__ gran,
                         const double* __restrict__ best_ever, unsigned long long seed, unsigned generation,
                         const unsigned* __restrict__ attempt, double* __restrict__ disc_mut) {
  const long long li = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (li >= samples) return;
  if (generation == kGenFromDevice) generation = (unsigned)sc->gen;
  const unsigned long long i = sample_begin + (unsigned long long)li;   // global sample index
  const unsigned att = attempt ? attempt[mirrored ? (li >> 1) : li] : 0u;
  const unsigned long long ndm = (unsigned long long)sc->n_disc_mut;
  const int n_mask = sc->n_mask;
  double* x = X + (size_t)li * ldx;
  double* dm = disc_mut + (size_t)li * ldx;
  for (int d = 0; d < n; d++) dm[d] = 0.0;
  unsigned k = 0;
  if ((i + 1) < ndm) {
    const double p_geom = pow(0.7, 1.0 / (double)n_mask);
    unsigned long long select = (unsigned long long)floor(discrete_uniform(seed, generation, att, i, k++) * (double)n_mask);
    for (int d = 0; d < n; ++d)
      if ((mask[d] == 1.0) && (select-- == 0)) {
        double dmutation = 1.0;
        while (discrete_uniform(seed, generation, att, i, k++) > p_geom) dmutation += 1.0;
        dmutation *= gran[d];
        if (discrete_uniform(seed, generation, att, i, k++) > 0.5) dmutation *= -1.0;
        dm[d] = dmutation;
        x[d] += dmutation;
      }
  } else if ((i + 1) == ndm) {
    for (int d = 0; d < n; ++d)
      if (gran[d] != 0.0) {
        const double dmutation = round(best_ever[d] / gran[d]) * gran[d] - x[d];
        dm[d] = dmutation;
        x[d] += dmutation;
      }
  }
  for (int d = 0; d < n; ++d)
    if (gran[d] != 0.0) x[d] = round(x[d] / gran[d]) * gran[d];
}

void launch_discrete_mutation(cudaStream_t st, double* X, int ldx, long long samples, int n, unsigned long long sample_begin, int mirrored,
                              const DevScalars* sc, const double* mask, const double* gran, const double* best_ever, unsigned long long seed,
                              unsigned generation, const unsigned* attempt, double* disc_mut) {
  if (samples <= 0) return;
  discrete_mutation_kernel<<<(unsigned)((samples + 127) / 128), 128, 0, st>>>(X, ldx, samples, n, sample_begin, mirrored, sc, mask, gran,
                                                                             best_ever, seed, generation, attempt, disc_mut);
}

// Raw Philox block (known-answer tests).
__global__ void philox_raw_kernel(const uint32_t* in, uint32_t* out) {
  uint32_t r[4];
  philox4x32_10(in[0], in[1], in[2], in[3], in[4], in[5], r);
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2]; out[3] = r[3];
}

void launch_philox_normal(cudaStream_t st, double* Z, int ldz, long long rows, int n, unsigned long long seed,
                          unsigned generation, unsigned long long row_begin, const unsigned* attempt,
                          const int* row_list, int num_sms, const DevScalars* gen_src) {
  if (rows <= 0) return;
  const long long total = rows * ((n + 1) / 2);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms * 16;
  if (blocks > cap) blocks = cap;
  const int npairs = (n + 1) / 2;
  if (npairs >= 128) {
    const unsigned g = (unsigned)std::min<long long>(rows, cap);
    if (npairs <= 160) philox_normal_kernel<true><<<g, 128, 0, st>>>(Z, ldz, rows, n, seed, generation, row_begin, attempt, row_list, gen_src);
    else philox_normal_kernel<true><<<g, 256, 0, st>>>(Z, ldz, rows, n, seed, generation, row_begin, attempt, row_list, gen_src);
    return;
  }
  philox_normal_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(Z, ldz, rows, n, seed, generation, row_begin, attempt, row_list, gen_src);
}

void launch_philox_raw(cudaStream_t st, const uint32_t* in6, uint32_t* out4) {
  philox_raw_kernel<<<1, 1, 0, st>>>(in6, out4);
}

}  // namespace kc
