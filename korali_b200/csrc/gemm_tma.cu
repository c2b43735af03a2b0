// gemm_tma.cu — version 2 of the FP64 tensor-core kernels: operand tiles are staged by the Tensor Memory Accelerator
// (cp.async.bulk.tensor, SASS UTMALDG) into 128-byte-swizzled shared memory, signalled through mbarriers, and consumed by
// DMMA.8x8x4 warps. One producer warp + eight consumer warps per CTA (warp specialisation), 6-stage ring.
//
//   gemm_tn_tma : Y[M x Nc] = Z[M x K] * A[Nc x K]^T     (K2, sampling product)
//   syrk_tt_tma : P = S^T S, lower-triangular tiles, split-K (K6, rank-mu sum)
//
// Conflict-free fragment reads under the hardware swizzle. A swizzled tile stores the 16-byte chunk j of 128-byte row r
// at chunk (j ^ (r & 7)). The DMMA contraction index is free to be permuted as long as both operands use the same
// permutation, so the k index each lane fetches is CHOSEN to make a half-warp's LDS.64 hit 16 distinct bank pairs:
//   K-contiguous tiles (TN):  lane (g = lane>>2, t = lane&3), step s:  k = 2s + (t&1) + 8(t>>1)
//   N-contiguous tiles (SYRK): lane (g, t), step s:                    k = 8(s>>1) + 2t + (s&1)
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "kernels.h"

namespace kc {

namespace {
constexpr int TBM = 128, TBN = 128, TBK = 16;
constexpr int TSTAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int TTHREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int TILE_BYTES = TBM * TBK * 8;          // 16 KB per operand tile
constexpr int STAGE_BYTES = 2 * TILE_BYTES;        // 32 KB
constexpr size_t TMA_SMEM = (size_t)TSTAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void store_pair2(double* p, bool ok0, bool ok1, double c0, double c1) {
  if (ok1) *reinterpret_cast<double2*>(p) = make_double2(c0, c1);
  else if (ok0) p[0] = c0;
}
}  // namespace

// -----------------------------------------------------------------------------------------------------------------
// tmA: tensor {K (ld), M rows}, box {16, 128}; tmB: tensor {K (ld), Nc rows}, box {16, 128}; both SWIZZLE_128B.
// -----------------------------------------------------------------------------------------------------------------
// PERSISTENT: grid = #SMs; every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (column tile fastest, so the
// eight column tiles of one Z row panel are in flight together and share it in L2). The stage ring and its phases run
// continuously across tiles, so the producer warp is already loading the next tile while the DMMA warps store this one.
__global__ void __launch_bounds__(TTHREADS, 1)
gemm_tn_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int Nc, int K,
                   double* __restrict__ C, int ldc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)TSTAGES * STAGE_BYTES);
  uint64_t* empty = full + TSTAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = (K + TBK - 1) / TBK;
  const int ntn = (Nc + TBN - 1) / TBN, ntm = (M + TBM - 1) / TBM;
  const int ntiles = ntn * ntm;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ===== TMA producer warp (one elected lane) =====
    if (lane == 0) {
      int it = 0;   // global k-tile counter across output tiles
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = (tile / ntn) * TBM, col0 = (tile % ntn) * TBN;
        for (int kt = 0; kt < nk; kt++, it++) {
          const int s = it % TSTAGES;
          if (it >= TSTAGES) mbar_wait(&empty[s], ((it / TSTAGES) - 1) & 1);
          mbar_expect_tx(&full[s], STAGE_BYTES);
          tma_load_2d(tiles + (size_t)s * STAGE_BYTES, &tmA, kt * TBK, row0, &full[s]);
          tma_load_2d(tiles + (size_t)s * STAGE_BYTES + TILE_BYTES, &tmB, kt * TBK, col0, &full[s]);
        }
      }
    }
    return;
  }
  // ===== DMMA consumer warps: 2 x 4 warps, warp tile 64 x 32 =====
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  // byte offset inside a tile of the element (row = base + 8 i + g, k = 2s + (t&1) + 8 (t>>1)):
  //   row*128 + (((s + 4 (t>>1)) ^ g) << 4) + (t&1)*8          (row & 7 == g)
  int koff[4];
#pragma unroll
  for (int s = 0; s < 4; s++) koff[s] = (((s + 4 * (t >> 1)) ^ g) << 4) + (t & 1) * 8;
  const int a_row = (wm * 64 + g) * 128, b_row = (wn * 32 + g) * 128;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int row0 = (tile / ntn) * TBM, col0 = (tile % ntn) * TBN;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kt = 0; kt < nk; kt++, it++) {
      const int s = it % TSTAGES;
      mbar_wait(&full[s], (it / TSTAGES) & 1);
      const uint8_t* as = tiles + (size_t)s * STAGE_BYTES + a_row;
      const uint8_t* bs = tiles + (size_t)s * STAGE_BYTES + TILE_BYTES + b_row;
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const double*>(as + i * 1024 + koff[kk]);
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = *reinterpret_cast<const double*>(bs + j * 1024 + koff[kk]);
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int row = row0 + wm * 64 + i * 8 + g;
      if (row >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int col = col0 + wn * 32 + j * 8 + 2 * t;
        store_pair2(C + (size_t)row * ldc + col, col < Nc, col + 1 < Nc, acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}

// -----------------------------------------------------------------------------------------------------------------
// tmS: tensor {n columns (ld), K rows}, box {16 columns, 16 rows}, SWIZZLE_128B. A 128-column operand tile = 8 boxes.
// -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TTHREADS, 1)
syrk_tt_tma_kernel(const __grid_constant__ CUtensorMap tmS, int n, int K, const int* __restrict__ kptr, double* __restrict__ W, int ldw) {
  if (kptr) K = min(K, *kptr);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)TSTAGES * STAGE_BYTES);
  uint64_t* empty = full + TSTAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bi = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) bi++;
  while (bi * (bi + 1) / 2 > (int)blockIdx.x) bi--;
  const int bj = blockIdx.x - bi * (bi + 1) / 2;
  const bool diag = (bi == bj);
  const int d0 = bi * TBM, e0 = bj * TBN;
  const int nk_total = (K + TBK - 1) / TBK;
  const int ktiles_per_split = (nk_total + (int)gridDim.y - 1) / (int)gridDim.y;
  const int kt_begin = blockIdx.y * ktiles_per_split;
  const int kt_end = min(nk_total, kt_begin + ktiles_per_split);
  const int nk = max(0, kt_end - kt_begin);
  if (threadIdx.x == 0) {
    for (int s = 0; s < TSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    if (lane == 0) {
      for (int kt = 0; kt < nk; kt++) {
        const int s = kt % TSTAGES;
        if (kt >= TSTAGES) mbar_wait(&empty[s], ((kt / TSTAGES) - 1) & 1);
        mbar_expect_tx(&full[s], diag ? TILE_BYTES : STAGE_BYTES);
        const int k0 = (kt_begin + kt) * TBK;
        uint8_t* sa = tiles + (size_t)s * STAGE_BYTES;
#pragma unroll
        for (int cb = 0; cb < 8; cb++) tma_load_2d(sa + cb * 2048, &tmS, d0 + cb * 16, k0, &full[s]);
        if (!diag) {
#pragma unroll
          for (int cb = 0; cb < 8; cb++) tma_load_2d(sa + TILE_BYTES + cb * 2048, &tmS, e0 + cb * 16, k0, &full[s]);
        }
      }
    }
    return;
  }
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  // element (k, m) of an operand tile: box cb = m>>4, mm = m&15: cb*2048 + k*128 + (((mm>>1) ^ (k&7)) << 4) + (mm&1)*8
  // with k = 8 (s>>1) + 2 t + (s&1) for MMA step s.  m = base + 8 i + g  ->  cb = base/16 + (i>>1), mm = 8 (i&1) + g.
  int koff[4][2];   // [s][i&1] : k*128 + swizzled chunk + half, without the box offset
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int k = 8 * (s >> 1) + 2 * t + (s & 1);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int mm = 8 * h + g;
      koff[s][h] = k * 128 + ((((mm >> 1) ^ (k & 7))) << 4) + (mm & 1) * 8;
    }
  }
  const int a_box = (wm * 4) * 2048, b_box = (wn * 2) * 2048;

  for (int kt = 0; kt < nk; kt++) {
    const int s = kt % TSTAGES;
    mbar_wait(&full[s], (kt / TSTAGES) & 1);
    const uint8_t* sa = tiles + (size_t)s * STAGE_BYTES + a_box;
    const uint8_t* sb = tiles + (size_t)s * STAGE_BYTES + (diag ? 0 : TILE_BYTES) + b_box;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const double*>(sa + (i >> 1) * 2048 + koff[kk][i & 1]);
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = *reinterpret_cast<const double*>(sb + (j >> 1) * 2048 + koff[kk][j & 1]);
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  double* Wp = W + (size_t)blockIdx.y * n * ldw;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int row = d0 + wm * 64 + i * 8 + g;
    if (row >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int col = e0 + wn * 32 + j * 8 + 2 * t;
      store_pair2(Wp + (size_t)row * ldw + col, col < n, col + 1 < n, acc[i][j][0], acc[i][j][1]);
    }
  }
}


// -----------------------------------------------------------------------------------------------------------------
// syrk_sk_tma_kernel — the rank-mu product with BALANCED work: stream-K over the lower-triangular tiles.
//   * a diagonal tile needs only its lower triangle: its 16 x 16 grid of 8x8 DMMA blocks is dealt to the eight consumer warps as
//     six full 32x32 squares (16 blocks each) and two pairs of 32x32 diagonal triangles (2 x 10 blocks), so a k-step of a
//     diagonal tile costs 20/32 of a full one (and loads one operand tile instead of two);
//   * the (tile, k) iteration space, weighted 8 : 5, is cut into gridDim.x equal contiguous spans, one per CTA (one CTA per
//     SM); a span that crosses a tile boundary gives two (or more) partial tiles. Part p of a tile goes to slab p of W
//     (p = CTA index - index of the CTA that holds the tile's first k-step), the slabs are zeroed by the launcher and summed in
//     slab order by adapt_c / reduce_splits: a fixed order, no atomics.
// The split-K predecessor (syrk_tt_tma_kernel) ran 36 tiles x 4 splits = 144 equal CTAs at N = 1000 and spent 15 % of its DMMAs on
// the upper halves of the diagonal tiles.
// -----------------------------------------------------------------------------------------------------------------
namespace {
constexpr int SK_WFULL = 8, SK_WDIAG = 5;
struct SkWalk {      // position in the weighted (tile, k) space
  int nt, bi, bj;    // tile rows; current tile
  long long cum;     // weight position of the current tile's first k-step
  long long nk;
  __device__ __forceinline__ int w() const { return bi == bj ? SK_WDIAG : SK_WFULL; }
  __device__ __forceinline__ long long tile_end() const { return cum + (long long)w() * nk; }
  __device__ __forceinline__ void next() { cum = tile_end(); if (bj == bi) { bi++; bj = 0; } else bj++; }
  __device__ __forceinline__ void seek(long long pos) { while (bi < nt && tile_end() <= pos) next(); }
};
__device__ __forceinline__ long long sk_bound(long long c, long long total, long long G) { return c * total / G; }
}  // namespace

// Byte offset of the fragment element for MMA step s (0..3) and row half h (0, 1) from the lane's base offset (s = 0, h = 0):
// k = 8 (s>>1) + 2 t + (s&1) and m = 8 h + g give k*128 + (((m>>1) ^ (k&7)) << 4) + (m&1)*8; the terms of s and h touch disjoint
// bits, so the eight offsets of syrk_tt_tma_kernel's table follow from one register.
__device__ __forceinline__ int sk_off(int base, int s, int h) { return ((base + (s & 1) * 128) ^ ((s & 1) * 16) ^ (h * 64)) + (s >> 1) * 1024; }

// One segment (k-steps [k_lo, k_hi) of one tile) for one consumer warp in one role: accumulate, then store the warp's blocks into slab Wp.
// rb0 / cb0: the warp's first row / column inside the tile (role 2: first row = first column of its first diagonal square).
template <int ROLE>
__device__ __forceinline__ void sk_segment(const uint8_t* tiles, uint64_t* full, uint64_t* empty, int& it, int k_lo, int k_hi, int a_off, int b_off,
                                           int base, int lane, double* __restrict__ Wp, int ldw, int n, int row0, int col0) {
  constexpr int RI = ROLE == 1 ? 4 : 8;
  double acc[RI][4][2];
#pragma unroll
  for (int i = 0; i < RI; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int kt = k_lo; kt < k_hi; kt++, it++) {
    const int s = it % TSTAGES;
    mbar_wait(&full[s], (it / TSTAGES) & 1);
    const uint8_t* sa = tiles + (size_t)s * STAGE_BYTES + a_off;
    const uint8_t* sb = tiles + (size_t)s * STAGE_BYTES + b_off;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      if (ROLE == 0) {
        double a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const double*>(sa + (i >> 1) * 2048 + sk_off(base, kk, i & 1));
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = *reinterpret_cast<const double*>(sb + (j >> 1) * 2048 + sk_off(base, kk, j & 1));
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      } else if (ROLE == 1) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const double*>(sa + (i >> 1) * 2048 + sk_off(base, kk, i & 1));
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = *reinterpret_cast<const double*>(sb + (j >> 1) * 2048 + sk_off(base, kk, j & 1));
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      } else {
#pragma unroll
        for (int q = 0; q < 2; q++) {   // rows and columns of a diagonal square are the same 32 matrix columns: a[j] serves as the B fragment
          double a[4];
#pragma unroll
          for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const double*>(sa + (2 * q + (i >> 1)) * 2048 + sk_off(base, kk, i & 1));
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j <= i; j++) dmma884(acc[q * 4 + i][j][0], acc[q * 4 + i][j][1], a[i], a[j]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < RI; i++) {
    const int sq = ROLE == 2 ? (i >> 2) * 32 : 0;       // role 2: second diagonal square, 32 rows and columns further
    const int row = row0 + sq + (ROLE == 2 ? (i & 3) : i) * 8 + g;
    if (row >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (ROLE == 2 && j > (i & 3)) continue;
      const int col = col0 + sq + j * 8 + 2 * t;
      store_pair2(Wp + (size_t)row * ldw + col, col < n, col + 1 < n, acc[i][j][0], acc[i][j][1]);
    }
  }
}

constexpr int SK_MAXSEG = 64;

__global__ void __launch_bounds__(TTHREADS, 1)
syrk_sk_tma_kernel(const __grid_constant__ CUtensorMap tmS, int n, int K, const int* __restrict__ kptr, double* __restrict__ W, int ldw) {
  if (kptr) K = min(K, *kptr);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)TSTAGES * STAGE_BYTES);
  uint64_t* empty = full + TSTAGES;
  __shared__ int seg[SK_MAXSEG][5];   // {tile row, tile column, first k-step, end k-step, slab}
  __shared__ int nseg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    const int nt = (n + TBM - 1) / TBM;
    const long long nk = (K + TBK - 1) / TBK;
    const long long G = gridDim.x, c = blockIdx.x;
    const long long total = nk * ((long long)SK_WFULL * (nt * (nt - 1) / 2) + (long long)SK_WDIAG * nt);
    const long long b0 = sk_bound(c, total, G), b1 = sk_bound(c + 1, total, G);
    int ns = 0;
    if (b1 > b0) {
      SkWalk wk{nt, 0, 0, 0, nk};
      wk.seek(b0);
      for (long long pos = b0; pos < b1 && wk.bi < nt && ns < SK_MAXSEG; wk.next(), pos = wk.cum) {
        const int wgt = wk.w();
        const long long k_lo = (pos - wk.cum) / wgt, k_hi = b1 >= wk.tile_end() ? nk : (b1 - wk.cum) / wgt;
        if (k_hi <= k_lo) continue;
        // slab of this part: CTA index minus the index of the CTA whose span holds the tile's first k-step
        long long cf = wk.cum * G / total;
        while (sk_bound(cf + 1, total, G) <= wk.cum) cf++;
        while (sk_bound(cf, total, G) > wk.cum) cf--;
        seg[ns][0] = wk.bi; seg[ns][1] = wk.bj; seg[ns][2] = (int)k_lo; seg[ns][3] = (int)k_hi; seg[ns][4] = (int)(c - cf);
        ns++;
      }
    }
    nseg = ns;
    for (int s = 0; s < TSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ns = nseg;

  if (warp == CONSUMER_WARPS) {
    if (lane == 0) {
      int it = 0;
      for (int q = 0; q < ns; q++) {
        const int bi = seg[q][0], bj = seg[q][1], k_lo = seg[q][2], k_hi = seg[q][3];
        const bool diag = bi == bj;
        const int d0 = bi * TBM, e0 = bj * TBN;
        for (int kt = k_lo; kt < k_hi; kt++, it++) {
          const int s = it % TSTAGES;
          if (it >= TSTAGES) mbar_wait(&empty[s], ((it / TSTAGES) - 1) & 1);
          mbar_expect_tx(&full[s], diag ? TILE_BYTES : STAGE_BYTES);
          const int k0 = kt * TBK;
          uint8_t* sa = tiles + (size_t)s * STAGE_BYTES;
#pragma unroll
          for (int cb = 0; cb < 8; cb++) tma_load_2d(sa + cb * 2048, &tmS, d0 + cb * 16, k0, &full[s]);
          if (!diag) {
#pragma unroll
            for (int cb = 0; cb < 8; cb++) tma_load_2d(sa + TILE_BYTES + cb * 2048, &tmS, e0 + cb * 16, k0, &full[s]);
          }
        }
      }
    }
    return;
  }
  const int g = lane >> 2, t = lane & 3;
  const int base = (2 * t) * 128 + (((g >> 1) ^ (2 * t)) << 4) + (g & 1) * 8;
  // diagonal tiles: warps 0..5 take the strictly lower 32x32 squares (si, sj), warps 6 and 7 the diagonal triangles {0, 1} and {2, 3}
  const int si = warp < 1 ? 1 : warp < 3 ? 2 : 3, sj = warp < 1 ? 0 : warp < 3 ? warp - 1 : warp - 3;
  const int wm = warp >> 2, wn = warp & 3;
  int it = 0;
  for (int q = 0; q < ns; q++) {
    const int bi = seg[q][0], bj = seg[q][1], k_lo = seg[q][2], k_hi = seg[q][3];
    double* Wp = W + (size_t)seg[q][4] * n * ldw;
    const int d0 = bi * TBM, e0 = bj * TBN;
    if (bi != bj)
      sk_segment<0>(tiles, full, empty, it, k_lo, k_hi, (wm * 4) * 2048, TILE_BYTES + (wn * 2) * 2048, base, lane, Wp, ldw, n, d0 + wm * 64, e0 + wn * 32);
    else if (warp < 6)
      sk_segment<1>(tiles, full, empty, it, k_lo, k_hi, (si * 2) * 2048, (sj * 2) * 2048, base, lane, Wp, ldw, n, d0 + si * 32, e0 + sj * 32);
    else
      sk_segment<2>(tiles, full, empty, it, k_lo, k_hi, ((warp - 6) * 4) * 2048, 0, base, lane, Wp, ldw, n, d0 + (warp - 6) * 64, e0 + (warp - 6) * 64);
  }
}

// ---- host: tensor maps through the driver entry point (no -lcuda link dependency) ----------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_tma_state = 0;  // 0 unknown, 1 ok, -1 unavailable

static bool tma_ready() {
  static std::atomic<unsigned long long> attr{0};
  if (first_call_on_device(attr)) {   // the shared-memory limits are per device; the driver entry point is per process
    if (cudaFuncSetAttribute(gemm_tn_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(syrk_tt_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(syrk_sk_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM) != cudaSuccess)
      g_tma_state = -1;
  }
  if (g_tma_state) return g_tma_state > 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
    g_tma_state = -1;
    return false;
  }
  g_encode = (EncodeTiledFn)fn;
  g_tma_state = 1;
  return true;
}

// row-major matrix [rows][ld] of doubles; box = box_cols x box_rows
static bool make_map(CUtensorMap* m, const double* base, long long rows, int ld, int box_cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool launch_gemm_tn_tma(cudaStream_t st, int M, int Nc, int K, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
  if (M <= 0 || Nc <= 0) return true;
  if (!tma_ready()) return false;
  if ((lda % 2) || (ldb % 2) || ((uintptr_t)A & 15) || ((uintptr_t)B & 15)) return false;
  CUtensorMap ta, tb;
  if (!make_map(&ta, A, M, lda, TBK, TBM) || !make_map(&tb, B, Nc, ldb, TBK, TBN)) return false;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int ntiles = ((Nc + TBN - 1) / TBN) * ((M + TBM - 1) / TBM);
  gemm_tn_tma_kernel<<<ntiles < sms ? ntiles : sms, TTHREADS, TMA_SMEM, st>>>(ta, tb, M, Nc, K, C, ldc);
  return true;
}

bool launch_syrk_tt_tma(cudaStream_t st, int n, int K, const int* kptr, const double* S, int lds, long long s_rows, double* W, int ldw,
                        int splits) {
  if (!tma_ready()) return false;
  if ((lds % 2) || ((uintptr_t)S & 15)) return false;
  CUtensorMap ts;
  if (!make_map(&ts, S, s_rows, lds, 16, TBK)) return false;
  dim3 grid(syrk_tiles(n), splits);
  syrk_tt_tma_kernel<<<grid, TTHREADS, TMA_SMEM, st>>>(ts, n, K, kptr, W, ldw);
  return true;
}

// Stream-K launch: returns the number of slabs of W the parts were written to (they are zeroed here), 0 if the TMA path is unavailable.
int launch_syrk_sk_tma(cudaStream_t st, int n, int K, const int* kptr, const double* S, int lds, long long s_rows, double* W, int ldw,
                       int num_sms, int max_splits) {
  if (!tma_ready()) return 0;
  if ((lds % 2) || ((uintptr_t)S & 15)) return 0;
  CUtensorMap ts;
  if (!make_map(&ts, S, s_rows, lds, 16, TBK)) return 0;
  const int nt = (n + TBM - 1) / TBM;
  const long long sw = (long long)SK_WFULL * (nt * (nt - 1) / 2) + (long long)SK_WDIAG * nt;
  const long long nk = (K + TBK - 1) / TBK;
  const int wmax = nt > 1 ? SK_WFULL : SK_WDIAG;
  long long G = num_sms;
  G = std::min<long long>(G, std::max<long long>(1, sw * nk / 64));             // at least ~8 full k-steps per CTA
  G = std::min<long long>(G, std::max<long long>(1, (long long)(max_splits - 1) * sw / wmax));   // parts of a tile <= max_splits
  const int parts = (int)std::min<long long>(max_splits, (wmax * G + sw - 1) / sw + 1);
  if ((long long)nt * (nt + 1) / 2 / G + 3 > SK_MAXSEG) return 0;   // more tiles per CTA than the segment table holds
  cudaMemsetAsync(W, 0, sizeof(double) * (size_t)parts * n * ldw, st);
  syrk_sk_tma_kernel<<<(unsigned)G, TTHREADS, TMA_SMEM, st>>>(ts, n, K, kptr, W, ldw);
  return parts;
}

}  // namespace kc
