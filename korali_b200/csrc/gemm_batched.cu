// gemm_batched.cu — descriptor-driven FP64 tensor-core GEMM for the tridiagonalisation-based eigensolver (tridiag.cu, dc.cu).
//
//   C[M x Nc] = beta * C + alpha * A[M x K] * B[Nc x K]^T         (both operands K-contiguous, like gemm_tn_kernel)
//
// One launch covers a whole batch (all merges of a divide & conquer level, all panels of the back-transform): blockIdx.z picks
// the descriptor (and the k-split), blockIdx.x/y the 128 x 128 tile; CTAs outside a descriptor's tile range exit. Extras the
// eigensolver needs:
//   * split-K with plain slabs (deterministic second-stage sum by reduce_slabs_kernel) for the skinny products W = X^T V;
//   * a block-diagonal k-range: the eigenvector matrix of a merge is blockdiag(Q1, Q2) unless a deflating rotation paired a
//     column of Q1 with one of Q2 (device flag `mixed`), so a tile whose rows of the block-diagonal operand lie inside one
//     block only walks that block's k-range (half the flops of the merge GEMMs).
// Same DMMA.8x8x4 warp tiling and cp.async staging as gemm.cu (version 1 of K2): 8 warps (2 x 4), warp tile 64 x 32, BK = 16,
// 4 stages, padded rows (stride == 32 B mod 128 B: conflict-free LDS.64). Requirements: even lda/ldb/ldc and even k offsets
// (16-byte cp.async), operand pads beyond K readable and zero in at least one operand when K is odd.
#include "common.cuh"
#include "kernels.h"

namespace kc {

namespace {
constexpr int BM = 128, BN = 128, BK = 16;
constexpr int THREADS = 256;
constexpr int STAGES = 4;
constexpr int PADK = BK + 4;
}  // namespace

size_t gemm_batched_smem_bytes() { return sizeof(double) * STAGES * (BM + BN) * PADK; }

__global__ void __launch_bounds__(THREADS, 1)
gemm_tn_batched_kernel(const GemmDesc* __restrict__ descs, int max_splits) {
  const int bz = blockIdx.z / max_splits, split = blockIdx.z % max_splits;
  const GemmDesc ds = descs[bz];
  const int M = ds.M, Nc = ds.Nc;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  if (row0 >= M || col0 >= Nc || split >= ds.splits) return;
  // k-range of this CTA
  int kb = 0, ke = ds.K;
  if (ds.krule && !(ds.mixed && *ds.mixed)) {
    const int t0 = ds.krule == 1 ? row0 : col0, t1 = min(t0 + BM, ds.krule == 1 ? M : Nc);
    if (t1 <= ds.n1) ke = ds.n1;           // tile inside the first diagonal block
    else if (t0 >= ds.n1) kb = ds.n1;      // inside the second
  }
  if (ds.splits > 1) {
    const int nkt = (ke - kb + BK - 1) / BK, per = (nkt + ds.splits - 1) / ds.splits;
    const int b0 = kb + split * per * BK;
    ke = min(ke, b0 + per * BK);
    kb = b0;
  }
  const double* __restrict__ A = ds.A;
  const double* __restrict__ B = ds.B;
  double* __restrict__ C = ds.C + (size_t)split * ds.split_stride;
  const int lda = ds.lda, ldb = ds.ldb, ldc = ds.ldc;

  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * BM * PADK;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int nk = ke > kb ? (ke - kb + BK - 1) / BK : 0;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kb + kt * BK;
    double* as = As + stage * BM * PADK;
    double* bs = Bs + stage * BN * PADK;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int idx = tid + i * THREADS;
      const int r = idx >> 3, ch = idx & 7;
      const int gk = k0 + 2 * ch;
      {
        const int gr = row0 + r;
        const bool ok = (gr < M) && (gk < ke);
        const double* src = ok ? (A + (size_t)gr * lda + gk) : A;
        cp_async16(as + r * PADK + 2 * ch, src, ok ? 16 : 0);
      }
      {
        const int gr = col0 + r;
        const bool ok = (gr < Nc) && (gk < ke);
        const double* src = ok ? (B + (size_t)gr * ldb + gk) : B;
        cp_async16(bs + r * PADK + 2 * ch, src, ok ? 16 : 0);
      }
    }
  };

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    const double* as = As + (kt % STAGES) * BM * PADK + (wm * 64 + g) * PADK + t;
    const double* bs = Bs + (kt % STAGES) * BN * PADK + (wn * 32 + g) * PADK + t;
#pragma unroll
    for (int kk = 0; kk < BK / 4; kk++) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = as[i * 8 * PADK + kk * 4];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = bs[j * 8 * PADK + kk * 4];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  const double alpha = ds.alpha, beta = (ds.splits > 1) ? 0.0 : ds.beta;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int row = row0 + wm * 64 + i * 8 + g;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int col = col0 + wn * 32 + j * 8 + 2 * t;
      double* p = C + (size_t)row * ldc + col;
      if (col + 1 < Nc) {
        double2 o = make_double2(alpha * acc[i][j][0], alpha * acc[i][j][1]);
        if (beta != 0.0) { const double2 c = *reinterpret_cast<const double2*>(p); o.x += beta * c.x; o.y += beta * c.y; }
        *reinterpret_cast<double2*>(p) = o;
      } else if (col < Nc) {
        double o = alpha * acc[i][j][0];
        if (beta != 0.0) o += beta * p[0];
        p[0] = o;
      }
    }
  }
}

// out[r][c] = sum_s slabs[s][r][c] (fixed order: bitwise reproducible), rows x cols with leading dimension ld.
__global__ void __launch_bounds__(256)
reduce_slabs_kernel(const double* __restrict__ slabs, long long stride, int splits, int rows, int cols, int ld, double* __restrict__ out) {
  const long long total = (long long)rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if ((int)(i % ld) >= cols) { out[i] = 0.0; continue; }
    double s = 0.0;
    for (int k = 0; k < splits; k++) s += slabs[(size_t)k * stride + i];
    out[i] = s;
  }
}

void launch_gemm_batched(cudaStream_t st, const GemmDesc* d_descs, int batch, int max_m, int max_nc, int max_splits) {
  static std::atomic<unsigned long long> attr{0};
  if (first_call_on_device(attr)) cudaFuncSetAttribute(gemm_tn_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_batched_smem_bytes());
  if (batch <= 0 || max_m <= 0 || max_nc <= 0) return;
  dim3 grid((max_nc + BN - 1) / BN, (max_m + BM - 1) / BM, batch * max_splits);
  gemm_tn_batched_kernel<<<grid, THREADS, gemm_batched_smem_bytes(), st>>>(d_descs, max_splits);
}

void launch_reduce_slabs(cudaStream_t st, const double* slabs, long long stride, int splits, int rows, int cols, int ld, double* out,
                         int num_sms) {
  reduce_slabs_kernel<<<num_sms * 2, 256, 0, st>>>(slabs, stride, splits, rows, cols, ld, out);
}

}  // namespace kc
