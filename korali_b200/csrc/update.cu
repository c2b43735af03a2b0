// update.cu — K5 / K7 and the scalar bookkeeping of updateDistribution (CMAES.cpp.base:547-688), adaptC
// (:690-718), updateSigma (:720-761), numericalErrorTreatment (:763-772), updateViabilityBoundaries (:426-437).
// All scalars live in DevScalars in HBM so a generation needs no host round trip.
// Compiled with --fmad=false: products and sums round like the reference's scalar code.
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace kc {

// A[d][e] = B[d][e] * D[e]  (operand of the sampling GEMM; sampleSingle applies D to z first, :504)
__global__ void __launch_bounds__(256) scale_bd_kernel(const double* __restrict__ B, int ldb, const double* __restrict__ D,
                                                       double* __restrict__ A, int lda, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = blockIdx.y;
  if (e < n) A[(size_t)d * lda + e] = B[(size_t)d * ldb + e] * D[e];
}

// ---- ranking bookkeeping (:552-566, :743) ---------------------------------------------------------
// Single block. viol (nullable) = per-sample constraint violation counts (global sample order).
__global__ void __launch_bounds__(256)
rank_bookkeeping_kernel(const double* __restrict__ f, const unsigned* __restrict__ idx, int lambda, int mu,
                        const unsigned long long* __restrict__ viol, int best_is_first, DevScalars* __restrict__ sc) {
  __shared__ int best_rank;
  if (threadIdx.x == 0) best_rank = best_is_first ? 0 : -1;
  __syncthreads();
  if (!best_is_first) {
    // reference loop has no break: it ends on the LOWEST-ranked sample without violations
    int local = -1;
    for (int r = threadIdx.x; r < lambda; r += blockDim.x)
      if (viol[idx[r]] == 0) local = max(local, r);
    atomicMax(&best_rank, local);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sc->previous_best_value = sc->current_best_value;
    // No sample without violations: the reference indexes _valueVector[-1] here (undefined behaviour, CMAES.cpp.base:556-565).
    // Defined behaviour instead: fall back to the best-ranked sample and raise a warning.
    if (best_rank < 0) { best_rank = 0; sc->warn_no_valid = 1; }
    const unsigned s = idx[best_rank];
    sc->best_valid_sample = s;
    sc->current_best_value = f[s];
    sc->value_at_mu = f[idx[mu - 1]];
  }
}

// Mu Type "Proportional" (:584-600): w_i = f_i / sum_{top mu} f.
__global__ void __launch_bounds__(1024)
proportional_weights_kernel(const double* __restrict__ f, const unsigned* __restrict__ idx, int mu, double* __restrict__ w) {
  __shared__ double total;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < mu; i++) s += f[idx[i]];  // rank order, as the reference
    total = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < mu; i += blockDim.x) w[i] = f[idx[i]] / total;
}

// Selection list of THIS rank: every rank r < mu whose sample idx[r] lies in [lo, hi) -> (local sample, weight), in rank order.
// Block b compacts ranks [1024 b, 1024 (b+1)); its output offset = number of selected ranks in front of its chunk, which it counts
// itself (b block-wide counts over the L2-resident index array) — deterministic, no inter-block communication. The single-block
// predecessor walked all mu ranks with three barriers per 1024 (47 us at mu = 32768).
__global__ void __launch_bounds__(1024)
select_local_kernel(const unsigned* __restrict__ idx, const double* __restrict__ w, int mu, unsigned lo, unsigned hi,
                    int* __restrict__ sel_sample, double* __restrict__ sel_weight, int* __restrict__ count_out) {
  __shared__ int warp_tot[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int before = 0;
  for (int c = 0; c < (int)blockIdx.x; c++) {
    const unsigned s = idx[c * 1024 + threadIdx.x];     // chunks in front of this block are full
    before += __syncthreads_count(s >= lo && s < hi);
  }
  const int r = blockIdx.x * 1024 + threadIdx.x;
  unsigned s = 0;
  bool take = false;
  if (r < mu) { s = idx[r]; take = (s >= lo && s < hi); }
  const unsigned m = __ballot_sync(0xffffffffu, take);
  if (lane == 0) warp_tot[warp] = __popc(m);
  __syncthreads();
  if (warp == 0) {
    int x = warp_tot[lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += y;
    }
    warp_tot[lane] = x;
  }
  __syncthreads();
  if (take) {
    const int pos = before + (warp ? warp_tot[warp - 1] : 0) + __popc(m & ((1u << lane) - 1u));
    sel_sample[pos] = (int)(s - lo);
    sel_weight[pos] = w[r];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *count_out = before + warp_tot[31];
}

// ---- K5: gather selected rows, weighted mean partials, and the rank-mu operand ------------------------
// For selected entry j: sample s, weight w: x = m + (+-sigma) y_s ; t = x - m ; S[j][:] = sqrt(w) * t ;
// partial[cta][:] += w * x  (mean, :603-609).  from_x: x is read from the injected "Sample Population".
constexpr int GM_ROWS_MAX = 64;   // rows_per_cta of the callers
constexpr int GM_UNROLL = 8;      // rows in flight per thread (the kernel is a 262 MB read + 262 MB write stream at config 3)
__global__ void __launch_bounds__(256)
gather_mean_kernel(const double* __restrict__ Y, int ldy, int mirrored, int from_x, const int* __restrict__ sel_sample,
                   const double* __restrict__ sel_weight, const int* __restrict__ count_ptr, int rows_per_cta, int n, int ld,
                   const double* __restrict__ mean, const DevScalars* __restrict__ sc, double* __restrict__ S, int lds,
                   double* __restrict__ partial) {
  __shared__ long long row_off[GM_ROWS_MAX];   // element offset of the sample's row in Y
  __shared__ double row_w[GM_ROWS_MAX], row_rw[GM_ROWS_MAX], row_ss[GM_ROWS_MAX];
  const int count = *count_ptr;
  const int j0 = blockIdx.x * rows_per_cta;
  const int j1 = min(count, j0 + rows_per_cta);
  const int nr = j1 - j0;
  const double sigma = sc->sigma;
  for (int r = threadIdx.x; r < nr; r += blockDim.x) {   // per-row quantities once per CTA instead of once per thread
    const int s = sel_sample[j0 + r];
    const double w = sel_weight[j0 + r];
    row_off[r] = (long long)((from_x || !mirrored) ? s : (s >> 1)) * ldy;
    row_w[r] = w;
    row_rw[r] = sqrt(fabs(w));   // Proportional weights can be negative: the sign is applied by signed_rank_mu_kernel
    row_ss[r] = (mirrored && (s & 1)) ? -sigma : sigma;
  }
  __syncthreads();
  for (int c = threadIdx.x; 2 * c < ld; c += blockDim.x) {
    const int d = 2 * c;
    const bool in0 = d < n, in1 = d + 1 < n;
    const double m0 = in0 ? mean[d] : 0.0, m1 = in1 ? mean[d + 1] : 0.0;
    double a0 = 0.0, a1 = 0.0;
    for (int r0 = 0; r0 < nr; r0 += GM_UNROLL) {
      double2 v[GM_UNROLL];
#pragma unroll
      for (int u = 0; u < GM_UNROLL; u++)
        if (r0 + u < nr) v[u] = __ldcs(reinterpret_cast<const double2*>(Y + row_off[r0 + u] + d));
#pragma unroll
      for (int u = 0; u < GM_UNROLL; u++) {
        const int r = r0 + u;
        if (r < nr) {
          double x0, x1;
          if (from_x) { x0 = v[u].x; x1 = v[u].y; }
          else { const double ss = row_ss[r]; x0 = m0 + ss * v[u].x; x1 = m1 + ss * v[u].y; }
          const double rw = row_rw[r], w = row_w[r];
          const double t0 = in0 ? rw * (x0 - m0) : 0.0, t1 = in1 ? rw * (x1 - m1) : 0.0;
          *reinterpret_cast<double2*>(S + (size_t)(j0 + r) * lds + d) = make_double2(t0, t1);
          a0 += w * x0; a1 += w * x1;      // in row order: the partial mean does not depend on the unrolling
        }
      }
    }
    *reinterpret_cast<double2*>(partial + (size_t)blockIdx.x * ld + d) = make_double2(in0 ? a0 : 0.0, in1 ? a1 : 0.0);
  }
}

// Rows [count, rows_padded) of S must be zero for the k-tail of the SYRK.
__global__ void __launch_bounds__(256)
zero_tail_kernel(double* __restrict__ S, int lds, const int* __restrict__ count_ptr, int rows_padded) {
  const int count = *count_ptr;
  const long long total = (long long)(rows_padded - count) * lds;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    S[(size_t)count * lds + i] = 0.0;
}

// mean_out[d] = sum over CTAs (fixed order); x_best[d] = variables of the current best sample if it is local.
__global__ void __launch_bounds__(256)
mean_reduce_kernel(const double* __restrict__ partial, const int* __restrict__ count_ptr, int rows_per_cta, int n, int ld,
                   double* __restrict__ mean_out, const double* __restrict__ Y, int ldy, int mirrored, int from_x,
                   const double* __restrict__ mean, const DevScalars* __restrict__ sc, unsigned lo, unsigned hi,
                   double* __restrict__ best_x) {
  // block = 32 dimensions x 8 chunks of the partial list: a chunk is summed in order (16 loads in flight), the 8 chunk sums in order
  __shared__ double red[8][33];
  const int dx = threadIdx.x & 31, cy = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + dx;
  const int nparts = (*count_ptr + rows_per_cta - 1) / rows_per_cta;
  const int chunk = (nparts + 7) / 8;
  const int p1 = min(nparts, (cy + 1) * chunk);
  double a = 0.0;
  if (d < n) {
    int p = cy * chunk;
    for (; p + 16 <= p1; p += 16) {
      double v[16];
#pragma unroll
      for (int k = 0; k < 16; k++) v[k] = partial[(size_t)(p + k) * ld + d];
#pragma unroll
      for (int k = 0; k < 16; k++) a += v[k];
    }
    for (; p < p1; p++) a += partial[(size_t)p * ld + d];
  }
  red[cy][dx] = a;
  __syncthreads();
  if (cy != 0 || d >= n) return;
  a = red[0][dx];
#pragma unroll
  for (int k = 1; k < 8; k++) a += red[k][dx];
  mean_out[d] = a;
  const unsigned long long b = sc->best_valid_sample;
  double bx = 0.0;
  if (b >= lo && b < hi) {
    const unsigned s = (unsigned)(b - lo);
    if (from_x) bx = Y[(size_t)s * ldy + d];
    else {
      const double ss = (mirrored && (s & 1)) ? -sc->sigma : sc->sigma;
      bx = mean[d] + ss * Y[(size_t)(mirrored ? (s >> 1) : s) * ldy + d];
    }
  }
  best_x[d] = bx;
}

// ---- best-ever bookkeeping (:567-581) ------------------------------------------------------------------
__global__ void __launch_bounds__(256)
best_update_kernel(const double* __restrict__ best_x, int n, unsigned generation, double* __restrict__ cur_best_vars,
                   double* __restrict__ best_ever_vars, DevScalars* __restrict__ sc, const double* __restrict__ con_evals,
                   long long ldg, int n_con, double* __restrict__ best_con_evals) {
  __shared__ int upd;
  if (threadIdx.x == 0) {
    if (generation == kGenFromDevice) generation = (unsigned)sc->gen;   // CUDA-graph replay
    upd = (sc->current_best_value > sc->best_ever_value) || generation == 1;
    sc->best_updated = upd;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < n; d += blockDim.x) {
    const double v = best_x[d];
    cur_best_vars[d] = v;
    if (upd) best_ever_vars[d] = v;
  }
  if (upd && con_evals)
    for (int c = threadIdx.x; c < n_con; c += blockDim.x) best_con_evals[c] = con_evals[(size_t)c * ldg + sc->best_valid_sample];
  __syncthreads();
  if (threadIdx.x == 0 && upd) {
    sc->previous_best_ever_value = sc->best_ever_value;
    sc->best_ever_value = sc->current_best_value;
  }
}

// ---- K7: mean / evolution paths (:603-662) ------------------------------------------------------------
// previous mean <- mean ; mean <- new ; y = (mean - previous)/sigma
__global__ void __launch_bounds__(256)
mean_step_kernel(const double* __restrict__ mean_new, double* __restrict__ mean, double* __restrict__ mean_old,
                 double* __restrict__ y, int n, const DevScalars* __restrict__ sc) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const double mo = mean[d], mn = mean_new[d];
  mean_old[d] = mo;
  mean[d] = mn;
  y[d] = (mn - mo) / sc->sigma;
}

// t[d] = (sum_e B[e][d] * y[e]) / D[d]      (z := D^-1 B^T y, :627-636). block = (32 columns, 32 e-slices)
constexpr int BTY_SLICES = 32;
__global__ void __launch_bounds__(32 * BTY_SLICES)
bt_y_kernel(const double* __restrict__ B, int ldb, const double* __restrict__ y, const double* __restrict__ D,
            double* __restrict__ tvec, int n, int diagonal) {
  __shared__ double red[BTY_SLICES][33];
  const int cx = threadIdx.x & 31, ey = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + cx;
  double a = 0.0;
  if (d < n) {
    if (diagonal) { if (ey == 0) a = y[d]; }
    else for (int e = ey; e < n; e += BTY_SLICES) a += B[(size_t)e * ldb + d] * y[e];
  }
  red[ey][cx] = a;
  __syncthreads();
  if (ey == 0 && d < n) {
    double s = red[0][cx];
#pragma unroll
    for (int k = 1; k < BTY_SLICES; k++) s += red[k][cx];
    tvec[d] = s / D[d];
  }
}

// ps[d] = (1-cs) ps[d] + sqrt(cs (2-cs) mueff) * (sum_e B[d][e] t[e])   (:641-653). One warp per row.
__global__ void __launch_bounds__(256)
b_t_ps_kernel(const double* __restrict__ B, int ldb, const double* __restrict__ tvec, double* __restrict__ ps, int n,
              int diagonal, double cs, double mueff) {
  const int lane = threadIdx.x & 31;
  const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (d >= n) return;
  double a = 0.0;
  if (diagonal) { if (lane == 0) a = tvec[d]; }
  else for (int e = lane; e < n; e += 32) a += B[(size_t)d * ldb + e] * tvec[e];
  a = warp_sum_butterfly(a);
  if (lane == 0) ps[d] = (1. - cs) * ps[d] + sqrt(cs * (2. - cs) * mueff) * a;
}

// |ps|, hsig, pc (:638-662). Single block.
__global__ void __launch_bounds__(1024)
hsig_pc_kernel(const double* __restrict__ ps, const double* __restrict__ y, double* __restrict__ pc, int n, double cs, double cc,
               double mueff, double chi_n, unsigned generation, DevScalars* __restrict__ sc) {
  __shared__ double wsum[32];
  __shared__ double hs;
  double a = 0.0;
  for (int d = threadIdx.x; d < n; d += blockDim.x) a += ps[d] * ps[d];
  a = warp_sum_butterfly(a);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += wsum[w];
    const double norm = sqrt(s);
    sc->ps_l2norm = norm;
    if (generation == kGenFromDevice) generation = (unsigned)sc->gen;   // CUDA-graph replay
    const int hsig = (1.4 + 2.0 / (n + 1) > norm / sqrt(1. - pow(1. - cs, 2.0 * (1.0 + generation))) / chi_n);
    sc->hsig = hsig;
    hs = hsig;
  }
  __syncthreads();
  const double h = hs;
  for (int d = threadIdx.x; d < n; d += blockDim.x) pc[d] = (1. - cc) * pc[d] + h * sqrt(cc * (2. - cc) * mueff) * y[d];
}

// ---- adaptC combine (:690-707): C <- (1-c1-cmu) C + c1 (pc pc^T + (1-hsig) cc (2-cc) C) + cmu/sigma^2 * P ---
// P = sum over `splits` slabs of W (fixed order). Lower triangle computed, mirrored to the upper one.
__global__ void __launch_bounds__(256)
adapt_c_kernel(double* __restrict__ C, int ldc, const double* __restrict__ W, int ldw, int splits, int n,
               const double* __restrict__ pc, double c1, double cmu, double cc, int diagonal,
               const DevScalars* __restrict__ sc) {
  const int e = blockIdx.x * 16 + (threadIdx.x & 15);
  const int d = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (d >= n || e > d) return;
  if (diagonal && e != d) return;
  double p = 0.0;
  for (int s = 0; s < splits; s++) p += W[(size_t)s * n * ldw + (size_t)d * ldw + e];
  const double sigma = sc->sigma;
  const double hsig = sc->hsig;
  const double cold = C[(size_t)d * ldc + e];
  double cn = (1 - c1 - cmu) * cold + c1 * (pc[d] * pc[e] + (1 - hsig) * cc * (2. - cc) * cold);
  cn += cmu * p / (sigma * sigma);
  C[(size_t)d * ldc + e] = cn;
  if (e < d) C[(size_t)e * ldc + d] = cn;
}

// Sum split-K slabs into one matrix (multi-GPU path: the all-reduce input).
__global__ void __launch_bounds__(256)
reduce_splits_kernel(const double* __restrict__ W, int ldw, int splits, int n, double* __restrict__ P) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = blockIdx.y;
  if (e >= ldw) return;
  double p = 0.0;
  if (e <= d && e < n)
    for (int s = 0; s < splits; s++) p += W[(size_t)s * n * ldw + (size_t)d * ldw + e];
  P[(size_t)d * ldw + e] = p;
}

// Diagonal-covariance rank-mu term: P[d][d] = sum_j S[j][d]^2 (column sums of squares), slabs = CTAs.
__global__ void __launch_bounds__(256)
diag_rank_mu_kernel(const double* __restrict__ S, int lds, const int* __restrict__ count_ptr, int rows_per_cta, int n,
                    double* __restrict__ W, int ldw) {
  const int count = *count_ptr;
  const int j0 = blockIdx.y * rows_per_cta, j1 = min(count, j0 + rows_per_cta);
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  double a = 0.0;
  for (int j = j0; j < j1; j++) { const double v = S[(size_t)j * lds + d]; a += v * v; }
  W[(size_t)blockIdx.y * n * ldw + (size_t)d * ldw + d] = a;
}

// Mu Type "Proportional" (:584-600) can produce negative weights (F(x) of mixed sign), which the S^T S form cannot
// express. P[d][e] = sum_j sign(w_j) S[j][d] S[j][e] with S = sqrt|w| t, one thread per lower-triangular entry.
// Plain FP64 FMAs: this mode is a small-N convenience in the reference's tests, not a throughput path.
__global__ void __launch_bounds__(256)
signed_rank_mu_kernel(const double* __restrict__ S, int lds, const int* __restrict__ count_ptr, const double* __restrict__ sel_weight,
                      int n, double* __restrict__ W, int ldw) {
  const int e = blockIdx.x * 16 + (threadIdx.x & 15);
  const int d = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (d >= n || e > d) return;
  const int count = *count_ptr;
  double a = 0.0;
  for (int j = 0; j < count; j++) {
    const double v = S[(size_t)j * lds + d] * S[(size_t)j * lds + e];
    a += (sel_weight[j] < 0.0) ? -v : v;
  }
  W[(size_t)d * ldw + e] = a;
}

// ---- scalar tail: min/max diag (:709-717), viability boundaries (:426-437), updateSigma (:720-761),
// numericalErrorTreatment (:763-772), min/max standard deviation (:679-687). Single block.
__global__ void __launch_bounds__(1024)
sigma_kernel(const double* __restrict__ C, int ldc, int n, const double* __restrict__ min_sd_update, int any_min_sd,
             double cs, double damp, double chi_n, double trace, int is_sigma_bounded, int mu_value_gt1,
             int viability_regime, double global_success_lr, double target_success_rate, int has_discrete,
             DevScalars* __restrict__ sc) {
  __shared__ double smax[32], smin[32];
  double mx = -INFINITY, mn = INFINITY;
  for (int d = threadIdx.x; d < n; d += blockDim.x) {
    const double v = C[(size_t)d * ldc + d];
    mx = fmax(mx, v); mn = fmin(mn, v);
  }
  mx = warp_max(mx); mn = warp_min(mn);
  if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = mx; smin[threadIdx.x >> 5] = mn; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int w = 1; w < (int)(blockDim.x >> 5); w++) { mx = fmax(mx, smax[w]); mn = fmin(mn, smin[w]); }
  mx = fmax(mx, smax[0]); mn = fmin(mn, smin[0]);
  sc->max_diag_c = mx;
  sc->min_diag_c = mn;
  double sigma = sc->sigma;
  if (viability_regime) {
    const double gsr = (1 - global_success_lr) * sc->global_success_rate;
    sc->global_success_rate = gsr;
    sigma *= exp((gsr - (target_success_rate / (1.0 - target_success_rate)) * (1 - gsr)) / damp);
  } else if (has_discrete) {   // :730-734
    sigma *= exp(cs / damp * (sqrt(sc->disc_path_l2) / sc->chi_dm - 1.));
  } else {
    sigma *= exp(cs / damp * (sc->ps_l2norm / chi_n - 1.));
  }
  if (mu_value_gt1 && sc->current_best_value == sc->value_at_mu) {
    sigma *= exp(0.2 + cs / damp);
    sc->warn_flat = 1;
  }
  const double upper = sqrt(trace / n);
  if (sigma > upper && is_sigma_bounded) sigma = upper;
  if (any_min_sd) {
    for (int d = 0; d < n; ++d) {
      const double cdd = C[(size_t)d * ldc + d];
      if (sigma * sqrt(cdd) < min_sd_update[d]) {
        sigma = (min_sd_update[d]) / sqrt(cdd) * exp(0.05 + cs / damp);
        sc->warn_minsd = 1;
      }
    }
  }
  sc->sigma = sigma;
  sc->cur_min_sd = sigma * sqrt(mn);
  sc->cur_max_sd = sigma * sqrt(mx);
}

// updateViabilityBoundaries (:426-437). One thread per constraint (n_con is tiny).
__global__ void viability_boundaries_kernel(const double* __restrict__ G, long long ldg, int n_con, const unsigned* __restrict__ idx,
                                            int mu, double* __restrict__ bounds) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_con) return;
  double maxviolation = 0.0;
  for (int i = 0; i < mu; ++i) {
    const double g = G[(size_t)c * ldg + idx[i]];
    if (g > maxviolation) maxviolation = g;
  }
  bounds[c] = fmax(0.0, fmin(bounds[c], 0.5 * (maxviolation + bounds[c])));
}

// ---- launchers -------------------------------------------------------------------------------------------
void launch_scale_bd(cudaStream_t st, const double* B, int ldb, const double* D, double* A, int lda, int n) {
  dim3 grid((n + 255) / 256, n);
  scale_bd_kernel<<<grid, 256, 0, st>>>(B, ldb, D, A, lda, n);
}
void launch_rank_bookkeeping(cudaStream_t st, const double* f, const unsigned* idx, int lambda, int mu,
                             const unsigned long long* viol, int best_is_first, DevScalars* sc) {
  rank_bookkeeping_kernel<<<1, 256, 0, st>>>(f, idx, lambda, mu, viol, best_is_first, sc);
}
void launch_proportional_weights(cudaStream_t st, const double* f, const unsigned* idx, int mu, double* w) {
  proportional_weights_kernel<<<1, 1024, 0, st>>>(f, idx, mu, w);
}
void launch_select_local(cudaStream_t st, const unsigned* idx, const double* w, int mu, unsigned lo, unsigned hi,
                         int* sel_sample, double* sel_weight, int* count_out) {
  select_local_kernel<<<mu > 0 ? (mu + 1023) / 1024 : 1, 1024, 0, st>>>(idx, w, mu, lo, hi, sel_sample, sel_weight, count_out);
}
void launch_gather_mean(cudaStream_t st, const double* Y, int ldy, int mirrored, int from_x, const int* sel_sample,
                        const double* sel_weight, const int* count_ptr, int max_count, int rows_per_cta, int n, int ld,
                        const double* mean, const DevScalars* sc, double* S, int lds, int rows_padded, double* partial) {
  const int ctas = (max_count + rows_per_cta - 1) / rows_per_cta;
  if (rows_per_cta > GM_ROWS_MAX) { fprintf(stderr, "launch_gather_mean: rows_per_cta %d > %d\n", rows_per_cta, GM_ROWS_MAX); abort(); }
  if (ctas > 0)
    gather_mean_kernel<<<ctas, 256, 0, st>>>(Y, ldy, mirrored, from_x, sel_sample, sel_weight, count_ptr, rows_per_cta, n, ld, mean,
                                             sc, S, lds, partial);
  zero_tail_kernel<<<64, 256, 0, st>>>(S, lds, count_ptr, rows_padded);
}
void launch_mean_reduce(cudaStream_t st, const double* partial, const int* count_ptr, int rows_per_cta, int n, int ld,
                        double* mean_out, const double* Y, int ldy, int mirrored, int from_x, const double* mean,
                        const DevScalars* sc, unsigned lo, unsigned hi, double* best_x) {
  mean_reduce_kernel<<<(n + 31) / 32, 256, 0, st>>>(partial, count_ptr, rows_per_cta, n, ld, mean_out, Y, ldy, mirrored, from_x,
                                                      mean, sc, lo, hi, best_x);
}
// "Use Gradient Information" (CMAES.cpp.base:611-621): mean[d] += sum_i w_i * step / sqrt(N) * gradient[sel_i][d] over the selected
// samples this rank owns (the partial means are summed by the all-reduce). One thread per dimension, ranks in order.
__global__ void __launch_bounds__(256)
gradient_mean_kernel(const double* __restrict__ G, int ldg, const int* __restrict__ sel_sample, const double* __restrict__ sel_weight,
                     const int* __restrict__ count_ptr, int n, double step, double* __restrict__ mean_new) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  const int count = *count_ptr;
  const double scale = 1.0 / sqrt((double)n);
  double m = mean_new[d];
  for (int i = 0; i < count; i++) m += sel_weight[i] * step * scale * G[(size_t)sel_sample[i] * ldg + d];
  mean_new[d] = m;
}
void launch_gradient_mean(cudaStream_t st, const double* G, int ldg, const int* sel_sample, const double* sel_weight, const int* count_ptr,
                          int n, double step, double* mean_new) {
  gradient_mean_kernel<<<(n + 255) / 256, 256, 0, st>>>(G, ldg, sel_sample, sel_weight, count_ptr, n, step, mean_new);
}

// updateDiscreteMutationMatrix (CMAES.cpp.base:834-860) + the masked path length of updateSigma (:733). Single block.
// Uses the sigma of BEFORE updateSigma and the covariance of AFTER adaptC, like the reference (:665-675).
__global__ void __launch_bounds__(256)
discrete_matrix_kernel(const double* __restrict__ C, int ldc, int n, const double* __restrict__ gran, const double* __restrict__ ps, double cs,
                       double population_size, double* __restrict__ mask, double* __restrict__ mask_sigma, DevScalars* __restrict__ sc) {
  __shared__ int s_removed, s_mask;
  __shared__ double s_path[8];
  if (threadIdx.x == 0) { s_removed = 0; s_mask = 0; }
  __syncthreads();
  const double sigma = sc->sigma;
  int removed = 0, masked = 0;
  double path = 0.0;
  for (int d = threadIdx.x; d < n; d += blockDim.x) {
    const double sd = sigma * sqrt(C[(size_t)d * ldc + d]);
    double ms = 1.0;
    if (sd / sqrt(cs) < 0.2 * gran[d]) { ms = 0.0; removed++; }
    mask_sigma[d] = ms;
    double mk = 0.0;
    if (2.0 * sd < gran[d]) { mk = 1.0; masked++; }
    mask[d] = mk;
    path += ms * ps[d] * ps[d];
  }
  path = warp_sum_butterfly(path);
  if ((threadIdx.x & 31) == 0) s_path[threadIdx.x >> 5] = path;
  if (removed) atomicAdd(&s_removed, removed);
  if (masked) atomicAdd(&s_mask, masked);
  __syncthreads();
  if (threadIdx.x == 0) {
    double p = 0.0;
    for (int w = 0; w < 8; w++) p += s_path[w];
    const double entries = (double)(n + 1 - s_removed);   // +1 to prevent 0-ness
    sc->chi_dm = sqrt(entries) * (1. - 1. / (4. * entries) + 1. / (21. * entries * entries));
    sc->disc_path_l2 = p;
    sc->n_mask = s_mask;
    sc->n_disc_mut = (int)fmin(round(population_size / 10.0 + s_mask + 1), floor(population_size / 2.0) - 1);
  }
}
void launch_discrete_matrix(cudaStream_t st, const double* C, int ldc, int n, const double* gran, const double* ps, double cs,
                            double population_size, double* mask, double* mask_sigma, DevScalars* sc) {
  discrete_matrix_kernel<<<1, 256, 0, st>>>(C, ldc, n, gran, ps, cs, population_size, mask, mask_sigma, sc);
}

// isSampleFeasible on materialised samples (after the discrete mutations X is no longer mean + sigma * y).
__global__ void __launch_bounds__(256)
feasibility_x_kernel(const double* __restrict__ X, int ldx, long long samples, int n, const double* __restrict__ lower,
                     const double* __restrict__ upper, unsigned char* __restrict__ infeasible) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long s = warp; s < samples; s += nwarps) {
    bool bad = false;
    for (int d = lane; d < n; d += 32) {
      const double v = X[(size_t)s * ldx + d];
      bad |= !isfinite(v) || v < lower[d] || v > upper[d];
    }
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) infeasible[s] = m ? 1 : 0;
  }
}
void launch_feasibility_x(cudaStream_t st, const double* X, int ldx, long long samples, int n, const double* lower, const double* upper,
                          unsigned char* infeasible, int num_sms) {
  if (samples <= 0) return;
  long long blocks = (samples + 7) / 8;
  if (blocks > (long long)num_sms * 8) blocks = (long long)num_sms * 8;
  feasibility_x_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, ldx, samples, n, lower, upper, infeasible);
}

void launch_best_update(cudaStream_t st, const double* best_x, int n, unsigned generation, double* cur_best_vars,
                        double* best_ever_vars, DevScalars* sc, const double* con_evals, long long ldg, int n_con,
                        double* best_con_evals) {
  best_update_kernel<<<1, 256, 0, st>>>(best_x, n, generation, cur_best_vars, best_ever_vars, sc, con_evals, ldg, n_con, best_con_evals);
}
void launch_paths(cudaStream_t st, const double* mean_new, double* mean, double* mean_old, double* y, double* tvec,
                  double* ps, double* pc, const double* B, int ldb, const double* D, int n, int diagonal, double cs,
                  double cc, double mueff, double chi_n, unsigned generation, DevScalars* sc) {
  mean_step_kernel<<<(n + 255) / 256, 256, 0, st>>>(mean_new, mean, mean_old, y, n, sc);
  bt_y_kernel<<<(n + 31) / 32, 32 * BTY_SLICES, 0, st>>>(B, ldb, y, D, tvec, n, diagonal);
  b_t_ps_kernel<<<(n + 7) / 8, 256, 0, st>>>(B, ldb, tvec, ps, n, diagonal, cs, mueff);
  hsig_pc_kernel<<<1, 1024, 0, st>>>(ps, y, pc, n, cs, cc, mueff, chi_n, generation, sc);
}
void launch_adapt_c(cudaStream_t st, double* C, int ldc, const double* W, int ldw, int splits, int n, const double* pc,
                    double c1, double cmu, double cc, int diagonal, const DevScalars* sc) {
  dim3 grid((n + 15) / 16, (n + 15) / 16);
  adapt_c_kernel<<<grid, 256, 0, st>>>(C, ldc, W, ldw, splits, n, pc, c1, cmu, cc, diagonal, sc);
}
void launch_reduce_splits(cudaStream_t st, const double* W, int ldw, int splits, int n, double* P) {
  dim3 grid((ldw + 255) / 256, n);
  reduce_splits_kernel<<<grid, 256, 0, st>>>(W, ldw, splits, n, P);
}
void launch_diag_rank_mu(cudaStream_t st, const double* S, int lds, const int* count_ptr, int max_count, int rows_per_cta,
                         int n, double* W, int ldw, int slabs) {
  dim3 grid((n + 255) / 256, slabs);
  diag_rank_mu_kernel<<<grid, 256, 0, st>>>(S, lds, count_ptr, rows_per_cta, n, W, ldw);
}
void launch_signed_rank_mu(cudaStream_t st, const double* S, int lds, const int* count_ptr, const double* sel_weight, int n, double* W,
                           int ldw) {
  dim3 grid((n + 15) / 16, (n + 15) / 16);
  signed_rank_mu_kernel<<<grid, 256, 0, st>>>(S, lds, count_ptr, sel_weight, n, W, ldw);
}
void launch_sigma(cudaStream_t st, const double* C, int ldc, int n, const double* min_sd_update, int any_min_sd, double cs,
                  double damp, double chi_n, double trace, int is_sigma_bounded, int mu_value_gt1, int viability_regime,
                  double global_success_lr, double target_success_rate, int has_discrete, DevScalars* sc) {
  sigma_kernel<<<1, 1024, 0, st>>>(C, ldc, n, min_sd_update, any_min_sd, cs, damp, chi_n, trace, is_sigma_bounded, mu_value_gt1,
                                   viability_regime, global_success_lr, target_success_rate, has_discrete, sc);
}
void launch_viability_boundaries(cudaStream_t st, const double* G, long long ldg, int n_con, const unsigned* idx, int mu,
                                 double* bounds) {
  viability_boundaries_kernel<<<(n_con + 63) / 64, 64, 0, st>>>(G, ldg, n_con, idx, mu, bounds);
}

}  // namespace kc
