// constraints.cu — K9: the viability-regime constraint path of (C)CMA-ES:
//   checkMeanAndSetRegime (CMAES.cpp.base:315-345), updateConstraints (:347-385), handleConstraints (:774-832),
//   reEvaluateConstraints (:387-424).
// The reference walks samples sequentially; here every loop over samples is a kernel, and the one truly
// order-dependent piece — the exponential moving average of each constraint's normal vector over the violating
// samples in ascending order — is kept sequential per (constraint, dimension) thread, so v_c reproduces the
// reference's arithmetic exactly; the rank-1 corrections C_aux -= beta^2 v v^T / (|v|^2 cnt^2) are then applied
// together as one SYRK on the DMMA pipe (summation order differs: tolerance, not bit parity).
// Compiled with --fmad=false.
#include "common.cuh"
#include "kernels.h"

namespace kc {

// g_c(mean) for the built-in half-space family; sets sc->mean_feasible.
__global__ void constraints_mean_kernel(const double* __restrict__ mean, const double* __restrict__ shift, int n, int n_con,
                                        DevScalars* __restrict__ sc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int ok = 1;
    for (int c = 0; c < n_con; c++) {
      const double g = -(mean[c % n] - shift[c]);
      if (!isfinite(g)) sc->nonfinite = 1;
      if (g > 0.0) ok = 0;
    }
    sc->mean_feasible = ok;
    sc->constraint_evaluation_count += 1;
  }
}

// updateConstraints counting loop (:371-384). Single block.
//   set_bounds (generation 1 inside the viability regime): boundary_c = max(0, max_i g_ci); the reference sets it to
//   the running maximum inside the loop, so no sample is counted as violating in that generation.
__global__ void __launch_bounds__(1024)
constraint_count_kernel(const double* __restrict__ G, long long ldg, int lambda, int n_con, double* __restrict__ bounds,
                        int set_bounds, unsigned long long* __restrict__ viol, DevScalars* __restrict__ sc) {
  __shared__ double sb[64];
  __shared__ double red[32];
  __shared__ unsigned long long smax, snum;
  if (threadIdx.x == 0) { smax = 0; snum = 0; }
  if (set_bounds) {
    for (int c = 0; c < n_con; c++) {
      double m = 0.0;
      for (int i = threadIdx.x; i < lambda; i += blockDim.x) m = fmax(m, G[(size_t)c * ldg + i]);
      m = warp_max(m);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmax(m, red[w]);
        bounds[c] = m;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_con && c < 64; c += blockDim.x) sb[c] = bounds[c];
  __syncthreads();
  unsigned long long lmax = 0, lnum = 0;
  for (int i = threadIdx.x; i < lambda; i += blockDim.x) {
    unsigned long long cnt = 0;
    if (!set_bounds)
      for (int c = 0; c < n_con; c++) cnt += (G[(size_t)c * ldg + i] > (c < 64 ? sb[c] : bounds[c]) + 1e-12) ? 1 : 0;
    viol[i] = cnt;
    lmax = max(lmax, cnt);
    lnum += cnt ? 1 : 0;
  }
  atomicMax(&smax, lmax);
  atomicAdd(&snum, lnum);
  __syncthreads();
  if (threadIdx.x == 0) { sc->max_violation_count = smax; sc->violating_samples = snum; sc->constraint_evaluation_count += lambda; }
}

// Ordered list of (sample, constraint) correction events of one handleConstraints iteration (:780-786) and of the
// violating samples (:810-811), with the max-corrections cut-off (:787-793). Single block.
__global__ void __launch_bounds__(1024)
constraint_events_kernel(const unsigned long long* __restrict__ viol, const unsigned char* __restrict__ indicator, long long ldg,
                         int lambda, int n_con, unsigned long long max_corrections, int* __restrict__ ev_sample,
                         int* __restrict__ ev_con, int* __restrict__ vio_rows, int* __restrict__ counts_out /*[0]=events,[1]=violators*/,
                         DevScalars* __restrict__ sc) {
  __shared__ int wtot_e[32], wtot_v[32];
  __shared__ int carry_e, carry_v;
  if (threadIdx.x == 0) { carry_e = 0; carry_v = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < lambda; base += 1024) {
    const int i = base + threadIdx.x;
    int ke = 0, kv = 0;
    if (i < lambda && viol[i] > 0) {
      kv = 1;
      for (int c = 0; c < n_con; c++) ke += indicator[(size_t)c * ldg + i] ? 1 : 0;
    }
    int xe = ke, xv = kv;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int ye = __shfl_up_sync(0xffffffffu, xe, off), yv = __shfl_up_sync(0xffffffffu, xv, off);
      if (lane >= off) { xe += ye; xv += yv; }
    }
    if (lane == 31) { wtot_e[warp] = xe; wtot_v[warp] = xv; }
    __syncthreads();
    if (warp == 0) {
      int a = wtot_e[lane], b = wtot_v[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int ya = __shfl_up_sync(0xffffffffu, a, off), yb = __shfl_up_sync(0xffffffffu, b, off);
        if (lane >= off) { a += ya; b += yb; }
      }
      wtot_e[lane] = a; wtot_v[lane] = b;
    }
    __syncthreads();
    const int ce = carry_e, cv = carry_v;
    int pe = ce + (warp ? wtot_e[warp - 1] : 0) + xe - ke;
    const int pv = cv + (warp ? wtot_v[warp - 1] : 0) + xv - kv;
    if (kv) {
      vio_rows[pv] = i;
      for (int c = 0; c < n_con; c++)
        if (indicator[(size_t)c * ldg + i]) { ev_sample[pe] = i; ev_con[pe] = c; pe++; }
    }
    __syncthreads();
    if (threadIdx.x == 0) { carry_e = ce + wtot_e[31]; carry_v = cv + wtot_v[31]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const unsigned long long before = sc->cov_adaptation_count;
    const unsigned long long room = max_corrections > before ? max_corrections - before : 0ull;
    int J = carry_e;
    int abort_flag = 0;
    if ((unsigned long long)J > room) { J = (int)room; abort_flag = 1; }
    sc->cov_adaptation_count = before + (unsigned long long)J + (abort_flag ? 1ull : 0ull);
    sc->n_events = J;
    sc->adaptation_abort = abort_flag;
    counts_out[0] = J;
    counts_out[1] = carry_v;
  }
}

// v_c <- (1-lr) v_c + lr * y_i for the events of constraint c in order (:798); Uraw[j][:] = v_c after event j.
// grid = (ceil(n/256), n_con); each thread owns one (c, d) and walks the event list.
__global__ void __launch_bounds__(256)
constraint_normals_kernel(const int* __restrict__ ev_sample, const int* __restrict__ ev_con, const int* __restrict__ counts,
                          const double* __restrict__ Y, int ldy, double* __restrict__ normal, int ldn, double lr, double* __restrict__ U,
                          int ldu, int n) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (d >= n) return;
  const int J = counts[0];
  double v = normal[(size_t)c * ldn + d];
  for (int j = 0; j < J; j++) {
    if (ev_con[j] != c) continue;
    v = (1.0 - lr) * v + lr * Y[(size_t)ev_sample[j] * ldy + d];
    U[(size_t)j * ldu + d] = v;
  }
  normal[(size_t)c * ldn + d] = v;
}

// U[j][:] *= beta / (|v_j| * cnt_i)  so that sum_j U_j U_j^T = sum_j beta^2 v v^T / (v2 cnt^2) (:803). One warp per event.
__global__ void __launch_bounds__(256)
constraint_scale_kernel(double* __restrict__ U, int ldu, int n, const int* __restrict__ ev_sample, const int* __restrict__ counts,
                        const unsigned long long* __restrict__ viol, double beta, int rows_padded) {
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= rows_padded) return;
  const int J = counts[0];
  if (j >= J) {  // zero the k-tail of the SYRK operand
    for (int d = lane; d < ldu; d += 32) U[(size_t)j * ldu + d] = 0.0;
    return;
  }
  double a = 0.0;
  for (int d = lane; d < n; d += 32) { const double v = U[(size_t)j * ldu + d]; a += v * v; }
  a = warp_sum_butterfly(a);
  const double cnt = (double)viol[ev_sample[j]];
  const double f = beta / (sqrt(a) * cnt);
  for (int d = lane; d < n; d += 32) U[(size_t)j * ldu + d] *= f;
}

// C_aux = C - P, P = sum of split-K slabs (lower triangle), mirrored.
__global__ void __launch_bounds__(256)
caux_kernel(const double* __restrict__ C, double* __restrict__ Caux, int ldc, const double* __restrict__ W, int ldw, int splits, int n,
            const int* __restrict__ counts) {
  const int e = blockIdx.x * 16 + (threadIdx.x & 15);
  const int d = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (d >= n || e > d) return;
  double p = 0.0;
  if (counts[0] > 0)
    for (int s = 0; s < splits; s++) p += W[(size_t)s * n * ldw + (size_t)d * ldw + e];
  const double v = C[(size_t)d * ldc + e] - p;
  Caux[(size_t)d * ldc + e] = v;
  if (e < d) Caux[(size_t)e * ldc + d] = v;
}

// reEvaluateConstraints bookkeeping (:406-422) for the listed (violating) samples; G already refreshed for them.
__global__ void __launch_bounds__(256)
constraint_recount_kernel(const double* __restrict__ G, long long ldg, int n_con, const double* __restrict__ bounds,
                          const int* __restrict__ rows, const int* __restrict__ counts, unsigned long long* __restrict__ viol,
                          unsigned char* __restrict__ indicator) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= counts[1]) return;
  const int i = rows[li];
  unsigned long long cnt = 0;
  for (int c = 0; c < n_con; c++) {
    const bool bad = G[(size_t)c * ldg + i] > bounds[c] + 1e-12;
    indicator[(size_t)c * ldg + i] = bad ? 1 : 0;
    cnt += bad ? 1 : 0;
  }
  viol[i] = cnt;
}

// max violation count over all samples + counters after a re-evaluation. Single block.
__global__ void __launch_bounds__(1024)
constraint_max_kernel(const unsigned long long* __restrict__ viol, int lambda, const int* __restrict__ counts, DevScalars* __restrict__ sc) {
  __shared__ unsigned long long smax, snum;
  if (threadIdx.x == 0) { smax = 0; snum = 0; }
  __syncthreads();
  unsigned long long lmax = 0, lnum = 0;
  for (int i = threadIdx.x; i < lambda; i += blockDim.x) { lmax = max(lmax, viol[i]); lnum += viol[i] ? 1 : 0; }
  atomicMax(&smax, lmax);
  atomicAdd(&snum, lnum);
  __syncthreads();
  if (threadIdx.x == 0) {
    sc->max_violation_count = smax;
    sc->violating_samples = snum;
    sc->constraint_evaluation_count += (unsigned long long)counts[1];
  }
}

__global__ void add_resampled_kernel(DevScalars* sc, const int* counts) { sc->resampled_parameter_count += (unsigned long long)counts[1]; }

void launch_constraints_mean(cudaStream_t st, const double* mean, const double* shift, int n, int n_con, DevScalars* sc) {
  constraints_mean_kernel<<<1, 32, 0, st>>>(mean, shift, n, n_con, sc);
}
void launch_constraint_count(cudaStream_t st, const double* G, long long ldg, int lambda, int n_con, double* bounds, int set_bounds,
                             unsigned long long* viol, DevScalars* sc) {
  constraint_count_kernel<<<1, 1024, 0, st>>>(G, ldg, lambda, n_con, bounds, set_bounds, viol, sc);
}
void launch_constraint_events(cudaStream_t st, const unsigned long long* viol, const unsigned char* indicator, long long ldg, int lambda,
                              int n_con, unsigned long long max_corrections, int* ev_sample, int* ev_con, int* vio_rows, int* counts_out,
                              DevScalars* sc) {
  constraint_events_kernel<<<1, 1024, 0, st>>>(viol, indicator, ldg, lambda, n_con, max_corrections, ev_sample, ev_con, vio_rows, counts_out, sc);
}
void launch_constraint_normals(cudaStream_t st, const int* ev_sample, const int* ev_con, const int* counts, const double* Y, int ldy,
                               double* normal, int ldn, double lr, double* U, int ldu, int n, int n_con) {
  dim3 grid((n + 255) / 256, n_con);
  constraint_normals_kernel<<<grid, 256, 0, st>>>(ev_sample, ev_con, counts, Y, ldy, normal, ldn, lr, U, ldu, n);
}
void launch_constraint_scale(cudaStream_t st, double* U, int ldu, int n, const int* ev_sample, const int* counts,
                             const unsigned long long* viol, double beta, int rows_padded) {
  constraint_scale_kernel<<<(rows_padded + 7) / 8, 256, 0, st>>>(U, ldu, n, ev_sample, counts, viol, beta, rows_padded);
}
void launch_caux(cudaStream_t st, const double* C, double* Caux, int ldc, const double* W, int ldw, int splits, int n, const int* counts) {
  dim3 grid((n + 15) / 16, (n + 15) / 16);
  caux_kernel<<<grid, 256, 0, st>>>(C, Caux, ldc, W, ldw, splits, n, counts);
}
void launch_constraint_recount(cudaStream_t st, const double* G, long long ldg, int n_con, const double* bounds, const int* rows,
                               const int* counts, int max_rows, unsigned long long* viol, unsigned char* indicator) {
  if (max_rows <= 0) return;
  constraint_recount_kernel<<<(max_rows + 255) / 256, 256, 0, st>>>(G, ldg, n_con, bounds, rows, counts, viol, indicator);
}
void launch_constraint_max(cudaStream_t st, const unsigned long long* viol, int lambda, const int* counts, DevScalars* sc) {
  constraint_max_kernel<<<1, 1024, 0, st>>>(viol, lambda, counts, sc);
}
void launch_add_resampled(cudaStream_t st, DevScalars* sc, const int* counts) { add_resampled_kernel<<<1, 1, 0, st>>>(sc, counts); }

}  // namespace kc
