// dc.cu — K8b: divide & conquer eigensolver of the symmetric tridiagonal T = (dT, eT) produced by sytrd_kernel (tridiag.cu).
// Cuppen's method with Gu & Eisenstat's stable eigenvectors, laid out for the GPU:
//
//   tear     T = blockdiag(T_1 .. T_L) + sum_b |e_b| u_b u_b^T at every leaf boundary b (dc_tear_kernel)
//   leaves   <= 32 x 32 each, one warp per leaf: cyclic two-sided Jacobi in shared memory (dc_leaf_kernel)
//   merges   level by level, ALL merges of a level per launch, no host round trip:
//            dc_setup_kernel    z = [last row of Q1 | +-first row of Q2]/sqrt2, rho = 2|e_b|, merge the two sorted spectra,
//                               deflate (tiny z_i; close poles by a Givens rotation of two eigenvector columns)
//            dc_secular_kernel  one warp per root: secular_root() (dc_inner.cuh), pole differences DELTA[j][i] = dl_i - lam_j
//            dc_loewner_kernel  w_hat_i^2 = prod_j (lam_j - dl_i) / prod_{j != i} (dl_j - dl_i)  (orthogonal vectors without
//                               extended precision)
//            dc_vectors_kernel  final sorted position of every root / deflated value, row of U^T (n_merge wide, zero or a
//                               unit vector where deflated): the merge is "deflation oblivious" — it always multiplies at
//                               full size, so its shapes are static and the level needs no device -> host count
//            gemm_tn_batched    Q_new = blockdiag(Q1, Q2) * U on the FP64 tensor cores (k-range halved while Q is block
//                               diagonal); the last merge is written transposed (rows = eigenvectors) for the back-transform
// NumPy statement of the same algorithm: profiles/microbench/tridiag_dc_proto.py. Replaces the QL iteration half of
// gsl_eigen_symmv (CMAES.cpp.base:917-936).
#include <stdlib.h>

#include "common.cuh"
#include "dc_inner.cuh"
#include "jacobi_inner.cuh"
#include "tridiag.h"

namespace kc {

namespace {

constexpr double kEps = 2.220446049250313e-16;

struct WarpLanes {
  __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int width() const { return 32; }
  __device__ __forceinline__ double sum(double v) const { return warp_sum_butterfly(v); }
};

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// ---- tear: d_out[i] = dT[i] - |e| of the leaf boundaries next to row i --------------------------------------------
__global__ void __launch_bounds__(256)
dc_tear_kernel(const double* __restrict__ dT, const double* __restrict__ eT, const int* __restrict__ bounds, int leaves, int n,
               double* __restrict__ dout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // leaf of row i: largest k with bounds[k] <= i
  int lo = 0, hi = leaves;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (bounds[mid] <= i) lo = mid; else hi = mid; }
  double v = dT[i];
  if (lo > 0 && i == bounds[lo]) v -= fabs(eT[i - 1]);                 // first row of a leaf that has a left neighbour
  if (lo + 1 < leaves && i == bounds[lo + 1] - 1) v -= fabs(eT[i]);     // last row of a leaf that has a right neighbour
  dout[i] = v;
}

// ---- leaves: one CTA per leaf, two-sided cyclic Jacobi (round-robin order) on the dense <= 32 x 32 block ------------------
// A round = np/2 disjoint rotations: parameters (one thread per pair), then A <- A J and V <- V J as (row, pair) work items, then
// A <- J^T A as (column, pair) work items: three barriers per round, all 256 threads busy.
constexpr int LEAF_MAX = 32;
constexpr int LEAF_LD = LEAF_MAX + 1;
constexpr int LEAF_NT = 256;

__global__ void __launch_bounds__(LEAF_NT)
dc_leaf_kernel(const int* __restrict__ bounds, int leaves, const double* __restrict__ dtorn, const double* __restrict__ eT,
               double* __restrict__ dout, double* __restrict__ Q, int ld) {
  __shared__ double A[LEAF_MAX * LEAF_LD], V[LEAF_MAX * LEAF_LD];
  __shared__ double cs_c[LEAF_MAX / 2], cs_s[LEAF_MAX / 2], diag[LEAF_MAX], nrm_s[LEAF_MAX];
  __shared__ int cs_p[LEAF_MAX / 2], cs_q[LEAF_MAX / 2], rnk[LEAF_MAX], rotated;
  const int tid = threadIdx.x;
  const int leaf = blockIdx.x;
  if (leaf >= leaves) return;
  const int off = bounds[leaf], m = bounds[leaf + 1] - off;
  for (int i = tid; i < LEAF_MAX * LEAF_LD; i += LEAF_NT) { A[i] = 0.0; V[i] = 0.0; }
  __syncthreads();
  if (tid < m) {
    const double dd = dtorn[off + tid];
    A[tid * LEAF_LD + tid] = dd;
    V[tid * LEAF_LD + tid] = 1.0;
    double nv = fabs(dd);
    if (tid + 1 < m) {
      const double ee = eT[off + tid];
      A[tid * LEAF_LD + tid + 1] = ee;
      A[(tid + 1) * LEAF_LD + tid] = ee;
      nv = fmax(nv, fabs(ee));
    }
    nrm_s[tid] = nv;
  }
  __syncthreads();
  double nrm = 0.0;
  for (int k = 0; k < m; k++) nrm = fmax(nrm, nrm_s[k]);
  const double thr = 0.25 * kEps * nrm;
  const int np = (m + 1) & ~1, hp = np / 2;
  for (int sweep = 0; sweep < 60 && m > 1; sweep++) {
    if (tid == 0) rotated = 0;
    __syncthreads();
    for (int round = 0; round < np - 1; round++) {
      if (tid < hp) {
        int p, q;
        rr_pair(np, round, tid, p, q);
        if (p > q) { const int t = p; p = q; q = t; }
        double c = 1.0, sn = 0.0;
        if (q < m) {
          const double apq = A[p * LEAF_LD + q];
          if (fabs(apq) > thr) {
            const double theta = (A[q * LEAF_LD + q] - A[p * LEAF_LD + p]) / (2.0 * apq);
            const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            sn = t * c;
            rotated = 1;
          }
        } else { p = 0; q = 0; }
        cs_c[tid] = c; cs_s[tid] = sn; cs_p[tid] = p; cs_q[tid] = q;
      }
      __syncthreads();
      for (int it = tid; it < m * hp; it += LEAF_NT) {   // A <- A J, V <- V J: (row, pair)
        const int row = it / hp, k = it - row * hp;
        const double sn = cs_s[k];
        if (sn == 0.0) continue;
        const double c = cs_c[k];
        const int p = cs_p[k], q = cs_q[k];
        const double ap = A[row * LEAF_LD + p], aq = A[row * LEAF_LD + q];
        A[row * LEAF_LD + p] = c * ap - sn * aq;
        A[row * LEAF_LD + q] = sn * ap + c * aq;
        const double vp = V[row * LEAF_LD + p], vq = V[row * LEAF_LD + q];
        V[row * LEAF_LD + p] = c * vp - sn * vq;
        V[row * LEAF_LD + q] = sn * vp + c * vq;
      }
      __syncthreads();
      for (int it = tid; it < m * hp; it += LEAF_NT) {   // A <- J^T A: (column, pair)
        const int k = it / m, col = it - k * m;
        const double sn = cs_s[k];
        if (sn == 0.0) continue;
        const double c = cs_c[k];
        const int p = cs_p[k], q = cs_q[k];
        const double ap = A[p * LEAF_LD + col], aq = A[q * LEAF_LD + col];
        A[p * LEAF_LD + col] = c * ap - sn * aq;
        A[q * LEAF_LD + col] = sn * ap + c * aq;
      }
      __syncthreads();
    }
    const int any = rotated;   // written before the last barrier of the sweep
    __syncthreads();           // every thread has read it before the next sweep re-arms it
    if (!any) break;
  }
  // ascending order, columns of Q = eigenvectors
  if (tid < m) diag[tid] = A[tid * LEAF_LD + tid];
  __syncthreads();
  if (tid < m) {
    const double val = diag[tid];
    int r = 0;
    for (int j = 0; j < m; j++) { const double o = diag[j]; r += (o < val) || (o == val && j < tid); }
    rnk[tid] = r;
    dout[off + r] = val;
  }
  __syncthreads();
  for (int it = tid; it < m * m; it += LEAF_NT) {
    const int row = it / m, c = it - row * m;
    Q[(size_t)(off + row) * ld + off + rnk[c]] = V[row * LEAF_LD + c];
  }
}

// ---- merge set-up + deflation: one CTA per merge ------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dc_setup_kernel(const DcNode* __restrict__ nodes, int node0, const double* __restrict__ dcur, const double* __restrict__ eT,
                double* __restrict__ Q, int ld, double* __restrict__ dl, double* __restrict__ w, double* __restrict__ defl_val,
                int* __restrict__ col2k, int* __restrict__ nd_col, int* __restrict__ defl_col, double* __restrict__ rho_out,
                int* __restrict__ Kcnt, int* __restrict__ mixed, int* __restrict__ rot_p, int* __restrict__ rot_q,
                double* __restrict__ rot_c, double* __restrict__ rot_s) {
  extern __shared__ __align__(16) double sm[];
  const int node = node0 + blockIdx.x;
  const DcNode nd = nodes[node];
  const int off = nd.off, n1 = nd.n1, n = nd.n, n2 = n - n1, tid = threadIdx.x, nt = blockDim.x;
  double* d_s = sm;
  double* z_s = sm + n;
  int* ord = reinterpret_cast<int*>(sm + 2 * n);
  __shared__ double red[16];
  __shared__ int s_K, s_M, s_nrot, s_mixed;
  const double beta = eT[off + n1 - 1];
  const double sgn = beta < 0.0 ? -1.0 : 1.0, rho = 2.0 * fabs(beta);
  double dmax = 0.0, zmax = 0.0;
  for (int i = tid; i < n; i += nt) {
    const double dv = dcur[off + i];
    const double zv = (i < n1 ? Q[(size_t)(off + n1 - 1) * ld + off + i] : sgn * Q[(size_t)(off + n1) * ld + off + i]) * 0.70710678118654752440;
    d_s[i] = dv; z_s[i] = zv;
    dmax = fmax(dmax, fabs(dv)); zmax = fmax(zmax, fabs(zv));
  }
  dmax = warp_max(dmax); zmax = warp_max(zmax);
  if ((tid & 31) == 0) { red[tid >> 5] = dmax; red[8 + (tid >> 5)] = zmax; }
  __syncthreads();
  dmax = 0.0; zmax = 0.0;
  for (int k = 0; k < (nt >> 5); k++) { dmax = fmax(dmax, red[k]); zmax = fmax(zmax, red[8 + k]); }
  const double tol = 8.0 * kEps * fmax(dmax, zmax);
  // merge of the two ascending halves (ties: first half first)
  for (int i = tid; i < n; i += nt) {
    const double v = d_s[i];
    int lo, hi, pos;
    if (i < n1) {   // number of second-half values < v
      lo = 0; hi = n2;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (d_s[n1 + mid] < v) lo = mid + 1; else hi = mid; }
      pos = i + lo;
    } else {        // number of first-half values <= v
      lo = 0; hi = n1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (d_s[mid] <= v) lo = mid + 1; else hi = mid; }
      pos = (i - n1) + lo;
    }
    ord[pos] = i;
  }
  __syncthreads();
  // Deflation. The scan of dlaed2 is sequential only through close-pole deflations (a Givens rotation changes the next comparison);
  // those are rare, so the CTA first classifies every pole in parallel — small z_i deflates, every survivor is tested against the
  // previous survivor — and compacts in parallel when no pair is close; otherwise (and when everything deflates) thread 0 scans.
  int* surv = ord + n;
  __shared__ int s_fast, cnt_s[256];
  if (tid == 0) s_fast = (rho * zmax > tol) ? 1 : 0;
  __syncthreads();
  if (s_fast) {
    for (int pos = tid; pos < n; pos += nt) surv[pos] = (rho * fabs(z_s[ord[pos]]) > tol) ? 1 : 0;
    __syncthreads();
    for (int pos = tid; pos < n; pos += nt) {
      if (!surv[pos]) continue;
      int q = pos - 1;
      while (q >= 0 && !surv[q]) q--;
      if (q < 0) continue;
      const int i = ord[pos], pj = ord[q];
      const double zi = z_s[i], zpj = z_s[pj], tt = d_s[i] - d_s[pj];
      if (fabs(tt * zi * zpj) <= tol * (zi * zi + zpj * zpj)) s_fast = 0;
    }
    __syncthreads();
  }
  if (s_fast) {
    const int chunk = (n + nt - 1) / nt, lo = min(n, tid * chunk), hi = min(n, lo + chunk);
    int c = 0;
    for (int pos = lo; pos < hi; pos++) c += surv[pos];
    cnt_s[tid] = c;
    __syncthreads();
    int base = 0, total = 0;
    for (int t = 0; t < nt; t++) { const int v = cnt_s[t]; total += v; if (t < tid) base += v; }
    for (int pos = lo; pos < hi; pos++) {
      const int i = ord[pos];
      if (surv[pos]) nd_col[off + base++] = i;
      else defl_col[off + (pos - base)] = i;
    }
    if (tid == 0) {
      s_K = total; s_M = n - total; s_nrot = 0; s_mixed = 0;
      Kcnt[node] = total; mixed[node] = 0; rho_out[node] = rho;
    }
  } else if (tid == 0) {   // sequential scan; rotations are recorded and applied by the whole CTA below
    int K = 0, M = 0, nrot = 0, mix = 0;
    if (rho * zmax <= tol) {
      for (int pos = 0; pos < n; pos++) defl_col[off + M++] = ord[pos];
    } else {
      // (pj, z_pj, d_pj) = the previous survivor, kept in registers; the close-pole test |tt c s| <= tol with c = z_i / t,
      // s = -z_pj / t, t = hypot(z_i, z_pj) is evaluated without square root or division: |tt z_i z_pj| <= tol (z_i^2 + z_pj^2)
      int pj = -1;
      double zpj = 0.0, dpj = 0.0;
      int inext = ord[0];
      double znext = z_s[inext], dnext = d_s[inext];
      for (int pos = 0; pos < n; pos++) {
        const int i = inext;
        const double zi = znext, dv = dnext;
        if (pos + 1 < n) { inext = ord[pos + 1]; znext = z_s[inext]; dnext = d_s[inext]; }
        if (rho * fabs(zi) <= tol) { defl_col[off + M++] = i; continue; }
        if (pj < 0) { pj = i; zpj = zi; dpj = dv; continue; }
        const double tt = dv - dpj, t2 = zi * zi + zpj * zpj;
        if (fabs(tt * zi * zpj) <= tol * t2) {
          const double t = sqrt(t2);
          const double c = zi / t, sn = -zpj / t;
          z_s[i] = t; z_s[pj] = 0.0;
          rot_p[off + nrot] = pj; rot_q[off + nrot] = i; rot_c[off + nrot] = c; rot_s[off + nrot] = sn; nrot++;
          const double tn = dpj * c * c + dv * sn * sn;
          const double di = dpj * sn * sn + dv * c * c;
          d_s[i] = di;
          d_s[pj] = tn;
          defl_col[off + M++] = pj;
          if ((pj < n1) != (i < n1)) mix = 1;
          pj = i; zpj = t; dpj = di;
        } else {
          nd_col[off + K++] = pj;
          pj = i; zpj = zi; dpj = dv;
        }
      }
      if (pj >= 0) nd_col[off + K++] = pj;
    }
    s_K = K; s_M = M; s_nrot = nrot; s_mixed = mix;
    Kcnt[node] = K; mixed[node] = mix; rho_out[node] = rho;
  }
  __syncthreads();
  const int K = s_K, M = s_M, nrot = s_nrot;
  for (int i = tid; i < n; i += nt) col2k[off + i] = -1;
  __syncthreads();
  for (int k = tid; k < K; k += nt) {
    const int c = nd_col[off + k];
    col2k[off + c] = k;
    dl[off + k] = d_s[c];
    w[off + k] = z_s[c];
  }
  for (int m = tid; m < M; m += nt) defl_val[off + m] = d_s[defl_col[off + m]];
  // Givens rotations of eigenvector columns (x' = c x + s y, y' = c y - s x); a thread owns its rows, so the chain needs no barrier
  for (int t = 0; t < nrot; t++) {
    const int p = rot_p[off + t], q = rot_q[off + t];
    const double c = rot_c[off + t], s = rot_s[off + t];
    for (int r = tid; r < n; r += nt) {
      double* row = Q + (size_t)(off + r) * ld + off;
      const double x = row[p], y = row[q];
      row[p] = c * x + s * y;
      row[q] = c * y - s * x;
    }
  }
}

// ---- secular equation: one warp per root -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dc_secular_kernel(const DcNode* __restrict__ nodes, const int* __restrict__ row2node, int n_total, const double* __restrict__ dl,
                  const double* __restrict__ w, const double* __restrict__ rho, const int* __restrict__ Kcnt,
                  double* __restrict__ lam, double* __restrict__ DELTA, int ld) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= n_total) return;
  const int node = row2node[g];
  const int off = nodes[node].off, j = g - off, K = Kcnt[node];
  if (j >= K) return;
  const double* dlp = dl + off;
  const double* wp = w + off;
  int o; double mu;
  WarpLanes cx;
  secular_root(cx, j, K, dlp, wp, rho[node], o, mu);
  const double dorg = dlp[o];
  if (lane == 0) lam[off + j] = dorg + mu;
  double* row = DELTA + (size_t)(off + j) * ld + off;
  for (int i = lane; i < K; i += 32) row[i] = (dlp[i] - dorg) - mu;
}

// ---- Loewner weights: block = 32 consecutive poles x 32 slices of the product over the roots (a slice is a dependent chain of
// K / 32 divisions: 32 steps at the top merge of N = 1000 instead of the 125 of the 8-slice version) ---------------------------------
constexpr int LW_SLICES = 32;
__global__ void __launch_bounds__(32 * LW_SLICES)
dc_loewner_kernel(const DcNode* __restrict__ nodes, const int* __restrict__ row2node, int n_total, const double* __restrict__ dl,
                  const double* __restrict__ w, const int* __restrict__ Kcnt, const double* __restrict__ DELTA, int ld,
                  double* __restrict__ what) {
  __shared__ double part[LW_SLICES][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int g = blockIdx.x * 32 + tx;
  double prod = 1.0;
  bool live = false;
  int off = 0, i = 0;
  if (g < n_total) {
    const int node = row2node[g];
    off = nodes[node].off; i = g - off;
    const int K = Kcnt[node];
    if (i < K) {
      live = true;
      const double di = dl[off + i];
      for (int j = ty; j < K; j += LW_SLICES) {
        const double dji = DELTA[(size_t)(off + j) * ld + off + i];
        prod *= (j == i) ? dji : dji / (di - dl[off + j]);
      }
    }
  }
  part[ty][tx] = prod;
  __syncthreads();
  if (ty == 0 && live) {
    double p = part[0][tx];
#pragma unroll
    for (int k = 1; k < LW_SLICES; k++) p *= part[k][tx];
    what[off + i] = copysign(sqrt(fabs(p)), w[off + i]);
  }
}

// ---- split-K slabs of a level's merge products -> the level's output, diagonal blocks only (fixed order) ----------------------
__global__ void __launch_bounds__(256)
dc_reduce_blocks_kernel(const double* __restrict__ slabs, long long stride, int splits, const DcNode* __restrict__ nodes,
                        const int* __restrict__ row2node, int ld, double* __restrict__ out) {
  const int r = blockIdx.y;
  const DcNode nd = nodes[row2node[r]];
  const int cc = blockIdx.x * blockDim.x + threadIdx.x;
  if (cc >= nd.n) return;
  const size_t i = (size_t)r * ld + nd.off + cc;
  double s = 0.0;
  for (int k = 0; k < splits; k++) s += slabs[(size_t)k * stride + i];
  out[i] = s;
}

// ---- rows of U^T at their final (ascending) positions, new eigenvalues: one warp per root / deflated value ----------
__global__ void __launch_bounds__(256)
dc_vectors_kernel(const DcNode* __restrict__ nodes, const int* __restrict__ row2node, int n_total, const double* __restrict__ lam,
                  const double* __restrict__ defl_val, const int* __restrict__ defl_col, const int* __restrict__ col2k,
                  const int* __restrict__ Kcnt, const double* __restrict__ what, const double* __restrict__ DELTA, int ld,
                  double* __restrict__ UT, double* __restrict__ dnext) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= n_total) return;
  const int node = row2node[g];
  const int off = nodes[node].off, n = nodes[node].n, t = g - off, K = Kcnt[node], M = n - K;
  const bool is_root = t < K;
  const int m = t - K;
  const double val = is_root ? lam[off + t] : defl_val[off + m];
  // total order: value, roots before deflated values, then index
  int cnt = 0;
  if (is_root) {
    for (int k = lane; k < M; k += 32) cnt += defl_val[off + k] < val;
    cnt = warp_sum_int(cnt) + t;
  } else {
    for (int k = lane; k < K; k += 32) cnt += lam[off + k] <= val;
    for (int k = lane; k < M; k += 32) { const double o = defl_val[off + k]; cnt += (o < val) || (o == val && k < m); }
    cnt = warp_sum_int(cnt);
  }
  const int pos = cnt;
  if (lane == 0) dnext[off + pos] = val;
  double* urow = UT + (size_t)(off + pos) * ld + off;
  if (is_root) {
    const double* drow = DELTA + (size_t)(off + t) * ld + off;
    double nn = 0.0;
    for (int k = lane; k < K; k += 32) { const double u = what[off + k] / drow[k]; nn += u * u; }
    nn = warp_sum_butterfly(nn);
    const double inv = 1.0 / sqrt(nn);
    for (int c = lane; c < n; c += 32) {
      const int k = col2k[off + c];
      urow[c] = (k >= 0) ? what[off + k] / drow[k] * inv : 0.0;
    }
  } else {
    const int dc = defl_col[off + m];
    for (int c = lane; c < n; c += 32) urow[c] = (c == dc) ? 1.0 : 0.0;
  }
}

__global__ void __launch_bounds__(256) dc_transpose_kernel(const double* __restrict__ in, double* __restrict__ out, int ld, int n) {
  __shared__ double tile[32][33];
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = (y0 + r < n && x0 + tx < n) ? in[(size_t)(y0 + r) * ld + x0 + tx] : 0.0;
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (x0 + r < n && y0 + tx < n) out[(size_t)(x0 + r) * ld + y0 + tx] = tile[tx][r];
}

}  // namespace

void launch_transpose(cudaStream_t st, const double* in, double* out, int ld, int n) {
  dim3 grid((n + 31) / 32, (n + 31) / 32);
  dc_transpose_kernel<<<grid, 256, 0, st>>>(in, out, ld, n);
}

// Eigen-decomposition of the tridiagonal (ws->dT, ws->eT): rows of ws->XT = eigenvectors, ws->ev_final = ascending eigenvalues.
bool dc_solve(cudaStream_t st, TridiagWs* ws, int* launches) {
  const int n = ws->n, ld = ws->ld;
  static std::atomic<unsigned long long> attr{0};
  if (first_call_on_device(attr)) cudaFuncSetAttribute(dc_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const size_t mat = sizeof(double) * (size_t)n * ld;
  cudaMemsetAsync(ws->Qa, 0, mat, st);
  if (ws->levels > 1) cudaMemsetAsync(ws->Qb, 0, mat, st);
  dc_tear_kernel<<<(n + 255) / 256, 256, 0, st>>>(ws->dT, ws->eT, ws->d_bounds, ws->leaf_count, n, ws->dB);
  dc_leaf_kernel<<<ws->leaf_count, LEAF_NT, 0, st>>>(ws->d_bounds, ws->leaf_count, ws->dB, ws->eT, ws->dA, ws->Qa, ld);
  *launches += 2;
  double *dcur = ws->dA, *dnext = ws->dB;
  double *qsrc = ws->Qa;
  if (ws->levels == 0) {
    launch_transpose(st, ws->Qa, ws->XT, ld, n);
    *launches += 1;
    ws->ev_final = dcur;
    return cudaGetLastError() == cudaSuccess;
  }
  const int warp_blocks = (n * 32 + 255) / 256;
  for (int l = 1; l <= ws->levels; l++) {
    const std::vector<DcNode>& nodes = ws->lvl[l - 1];
    const int node0 = ws->lvl_node_begin[l - 1], cnt = (int)nodes.size();
    int nmax = 0;
    for (const DcNode& nd : nodes) nmax = nd.n > nmax ? nd.n : nmax;
    const int* r2n = ws->d_row2node + (size_t)(l - 1) * n;
    const size_t smem = sizeof(double) * 2 * nmax + sizeof(int) * 2 * nmax;
    dc_setup_kernel<<<cnt, 256, smem, st>>>(ws->d_nodes, node0, dcur, ws->eT, qsrc, ld, ws->dl, ws->w, ws->defl_val, ws->col2k,
                                            ws->nd_col, ws->defl_col, ws->rho, ws->Kcnt, ws->mixed, ws->rot_p, ws->rot_q, ws->rot_c,
                                            ws->rot_s);
    dc_secular_kernel<<<warp_blocks, 256, 0, st>>>(ws->d_nodes, r2n, n, ws->dl, ws->w, ws->rho, ws->Kcnt, ws->lam, ws->DELTA, ld);
    dc_loewner_kernel<<<(n + 31) / 32, 32 * LW_SLICES, 0, st>>>(ws->d_nodes, r2n, n, ws->dl, ws->w, ws->Kcnt, ws->DELTA, ld, ws->what);
    dc_vectors_kernel<<<warp_blocks, 256, 0, st>>>(ws->d_nodes, r2n, n, ws->lam, ws->defl_val, ws->defl_col, ws->col2k, ws->Kcnt,
                                                   ws->what, ws->DELTA, ld, ws->UT, dnext);
    const int msplit = (l == ws->levels) ? ws->split_top : (ws->lvl_split.empty() || !ws->slabsF ? 1 : ws->lvl_split[l - 1]);
    launch_gemm_batched(st, ws->d_desc + ws->desc_level_begin[l - 1], cnt, nmax, nmax, msplit);
    if (msplit > 1 && l == ws->levels) { launch_reduce_slabs(st, ws->slabsF, (long long)n * ld, msplit, n, n, ld, ws->XT, ws->num_sms); *launches += 1; }
    else if (msplit > 1) {   // sum the slabs of the diagonal blocks into the level's output (level l writes Qb when l is odd, Qa when even)
      dc_reduce_blocks_kernel<<<dim3((nmax + 255) / 256, n), 256, 0, st>>>(ws->slabsF, (long long)n * ld, msplit, ws->d_nodes, r2n, ld, (l & 1) ? ws->Qb : ws->Qa);
      *launches += 1;
    }
    *launches += 5;
    double* t = dcur; dcur = dnext; dnext = t;
    qsrc = (qsrc == ws->Qa) ? ws->Qb : ws->Qa;
  }
  ws->ev_final = dcur;
  return cudaGetLastError() == cudaSuccess;
}

}  // namespace kc
