// eigen.cu — K8: symmetric eigendecomposition of the covariance matrix, once per generation.
// Replaces eigen() (CMAES.cpp.base:896-938: gsl_eigen_symmv + sort ABS_ASC) — the one step of the loop that
// is not population-parallel; timed separately ("eigen" phase).
//
// Method (version 1): WARM-STARTED one-sided (Hestenes) Jacobi. With V0 = the eigenvectors of the previous
// generation (C changes by ~c1+cmu per generation, so V0 nearly diagonalises it), form G = C V0 with the FP64
// tensor-core GEMM and orthogonalise the columns of G by plane rotations accumulated into V. At convergence
// C V = G has orthogonal columns, i.e. V holds the eigenvectors and lambda_i = v_i . g_i (Rayleigh quotient,
// signed). Vectors are stored as ROWS (VT, GT) so each rotation touches contiguous memory.
// One launch per round-robin step (n/2 disjoint pairs), one CTA per pair.
#include "common.cuh"
#include "kernels.h"

namespace kc {

// Round-robin (chess tournament) schedule over np = even number of players; step in [0, np-1), k in [0, np/2).
__device__ __forceinline__ void rr_pair(int np, int step, int k, int& p, int& q) {
  const int m = np - 1;
  int a, b;
  if (k == 0) { a = m; b = step; }
  else { a = (step + k) % m; b = (step - k + m) % m; }
  p = min(a, b); q = max(a, b);
}

__global__ void __launch_bounds__(128)
jacobi_step_kernel(double* __restrict__ GT, double* __restrict__ VT, int ld, int n, int np, int step, double tol,
                   DevScalars* __restrict__ sc) {
  int p, q;
  rr_pair(np, step, blockIdx.x, p, q);
  if (q >= n) return;  // dummy player when n is odd
  double* gp = GT + (size_t)p * ld; double* gq = GT + (size_t)q * ld;
  double* vp = VT + (size_t)p * ld; double* vq = VT + (size_t)q * ld;
  __shared__ double red[3][4];
  __shared__ double cs_s[2];
  double a = 0, b = 0, g = 0;
  for (int i = threadIdx.x; i < n; i += 128) {
    const double x = gp[i], y = gq[i];
    a += x * x; b += y * y; g += x * y;
  }
  a = warp_sum_butterfly(a); b = warp_sum_butterfly(b); g = warp_sum_butterfly(g);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = a; red[1][w] = b; red[2][w] = g; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double alpha = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    const double beta = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    const double gamma = red[2][0] + red[2][1] + red[2][2] + red[2][3];
    double c = 1.0, s = 0.0;
    if (fabs(gamma) > tol * sqrt(alpha * beta) && gamma != 0.0) {
      const double zeta = (beta - alpha) / (2.0 * gamma);
      const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      c = 1.0 / sqrt(1.0 + t * t);
      s = c * t;
      atomicAdd(&sc->jacobi_rotations, 1);
    }
    cs_s[0] = c; cs_s[1] = s;
  }
  __syncthreads();
  const double c = cs_s[0], s = cs_s[1];
  if (s == 0.0) return;
  for (int i = threadIdx.x; i < n; i += 128) {
    const double x = gp[i], y = gq[i];
    gp[i] = c * x - s * y; gq[i] = s * x + c * y;
    const double u = vp[i], v = vq[i];
    vp[i] = c * u - s * v; vq[i] = s * u + c * v;
  }
}

// ev[i] = v_i . g_i ; sign[i] = -1 if the component of largest magnitude (first on ties) is negative
// (sign convention shared with the CPU oracle). One warp per vector.
__global__ void __launch_bounds__(256)
rayleigh_kernel(const double* __restrict__ GT, const double* __restrict__ VT, int ld, int n, double* __restrict__ ev,
                double* __restrict__ sign) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  double a = 0.0, nv = 0.0, best = -1.0, bval = 0.0;
  int bidx = 0x7fffffff;
  for (int k = lane; k < n; k += 32) {
    const double v = VT[(size_t)i * ld + k];
    a += v * GT[(size_t)i * ld + k];
    nv += v * v;
    if (fabs(v) > best) { best = fabs(v); bval = v; bidx = k; }
  }
  a = warp_sum_butterfly(a); nv = warp_sum_butterfly(nv);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, off), ov = __shfl_xor_sync(0xffffffffu, bval, off);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
    if (ob > best || (ob == best && oi < bidx)) { best = ob; bval = ov; bidx = oi; }
  }
  if (lane == 0) { ev[i] = a / nv; sign[i] = bval < 0.0 ? -1.0 : 1.0; }
}

// perm = ascending order of |ev| (GSL_EIGEN_SORT_ABS_ASC; ties by index), min/max eigenvalue, acceptance test
// (min <= 0 keeps the previous B, D: CMAES.cpp.base:876-880). Single block; n <= a few thousand.
__global__ void __launch_bounds__(1024)
eig_order_kernel(const double* __restrict__ ev, int n, int* __restrict__ perm, DevScalars* __restrict__ sc) {
  __shared__ double smin[32], smax[32];
  double mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double a = fabs(ev[i]);
    int r = 0;
    for (int j = 0; j < n; j++) {
      const double b = fabs(ev[j]);
      r += (b < a) || (b == a && j < i);
    }
    perm[r] = i;
    mn = fmin(mn, ev[i]); mx = fmax(mx, ev[i]);
  }
  mn = warp_min(mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) { mn = fmin(mn, smin[w]); mx = fmax(mx, smax[w]); }
    if (mn <= 0.0 || !(mn == mn)) {
      sc->eig_rejected = 1;
    } else {
      sc->eig_rejected = 0;
      sc->min_eig = mn;
      sc->max_eig = mx;
    }
  }
}

// On acceptance: B[d][e] = VTw[perm[e]][d], D[e] = sqrt(ev[perm[e]]), A[d][e] = B[d][e]*D[e], VT <- VTw (permuted).
// 32x32 smem transpose tiles. On rejection nothing is written.
__global__ void __launch_bounds__(256)
eig_commit_kernel(const double* __restrict__ VTw, int ld, int n, const int* __restrict__ perm, const double* __restrict__ ev,
                  const double* __restrict__ sign, double* __restrict__ B, double* __restrict__ A, double* __restrict__ D, double* __restrict__ VT,
                  const DevScalars* __restrict__ sc) {
  if (sc->eig_rejected) return;
  __shared__ double tile[32][33];
  const int e0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int e = e0 + r, d = d0 + tx;
    double v = 0.0;
    if (e < n && d < n) {
      v = sign[perm[e]] * VTw[(size_t)perm[e] * ld + d];
      VT[(size_t)e * ld + d] = v;
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int d = d0 + r, e = e0 + tx;
    if (d < n && e < n) {
      const double v = tile[tx][r];
      const double dd = sqrt(ev[perm[e]]);
      B[(size_t)d * ld + e] = v;
      A[(size_t)d * ld + e] = v * dd;
      if (d == 0) D[e] = dd;
    }
  }
}

// Diagonal Covariance mode (:898-903): Q = I, eigenvalues = diag(C) UNSORTED.
__global__ void __launch_bounds__(256)
eig_diagonal_kernel(const double* __restrict__ C, int ldc, int n, double* __restrict__ D, DevScalars* __restrict__ sc) {
  __shared__ double smin[8], smax[8];
  double mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const double v = C[(size_t)i * ldc + i]; mn = fmin(mn, v); mx = fmax(mx, v); }
  mn = warp_min(mn); mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = mn; smax[threadIdx.x >> 5] = mx; }
  __syncthreads();
  __shared__ int rej;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 8; w++) { mn = fmin(mn, smin[w]); mx = fmax(mx, smax[w]); }
    rej = (mn <= 0.0 || !(mn == mn));
    sc->eig_rejected = rej;
    if (!rej) { sc->min_eig = mn; sc->max_eig = mx; }
  }
  __syncthreads();
  if (rej) return;
  for (int i = threadIdx.x; i < n; i += blockDim.x) D[i] = sqrt(C[(size_t)i * ldc + i]);
}

__global__ void __launch_bounds__(256) set_identity_kernel(double* __restrict__ M, int ld, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j < ld) M[(size_t)i * ld + j] = (i == j && j < n) ? 1.0 : 0.0;
}

__global__ void reset_rotations_kernel(DevScalars* sc) { sc->jacobi_rotations = 0; }

void launch_set_identity(cudaStream_t st, double* M, int ld, int n) {
  dim3 grid((ld + 255) / 256, n);
  set_identity_kernel<<<grid, 256, 0, st>>>(M, ld, n);
}
void launch_jacobi_sweep(cudaStream_t st, double* GT, double* VT, int ld, int n, double tol, DevScalars* sc, int* launches) {
  const int np = (n + 1) & ~1;
  reset_rotations_kernel<<<1, 1, 0, st>>>(sc);
  for (int step = 0; step < np - 1; step++) jacobi_step_kernel<<<np / 2, 128, 0, st>>>(GT, VT, ld, n, np, step, tol, sc);
  if (launches) *launches += np;
}
void launch_rayleigh(cudaStream_t st, const double* GT, const double* VT, int ld, int n, double* ev, double* sign) {
  rayleigh_kernel<<<(n + 7) / 8, 256, 0, st>>>(GT, VT, ld, n, ev, sign);
}
void launch_eig_order(cudaStream_t st, const double* ev, int n, int* perm, DevScalars* sc) {
  eig_order_kernel<<<1, 1024, 0, st>>>(ev, n, perm, sc);
}
void launch_eig_commit(cudaStream_t st, const double* VTw, int ld, int n, const int* perm, const double* ev, const double* sign,
                       double* B, double* A, double* D, double* VT, const DevScalars* sc) {
  dim3 grid((n + 31) / 32, (n + 31) / 32);
  eig_commit_kernel<<<grid, 256, 0, st>>>(VTw, ld, n, perm, ev, sign, B, A, D, VT, sc);
}
void launch_eig_diagonal(cudaStream_t st, const double* C, int ldc, int n, double* D, DevScalars* sc) {
  eig_diagonal_kernel<<<1, 256, 0, st>>>(C, ldc, n, D, sc);
}

}  // namespace kc
