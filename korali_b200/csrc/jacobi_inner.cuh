// jacobi_inner.cuh — the 8x8 inner problem of a Gram-update Jacobi step, entirely in registers.
//
// A step of the blocked one-sided Jacobi (eigen.cu) owns 8 rows of G. Their Gram matrix Gamma (8x8, symmetric) is all the
// rotations need: rotating rows p, q of G by (c, s) maps Gamma -> J Gamma J^T, so the whole sequence of plane rotations of
// the step is carried out on Gamma while R = prod J (8x8) accumulates; the long rows are then updated once, rows <- R rows.
//
// This sequence is the serial heart of the eigensolver (N-1 dependent rounds per sweep). Measured on B200
// (profiles/microbench/fp64_latency.cu): DFMA/DMUL/DADD 8 cycles dependent-issue, MUFU.RSQ64H ~19, but a shared-memory or
// shuffle exchange costs >= 35 cycles and, worse, queues behind the step's global loads in the MIO pipe (the version that
// kept Gamma in shared memory spent 1850 cycles per round, 40 % of the step). So every lane of the rotation warp holds ALL
// of Gamma (36 unique entries) and performs the identical arithmetic: no exchange, no memory instruction, the four disjoint
// rotations of a round give the scheduler four independent chains. Lane j keeps column j of R.
//
// All indices are compile-time constants (templates + full unrolling) so Gamma lives in registers.
// The same code compiles for the host (tests/test_jacobi_inner.py builds it with g++ and checks R Gamma R^T).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define KC_HD __host__ __device__ __forceinline__
#else
#define KC_HD inline
#endif

namespace kc {

// Round-robin (chess tournament) schedule over np = even number of players; step in [0, np-1), k in [0, np/2).
KC_HD void rr_pair(int np, int step, int k, int& p, int& q) {
  const int m = np - 1;
  int a, b;
  if (k == 0) { a = m; b = step; }
  else { a = (step + k) % m; b = (step - k + m) % m; }
  p = a < b ? a : b; q = a < b ? b : a;
}

// Ring tournament over np = power-of-two players: the first half A anchors, the second half B travels round a ring — in step s
// slot k pairs A[k] with B[(k + s) mod np/2] — and after np/2 steps the tournaments inside A and inside B run side by side with
// the same construction (slots 0..np/4-1 and np/4..np/2-1). np - 1 steps of np/2 disjoint pairs, every pair exactly once
// (profiles/microbench/jacobi_orderings_sim.py checks the 1-factorisation and the sweep counts: one sweep fewer than the
// round-robin order on CMA-ES covariances, warm-started or not). Players >= the real block count are phantom blocks (rows >= n).
KC_HD void ring_pair(int np, int step, int k, int& p, int& q) {
  int base = 0;
  for (;;) {
    const int h = np >> 1;
    if (step < h) { p = base + k; q = base + h + ((k + step) & (h - 1)); return; }
    step -= h;
    if (k >= (h >> 1)) { base += h; k -= h >> 1; }
    np = h;
  }
}
// A sweep in which every rotated pair already had cos^2 below this value is the last one: the rotations of that sweep leave
// cos ~ (cos_before)^2 * (lambda / gap), so the confirmation sweep (a full pass that rotates nothing) would be redundant.
#ifndef KC_QUAD_TAIL
#define KC_QUAD_TAIL 1e-16
#endif

KC_HD double rsqrt_seed(double a) {   // ~2^-22 relative
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  return r;
#else
  return (double)(float)(1.0 / sqrt(a));
#endif
}
KC_HD double rcp_seed(double a) {   // ~2^-22 relative
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  return r;
#else
  return (double)(float)(1.0 / a);
#endif
}

// Rotation (c, s) that orthogonalises two rows with |g_p|^2 = alpha, |g_q|^2 = beta, g_p.g_q = gamma != 0:
//   rows_p' = c rows_p - s rows_q,  rows_q' = s rows_p + c rows_q,  t = s/c = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)),
//   zeta = (beta - alpha) / (2 gamma)   <=>   |t| = 2|gamma| / (|d| + sqrt(d^2 + 4 gamma^2)),  d = beta - alpha.
// The ANGLE only needs a few digits (an error e_t leaves a residual e_t |gamma|: quadratic convergence is untouched down
// to 1e-7 per sweep) but (c, s) must be orthonormal to FP64 precision, so t comes from the 22-bit MUFU seeds and
// c = 1/sqrt(1 + t^2) is refined by two Newton steps; s = c t. 17 dependent FP64 operations, no divide, no sqrt call,
// and no branch: the four rotations of a round are scheduled as four interleaved chains.
// Returns false when d^2 + 4 gamma^2 left the safe exponent range (the caller then uses jacobi_cs_scaled).
KC_HD bool jacobi_cs_fast(double alpha, double beta, double gamma, double& c, double& s, double& t) {
  const double d = beta - alpha, g2 = gamma + gamma;
  const double h2 = fma(d, d, g2 * g2);
  const double h = h2 * rsqrt_seed(h2);                 // sqrt(d^2 + 4 gamma^2), 22 bits
  const double at = fabs(g2) * rcp_seed(fabs(d) + h);   // |t| in (0, 1]
  // sign(t) = sign(d) sign(gamma), with d = +-0 counted as positive/negative alike (|t| = 1 orthogonalises either way);
  // done on the integer pipe: the FP64 pipe is the bottleneck of this routine
#if defined(__CUDA_ARCH__)
  const int sgn = (__double2hiint(d) ^ __double2hiint(g2)) & 0x80000000;
  t = __hiloint2double(__double2hiint(at) | sgn, __double2loint(at));
  const unsigned ex = ((unsigned)__double2hiint(h2) >> 20) & 0x7ffu;
  const bool safe = (ex - 64u) < 1920u;                 // 2^-959 < h2 < 2^961
#else
  t = (signbit(d) != signbit(g2)) ? -at : at;
  const bool safe = h2 > 1e-288 && h2 < 1e288;
#endif
  const double x = fma(t, t, 1.0);
  const double mhx = -0.5 * x;
  double c0 = rsqrt_seed(x);
  c0 = c0 * fma(mhx, c0 * c0, 1.5);
  c0 = c0 * fma(mhx, c0 * c0, 1.5);
  c = c0;
  s = c0 * t;
  return safe;
}
KC_HD bool jacobi_cs_fast(double alpha, double beta, double gamma, double& c, double& s) {
  double t;
  return jacobi_cs_fast(alpha, beta, gamma, c, s, t);
}
// The same rotation after scaling (alpha, beta, gamma) by an exact power of two into the safe range.
KC_HD void jacobi_cs_scaled(double alpha, double beta, double gamma, double& c, double& s) {
  const double m = fmax(fmax(fabs(alpha), fabs(beta)), fabs(gamma));
#if defined(__CUDA_ARCH__)
  const int e = (__double2hiint(m) >> 20) & 0x7ff;                 // biased exponent of the largest magnitude
  const double sc = __hiloint2double((2046 - min(e, 2045)) << 20, 0);   // 2^-(e - 1023), exact
#else
  int e;
  frexp(m, &e);
  const double sc = ldexp(1.0, 1 - e);
#endif
  jacobi_cs_fast(alpha * sc, beta * sc, gamma * sc, c, s);
}
KC_HD void jacobi_cs_scaled(double alpha, double beta, double gamma, double& c, double& s, double& t) {
  jacobi_cs_scaled(alpha, beta, gamma, c, s);
  t = s / c;
}

// Gamma as a full 8x8 register array of which only the upper triangle (i <= j) is live.
struct Inner8 {
  double g[8][8];
  double rc[8];       // column (lane & 7) of R
  int rotations;      // plane rotations applied
  int big;            // some pair still had cos^2 >= KC_QUAD_TAIL (not yet in the quadratic tail)
};

template <int I, int J>
KC_HD double& gsym(Inner8& m) {
  return m.g[I < J ? I : J][I < J ? J : I];
}

template <int P, int Q, int J>
KC_HD void rot_offdiag(Inner8& m, double c, double s) {
  if (J != P && J != Q) {
    const double x = gsym<P, J>(m), y = gsym<Q, J>(m);
    gsym<P, J>(m) = fma(c, x, -(s * y));
    gsym<Q, J>(m) = fma(s, x, c * y);
  }
}

// Parameters of the plane rotation of rows P < Q: identity (c = 1, s = 0, t = 0) when cos^2 <= tol2. Branch-free.
template <int P, int Q>
KC_HD bool pair_params(Inner8& m, double tol2, double& c, double& s, double& t) {
  const double alpha = m.g[P][P], beta = m.g[Q][Q], gamma = m.g[P][Q];
  const bool on = gamma * gamma > tol2 * (alpha * beta);
  double cf, sf, tf;
  const bool safe = jacobi_cs_fast(alpha, beta, gamma, cf, sf, tf);
  c = on ? cf : 1.0;
  s = on ? sf : 0.0;
  t = on ? tf : 0.0;
  m.rotations += on ? 1 : 0;
  return on && !safe;
}

// Gamma <- J Gamma J^T and R <- J R for the rotation (c, s = c t) of rows/columns P < Q (exact identity for c = 1, s = t = 0).
// The pivot entries use the closed forms alpha' = alpha - t gamma, beta' = beta + t gamma, gamma' = 0: with the 22-bit angle
// they are off by ~1e-7 |t gamma|, which only perturbs the ANGLES of the later rounds of this step (Gamma is recomputed
// from the rows at every step).
template <int P, int Q>
KC_HD void rot_apply(Inner8& m, double c, double s, double t) {
  const double tg = t * m.g[P][Q];
  rot_offdiag<P, Q, 0>(m, c, s); rot_offdiag<P, Q, 1>(m, c, s); rot_offdiag<P, Q, 2>(m, c, s); rot_offdiag<P, Q, 3>(m, c, s);
  rot_offdiag<P, Q, 4>(m, c, s); rot_offdiag<P, Q, 5>(m, c, s); rot_offdiag<P, Q, 6>(m, c, s); rot_offdiag<P, Q, 7>(m, c, s);
  m.g[P][P] -= tg;
  m.g[Q][Q] += tg;
  m.g[P][Q] = 0.0;
  const double u = m.rc[P], v = m.rc[Q];
  m.rc[P] = fma(c, u, -(s * v));
  m.rc[Q] = fma(s, u, c * v);
}

// One round: four DISJOINT pairs. Their parameters are independent (a rotation of rows p, q leaves alpha, beta, gamma of a
// disjoint pair untouched), so the four chains are computed side by side, then the four updates.
template <int P0, int Q0, int P1, int Q1, int P2, int Q2, int P3, int Q3>
KC_HD void round4(Inner8& m, double tol2) {
  double c0, s0, t0, c1, s1, t1, c2, s2, t2, c3, s3, t3;
  const bool b0 = pair_params<P0, Q0>(m, tol2, c0, s0, t0);
  const bool b1 = pair_params<P1, Q1>(m, tol2, c1, s1, t1);
  const bool b2 = pair_params<P2, Q2>(m, tol2, c2, s2, t2);
  const bool b3 = pair_params<P3, Q3>(m, tol2, c3, s3, t3);
  if (b0 | b1 | b2 | b3) {   // exponent range of the fast path exceeded (|C| ~ 1e+-72): never in practice
    if (b0) jacobi_cs_scaled(m.g[P0][P0], m.g[Q0][Q0], m.g[P0][Q0], c0, s0, t0);
    if (b1) jacobi_cs_scaled(m.g[P1][P1], m.g[Q1][Q1], m.g[P1][Q1], c1, s1, t1);
    if (b2) jacobi_cs_scaled(m.g[P2][P2], m.g[Q2][Q2], m.g[P2][Q2], c2, s2, t2);
    if (b3) jacobi_cs_scaled(m.g[P3][P3], m.g[Q3][Q3], m.g[P3][Q3], c3, s3, t3);
  }
  rot_apply<P0, Q0>(m, c0, s0, t0);
  rot_apply<P1, Q1>(m, c1, s1, t1);
  rot_apply<P2, Q2>(m, c2, s2, t2);
  rot_apply<P3, Q3>(m, c3, s3, t3);
}

// cos^2 of the pair (P, Q) above tol2? Also records whether it is still >= KC_QUAD_TAIL (not yet in the quadratic tail).
template <int P, int Q>
KC_HD bool pair_on(Inner8& m, double tol2) {
  const double gamma = m.g[P][Q];
  const double g2 = gamma * gamma, ab = m.g[P][P] * m.g[Q][Q];
  m.big |= (g2 > KC_QUAD_TAIL * ab) ? 1 : 0;
  return g2 > tol2 * ab;
}

// Rows 0-3 against rows 4-7 (the two 4-row blocks were orthogonalised internally at the first step of the sweep):
// round r pairs k with 4 + (k + r) mod 4. Returns immediately when no cross pair is above tolerance (the common case in the
// last sweeps). m.big is judged on the Gamma the step starts from.
KC_HD void inner_cross(Inner8& m, double tol2) {
  const bool any = pair_on<0, 4>(m, tol2) | pair_on<0, 5>(m, tol2) | pair_on<0, 6>(m, tol2) | pair_on<0, 7>(m, tol2) |
                   pair_on<1, 4>(m, tol2) | pair_on<1, 5>(m, tol2) | pair_on<1, 6>(m, tol2) | pair_on<1, 7>(m, tol2) |
                   pair_on<2, 4>(m, tol2) | pair_on<2, 5>(m, tol2) | pair_on<2, 6>(m, tol2) | pair_on<2, 7>(m, tol2) |
                   pair_on<3, 4>(m, tol2) | pair_on<3, 5>(m, tol2) | pair_on<3, 6>(m, tol2) | pair_on<3, 7>(m, tol2);
  if (!any) { m.big = 0; return; }
  round4<0, 4, 1, 5, 2, 6, 3, 7>(m, tol2);
  round4<0, 5, 1, 6, 2, 7, 3, 4>(m, tol2);
  round4<0, 6, 1, 7, 2, 4, 3, 5>(m, tol2);
  round4<0, 7, 1, 4, 2, 5, 3, 6>(m, tol2);
}

// All 28 pairs of the 8 rows as 7 rounds of a round-robin tournament (first step of a sweep: intra-block pairs included).
KC_HD void inner_full(Inner8& m, double tol2) {
  (void)(pair_on<0, 1>(m, tol2) | pair_on<0, 2>(m, tol2) | pair_on<0, 3>(m, tol2) | pair_on<1, 2>(m, tol2) | pair_on<1, 3>(m, tol2) |
         pair_on<2, 3>(m, tol2) | pair_on<4, 5>(m, tol2) | pair_on<4, 6>(m, tol2) | pair_on<4, 7>(m, tol2) | pair_on<5, 6>(m, tol2) |
         pair_on<5, 7>(m, tol2) | pair_on<6, 7>(m, tol2) | pair_on<0, 4>(m, tol2) | pair_on<0, 5>(m, tol2) | pair_on<0, 6>(m, tol2) |
         pair_on<0, 7>(m, tol2) | pair_on<1, 4>(m, tol2) | pair_on<1, 5>(m, tol2) | pair_on<1, 6>(m, tol2) | pair_on<1, 7>(m, tol2) |
         pair_on<2, 4>(m, tol2) | pair_on<2, 5>(m, tol2) | pair_on<2, 6>(m, tol2) | pair_on<2, 7>(m, tol2) | pair_on<3, 4>(m, tol2) |
         pair_on<3, 5>(m, tol2) | pair_on<3, 6>(m, tol2) | pair_on<3, 7>(m, tol2));
  round4<0, 7, 1, 6, 2, 5, 3, 4>(m, tol2);
  round4<1, 7, 0, 2, 3, 6, 4, 5>(m, tol2);
  round4<2, 7, 1, 3, 0, 4, 5, 6>(m, tol2);
  round4<3, 7, 2, 4, 1, 5, 0, 6>(m, tol2);
  round4<4, 7, 3, 5, 2, 6, 0, 1>(m, tol2);
  round4<5, 7, 4, 6, 0, 3, 1, 2>(m, tol2);
  round4<6, 7, 0, 5, 1, 4, 2, 3>(m, tol2);
}


}  // namespace kc
