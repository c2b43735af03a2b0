// dc_inner.cuh — scalar heart of the divide & conquer tridiagonal eigensolver (dc.cu): one root of the secular equation
//
//     f(lam) = 1/rho + sum_i w_i^2 / (dl_i - lam) = 0,   dl ascending, rho > 0
//
// Root j lies in (dl_j, dl_{j+1}) (the last one in (dl_{K-1}, dl_{K-1} + rho*|w|^2)). Following Gu & Eisenstat the root is
// represented as (origin o, offset mu), lam = dl_o + mu with o the CLOSER pole, so that the differences
// delta_i = (dl_i - dl_o) - mu that build the eigenvector are accurate to a few ulp even when lam is within 1e-28 of a pole.
// Iteration: Bunch-Nielsen-Sorensen rational interpolation from both sides (psi = poles <= j, phi = poles > j, each replaced by
// s + S/(pole - mu) matching value and slope), inside a bracket that every iterate tightens; an iterate that leaves the
// bracket is replaced by its midpoint. Stopping rule = LAPACK's dlaed4: |f| <= eps * (error bound of the evaluated sum).
//
// The K-term sums are lane-parallel: `Lanes` supplies lane(), width() and sum() (device: the 32 lanes of a warp with a
// butterfly reduction so that every lane sees the same bits and takes the same branches; host: one lane). The same code is
// built with g++ by tests/test_dc_inner.py and checked against numpy; profiles/microbench/tridiag_dc_proto.py is the NumPy
// statement of the whole solver.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define KC_DC_HD __host__ __device__ __forceinline__
#else
#define KC_DC_HD inline
#endif

namespace kc {

struct SerialLanes {
  KC_DC_HD int lane() const { return 0; }
  KC_DC_HD int width() const { return 1; }
  KC_DC_HD double sum(double v) const { return v; }
};

constexpr double kDcEps = 2.220446049250313e-16;

// Returns the number of iterations; o_out / mu_out as described above. w holds the weights (not squared).
template <class Lanes>
KC_DC_HD int secular_root(const Lanes& cx, int j, int K, const double* dl, const double* w, double rho, int& o_out, double& mu_out) {
  const int lane = cx.lane(), width = cx.width();
  const double rinv = 1.0 / rho;
  const bool last = (j == K - 1);
  const double dj = dl[j];
  double lo, hi;
  int o;
  if (last) {
    double s = 0.0;
    for (int i = lane; i < K; i += width) s += w[i] * w[i];
    s = cx.sum(s);
    o = j; lo = 0.0; hi = rho * s;
  } else {
    const double mid = 0.5 * (dl[j + 1] - dj);
    double s = 0.0;
    for (int i = lane; i < K; i += width) s += w[i] * w[i] / ((dl[i] - dj) - mid);
    s = rinv + cx.sum(s);
    if (s >= 0.0) { o = j; lo = 0.0; hi = mid; }
    else { o = j + 1; lo = -mid; hi = 0.0; }
  }
  const double dorg = dl[o];
  // ---- initial guess: the nearest pole(s) exact, the other poles frozen at the far end of the bracket
  double mu;
  if (last) {
    const double at = 0.5 * hi;
    double s = 0.0;
    for (int i = lane; i < j; i += width) s += w[i] * w[i] / ((dl[i] - dorg) - at);
    const double rest = rinv + cx.sum(s);
    mu = (rest <= 0.0) ? hi : fmin(hi, fmax(w[j] * w[j] / rest, 0.0));
    if (!(lo < mu && mu < hi)) mu = 0.5 * hi;
  } else {
    const double far = (o == j) ? hi : lo;
    double s = 0.0;
    for (int i = lane; i < K; i += width)
      if (i != j && i != j + 1) s += w[i] * w[i] / ((dl[i] - dorg) - far);
    const double c = rinv + cx.sum(s);
    const double d1 = dl[j] - dorg, d2 = dl[j + 1] - dorg, aa = w[j] * w[j], bb = w[j + 1] * w[j + 1];
    // c + aa/(d1 - mu) + bb/(d2 - mu) = 0  <=>  qa mu^2 + qb mu + qc = 0
    const double qa = c, qb = -(c * (d1 + d2) + aa + bb), qc = c * d1 * d2 + aa * d2 + bb * d1;
    mu = 0.5 * (lo + hi);
    if (qa != 0.0) {
      const double disc = qb * qb - 4.0 * qa * qc;
      if (disc >= 0.0) {
        const double q = -0.5 * (qb + copysign(sqrt(disc), qb));
        const double r1 = q / qa, r2 = (q != 0.0) ? qc / q : INFINITY;
        if (lo < r1 && r1 < hi) mu = r1;
        if (lo < r2 && r2 < hi) mu = r2;
      }
    } else if (qb != 0.0) {
      const double r = -qc / qb;
      if (lo < r && r < hi) mu = r;
    }
  }
  // ---- safeguarded rational iteration
  int it = 0;
  for (; it < 100; it++) {
    double psi = 0.0, phi = 0.0, dpsi = 0.0, dphi = 0.0, asum = 0.0;
    for (int i = lane; i < K; i += width) {
      const double dlt = (dl[i] - dorg) - mu;
      const double t = w[i] / dlt;
      const double term = w[i] * t;
      if (i <= j) { psi += term; dpsi += t * t; }
      else { phi += term; dphi += t * t; }
      asum += fabs(term);
    }
    psi = cx.sum(psi); phi = cx.sum(phi); dpsi = cx.sum(dpsi); dphi = cx.sum(dphi); asum = cx.sum(asum);
    const double f = rinv + psi + phi;
    const double err = kDcEps * (8.0 * asum + rinv + fabs(mu) * (dpsi + dphi));
    if (fabs(f) <= err) break;
    if (f < 0.0) lo = fmax(lo, mu);
    else hi = fmin(hi, mu);
    if (hi - lo <= 2.0 * kDcEps * fmax(fabs(lo), fabs(hi))) { mu = 0.5 * (lo + hi); break; }
    const double D1 = (dl[j] - dorg) - mu;
    double eta;
    if (last) {
      const double S = dpsi * D1 * D1, s = psi - dpsi * D1;
      const double c = rinv + s;
      eta = (c != 0.0) ? D1 + S / c : INFINITY;   // c + S/(D1 - eta) = 0
    } else {
      const double D2 = (dl[j + 1] - dorg) - mu;
      const double S = dpsi * D1 * D1, s = psi - dpsi * D1;
      const double R = dphi * D2 * D2, r = phi - dphi * D2;
      const double c = rinv + s + r;
      const double qa = c, qb = c * (D1 + D2) + S + R, qc = D1 * D2 * f;
      const double disc = sqrt(fabs(qb * qb - 4.0 * qa * qc));
      if (qa == 0.0) eta = (qb != 0.0) ? qc / qb : INFINITY;
      else if (qb <= 0.0) eta = (qb - disc) / (2.0 * qa);
      else eta = 2.0 * qc / (qb + disc);
    }
    double nw = mu + eta;
    if (!(lo < nw && nw < hi)) nw = 0.5 * (lo + hi);   // also catches NaN / inf
    mu = nw;
  }
  o_out = o;
  mu_out = mu;
  return it;
}

}  // namespace kc
