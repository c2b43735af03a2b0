/*
 * okcma.c — CPU ORACLE (test infrastructure only; see okcma.h).
 *
 * Restates /root/reference/source/modules/solver/optimizer/CMAES/CMAES.cpp.base loop by loop.
 * "ref:" comments give the reference lines each block follows. Build: see oracle/Makefile
 * (gcc -O2 -ffp-contract=off, so a*b+c is never fused — the reference's release build targets
 * generic x86-64 without FMA, pyproject.toml:24-30).
 */
#define _GNU_SOURCE
#include "okcma.h"
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * gsl_rng_mt19937 (GSL 2.6 rng/mt.c — not in tree; the published MT19937 of Matsumoto & Nishimura,
 * seeded with init_genrand as gsl_rng_set does) and gsl_ran_gaussian (randist/gauss.c: polar
 * Box-Muller). Call sites: normal.cpp.base:32-35, distribution.cpp.base:32-39.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t mt[624];
  int mti;
} mt19937_t;

static void mt_seed(mt19937_t* s, uint32_t seed) {
  if (seed == 0) seed = 4357; /* gsl: the default seed is 4357 */
  s->mt[0] = seed;
  for (int i = 1; i < 624; i++) s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
  s->mti = 624;
}

static uint32_t mt_next(mt19937_t* s) {
  if (s->mti >= 624) {
    uint32_t* mt = s->mt;
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
      uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < 623; kk++) {
      uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s->mti = 0;
  }
  uint32_t k = s->mt[s->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680u;
  k ^= (k << 15) & 0xefc60000u;
  k ^= (k >> 18);
  return k;
}

static double mt_uniform(mt19937_t* s) { return mt_next(s) / 4294967296.0; }
static double mt_uniform_pos(mt19937_t* s) {
  double x;
  do x = mt_uniform(s);
  while (x == 0);
  return x;
}

/* gsl_ran_gaussian(rng, sigma) */
static double mt_gaussian(mt19937_t* s, double sigma) {
  double x, y, r2;
  do {
    x = -1 + 2 * mt_uniform_pos(s);
    y = -1 + 2 * mt_uniform_pos(s);
    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0);
  return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

void okcma_mt19937_gaussian(uint64_t seed, uint64_t skip, uint64_t count, double* out) {
  mt19937_t s;
  mt_seed(&s, (uint32_t)seed);
  for (uint64_t i = 0; i < skip; i++) (void)mt_gaussian(&s, 1.0);
  for (uint64_t i = 0; i < count; i++) out[i] = mt_gaussian(&s, 1.0);
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11) — restatement of the DEVICE generator (K1) so tests can
 * check its integer stream bit-for-bit and its normals to a few ulp. Counter layout:
 *   ctr = { column pair p = d/2, z-row index, resampling attempt, generation }, key = seed.
 * ---------------------------------------------------------------------------------------- */
void okcma_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u64_to_unit_open(uint32_t lo, uint32_t hi) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 12) * 0x1.0p-52 + 0x1.0p-53; /* exact, in [2^-53, 1-2^-53] */
}

static void philox_normal_pair(uint64_t seed, uint32_t generation, uint32_t attempt, uint64_t row, uint32_t pair, double* z0, double* z1) {
  uint32_t ctr[4] = {pair, (uint32_t)row, attempt, generation};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t r[4];
  okcma_philox4x32_10(ctr, key, r);
  double u1 = u64_to_unit_open(r[0], r[1]);
  double u2 = u64_to_unit_open(r[2], r[3]);
  double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincos(2.0 * M_PI * u2, &s, &c);
  *z0 = rad * c;
  *z1 = rad * s;
}

/* Uniform stream of the discrete mutations (the reference draws them from its _uniformGenerator, CMAES.cpp.base:520-529):
 * a second Philox stream, key = { seed_lo, seed_hi ^ "DISC" }, ctr = { block, sample index, resampling attempt, generation };
 * draw k is half (k & 1) of block k >> 1. Same function on the device (constraints.cu / update.cu discrete kernels). */
static double philox_uniform(uint64_t seed, uint32_t generation, uint32_t attempt, uint64_t sample, uint32_t k) {
  uint32_t ctr[4] = {k >> 1, (uint32_t)sample, attempt, generation};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x44495343u};
  uint32_t r[4];
  okcma_philox4x32_10(ctr, key, r);
  return (k & 1u) ? u64_to_unit_open(r[2], r[3]) : u64_to_unit_open(r[0], r[1]);
}

static void philox_normal_row(uint64_t seed, uint32_t generation, uint32_t attempt, uint64_t row, uint64_t n, double* z) {
  for (uint64_t p = 0; 2 * p < n; p++) {
    double a, b;
    philox_normal_pair(seed, generation, attempt, row, (uint32_t)p, &a, &b);
    z[2 * p] = a;
    if (2 * p + 1 < n) z[2 * p + 1] = b;
  }
}

void okcma_philox_normal(uint64_t seed, uint64_t generation, uint64_t row_begin, uint64_t rows, uint64_t n, double* z_out) {
  for (uint64_t i = 0; i < rows; i++) philox_normal_row(seed, (uint32_t)generation, 0, row_begin + i, n, z_out + i * n);
}

/* ------------------------------------------------------------------------------------------
 * Built-in objectives (the batched device conduit's functions; definitions from
 * examples/optimization/stochastic/_model/model.py). Summation order is the device's
 * canonical one: 32 lane-strided partial sums, then an xor-butterfly (16,8,4,2,1), so that
 * polynomial objectives are bit-identical between this oracle and the CUDA kernel.
 * ---------------------------------------------------------------------------------------- */
static double butterfly32(double p[32]) {
  for (int off = 16; off >= 1; off >>= 1) {
    double q[32];
    for (int l = 0; l < 32; l++) q[l] = p[l] + p[l ^ off];
    memcpy(p, q, sizeof(q));
  }
  return p[0];
}

static double objective_one(int objective, uint64_t n, const double* x, const double* coef) {
  double p[32], p2[32];
  for (int l = 0; l < 32; l++) p[l] = p2[l] = 0.0;
  switch (objective) {
    case KCMA_OBJ_NEG_SPHERE:
    case KCMA_OBJ_NEG_SUMSQ:
      for (uint64_t i = 0; i < n; i++) p[i & 31] += x[i] * x[i];
      return objective == KCMA_OBJ_NEG_SPHERE ? -0.5 * butterfly32(p) : -butterfly32(p);
    case KCMA_OBJ_NEG_ELLIPSOID:
      for (uint64_t i = 0; i < n; i++) {
        double t = x[i] * x[i];
        p[i & 31] += coef[i] * t;
      }
      return -butterfly32(p);
    case KCMA_OBJ_NEG_ROSENBROCK:
      for (uint64_t i = 0; i + 1 < n; i++) {
        double a = x[i] * x[i];
        double b = x[i + 1] - a;
        double c = b * b;
        double d = 100.0 * c;
        double e = 1.0 - x[i];
        double f = e * e;
        p[i & 31] += d + f;
      }
      return -butterfly32(p);
    case KCMA_OBJ_NEG_ACKLEY: {
      const double c = 2.0 * M_PI;
      for (uint64_t i = 0; i < n; i++) {
        p[i & 31] += x[i] * x[i];
        p2[i & 31] += cos(c * x[i]);
      }
      double sum1 = butterfly32(p) / (double)n;
      double sum2 = butterfly32(p2) / (double)n;
      double r1 = 20.0 * exp(-0.2 * sqrt(sum1));
      double r2 = exp(sum2);
      return r1 + r2 - 20.0 - exp(1.0);
    }
    case KCMA_OBJ_NEG_SPHERE_SIN2:
      for (uint64_t i = 0; i < n; i++) {
        double s = sin(x[i]);
        double a = x[i] * x[i];
        double b = s * s;
        p[i & 31] += a + b;
      }
      return -butterfly32(p);
    default: return NAN;
  }
}

/* dF/dx of the built-in objectives (the reference takes it from the user model: sample["Gradient"],
 * examples/optimization/stochastic/_model/model.py:10-63). */
static void objective_gradient_one(int objective, uint64_t n, const double* x, const double* coef, double* g) {
  switch (objective) {
    case KCMA_OBJ_NEG_SPHERE: for (uint64_t i = 0; i < n; i++) g[i] = -x[i]; return;
    case KCMA_OBJ_NEG_SUMSQ: for (uint64_t i = 0; i < n; i++) g[i] = -2.0 * x[i]; return;
    case KCMA_OBJ_NEG_ELLIPSOID: for (uint64_t i = 0; i < n; i++) g[i] = -2.0 * coef[i] * x[i]; return;
    case KCMA_OBJ_NEG_ROSENBROCK:
      for (uint64_t i = 0; i < n; i++) {
        double d = 0.0;
        if (i + 1 < n) d += -400.0 * x[i] * (x[i + 1] - x[i] * x[i]) - 2.0 * (1.0 - x[i]);
        if (i > 0) d += 200.0 * (x[i] - x[i - 1] * x[i - 1]);
        g[i] = -d;
      }
      return;
    case KCMA_OBJ_NEG_ACKLEY: {
      const double c = 2.0 * M_PI;
      double s1 = 0.0, s2 = 0.0;
      for (uint64_t i = 0; i < n; i++) { s1 += x[i] * x[i]; s2 += cos(c * x[i]); }
      const double r = sqrt(s1 / (double)n);
      const double e1 = 20.0 * exp(-0.2 * r), e2 = exp(s2 / (double)n);
      for (uint64_t i = 0; i < n; i++) {
        const double a = r > 0.0 ? e1 * (-0.2) * x[i] / ((double)n * r) : 0.0;
        const double b = e2 * (-c * sin(c * x[i])) / (double)n;
        g[i] = a + b;
      }
      return;
    }
    case KCMA_OBJ_NEG_SPHERE_SIN2: for (uint64_t i = 0; i < n; i++) g[i] = -(2.0 * x[i] + 2.0 * sin(x[i]) * cos(x[i])); return;
    default: for (uint64_t i = 0; i < n; i++) g[i] = NAN;
  }
}

void okcma_objective_gradient(int objective, uint64_t n, uint64_t rows, const double* x, const double* coef, double* g_out) {
  for (uint64_t i = 0; i < rows; i++) objective_gradient_one(objective, n, x + i * n, coef, g_out + i * n);
}

void okcma_objective(int objective, uint64_t n, uint64_t rows, const double* x, const double* coef, double* f_out) {
  for (uint64_t i = 0; i < rows; i++) f_out[i] = objective_one(objective, n, x + i * n, coef);
}

/* ------------------------------------------------------------------------------------------
 * sort_index — ref: CMAES.cpp.base:940-950. The reference uses an UNSTABLE std::sort with
 * comparator vec[i1] > vec[i2]; the order among equal values is a libstdc++ detail. Definition
 * used here and on the device: descending value, ascending index among equals
 * (= std::stable_sort with the same comparator). Bottom-up merge sort.
 * ---------------------------------------------------------------------------------------- */
void okcma_sort_index(const double* f, uint64_t n, uint64_t* idx) {
  uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1));
  for (uint64_t i = 0; i < n; i++) idx[i] = i;
  uint64_t *src = idx, *dst = tmp;
  for (uint64_t w = 1; w < n; w *= 2) {
    for (uint64_t lo = 0; lo < n; lo += 2 * w) {
      uint64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      uint64_t a = lo, b = mid, o = lo;
      while (a < mid && b < hi) {
        if (f[src[b]] > f[src[a]]) dst[o++] = src[b++]; /* strictly greater only: stable */
        else dst[o++] = src[a++];
      }
      while (a < mid) dst[o++] = src[a++];
      while (b < hi) dst[o++] = src[b++];
    }
    uint64_t* t = src; src = dst; dst = t;
  }
  if (src != idx) memcpy(idx, src, sizeof(uint64_t) * n);
  free(tmp);
}

/* ------------------------------------------------------------------------------------------
 * eigen — ref: CMAES.cpp.base:896-938 calls gsl_eigen_symmv + gsl_eigen_symmv_sort(ABS_ASC)
 * (GSL 2.6 eigen/symmv.c: Householder tridiagonalisation + implicit symmetric QR). Restated as
 * Householder reduction with accumulated transforms (EISPACK tred2) + implicit-shift QL (tql2).
 * NOT bit-identical to GSL (eigenvector signs / last bits differ); acceptance is by residuals
 * ||Q diag(w) Q^T - C||, ||Q^T Q - I|| and eigenvalue agreement with the fixture (tests).
 * Output: w ascending by |w|; q row-major with eigenvectors as COLUMNS (q[j*n+i] = j-th
 * component of i-th vector, ref :926-930).
 * ---------------------------------------------------------------------------------------- */
static void tred2(int n, double* V, double* d, double* e) {
  for (int j = 0; j < n; j++) d[j] = V[(n - 1) * n + j];
  for (int i = n - 1; i > 0; i--) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; k++) scale += fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; j++) {
        d[j] = V[(i - 1) * n + j];
        V[i * n + j] = 0.0;
        V[j * n + i] = 0.0;
      }
    } else {
      for (int k = 0; k < i; k++) {
        d[k] /= scale;
        h += d[k] * d[k];
      }
      double f = d[i - 1];
      double g = sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h = h - f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; j++) e[j] = 0.0;
      for (int j = 0; j < i; j++) {
        f = d[j];
        V[j * n + i] = f;
        g = e[j] + V[j * n + j] * f;
        for (int k = j + 1; k <= i - 1; k++) {
          g += V[k * n + j] * d[k];
          e[k] += V[k * n + j] * f;
        }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; j++) {
        e[j] /= h;
        f += e[j] * d[j];
      }
      double hh = f / (h + h);
      for (int j = 0; j < i; j++) e[j] -= hh * d[j];
      for (int j = 0; j < i; j++) {
        f = d[j];
        g = e[j];
        for (int k = j; k <= i - 1; k++) V[k * n + j] -= (f * e[k] + g * d[k]);
        d[j] = V[(i - 1) * n + j];
        V[i * n + j] = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; i++) {
    V[(n - 1) * n + i] = V[i * n + i];
    V[i * n + i] = 1.0;
    double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; k++) d[k] = V[k * n + i + 1] / h;
      for (int j = 0; j <= i; j++) {
        double g = 0.0;
        for (int k = 0; k <= i; k++) g += V[k * n + i + 1] * V[k * n + j];
        for (int k = 0; k <= i; k++) V[k * n + j] -= g * d[k];
      }
    }
    for (int k = 0; k <= i; k++) V[k * n + i + 1] = 0.0;
  }
  for (int j = 0; j < n; j++) {
    d[j] = V[(n - 1) * n + j];
    V[(n - 1) * n + j] = 0.0;
  }
  V[(n - 1) * n + n - 1] = 1.0;
  e[0] = 0.0;
}

static int tql2(int n, double* V, double* d, double* e) {
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = 0x1.0p-52;
  for (int l = 0; l < n; l++) {
    double t = fabs(d[l]) + fabs(e[l]);
    if (t > tst1) tst1 = t;
    int m = l;
    while (m < n) {
      if (fabs(e[m]) <= eps * tst1) break;
      m++;
    }
    if (m >= n) m = n - 1;
    if (m > l) {
      int iter = 0;
      do {
        if (++iter > 120) return 1;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c;
        double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; i--) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; k++) {
            h = V[k * n + i + 1];
            V[k * n + i + 1] = s * V[k * n + i] + c * h;
            V[k * n + i] = c * V[k * n + i] - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (fabs(e[l]) > eps * tst1);
    }
    d[l] = d[l] + f;
    e[l] = 0.0;
  }
  return 0;
}

int okcma_eigen(uint64_t n64, const double* c, double* w, double* q) {
  int n = (int)n64;
  double* V = (double*)malloc(sizeof(double) * n * n);
  double* e = (double*)malloc(sizeof(double) * n);
  double* d = (double*)malloc(sizeof(double) * n);
  /* ref :908-913 symmetrise from the lower triangle */
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) V[i * n + j] = V[j * n + i] = c[i * n + j];
  tred2(n, V, d, e);
  int rc = tql2(n, V, d, e);
  /* GSL_EIGEN_SORT_ABS_ASC (selection sort on |w|, columns swapped together) */
  int* ord = (int*)malloc(sizeof(int) * n);
  for (int i = 0; i < n; i++) ord[i] = i;
  for (int i = 0; i < n - 1; i++) {
    int k = i;
    for (int j = i + 1; j < n; j++)
      if (fabs(d[ord[j]]) < fabs(d[ord[k]])) k = j;
    int t = ord[i]; ord[i] = ord[k]; ord[k] = t;
  }
  for (int i = 0; i < n; i++) {
    w[i] = d[ord[i]];
    /* Sign convention (eigenvectors are defined up to sign; GSL's choice is an implementation detail): the
     * component of largest magnitude (first one on ties) is made positive. The device solver does the same, so
     * free-running device and oracle runs draw the same samples from the same z. */
    int jm = 0;
    for (int j = 1; j < n; j++)
      if (fabs(V[j * n + ord[i]]) > fabs(V[jm * n + ord[i]])) jm = j;
    const double sg = V[jm * n + ord[i]] < 0 ? -1.0 : 1.0;
    for (int j = 0; j < n; j++) q[j * n + i] = sg * V[j * n + ord[i]];
  }
  free(ord); free(d); free(e); free(V);
  return rc;
}

/* sampleSingle for a batch — ref: CMAES.cpp.base:494-513 (full-covariance branch). */
void okcma_sample(uint64_t n, uint64_t rows, const double* z, const double* b, const double* d,
                  const double* mean, double sigma, double* y_out, double* x_out) {
  double* aux = (double*)malloc(sizeof(double) * n);
  for (uint64_t i = 0; i < rows; i++) {
    for (uint64_t k = 0; k < n; k++) aux[k] = d[k] * z[i * n + k];
    for (uint64_t k = 0; k < n; k++) {
      double acc = 0.0;
      for (uint64_t e = 0; e < n; e++) acc += b[k * n + e] * aux[e];
      y_out[i * n + k] = acc;
      if (x_out) x_out[i * n + k] = mean[k] + sigma * acc;
    }
  }
  free(aux);
}

/* rank-mu accumulation of adaptC — ref: CMAES.cpp.base:703-704 without the scalar factors:
 * P[d][e] = sum_k w_k * t[k][d] * t[k][e], k ascending, lower triangle mirrored. */
void okcma_rank_mu(uint64_t n, uint64_t rows, const double* t, const double* w, double* p) {
  for (uint64_t d = 0; d < n; d++)
    for (uint64_t e = 0; e <= d; e++) {
      double acc = 0.0;
      for (uint64_t k = 0; k < rows; k++) acc += w[k] * t[k * n + d] * t[k * n + e];
      p[d * n + e] = p[e * n + d] = acc;
    }
}

/* ------------------------------------------------------------------------------------------
 * Solver state — field names mirror CMAES.hpp members (generated from CMAES.config).
 * ---------------------------------------------------------------------------------------- */
struct okcma {
  kcma_cfg cfg;
  uint64_t N;
  /* variables (optimizer.config:45-82) */
  double *lower, *upper, *initial_value, *initial_sd, *min_sd_update;
  double *obj_coef, *con_shift;
  /* termination criteria */
  double tc_max_infeasible_resamplings_d; /* as given */
  double tc_max_condition, tc_min_sd, tc_max_sd, tc_max_value, tc_min_value_diff;
  double tc_max_model_evaluations, tc_max_generations;
  /* generation counter: the reference's _k->_currentGeneration as seen by the NEXT runGeneration */
  uint64_t gen;
  uint64_t model_evaluation_count;
  /* internal settings */
  int is_viability_regime, has_constraints;
  uint64_t cur_lambda, cur_mu, s_max, mu_max;
  double *mu_weights;
  double effective_mu, sigma_cumulation_factor, damp_factor, cumulative_covariance, chi_square_number;
  double sigma, trace;
  double *X;  /* Sample Population s_max x N */
  double *BDZ; /* BDZ Matrix s_max x N */
  double *aux_bdz;
  double *value_vector;
  uint64_t *sorting_index;
  double *C, *C_aux, *B, *B_aux, *D, *D_aux;
  double *mean, *mean_old, *mean_update, *pc, *ps;
  double ps_l2norm;
  double *best_ever_variables, *current_best_variables;
  double best_ever_value, previous_best_ever_value, previous_best_value, current_best_value;
  double optimizer_previous_best_value; /* base-class copy, never written -> 0.0 */
  uint64_t infeasible_sample_count, resampled_parameter_count;
  double max_diag_c, min_diag_c, max_eig, min_eig;
  double cur_min_sd, cur_max_sd;
  /* constraints */
  uint64_t n_con;
  int64_t best_valid_sample;
  double global_success_rate, cov_adaption_factor, normal_vector_learning_rate;
  uint64_t cov_adaptation_count, max_violation_count, constraint_evaluation_count;
  double *viability_boundaries;
  uint64_t *violation_counts;
  double *con_evals;       /* n_con x s_max */
  unsigned char *viability_indicator; /* n_con x s_max */
  double *normal_approx;   /* n_con x N */
  double *best_con_evals;
  /* rng */
  int rng_kind; /* 0 = MT19937+GSL gaussian (reference stream), 1 = Philox (device stream) */
  mt19937_t mt;
  uint64_t* philox_attempt; /* per z-row attempt counter within the generation */
  /* injection */
  double *inj_z; uint64_t inj_z_rows, inj_z_used; int have_inj_z;
  int have_inj_bd;
  int have_inj_f;
  double* gradients; int have_inj_grad;   /* _gradients (CMAES.cpp.base:82-86), lambda x N */
  /* discrete variables (:44-50, 101-107) */
  int has_discrete;
  double *granularity, *masking_matrix, *masking_matrix_sigma, *discrete_mutations;
  uint64_t n_mask, n_discrete_mutations;
  double chi_square_number_discrete_mutations;
  int skip_sampling; /* BDZ or X injected for this generation */
  /* callbacks */
  okcma_objective_fn obj_fn; void* obj_user;
  okcma_constraints_fn con_fn; void* con_user;
  /* messages */
  char err[1024];
  char warn[4096];
  char reason[1024];
};

static char g_create_err[1024];

static int fail(okcma_t* h, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h ? h->err : g_create_err, 1024, fmt, ap);
  va_end(ap);
  return 1;
}

static void warnf(okcma_t* h, const char* fmt, ...) {
  size_t l = strlen(h->warn);
  if (l > sizeof(h->warn) - 256) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h->warn + l, sizeof(h->warn) - l, fmt, ap);
  va_end(ap);
}

const char* okcma_last_error(const okcma_t* h) { return h ? h->err : g_create_err; }
const char* okcma_take_warnings(okcma_t* h) {
  static char buf[4096];
  strcpy(buf, h->warn);
  h->warn[0] = 0;
  return buf;
}
void okcma_set_objective_callback(okcma_t* h, okcma_objective_fn fn, void* user) { h->obj_fn = fn; h->obj_user = user; }
void okcma_set_constraints_callback(okcma_t* h, okcma_constraints_fn fn, void* user) { h->con_fn = fn; h->con_user = user; }

void okcma_cfg_defaults(kcma_cfg* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = KCMA_ABI_VERSION;
  c->mu_type = KCMA_MU_LOGARITHMIC;
  c->initial_sigma_cumulation_factor = -1.0;
  c->initial_damp_factor = -1.0;
  c->initial_cumulative_covariance = -1.0;
  c->viability_population_size = 2;
  c->max_covariance_matrix_corrections = 1000000;
  c->target_success_rate = 0.1818;
  c->covariance_matrix_adaption_strength = 0.1;
  c->normal_vector_learning_rate = -1.0;
  c->global_success_learning_rate = 0.2;
  c->nranks = 1;
  c->use_gradient_information = 0;
  c->gradient_step_size = 0.01;
}

static double* dalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }
static double* dcopy(const double* src, size_t n, double fill) {
  double* p = dalloc(n);
  for (size_t i = 0; i < n; i++) p[i] = src ? src[i] : fill;
  return p;
}

/* ref: CMAES.cpp.base:233-284 */
static int init_mu_weights(okcma_t* h, uint64_t numsamplesmu) {
  const uint64_t N = h->N;
  double* w = h->mu_weights;
  switch (h->cfg.mu_type) {
    case KCMA_MU_LINEAR: for (uint64_t i = 0; i < numsamplesmu; i++) w[i] = (double)(numsamplesmu - i); break;
    case KCMA_MU_EQUAL: for (uint64_t i = 0; i < numsamplesmu; i++) w[i] = 1.; break;
    case KCMA_MU_LOGARITHMIC:
      for (uint64_t i = 0; i < numsamplesmu; i++) w[i] = log(fmax((double)numsamplesmu, 0.5 * h->cur_lambda) + 0.5) - log(i + 1.);
      break;
    case KCMA_MU_PROPORTIONAL: for (uint64_t i = 0; i < numsamplesmu; i++) w[i] = 1.; break;
    default: return fail(h, "Invalid setting of Mu Type (%d) (Linear, Equal, Logarithmic, or Proportional accepted).", h->cfg.mu_type);
  }
  double s1 = 0.0, s2 = 0.0;
  for (uint64_t i = 0; i < numsamplesmu; i++) {
    s1 += w[i];
    s2 += w[i] * w[i];
  }
  h->effective_mu = s1 * s1 / s2;
  for (uint64_t i = 0; i < numsamplesmu; i++) w[i] /= s1;

  if ((h->cfg.initial_cumulative_covariance <= 0) || (h->cfg.initial_cumulative_covariance > 1))
    h->cumulative_covariance = (4.0 + h->effective_mu / (1.0 * N)) / (N + 4.0 + 2.0 * h->effective_mu / (1.0 * N));
  else
    h->cumulative_covariance = h->cfg.initial_cumulative_covariance;

  h->sigma_cumulation_factor = h->cfg.initial_sigma_cumulation_factor;
  if (h->sigma_cumulation_factor <= 0 || h->sigma_cumulation_factor >= 1) {
    if (h->has_constraints)
      h->sigma_cumulation_factor = sqrt(h->effective_mu) / (sqrt(h->effective_mu) + sqrt((double)N));
    else
      h->sigma_cumulation_factor = (h->effective_mu + 2.0) / (N + h->effective_mu + 3.0);
  }
  h->damp_factor = h->cfg.initial_damp_factor;
  if (h->damp_factor <= 0.0)
    h->damp_factor = (1.0 + 2 * fmax(0.0, sqrt((h->effective_mu - 1.0) / (N + 1.0)) - 1)) + h->sigma_cumulation_factor;
  return 0;
}

/* ref: CMAES.cpp.base:286-313 (only diagonals are written) */
static void init_covariance(okcma_t* h) {
  const uint64_t N = h->N;
  h->trace = 0.0;
  for (uint64_t i = 0; i < N; ++i) h->trace += h->initial_sd[i] * h->initial_sd[i];
  h->sigma = sqrt(h->trace / N);
  for (uint64_t i = 0; i < N; ++i) {
    h->B[i * N + i] = 1.0;
    h->C[i * N + i] = h->D[i] = h->initial_sd[i] * sqrt(N / h->trace);
    h->C[i * N + i] *= h->C[i * N + i];
  }
  double mn = h->D[0], mx = h->D[0];
  for (uint64_t i = 1; i < N; i++) {
    if (h->D[i] < mn) mn = h->D[i];
    if (h->D[i] > mx) mx = h->D[i];
  }
  h->min_eig = mn * mn;
  h->max_eig = mx * mx;
  h->max_diag_c = h->C[0];
  for (uint64_t i = 1; i < N; ++i)
    if (h->max_diag_c < h->C[i * N + i]) h->max_diag_c = h->C[i * N + i];
  h->min_diag_c = h->C[0];
  for (uint64_t i = 1; i < N; ++i)
    if (h->min_diag_c > h->C[i * N + i]) h->min_diag_c = h->C[i * N + i];
}

/* ref: CMAES.cpp.base:14-184 */
int okcma_create(const kcma_cfg* cfg, okcma_t** out) {
  *out = NULL;
  if (!cfg || cfg->abi_version != KCMA_ABI_VERSION) return fail(NULL, "kcma_cfg ABI version mismatch");
  if (cfg->n == 0) return fail(NULL, "Optimization Evaluation problems require at least one variable.\n");
  okcma_t* h = (okcma_t*)calloc(1, sizeof(okcma_t));
  h->cfg = *cfg;
  const uint64_t N = h->N = cfg->n;
  h->lower = dcopy(cfg->lower_bound, N, -INFINITY);
  h->upper = dcopy(cfg->upper_bound, N, INFINITY);
  h->initial_value = dcopy(cfg->initial_value, N, NAN);
  h->initial_sd = dcopy(cfg->initial_stddev, N, NAN);
  h->min_sd_update = dcopy(cfg->min_stddev_update, N, 0.0);
  h->granularity = dcopy(cfg->granularity, N, 0.0);
  for (uint64_t i = 0; i < N; i++) { /* ref :44-50 */
    if (h->granularity[i] < 0.0) { fail(NULL, "Negative granularity for variable %zu.\n", (size_t)i); okcma_destroy(h); return 1; }
    if (h->granularity[i] > 0.0) h->has_discrete = 1;
  }
  if (h->has_discrete && cfg->n_constraints > 0) { fail(NULL, "discrete variables together with constraints are not supported"); okcma_destroy(h); return 1; }
  h->obj_coef = dalloc(N);
  for (uint64_t i = 0; i < N; i++)
    h->obj_coef[i] = cfg->objective_coef ? cfg->objective_coef[i] : (N > 1 ? pow(10.0, 6.0 * (double)i / (double)(N - 1)) : 1.0);
  h->n_con = cfg->constraint_family == KCMA_CON_NONE && !cfg->n_constraints ? 0 : cfg->n_constraints;
  h->con_shift = dcopy(cfg->constraint_shift, h->n_con, 0.0);
  h->cfg.lower_bound = h->cfg.upper_bound = h->cfg.initial_value = h->cfg.initial_stddev = h->cfg.min_stddev_update = NULL;
  h->cfg.objective_coef = h->cfg.constraint_shift = NULL;

  /* termination defaults: CMAES.config:503-509, optimizer.config:122-127, solver.config:42-47 */
  h->tc_max_condition = INFINITY; h->tc_min_sd = -INFINITY; h->tc_max_sd = INFINITY;
  h->tc_max_value = INFINITY; h->tc_min_value_diff = -INFINITY;
  h->tc_max_model_evaluations = 1e9; h->tc_max_generations = 1e10;
  h->gen = 1;

  h->best_ever_value = -INFINITY;
  h->previous_best_ever_value = h->previous_best_value = h->current_best_value = h->best_ever_value;
  h->optimizer_previous_best_value = 0.0;

  uint64_t lambda = cfg->population_size, mu = cfg->mu_value, vlambda = cfg->viability_population_size, vmu = cfg->viability_mu_value;
  if (cfg->use_gradient_information && cfg->gradient_step_size <= 0.) { /* ref :86 */
    fail(NULL, "Gradient Step Size must be larger than 0.0 (is %f)", cfg->gradient_step_size); goto bad;
  }
  if (lambda == 1) { fail(NULL, "'Population Size' must be larger 1."); goto bad; }
  if (lambda == 0) { fail(NULL, "'Population Size' must be larger 1."); goto bad; }
  if (mu == 0) mu = lambda / 2;
  if (vmu == 0) vmu = vlambda / 2;
  h->cfg.mu_value = mu; h->cfg.viability_mu_value = vmu;
  h->s_max = lambda > vlambda ? lambda : vlambda;
  h->mu_max = mu > vmu ? mu : vmu;
  h->chi_square_number = sqrt((double)N) * (1. - 1. / (4. * N) + 1. / (21. * N * N));
  h->has_constraints = h->n_con > 0;
  h->is_viability_regime = h->has_constraints;
  if (h->is_viability_regime) { h->cur_lambda = vlambda; h->cur_mu = vmu; }
  else { h->cur_lambda = lambda; h->cur_mu = mu; }

  h->X = dalloc(h->s_max * N); h->BDZ = dalloc(h->s_max * N); h->aux_bdz = dalloc(N);
  if (h->cfg.use_gradient_information) h->gradients = dalloc(h->s_max * N); /* ref :82-85 */
  h->cfg.gradient_step_size = (double)(float)h->cfg.gradient_step_size; /* _gradientStepSize is a float (CMAES.hpp:61) */
  if (h->has_discrete) { /* ref :101-107 */
    h->masking_matrix = dalloc(N); h->masking_matrix_sigma = dalloc(N); h->discrete_mutations = dalloc(h->s_max * N);
    h->n_mask = 0; h->n_discrete_mutations = 0;
  }
  h->chi_square_number_discrete_mutations = sqrt((double)N) * (1. - 1. / (4. * N) + 1. / (21. * N * N)); /* ref :34 */
  h->value_vector = dalloc(h->s_max);
  h->sorting_index = (uint64_t*)calloc(h->s_max, sizeof(uint64_t));
  h->C = dalloc(N * N); h->C_aux = dalloc(N * N); h->B = dalloc(N * N); h->B_aux = dalloc(N * N);
  h->D = dalloc(N); h->D_aux = dalloc(N);
  h->mean = dalloc(N); h->mean_old = dalloc(N); h->mean_update = dalloc(N); h->pc = dalloc(N); h->ps = dalloc(N);
  h->best_ever_variables = dalloc(N); h->current_best_variables = dalloc(N);
  h->mu_weights = dalloc(h->mu_max);
  h->philox_attempt = (uint64_t*)calloc(h->s_max, sizeof(uint64_t));

  if (cfg->mirrored_sampling) {
    if (lambda % 2 == 1) { fail(NULL, "Mirrored Sampling can only be applied with an even Sample Population (is %zu)", (size_t)lambda); goto bad; }
    if (h->has_constraints) { fail(NULL, "Mirrored Sampling not applicable to problems with constraints"); goto bad; }
  }
  /* variable defaults, ref :111-126 */
  for (uint64_t i = 0; i < N; ++i) {
    if (!isfinite(h->initial_value[i])) {
      if (!isfinite(h->lower[i])) { fail(NULL, "'Initial Value' of variable \'X%zu\' not defined, and cannot be inferred because variable lower bound is not finite.\n", (size_t)i); goto bad; }
      if (!isfinite(h->upper[i])) { fail(NULL, "'Initial Value' of variable \'X%zu\' not defined, and cannot be inferred because variable upper bound is not finite.\n", (size_t)i); goto bad; }
      h->initial_value[i] = (h->upper[i] + h->lower[i]) * 0.5;
    }
    if (!isfinite(h->initial_sd[i])) {
      if (!isfinite(h->lower[i])) { fail(NULL, "Initial (Mean) Value of variable \'X%zu\' not defined, and cannot be inferred because variable lower bound is not finite.\n", (size_t)i); goto bad; }
      if (!isfinite(h->upper[i])) { fail(NULL, "Initial Standard Deviation \'X%zu\' not defined, and cannot be inferred because variable upper bound is not finite.\n", (size_t)i); goto bad; }
      h->initial_sd[i] = (h->upper[i] - h->lower[i]) * 0.3;
    }
  }
  if (h->has_constraints) {
    if ((cfg->global_success_learning_rate <= 0.0) || (cfg->global_success_learning_rate > 1.0)) { fail(NULL, "Invalid Global Success Learning Rate (%f), must be greater than 0.0 and less than 1.0\n", cfg->global_success_learning_rate); goto bad; }
    if ((cfg->target_success_rate <= 0.0) || (cfg->target_success_rate > 1.0)) { fail(NULL, "Invalid Target Success Rate (%f), must be greater than 0.0 and less than 1.0\n", cfg->target_success_rate); goto bad; }
    if (cfg->covariance_matrix_adaption_strength <= 0.0) { fail(NULL, "Invalid Adaption Size (%f), must be greater than 0.0\n", cfg->covariance_matrix_adaption_strength); goto bad; }
    h->global_success_rate = 0.5;
    h->best_valid_sample = -1;
    h->violation_counts = (uint64_t*)calloc(h->s_max, sizeof(uint64_t));
    h->viability_boundaries = dalloc(h->n_con);
    h->viability_indicator = (unsigned char*)calloc(h->n_con * h->s_max, 1);
    h->con_evals = dalloc(h->n_con * h->s_max);
    h->normal_approx = dalloc(h->n_con * N);
    h->best_con_evals = dalloc(h->n_con);
    h->normal_vector_learning_rate = 1.0 / (2.0 + N);
    h->cov_adaption_factor = cfg->covariance_matrix_adaption_strength / (N + 2.);
  } else {
    h->global_success_rate = -1.0;
    h->cov_adaption_factor = -1.0;
    h->best_valid_sample = 0;
    h->normal_vector_learning_rate = cfg->normal_vector_learning_rate;
  }
  h->cov_adaptation_count = 0;
  h->max_violation_count = 0;
  if (init_mu_weights(h, h->has_constraints ? vmu : mu)) { strcpy(g_create_err, h->err); goto bad; }
  init_covariance(h);
  h->infeasible_sample_count = 0;
  h->resampled_parameter_count = 0;
  h->ps_l2norm = 0.0;
  for (uint64_t i = 0; i < N; i++) h->mean[i] = h->mean_old[i] = h->initial_value[i];
  h->cur_min_sd = INFINITY;
  h->cur_max_sd = -INFINITY;
  mt_seed(&h->mt, (uint32_t)cfg->seed);
  *out = h;
  return 0;
bad:
  okcma_destroy(h);
  return 1;
}

void okcma_destroy(okcma_t* h) {
  if (!h) return;
  free(h->lower); free(h->upper); free(h->initial_value); free(h->initial_sd); free(h->min_sd_update);
  free(h->obj_coef); free(h->con_shift); free(h->mu_weights); free(h->X); free(h->BDZ); free(h->aux_bdz);
  free(h->value_vector); free(h->sorting_index); free(h->C); free(h->C_aux); free(h->B); free(h->B_aux);
  free(h->D); free(h->D_aux); free(h->mean); free(h->mean_old); free(h->mean_update); free(h->pc); free(h->ps);
  free(h->best_ever_variables); free(h->current_best_variables); free(h->viability_boundaries);
  free(h->violation_counts); free(h->con_evals); free(h->viability_indicator); free(h->normal_approx);
  free(h->best_con_evals); free(h->philox_attempt); free(h->inj_z); free(h->gradients);
  free(h->granularity); free(h->masking_matrix); free(h->masking_matrix_sigma); free(h->discrete_mutations);
  free(h);
}

/* ref: optimizer.cpp.base:5-14 */
static int is_sample_feasible(const okcma_t* h, const double* x) {
  for (uint64_t i = 0; i < h->N; i++) {
    if (!isfinite(x[i])) return 0;
    if (x[i] < h->lower[i]) return 0;
    if (x[i] > h->upper[i]) return 0;
  }
  return 1;
}

/* One row of N(0,1) draws for z-row `zrow` (sample index, or pair index when mirrored). */
static void draw_normals(okcma_t* h, uint64_t zrow, double* z) {
  const uint64_t N = h->N;
  if (h->have_inj_z && h->inj_z_used < h->inj_z_rows) {
    memcpy(z, h->inj_z + h->inj_z_used * N, sizeof(double) * N);
    h->inj_z_used++;
    return;
  }
  if (h->rng_kind == 0) {
    /* ref :449-450: dimension-minor draw order from one MT19937 stream */
    for (uint64_t d = 0; d < N; d++) z[d] = 0.0 + mt_gaussian(&h->mt, 1.0);
  } else {
    philox_normal_row(h->cfg.seed, (uint32_t)h->gen, (uint32_t)h->philox_attempt[zrow], zrow, N, z);
    h->philox_attempt[zrow]++;
  }
}

/* ref: CMAES.cpp.base:862-867 */
static void discretize(const okcma_t* h, double* x) {
  for (uint64_t d = 0; d < h->N; ++d)
    if (h->granularity[d] != 0.0) x[d] = round(x[d] / h->granularity[d]) * h->granularity[d];
}

/* ref: CMAES.cpp.base:515-544, then the discretize() of :453 / :478-481. `attempt` = resampling attempt of this draw. */
static void discrete_mutation(okcma_t* h, uint64_t i, uint32_t attempt) {
  const uint64_t N = h->N;
  uint32_t k = 0;
#define U() philox_uniform(h->cfg.seed, (uint32_t)h->gen, attempt, i, k++)
  if ((i + 1) < h->n_discrete_mutations) {
    const double p_geom = pow(0.7, 1.0 / h->n_mask);
    uint64_t select = (uint64_t)floor(U() * h->n_mask);
    for (uint64_t d = 0; d < N; ++d)
      if ((h->masking_matrix[d] == 1.0) && (select-- == 0)) {
        double dmutation = 1.0;
        while (U() > p_geom) dmutation += 1.0;
        dmutation *= h->granularity[d];
        if (U() > 0.5) dmutation *= -1.0;
        h->discrete_mutations[i * N + d] = dmutation;
        h->X[i * N + d] += dmutation;
      }
  } else if ((i + 1) == h->n_discrete_mutations) {
    for (uint64_t d = 0; d < N; ++d)
      if (h->granularity[d] != 0.0) {
        const double dmutation = round(h->best_ever_variables[d] / h->granularity[d]) * h->granularity[d] - h->X[i * N + d];
        h->discrete_mutations[i * N + d] = dmutation;
        h->X[i * N + d] += dmutation;
      }
  }
#undef U
  discretize(h, h->X + i * N);
}

/* ref: CMAES.cpp.base:834-860 */
static void update_discrete_mutation_matrix(okcma_t* h) {
  const uint64_t N = h->N;
  uint64_t entries = N + 1; /* +1 to prevent 0-ness */
  for (uint64_t d = 0; d < N; ++d) h->masking_matrix_sigma[d] = 1.0;
  for (uint64_t d = 0; d < N; ++d)
    if (h->sigma * sqrt(h->C[d * N + d]) / sqrt(h->sigma_cumulation_factor) < 0.2 * h->granularity[d]) {
      h->masking_matrix_sigma[d] = 0.0;
      entries--;
    }
  h->chi_square_number_discrete_mutations = sqrt((double)entries) * (1. - 1. / (4. * entries) + 1. / (21. * entries * entries));
  h->n_mask = 0;
  for (uint64_t d = 0; d < N; ++d) h->masking_matrix[d] = 0.0;
  for (uint64_t d = 0; d < N; ++d)
    if (2.0 * h->sigma * sqrt(h->C[d * N + d]) < h->granularity[d]) {
      h->masking_matrix[d] = 1.0;
      h->n_mask++;
    }
  h->n_discrete_mutations = (uint64_t)fmin(round(h->cfg.population_size / 10.0 + h->n_mask + 1), floor(h->cfg.population_size / 2.0) - 1);
  memset(h->discrete_mutations, 0, sizeof(double) * h->s_max * N);
}

/* ref: CMAES.cpp.base:494-513 */
static void sample_single(okcma_t* h, uint64_t i, const double* z) {
  const uint64_t N = h->N;
  for (uint64_t d = 0; d < N; ++d) {
    if (h->cfg.diagonal_covariance) {
      h->BDZ[i * N + d] = h->D[d] * z[d];
      h->X[i * N + d] = h->mean[d] + h->sigma * h->BDZ[i * N + d];
    } else
      h->aux_bdz[d] = h->D[d] * z[d];
  }
  if (!h->cfg.diagonal_covariance)
    for (uint64_t d = 0; d < N; ++d) {
      h->BDZ[i * N + d] = 0.0;
      for (uint64_t e = 0; e < N; ++e) h->BDZ[i * N + d] += h->B[d * N + e] * h->aux_bdz[e];
      h->X[i * N + d] = h->mean[d] + h->sigma * h->BDZ[i * N + d];
    }
}

/* ref: CMAES.cpp.base:896-938 */
static int eigen(okcma_t* h, const double* M, double* diag, double* Q) {
  const uint64_t N = h->N;
  if (h->cfg.diagonal_covariance) {
    memset(Q, 0, sizeof(double) * N * N);
    for (uint64_t i = 0; i < N; ++i) Q[i * N + i] = 1.;
    for (uint64_t i = 0; i < N; ++i) diag[i] = M[i * N + i];
    return 0;
  }
  return okcma_eigen(N, M, diag, Q);
}

/* ref: CMAES.cpp.base:869-890 */
static void update_eigensystem(okcma_t* h, const double* M) {
  const uint64_t N = h->N;
  if (h->have_inj_bd) { /* parity hook: B, D were injected for this generation */
    h->have_inj_bd = 0;
    double mn = INFINITY, mx = -INFINITY;
    for (uint64_t d = 0; d < N; d++) { double ev = h->D[d] * h->D[d]; if (ev < mn) mn = ev; if (ev > mx) mx = ev; }
    h->min_eig = mn; h->max_eig = mx;
    return;
  }
  eigen(h, M, h->D_aux, h->B_aux);
  double mn = h->D_aux[0], mx = h->D_aux[0];
  for (uint64_t d = 1; d < N; d++) { if (h->D_aux[d] < mn) mn = h->D_aux[d]; if (h->D_aux[d] > mx) mx = h->D_aux[d]; }
  if (mn <= 0.0) {
    warnf(h, "Min Eigenvalue smaller or equal 0.0 (%+6.3e) after Eigen decomp (no update possible).\n", h->min_eig);
    return;
  }
  for (uint64_t d = 0; d < N; ++d) h->D_aux[d] = sqrt(h->D_aux[d]);
  h->min_eig = mn;
  h->max_eig = mx;
  for (uint64_t d = 0; d < N; ++d) h->D[d] = h->D_aux[d];
  memcpy(h->B, h->B_aux, sizeof(double) * N * N);
}

/* ref: CMAES.cpp.base:439-492 */
static void prepare_generation(okcma_t* h) {
  const uint64_t N = h->N;
  update_eigensystem(h, h->C);
  memset(h->philox_attempt, 0, sizeof(uint64_t) * h->s_max);
  if (h->skip_sampling) return;
  double* z = dalloc(N);
  double* z2 = dalloc(N);
  const uint64_t maxres = h->cfg.max_infeasible_resamplings;
  if (!h->cfg.mirrored_sampling)
    for (uint64_t i = 0; i < h->cur_lambda; ++i) {
      int feas;
      do {
        draw_normals(h, i, z);
        sample_single(h, i, z);
        if (h->has_discrete) discrete_mutation(h, i, h->rng_kind ? (uint32_t)(h->philox_attempt[i] - 1) : 0u);
        feas = is_sample_feasible(h, h->X + i * N);
        h->infeasible_sample_count += feas ? 0 : 1;
      } while (!feas && (h->infeasible_sample_count < maxres));
    }
  else
    for (uint64_t i = 0; i < h->cur_lambda; i += 2) {
      int feas;
      do {
        draw_normals(h, i / 2, z);
        for (uint64_t d = 0; d < N; ++d) z2[d] = -z[d];
        sample_single(h, i, z);
        sample_single(h, i + 1, z2);
        if (h->has_discrete) {
          const uint32_t att = h->rng_kind ? (uint32_t)(h->philox_attempt[i / 2] - 1) : 0u;
          discrete_mutation(h, i, att);
          discrete_mutation(h, i + 1, att);
        }
        int f1 = is_sample_feasible(h, h->X + i * N);
        if (!f1) h->infeasible_sample_count++;
        int f2 = is_sample_feasible(h, h->X + (i + 1) * N);
        if (!f2) h->infeasible_sample_count++;
        feas = f1 || f2;
      } while (!feas && (h->infeasible_sample_count < maxres));
    }
  free(z); free(z2);
}

/* Built-in constraint families (the batched device conduit's constraint functions). */
static int eval_constraints(okcma_t* h, const double* x, double* g) {
  if (h->con_fn) h->con_fn(h->con_user, x, h->N, g, h->n_con);
  else if (h->cfg.constraint_family == KCMA_CON_HALFSPACE)
    for (uint64_t c = 0; c < h->n_con; c++) g[c] = -(x[c % h->N] - h->con_shift[c]);
  else return fail(h, "no constraint functions defined");
  h->constraint_evaluation_count++;
  /* ref: optimization.cpp.base:19-20 */
  for (uint64_t c = 0; c < h->n_con; c++)
    if (!isfinite(g[c])) return fail(h, "Non finite value of constraint evaluation %lu detected: %f\n", (unsigned long)c, g[c]);
  return 0;
}

/* ref: CMAES.cpp.base:315-345 */
static int check_mean_and_set_regime(okcma_t* h) {
  if (!h->is_viability_regime) return 0;
  double* g = dalloc(h->n_con);
  if (eval_constraints(h, h->mean, g)) { free(g); return 1; }
  for (uint64_t c = 0; c < h->n_con; c++)
    if (g[c] > 0.0) { free(g); return 0; }
  free(g);
  h->is_viability_regime = 0;
  for (uint64_t c = 0; c < h->n_con; c++) h->viability_boundaries[c] = 0;
  h->cur_lambda = h->cfg.population_size;
  h->cur_mu = h->cfg.mu_value;
  if (init_mu_weights(h, h->cur_mu)) return 1;
  init_covariance(h);
  return 0;
}

/* ref: CMAES.cpp.base:347-385 */
static int update_constraints(okcma_t* h) {
  const uint64_t N = h->N, S = h->s_max;
  double* g = dalloc(h->n_con);
  for (uint64_t i = 0; i < h->cur_lambda; i++) {
    h->violation_counts[i] = 0;
    if (eval_constraints(h, h->X + i * N, g)) { free(g); return 1; }
    for (uint64_t c = 0; c < h->n_con; c++) h->con_evals[c * S + i] = g[c];
  }
  free(g);
  h->max_violation_count = 0;
  for (uint64_t c = 0; c < h->n_con; c++) {
    double maxviolation = 0.0;
    for (uint64_t i = 0; i < h->cur_lambda; ++i) {
      if (h->con_evals[c * S + i] > maxviolation) maxviolation = h->con_evals[c * S + i];
      if (h->gen == 1 && h->is_viability_regime) h->viability_boundaries[c] = maxviolation;
      if (h->con_evals[c * S + i] > h->viability_boundaries[c] + 1e-12) h->violation_counts[i]++;
      if (h->violation_counts[i] > h->max_violation_count) h->max_violation_count = h->violation_counts[i];
    }
  }
  return 0;
}

/* ref: CMAES.cpp.base:387-424 */
static int re_evaluate_constraints(okcma_t* h) {
  const uint64_t N = h->N, S = h->s_max;
  h->max_violation_count = 0;
  double* g = dalloc(h->n_con);
  for (uint64_t i = 0; i < h->cur_lambda; ++i)
    if (h->violation_counts[i] > 0) {
      if (eval_constraints(h, h->X + i * N, g)) { free(g); return 1; }
      h->violation_counts[i] = 0;
      for (uint64_t c = 0; c < h->n_con; c++) {
        h->con_evals[c * S + i] = g[c];
        if (g[c] > h->viability_boundaries[c] + 1e-12) {
          h->viability_indicator[c * S + i] = 1;
          h->violation_counts[i]++;
        } else
          h->viability_indicator[c * S + i] = 0;
      }
      if (h->violation_counts[i] > h->max_violation_count) h->max_violation_count = h->violation_counts[i];
    }
  free(g);
  return 0;
}

/* ref: CMAES.cpp.base:774-832 */
static int handle_constraints(okcma_t* h) {
  const uint64_t N = h->N, S = h->s_max;
  double* z = dalloc(N);
  while (h->max_violation_count > 0) {
    memcpy(h->C_aux, h->C, sizeof(double) * N * N);
    for (uint64_t i = 0; i < h->cur_lambda; ++i)
      if (h->violation_counts[i] > 0) {
        for (uint64_t c = 0; c < h->n_con; c++)
          if (h->viability_indicator[c * S + i]) {
            h->cov_adaptation_count++;
            if (h->cov_adaptation_count > h->cfg.max_covariance_matrix_corrections) {
              warnf(h, "Exiting adaption loop, max adaptions (%zu) reached.\n", (size_t)h->cfg.max_covariance_matrix_corrections);
              free(z);
              return 0;
            }
            double v2 = 0;
            double* v = h->normal_approx + c * N;
            const double lr = h->normal_vector_learning_rate;
            for (uint64_t d = 0; d < N; ++d) {
              v[d] = (1.0 - lr) * v[d] + lr * h->BDZ[i * N + d];
              v2 += v[d] * v[d];
            }
            const double beta = h->cov_adaption_factor;
            const double cnt = (double)h->violation_counts[i];
            for (uint64_t d = 0; d < N; ++d)
              for (uint64_t e = 0; e < N; ++e)
                h->C_aux[d * N + e] = h->C_aux[d * N + e] - ((beta * beta * v[d] * v[e]) / (v2 * cnt * cnt));
          }
      }
    update_eigensystem(h, h->C_aux);
    for (uint64_t i = 0; i < h->cur_lambda; ++i)
      if (h->violation_counts[i] > 0) {
        int feas;
        do {
          h->resampled_parameter_count++;
          draw_normals(h, i, z);
          sample_single(h, i, z);
          feas = is_sample_feasible(h, h->X + i * N);
        } while (!feas && h->resampled_parameter_count < h->cfg.max_infeasible_resamplings);
      }
    if (re_evaluate_constraints(h)) { free(z); return 1; }
  }
  free(z);
  return 0;
}

/* ref: CMAES.cpp.base:426-437 */
static void update_viability_boundaries(okcma_t* h) {
  const uint64_t S = h->s_max;
  for (uint64_t c = 0; c < h->n_con; c++) {
    double maxviolation = 0.0;
    for (uint64_t i = 0; i < h->cur_mu; ++i)
      if (h->con_evals[c * S + h->sorting_index[i]] > maxviolation) maxviolation = h->con_evals[c * S + h->sorting_index[i]];
    h->viability_boundaries[c] = fmax(0.0, fmin(h->viability_boundaries[c], 0.5 * (maxviolation + h->viability_boundaries[c])));
  }
}

/* ref: CMAES.cpp.base:690-718 */
static void adapt_c(okcma_t* h, int hsig) {
  const uint64_t N = h->N;
  const double ccov1 = 2.0 / (pow(N + 1.3, 2) + h->effective_mu);
  const double ccovmu = fmin(1.0 - ccov1, 2.0 * (h->effective_mu - 2. + 1. / h->effective_mu) / (pow(N + 2.0, 2) + h->effective_mu));
  const double sigmasquare = h->sigma * h->sigma;
  const double cc = h->cumulative_covariance;
  for (uint64_t d = 0; d < N; ++d)
    for (uint64_t e = h->cfg.diagonal_covariance ? d : 0; e <= d; ++e) {
      h->C[d * N + e] = (1 - ccov1 - ccovmu) * h->C[d * N + e] + ccov1 * (h->pc[d] * h->pc[e] + (1 - hsig) * cc * (2. - cc) * h->C[d * N + e]);
      for (uint64_t k = 0; k < h->cur_mu; ++k)
        h->C[d * N + e] += ccovmu * h->mu_weights[k] * (h->X[h->sorting_index[k] * N + d] - h->mean_old[d]) * (h->X[h->sorting_index[k] * N + e] - h->mean_old[e]) / sigmasquare;
      if (e < d) h->C[e * N + d] = h->C[d * N + e];
    }
  h->max_diag_c = h->min_diag_c = h->C[0];
  for (uint64_t d = 1; d < N; ++d) {
    if (h->max_diag_c < h->C[d * N + d]) h->max_diag_c = h->C[d * N + d];
    else if (h->min_diag_c > h->C[d * N + d]) h->min_diag_c = h->C[d * N + d];
  }
}

/* ref: CMAES.cpp.base:720-761 (discrete-variable branch :730-735 out of scope) */
static void update_sigma(okcma_t* h) {
  if (h->has_constraints && h->is_viability_regime) {
    h->global_success_rate = (1 - h->cfg.global_success_learning_rate) * h->global_success_rate;
    h->sigma *= exp((h->global_success_rate - (h->cfg.target_success_rate / (1.0 - h->cfg.target_success_rate)) * (1 - h->global_success_rate)) / h->damp_factor);
  } else if (h->has_discrete) { /* ref :730-734 */
    double pathL2 = 0.0;
    for (uint64_t d = 0; d < h->N; ++d) pathL2 += h->masking_matrix_sigma[d] * h->ps[d] * h->ps[d];
    h->sigma *= exp(h->sigma_cumulation_factor / h->damp_factor * (sqrt(pathL2) / h->chi_square_number_discrete_mutations - 1.));
  } else {
    h->sigma *= exp(h->sigma_cumulation_factor / h->damp_factor * (h->ps_l2norm / h->chi_square_number - 1.));
  }
  if (h->cfg.mu_value > 1 && h->current_best_value == h->value_vector[h->sorting_index[h->cur_mu - 1]]) {
    h->sigma *= exp(0.2 + h->sigma_cumulation_factor / h->damp_factor);
    warnf(h, "Sigma increased due to equal function values.\n");
  }
  const double upper = sqrt(h->trace / h->N);
  if (h->sigma > upper) {
    if (h->cfg.is_sigma_bounded) h->sigma = upper;
  }
}

/* ref: CMAES.cpp.base:763-772 */
static void numerical_error_treatment(okcma_t* h) {
  const uint64_t N = h->N;
  for (uint64_t d = 0; d < N; ++d)
    if (h->sigma * sqrt(h->C[d * N + d]) < h->min_sd_update[d]) {
      h->sigma = (h->min_sd_update[d]) / sqrt(h->C[d * N + d]) * exp(0.05 + h->sigma_cumulation_factor / h->damp_factor);
      warnf(h, "Sigma increased due to minimal standard deviation.\n");
    }
}

/* ref: CMAES.cpp.base:547-688 */
static int update_distribution(okcma_t* h) {
  const uint64_t N = h->N, S = h->s_max;
  okcma_sort_index(h->value_vector, h->cur_lambda, h->sorting_index);

  if (!h->has_constraints || h->is_viability_regime)
    h->best_valid_sample = (int64_t)h->sorting_index[0];
  else {
    h->best_valid_sample = -1;
    for (uint64_t i = 0; i < h->cur_lambda; i++)
      if (h->violation_counts[h->sorting_index[i]] == 0) h->best_valid_sample = (int64_t)h->sorting_index[i]; /* no break: SURVEY Q3 */
  }
  if (h->best_valid_sample < 0) return fail(h, "no valid sample in generation (reference reads out of bounds here, CMAES.cpp.base:565)");

  h->previous_best_value = h->current_best_value;
  h->current_best_value = h->value_vector[h->best_valid_sample];
  for (uint64_t d = 0; d < N; ++d) h->current_best_variables[d] = h->X[h->best_valid_sample * N + d];

  if (h->current_best_value > h->best_ever_value || h->gen == 1) {
    h->previous_best_ever_value = h->best_ever_value;
    h->best_ever_value = h->current_best_value;
    for (uint64_t d = 0; d < N; ++d) h->best_ever_variables[d] = h->current_best_variables[d];
    if (h->has_constraints)
      for (uint64_t c = 0; c < h->n_con; c++) h->best_con_evals[c] = h->con_evals[c * S + h->best_valid_sample];
  }

  if (h->cfg.mu_type == KCMA_MU_PROPORTIONAL) { /* ref :584-600 */
    double valueSum = 0.;
    for (uint64_t i = 0; i < h->cur_mu; ++i) {
      const double value = h->value_vector[h->sorting_index[i]];
      h->mu_weights[i] = value;
      valueSum += value;
    }
    for (uint64_t i = 0; i < h->cur_mu; ++i) h->mu_weights[i] /= valueSum;
  }

  for (uint64_t d = 0; d < N; ++d) {
    h->mean_old[d] = h->mean[d];
    h->mean[d] = 0.;
    for (uint64_t i = 0; i < h->cur_mu; ++i) h->mean[d] += h->mu_weights[i] * h->X[h->sorting_index[i] * N + d];
  }
  if (h->cfg.use_gradient_information) { /* ref :611-621 (l2update is computed there but never used) */
    for (uint64_t d = 0; d < N; ++d)
      for (uint64_t i = 0; i < h->cur_mu; ++i)
        h->mean[d] += h->mu_weights[i] * h->cfg.gradient_step_size / sqrt((double)N) * h->gradients[h->sorting_index[i] * N + d];
  }
  for (uint64_t d = 0; d < N; ++d) h->mean_update[d] = (h->mean[d] - h->mean_old[d]) / h->sigma;

  for (uint64_t d = 0; d < N; ++d) {
    double sum = 0.0;
    if (h->cfg.diagonal_covariance) sum = h->mean_update[d];
    else
      for (uint64_t e = 0; e < N; ++e) sum += h->B[e * N + d] * h->mean_update[e];
    h->aux_bdz[d] = sum / h->D[d];
  }
  h->ps_l2norm = 0.0;
  const double cs = h->sigma_cumulation_factor, cc = h->cumulative_covariance;
  for (uint64_t d = 0; d < N; ++d) {
    double sum = 0.0;
    if (h->cfg.diagonal_covariance) sum = h->aux_bdz[d];
    else
      for (uint64_t e = 0; e < N; ++e) sum += h->B[d * N + e] * h->aux_bdz[e];
    h->ps[d] = (1. - cs) * h->ps[d] + sqrt(cs * (2. - cs) * h->effective_mu) * sum;
    h->ps_l2norm += pow(h->ps[d], 2.0);
  }
  h->ps_l2norm = sqrt(h->ps_l2norm);

  const int hsig = (1.4 + 2.0 / (N + 1) > h->ps_l2norm / sqrt(1. - pow(1. - cs, 2.0 * (1.0 + h->gen))) / h->chi_square_number);
  for (uint64_t d = 0; d < N; ++d)
    h->pc[d] = (1. - cc) * h->pc[d] + hsig * sqrt(cc * (2. - cc) * h->effective_mu) * h->mean_update[d];

  adapt_c(h, hsig);
  if (h->has_discrete) update_discrete_mutation_matrix(h); /* ref :668 */
  if (h->has_constraints && h->is_viability_regime) update_viability_boundaries(h);
  update_sigma(h);
  numerical_error_treatment(h);

  h->cur_min_sd = INFINITY;
  h->cur_max_sd = -INFINITY;
  for (uint64_t i = 0; i < N; ++i) {
    h->cur_min_sd = fmin(h->cur_min_sd, h->sigma * sqrt(h->C[i * N + i]));
    h->cur_max_sd = fmax(h->cur_max_sd, h->sigma * sqrt(h->C[i * N + i]));
  }
  return 0;
}

/* ---- public generation-loop API ---------------------------------------------------------- */
int okcma_ask(okcma_t* h) {
  if (h->has_constraints && check_mean_and_set_regime(h)) return 1;
  prepare_generation(h);
  h->have_inj_z = 0;
  h->skip_sampling = 0;
  if (h->has_constraints) {
    if (update_constraints(h)) return 1;
    if (handle_constraints(h)) return 1;
  }
  return 0;
}

int okcma_eval(okcma_t* h) {
  const uint64_t N = h->N;
  h->model_evaluation_count += h->cur_lambda; /* ref :214 */
  if (h->cfg.use_gradient_information && !h->have_inj_grad) { /* ref :199-200, 226-228: "Evaluate With Gradients" */
    if (h->obj_fn || h->have_inj_f)
      return fail(h, "Use Gradient Information: inject the gradients (KCMA_INJ_GRAD) together with the values of an external model");
    for (uint64_t i = 0; i < h->cur_lambda; i++) objective_gradient_one(h->cfg.objective, N, h->X + i * N, h->obj_coef, h->gradients + i * N);
  }
  h->have_inj_grad = 0;
  if (h->have_inj_f) { h->have_inj_f = 0; return 0; }
  for (uint64_t i = 0; i < h->cur_lambda; i++) {
    double f;
    if (h->obj_fn) h->obj_fn(h->obj_user, h->X + i * N, N, &f);
    else f = objective_one(h->cfg.objective, N, h->X + i * N, h->obj_coef);
    /* ref: optimization.cpp.base:32-33 */
    if (!isfinite(f)) return fail(h, "Non finite value of function evaluation detected: %f\n", f);
    h->value_vector[i] = f;
  }
  return 0;
}

int okcma_tell(okcma_t* h) {
  int rc = update_distribution(h);
  if (!rc) h->gen++;
  return rc;
}

int okcma_run_generation(okcma_t* h) {
  if (okcma_ask(h)) return 1;
  if (okcma_eval(h)) return 1;
  return okcma_tell(h);
}

/* ref: CMAES.cpp:1903-1933 -> optimizer.cpp:188-206 -> solver.cpp:92-110 (short-circuit chain) */
int okcma_check_termination(okcma_t* h, int* finished, const char** reason) {
  int fin = 0;
  h->reason[0] = 0;
  const uint64_t gen = h->gen;
  const uint64_t maxres = h->cfg.max_infeasible_resamplings;
  if (gen > 1 && ((maxres > 0) && (h->infeasible_sample_count >= maxres))) { strcat(h->reason, "CMAES['Max Infeasible Resamplings'];"); fin = 1; }
  if (gen > 1 && (h->max_eig >= h->tc_max_condition * h->min_eig)) { strcat(h->reason, "CMAES['Max Condition Covariance Matrix'];"); fin = 1; }
  if (gen > 1 && (h->cur_min_sd <= h->tc_min_sd)) { strcat(h->reason, "CMAES['Min Standard Deviation'];"); fin = 1; }
  if (gen > 1 && (h->cur_max_sd >= h->tc_max_sd)) { strcat(h->reason, "CMAES['Max Standard Deviation'];"); fin = 1; }
  if (!fin) {
    if (gen > 1 && (+h->best_ever_value > h->tc_max_value)) { strcat(h->reason, "optimizer['Max Value'];"); fin = 1; }
    if (gen > 1 && (fabs(h->current_best_value - h->optimizer_previous_best_value) < h->tc_min_value_diff)) { strcat(h->reason, "optimizer['Min Value Difference Threshold'];"); fin = 1; }
    if (!fin) {
      if (h->tc_max_model_evaluations <= (double)h->model_evaluation_count) { strcat(h->reason, "solver['Max Model Evaluations'];"); fin = 1; }
      if ((double)gen > h->tc_max_generations) { strcat(h->reason, "solver['Max Generations'];"); fin = 1; }
    }
  }
  *finished = fin;
  if (reason) *reason = h->reason;
  return 0;
}

int okcma_run(okcma_t* h, uint64_t max_generations, uint64_t* done) {
  uint64_t n = 0;
  int fin = 0;
  while (n < max_generations) {
    okcma_check_termination(h, &fin, NULL);
    if (fin) break;
    if (okcma_run_generation(h)) { if (done) *done = n; return 1; }
    n++;
  }
  if (done) *done = n;
  return 0;
}

int okcma_inject(okcma_t* h, int kind, const double* src, size_t count) {
  const uint64_t N = h->N;
  switch (kind) {
    case KCMA_INJ_Z:
      if (count % N) return fail(h, "inject Z: count must be a multiple of N");
      free(h->inj_z);
      h->inj_z = dcopy(src, count, 0);
      h->inj_z_rows = count / N; h->inj_z_used = 0; h->have_inj_z = 1;
      return 0;
    case KCMA_INJ_BDZ:
      if (count != h->cur_lambda * N) return fail(h, "inject BDZ: expected %zu values", (size_t)(h->cur_lambda * N));
      memcpy(h->BDZ, src, sizeof(double) * count);
      for (uint64_t i = 0; i < count; i++) h->X[i] = h->mean[i % N] + h->sigma * h->BDZ[i];
      h->skip_sampling = 1;
      return 0;
    case KCMA_INJ_X:
      if (count != h->cur_lambda * N) return fail(h, "inject X: expected %zu values", (size_t)(h->cur_lambda * N));
      memcpy(h->X, src, sizeof(double) * count);
      h->skip_sampling = 1;
      return 0;
    case KCMA_INJ_F:
      if (count != h->cur_lambda) return fail(h, "inject F: expected %zu values", (size_t)h->cur_lambda);
      for (size_t i = 0; i < count; i++)
        if (!isfinite(src[i])) return fail(h, "Non finite value of function evaluation detected: %f\n", src[i]);
      memcpy(h->value_vector, src, sizeof(double) * count);
      h->have_inj_f = 1;
      return 0;
    case KCMA_INJ_GRAD:
      if (!h->cfg.use_gradient_information) return fail(h, "inject Gradients: Use Gradient Information is off");
      if (count != h->cur_lambda * N) return fail(h, "inject Gradients: expected %zu values", (size_t)(h->cur_lambda * N));
      memcpy(h->gradients, src, sizeof(double) * count);
      h->have_inj_grad = 1;
      return 0;
    case KCMA_INJ_BD:
      if (count != N * N + N) return fail(h, "inject BD: expected N*N+N values");
      memcpy(h->B, src, sizeof(double) * N * N);
      memcpy(h->D, src + N * N, sizeof(double) * N);
      h->have_inj_bd = 1;
      return 0;
  }
  return fail(h, "unknown injection kind %d", kind);
}

/* ---- key/value access --------------------------------------------------------------------- */
typedef struct { const char* key; double* p; size_t n; } arr_ref;

static int find_array(okcma_t* h, const char* key, arr_ref* r) {
  const uint64_t N = h->N;
#define A(K, P, CNT) if (!strcmp(key, K)) { r->key = K; r->p = (P); r->n = (CNT); return 1; }
  A("Covariance Matrix", h->C, N * N)
  A("Auxiliar Covariance Matrix", h->C_aux, N * N)
  A("Covariance Eigenvector Matrix", h->B, N * N)
  A("Auxiliar Covariance Eigenvector Matrix", h->B_aux, N * N)
  A("Axis Lengths", h->D, N)
  A("Auxiliar Axis Lengths", h->D_aux, N)
  A("Current Mean", h->mean, N)
  A("Previous Mean", h->mean_old, N)
  A("Mean Update", h->mean_update, N)
  A("Evolution Path", h->pc, N)
  A("Conjugate Evolution Path", h->ps, N)
  A("Auxiliar BDZ Matrix", h->aux_bdz, N)
  A("Mu Weights", h->mu_weights, h->cur_mu)
  A("Value Vector", h->value_vector, h->cur_lambda)
  A("BDZ Matrix", h->BDZ, h->cur_lambda * N)
  if (h->gradients) { A("Gradients", h->gradients, h->cur_lambda * N) }
  if (h->has_discrete) {
    A("Masking Matrix", h->masking_matrix, N) A("Masking Matrix Sigma", h->masking_matrix_sigma, N)
    A("Discrete Mutations", h->discrete_mutations, h->cur_lambda * N)
  }
  A("Sample Population", h->X, h->cur_lambda * N)
  A("Best Ever Variables", h->best_ever_variables, N)
  A("Current Best Variables", h->current_best_variables, N)
  A("Viability Boundaries", h->viability_boundaries, h->n_con)
  A("Normal Constraint Approximation", h->normal_approx, h->n_con * N)
  A("Best Constraint Evaluations", h->best_con_evals, h->n_con)
  A("Constraint Evaluations", h->con_evals, h->n_con * h->s_max)
  A("Objective Coefficients", h->obj_coef, N)
#undef A
  return 0;
}

int okcma_get_array(okcma_t* h, const char* key, double* out, size_t cap, size_t* count) {
  arr_ref r;
  if (!find_array(h, key, &r) || (!r.p && r.n)) return fail(h, "unknown array key '%s'", key);
  if (count) *count = r.n;
  if (!out) return 0;
  if (cap < r.n) return fail(h, "buffer too small for '%s' (%zu < %zu)", key, cap, r.n);
  memcpy(out, r.p, sizeof(double) * r.n);
  return 0;
}

int okcma_set_array(okcma_t* h, const char* key, const double* in, size_t count) {
  arr_ref r;
  if (!find_array(h, key, &r) || (!r.p && r.n)) return fail(h, "unknown array key '%s'", key);
  if (count != r.n) return fail(h, "size mismatch for '%s' (%zu != %zu)", key, count, r.n);
  memcpy(r.p, in, sizeof(double) * r.n);
  return 0;
}

int okcma_get_index_array(okcma_t* h, const char* key, uint64_t* out, size_t cap, size_t* count) {
  const uint64_t* p = NULL;
  size_t n = 0;
  if (!strcmp(key, "Sorting Index")) { p = h->sorting_index; n = h->cur_lambda; }
  else if (!strcmp(key, "Sample Constraint Violation Counts")) { p = h->violation_counts; n = h->has_constraints ? h->cur_lambda : 0; }
  else return fail(h, "unknown index key '%s'", key);
  if (count) *count = n;
  if (!out) return 0;
  if (cap < n) return fail(h, "buffer too small for '%s'", key);
  if (n) memcpy(out, p, sizeof(uint64_t) * n);
  return 0;
}

static double* find_scalar(okcma_t* h, const char* key) {
#define S(K, P) if (!strcmp(key, K)) return &(P);
  S("Sigma", h->sigma) S("Trace", h->trace) S("Effective Mu", h->effective_mu)
  S("Sigma Cumulation Factor", h->sigma_cumulation_factor) S("Damp Factor", h->damp_factor)
  S("Cumulative Covariance", h->cumulative_covariance) S("Chi Square Number", h->chi_square_number)
  S("Chi Square Number Discrete Mutations", h->chi_square_number_discrete_mutations)
  S("Conjugate Evolution Path L2 Norm", h->ps_l2norm)
  S("Best Ever Value", h->best_ever_value) S("Previous Best Ever Value", h->previous_best_ever_value)
  S("Previous Best Value", h->previous_best_value) S("Current Best Value", h->current_best_value)
  S("Maximum Diagonal Covariance Matrix Element", h->max_diag_c) S("Minimum Diagonal Covariance Matrix Element", h->min_diag_c)
  S("Maximum Covariance Eigenvalue", h->max_eig) S("Minimum Covariance Eigenvalue", h->min_eig)
  S("Current Min Standard Deviation", h->cur_min_sd) S("Current Max Standard Deviation", h->cur_max_sd)
  S("Global Success Rate", h->global_success_rate) S("Covariance Matrix Adaption Factor", h->cov_adaption_factor)
  S("Normal Vector Learning Rate", h->normal_vector_learning_rate)
  S("Termination Criteria/Max Condition Covariance Matrix", h->tc_max_condition)
  S("Termination Criteria/Min Standard Deviation", h->tc_min_sd)
  S("Termination Criteria/Max Standard Deviation", h->tc_max_sd)
  S("Termination Criteria/Max Value", h->tc_max_value)
  S("Termination Criteria/Min Value Difference Threshold", h->tc_min_value_diff)
  S("Termination Criteria/Max Model Evaluations", h->tc_max_model_evaluations)
  S("Termination Criteria/Max Generations", h->tc_max_generations)
#undef S
  return NULL;
}

int okcma_get_scalar(okcma_t* h, const char* key, double* out) {
  double* p = find_scalar(h, key);
  if (p) { *out = *p; return 0; }
#define U(K, V) if (!strcmp(key, K)) { *out = (double)(V); return 0; }
  U("Current Generation", h->gen - 1) U("Model Evaluation Count", h->model_evaluation_count)
  U("Number Of Discrete Mutations", h->n_discrete_mutations) U("Number Masking Matrix Entries", h->n_mask)
  U("Variable Count", h->N) U("Current Population Size", h->cur_lambda) U("Current Mu Value", h->cur_mu)
  U("Infeasible Sample Count", h->infeasible_sample_count) U("Resampled Parameter Count", h->resampled_parameter_count)
  U("Is Viability Regime", h->is_viability_regime) U("Has Constraints", h->has_constraints)
  U("Best Valid Sample", h->best_valid_sample) U("Covariance Matrix Adaptation Count", h->cov_adaptation_count)
  U("Max Constraint Violation Count", h->max_violation_count) U("Constraint Evaluation Count", h->constraint_evaluation_count)
  U("Termination Criteria/Max Infeasible Resamplings", h->cfg.max_infeasible_resamplings)
  U("Oracle/RNG Kind", h->rng_kind)
#undef U
  return fail(h, "unknown scalar key '%s'", key);
}

int okcma_set_scalar(okcma_t* h, const char* key, double v) {
  double* p = find_scalar(h, key);
  if (p) { *p = v; return 0; }
#define U(K, STMT) if (!strcmp(key, K)) { STMT; return 0; }
  U("Current Generation", h->gen = (uint64_t)v + 1)
  U("Model Evaluation Count", h->model_evaluation_count = (uint64_t)v)
  U("Infeasible Sample Count", h->infeasible_sample_count = (uint64_t)v)
  U("Resampled Parameter Count", h->resampled_parameter_count = (uint64_t)v)
  U("Covariance Matrix Adaptation Count", h->cov_adaptation_count = (uint64_t)v)
  U("Termination Criteria/Max Infeasible Resamplings", h->cfg.max_infeasible_resamplings = (uint64_t)v)
  U("Oracle/RNG Kind", h->rng_kind = (int)v)
  U("Oracle/MT19937 Skip Gaussians", for (uint64_t i = 0; i < (uint64_t)v; i++) (void)mt_gaussian(&h->mt, 1.0))
#undef U
  return fail(h, "unknown scalar key '%s'", key);
}
