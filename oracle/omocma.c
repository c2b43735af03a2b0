/*
 * omocma.c — CPU ORACLE of the multi-objective CMA-ES generation loop. TEST INFRASTRUCTURE ONLY (same rules as okcma.c / odea.c:
 * only tests/ may load it; the product never links or calls it).
 *
 * Plain-C restatement of /root/reference/source/modules/solver/optimizer/MOCMAES/MOCMAES.cpp.base, loop for loop (each function cites
 * the lines it follows; -O2 -ffp-contract=off), with the stated differences:
 *   - random numbers: the Philox streams of include/kmocma.h instead of the reference's sequential MT19937 streams;
 *   - GSL (an un-vendored dependency, subprojects/gsl.wrap: release-2-6) is restated: gsl_linalg_cholesky_decomp1 as the unblocked
 *     left-looking Cholesky (row sums in ascending index), gsl_ran_multivariate_gaussian as y = L z by rows (dtrmv, lower, no
 *     transpose: the strictly lower part summed in ascending index, the diagonal term added last), x = y + mean.
 * PARITY UNPINNED: the reference ships no MOCMAES trajectory, golden vector or statistical test (only a unit test of its
 * configuration checks, tests/unit/modules/solver/optimizers.cpp:2112-2180, and the example examples/optimization/multiobjective);
 * this oracle is anchored by the example's behaviour (the Pareto front of run-mocmaes.py, tests/test_oracle_mocma.py).
 */
#define _GNU_SOURCE /* sincos */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/kmocma.h"
#include "okcma.h"

typedef struct omocma {
  kmocma_cfg cfg;
  uint64_t N, lambda, mu, K, gen, model_evals, infeasible, nondom;
  double cc, ccov, cp, target;
  double *lower, *upper;
  /* current / previous / parent: X, sigma, C, pc, psucc; values */
  double *X[3], *S[3], *C[3], *P[3], *PS[3];
  double *F, *Fprev;
  uint64_t* parent_index;
  int* sorted;
  double *best_ever, *best_ever_x, *prev_best, *prev_best_x, *cur_best, *cur_best_x, *val_diff, *var_diff, *min_sd, *max_sd;
  double *coll_x, *coll_f; size_t coll_n, coll_cap;
  double tc_min_value_diff, tc_min_var_diff, tc_min_sd, tc_max_sd, tc_max_generations, tc_max_model_evaluations;
  kmocma_host_objective_fn obj_fn; void* obj_user;
  int have_inj_f;
  char err[512], reason[512];
} omocma_t;
enum { CUR = 0, PREV = 1, PAR = 2 };

static char g_err[512];
static int failf(omocma_t* h, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(h ? h->err : g_err, 512, fmt, ap);
  va_end(ap);
  return 1;
}
static double unit_open(uint32_t lo, uint32_t hi) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 12) * 0x1.0p-52 + 0x1.0p-53;
}
static void mo_block(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t r[4]) {
  uint32_t ctr[4] = {c0, c1, c2, c3};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x4d4f434du};   /* "MOCM" */
  okcma_philox4x32_10(ctr, key, r);
}

void omocma_cfg_defaults(kmocma_cfg* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = KMOCMA_ABI_VERSION;
  c->num_objectives = 2;
  c->evolution_path_adaption_strength = -1.0; c->covariance_learning_rate = -1.0;
  c->target_success_rate = 0.175; c->threshold_probability = 0.44; c->success_learning_rate = 0.08;
}

const char* omocma_last_error(const omocma_t* h) { return h ? h->err : g_err; }

void omocma_destroy(omocma_t* h) {
  if (!h) return;
  for (int s = 0; s < 3; s++) { free(h->X[s]); free(h->S[s]); free(h->C[s]); free(h->P[s]); free(h->PS[s]); }
  free(h->lower); free(h->upper); free(h->F); free(h->Fprev); free(h->parent_index); free(h->sorted);
  free(h->best_ever); free(h->best_ever_x); free(h->prev_best); free(h->prev_best_x); free(h->cur_best); free(h->cur_best_x);
  free(h->val_diff); free(h->var_diff); free(h->min_sd); free(h->max_sd); free(h->coll_x); free(h->coll_f);
  free(h);
}

/* setInitialConfiguration :10-144 */
int omocma_create(const kmocma_cfg* cfg, omocma_t** out) {
  if (!cfg || cfg->abi_version != KMOCMA_ABI_VERSION) return failf(NULL, "kmocma_cfg ABI version mismatch");
  if (cfg->n < 1) return failf(NULL, "no variables");
  if (cfg->num_objectives < 2 || cfg->num_objectives > KMOCMA_MAX_OBJECTIVES)
    return failf(NULL, "Problem requires multiple objectives, 'Num Objectives' is set to %zu\n.", (size_t)cfg->num_objectives);
  omocma_t* h = calloc(1, sizeof(*h));
  h->cfg = *cfg;
  const uint64_t N = h->N = cfg->n, K = h->K = cfg->num_objectives;
  uint64_t lambda = cfg->population_size, mu = cfg->mu_value;
  if (lambda == 0) lambda = (uint64_t)ceil(4. + floor(3. * log((double)N)));
  if (mu == 0) mu = (uint64_t)(lambda / 2.);
  if (mu > lambda) { free(h); return failf(NULL, "Number of parents ('Mu Value' %zu) must be smaller or equal with population size (%zu).\n", (size_t)mu, (size_t)lambda); }
  h->lambda = lambda; h->mu = mu; h->nondom = 1; h->gen = 1;
  h->cp = cfg->success_learning_rate; h->target = cfg->target_success_rate;
  if (h->cp <= 0. || h->cp > 1.) { free(h); return failf(NULL, "Invalid Global Success Learning Rate (%f), must be greater than 0.0 and less or equal to 1.0\n", cfg->success_learning_rate); }
  if (h->target <= 0. || h->target > 1.) { free(h); return failf(NULL, "Invalid Target Success Rate (%f), must be greater than 0.0 and less or equal to 1.0\n", cfg->target_success_rate); }
  h->lower = malloc(sizeof(double) * N); h->upper = malloc(sizeof(double) * N);
  double* sd = malloc(sizeof(double) * N);
  double trace = 0., min_sdev = INFINITY, max_sdev = -INFINITY;
  for (uint64_t i = 0; i < N; i++) {
    h->lower[i] = cfg->lower_bound ? cfg->lower_bound[i] : -INFINITY;
    h->upper[i] = cfg->upper_bound ? cfg->upper_bound[i] : INFINITY;
    const double iv = cfg->initial_value ? cfg->initial_value[i] : NAN;
    double s = cfg->initial_stddev ? cfg->initial_stddev[i] : NAN;
    if (!isfinite(iv) && (!isfinite(h->lower[i]) || !isfinite(h->upper[i]))) {
      free(sd); omocma_destroy(h);
      return failf(NULL, "Initial (Mean) Value of variable %zu not defined, and cannot be inferred because a variable bound is not finite.\n", (size_t)i);
    }
    if (!isfinite(s)) {
      if (!isfinite(h->lower[i]) || !isfinite(h->upper[i])) {
        free(sd); omocma_destroy(h);
        return failf(NULL, "Initial Standard Deviation of variable %zu not defined, and cannot be inferred because a variable bound is not finite.\n", (size_t)i);
      }
      s = (h->upper[i] - h->lower[i]) * 0.3;
    }
    sd[i] = s;
    trace += s * s;
    if (s < min_sdev) min_sdev = s;
    if (s > max_sdev) max_sdev = s;
  }
  const uint64_t rows[3] = {lambda, lambda, mu};
  for (int s = 0; s < 3; s++) {
    h->X[s] = calloc(rows[s] * N, sizeof(double)); h->S[s] = calloc(rows[s], sizeof(double)); h->C[s] = calloc(rows[s] * N * N, sizeof(double));
    h->P[s] = calloc(rows[s] * N, sizeof(double)); h->PS[s] = calloc(rows[s], sizeof(double));
  }
  h->F = malloc(sizeof(double) * lambda * K); h->Fprev = malloc(sizeof(double) * lambda * K);
  for (uint64_t i = 0; i < lambda * K; i++) h->F[i] = h->Fprev[i] = -INFINITY;
  h->parent_index = calloc(lambda, sizeof(uint64_t)); h->sorted = calloc(2 * lambda, sizeof(int));
  h->best_ever = malloc(sizeof(double) * K); h->prev_best = malloc(sizeof(double) * K); h->cur_best = malloc(sizeof(double) * K);
  h->best_ever_x = calloc(K * N, sizeof(double)); h->prev_best_x = calloc(K * N, sizeof(double)); h->cur_best_x = calloc(K * N, sizeof(double));
  h->val_diff = malloc(sizeof(double) * K); h->var_diff = malloc(sizeof(double) * K);
  for (uint64_t k = 0; k < K; k++) { h->best_ever[k] = h->prev_best[k] = h->cur_best[k] = -INFINITY; h->val_diff[k] = h->var_diff[k] = INFINITY; }
  h->min_sd = malloc(sizeof(double) * lambda); h->max_sd = malloc(sizeof(double) * lambda);
  for (uint64_t i = 0; i < lambda; i++) { h->min_sd[i] = min_sdev; h->max_sd[i] = max_sdev; }
  const double sigma0 = sqrt(trace / (double)N);
  for (uint64_t i = 0; i < mu; i++) {
    h->PS[PAR][i] = h->target;
    h->S[PAR][i] = sigma0;
    for (uint64_t d = 0; d < N; d++) h->C[PAR][i * N * N + d * N + d] = sd[d] * sd[d] / (sigma0 * sigma0);
  }
  free(sd);
  h->cc = cfg->evolution_path_adaption_strength < 0. ? 2. / ((double)N + 2.) : cfg->evolution_path_adaption_strength;
  h->ccov = cfg->covariance_learning_rate < 0. ? 2. / ((double)N * (double)N + 6.) : cfg->covariance_learning_rate;
  h->tc_min_value_diff = -INFINITY; h->tc_min_var_diff = -INFINITY; h->tc_min_sd = -INFINITY; h->tc_max_sd = INFINITY;
  h->tc_max_generations = 1e10; h->tc_max_model_evaluations = 1e9;
  *out = h;
  return 0;
}

static int feasible(const omocma_t* h, const double* x) {
  for (uint64_t d = 0; d < h->N; d++)
    if (x[d] < h->lower[d] || x[d] > h->upper[d]) return 0;
  return 1;
}

/* gsl_linalg_cholesky_decomp1 restated: L lower, row by row, sums in ascending index. Returns non-zero if not positive definite. */
static int cholesky(const double* A, uint64_t N, double* L) {
  memset(L, 0, sizeof(double) * N * N);
  for (uint64_t i = 0; i < N; i++)
    for (uint64_t j = 0; j <= i; j++) {
      double s = A[i * N + j];
      for (uint64_t k = 0; k < j; k++) s -= L[i * N + k] * L[j * N + k];
      if (i == j) { if (!(s > 0.)) return 1; L[i * N + i] = sqrt(s); }
      else L[i * N + j] = s / L[j * N + j];
    }
  return 0;
}

/* prepareGeneration :177-189 + sampleSingle :191-230 */
int omocma_ask(omocma_t* h) {
  const uint64_t N = h->N, lambda = h->lambda;
  memcpy(h->Fprev, h->F, sizeof(double) * lambda * h->K);
  memcpy(h->X[PREV], h->X[CUR], sizeof(double) * lambda * N); memcpy(h->S[PREV], h->S[CUR], sizeof(double) * lambda);
  memcpy(h->C[PREV], h->C[CUR], sizeof(double) * lambda * N * N); memcpy(h->P[PREV], h->P[CUR], sizeof(double) * lambda * N);
  memcpy(h->PS[PREV], h->PS[CUR], sizeof(double) * lambda);
  double* L = malloc(sizeof(double) * N * N);
  double* z = malloc(sizeof(double) * (N + 1));
  for (uint64_t i = 0; i < lambda; i++) {
    uint64_t p;
    if (h->mu == lambda) p = i;
    else {
      uint32_t r[4];
      mo_block(h->cfg.seed, 0u, (uint32_t)i, 0u, (uint32_t)h->gen, r);
      const double u = unit_open(r[0], r[1]);
      p = (uint64_t)floor((double)(h->mu < h->nondom ? h->mu : h->nondom) * u);
    }
    h->parent_index[i] = p;
    if (cholesky(h->C[PAR] + p * N * N, N, L)) { free(L); free(z); return failf(h, "Error during Cholesky decomposition of covariance matrix.\n"); }
    const double sig = h->S[PAR][p];
    double* x = h->X[CUR] + i * N;
    for (uint32_t attempt = 0;; attempt++) {
      for (uint64_t pr = 0; 2 * pr < N; pr++) {
        uint32_t r[4];
        mo_block(h->cfg.seed, (1u << 20) + (uint32_t)pr, (uint32_t)i, attempt, (uint32_t)h->gen, r);
        const double u1 = unit_open(r[0], r[1]), u2 = unit_open(r[2], r[3]);
        const double rad = sqrt(-2.0 * log(u1));
        double s, c;
        sincos(2.0 * M_PI * u2, &s, &c);
        z[2 * pr] = rad * c; z[2 * pr + 1] = rad * s;
      }
      for (uint64_t d = 0; d < N; d++) {
        double y = 0.;
        for (uint64_t e = 0; e < d; e++) y += (L[d * N + e] * sig) * z[e];
        y += z[d] * (L[d * N + d] * sig);
        x[d] = y + h->X[PAR][p * N + d];
      }
      if (feasible(h, x)) break;
      if (attempt > 1000000u) { free(L); free(z); return failf(h, "no feasible sample after 10^6 draws"); }
    }
    memcpy(h->C[CUR] + i * N * N, h->C[PAR] + p * N * N, sizeof(double) * N * N);
    h->S[CUR][i] = sig;
    memcpy(h->P[CUR] + i * N, h->P[PAR] + p * N, sizeof(double) * N);
    h->PS[CUR][i] = h->PS[PAR][p];
  }
  free(L); free(z);
  return 0;
}

/* examples/optimization/multiobjective/_model/model.py:5-38 */
static void builtin_objective(int id, const double* x, uint64_t n, double* f) {
  double r1 = 0., r2 = 0., r3 = 0.;
  for (uint64_t i = 0; i + 1 < n; i++) r1 += 100 * ((x[i + 1] - x[i] * x[i]) * (x[i + 1] - x[i] * x[i])) + (1 - x[i]) * (1 - x[i]);
  for (uint64_t i = 0; i < n; i++) r2 += x[i] * x[i];
  f[0] = -r1; f[1] = -r2;
  if (id == KMOCMA_OBJ_NEG_ROSENBROCK_AND_TWO_SPHERES) {
    for (uint64_t i = 0; i < n; i++) r3 += (x[i] - 2) * (x[i] - 2);
    f[2] = -r3;
  }
}

int omocma_set_host_objective(omocma_t* h, kmocma_host_objective_fn fn, void* user) { h->obj_fn = fn; h->obj_user = user; return 0; }
int omocma_inject_f(omocma_t* h, const double* f, size_t count) {
  if (count != h->lambda * h->K) return failf(h, "inject_f: %zu values for %zu x %zu", count, (size_t)h->lambda, (size_t)h->K);
  memcpy(h->F, f, sizeof(double) * count);
  h->have_inj_f = 1;
  return 0;
}

/* runGeneration :152-170 */
int omocma_eval(omocma_t* h) {
  h->model_evals += h->lambda;
  if (h->have_inj_f) { h->have_inj_f = 0; return 0; }
  if (h->obj_fn) h->obj_fn(h->obj_user, h->X[CUR], h->lambda, h->N, h->F, h->K);
  else if (h->cfg.objective == KMOCMA_OBJ_EXTERNAL) return failf(h, "objective is External: inject the values or set a host objective before eval");
  else {
    const uint64_t want = h->cfg.objective == KMOCMA_OBJ_NEG_ROSENBROCK_AND_TWO_SPHERES ? 3 : 2;
    if (h->K != want) return failf(h, "the built-in objective has %zu objectives, 'Num Objectives' is %zu", (size_t)want, (size_t)h->K);
    for (uint64_t i = 0; i < h->lambda; i++) builtin_objective(h->cfg.objective, h->X[CUR] + i * h->N, h->N, h->F + i * h->K);
  }
  for (uint64_t i = 0; i < h->lambda * h->K; i++)
    if (!isfinite(h->F[i])) return failf(h, "Non finite value of function evaluation detected: %f\n", h->F[i]);
  return 0;
}

/* sortSampleIndices :232-342 (values: n x K, larger is better) */
static void sort_sample_indices(const omocma_t* h, const double* values, size_t n, int* sorted) {
  const size_t K = h->K;
  size_t min_rank = n;
  size_t* rank = calloc(n, sizeof(size_t));
  size_t* max_nb = malloc(sizeof(size_t) * n);
  for (size_t r = n; r >= 1; --r) {
    size_t min_max_nb = K;
    memset(max_nb, 0, sizeof(size_t) * n);
    for (size_t i = 0; i < n; ++i)
      if (rank[i] == 0) {
        for (size_t j = 0; j < n; ++j)
          if (i != j && rank[j] == 0) {
            size_t nb = 0;
            for (size_t k = 0; k < K; ++k)
              if (values[i * K + k] < values[j * K + k]) nb++;
            if (nb > max_nb[i]) max_nb[i] = nb;
          }
        if (max_nb[i] < min_max_nb) min_max_nb = max_nb[i];
      }
    for (size_t i = 0; i < n; ++i)
      if (rank[i] == 0 && max_nb[i] == min_max_nb) {
        rank[i] = r;
        if (r < min_rank) min_rank = r;
      }
  }
  const size_t max_rank = n - min_rank;
  for (size_t i = 0; i < n; ++i) rank[i] -= min_rank;
  double reference[KMOCMA_MAX_OBJECTIVES];
  for (size_t k = 0; k < K; ++k) reference[k] = INFINITY;
  for (size_t i = 0; i < n; ++i)
    for (size_t k = 0; k < K; ++k)
      if (values[i * K + k] < reference[k]) reference[k] = values[i * K + k];
  for (size_t i = 0; i < n; ++i) sorted[i] = -1;
  size_t* unsorted = malloc(sizeof(size_t) * n);
  int order = 0;
  for (size_t r = 0; r <= max_rank; ++r) {
    for (;;) {
      size_t m = 0;
      for (size_t i = 0; i < n; ++i)
        if (rank[i] == r && sorted[i] == -1) unsorted[m++] = i;
      if (m == 0) break;
      size_t next = unsorted[0];
      double best = 0.;
      for (size_t a = 0; a < m; ++a) {
        double hv = 0.0;
        for (size_t k = 0; k < K; ++k) {
          double ub = -INFINITY;
          for (size_t b = 0; b < m; ++b)
            if (a != b && values[unsorted[b] * K + k] > ub) ub = values[unsorted[b] * K + k];
          hv += (ub - reference[k]);
        }
        if (a == 0 || hv > best) { best = hv; next = unsorted[a]; }
      }
      sorted[next] = order++;
    }
  }
  free(rank); free(max_nb); free(unsorted);
}

static void copy_individual(omocma_t* h, int dst, uint64_t di, int src, uint64_t si) {
  const uint64_t N = h->N;
  memcpy(h->X[dst] + di * N, h->X[src] + si * N, sizeof(double) * N);
  h->S[dst][di] = h->S[src][si];
  memcpy(h->C[dst] + di * N * N, h->C[src] + si * N * N, sizeof(double) * N * N);
  memcpy(h->P[dst] + di * N, h->P[src] + si * N, sizeof(double) * N);
  h->PS[dst][di] = h->PS[src][si];
}

/* updateDistribution :344-418 + updateStatistics :420-520 */
int omocma_tell(omocma_t* h) {
  const uint64_t N = h->N, lambda = h->lambda, mu = h->mu, K = h->K, n2 = 2 * lambda;
  double* values = malloc(sizeof(double) * n2 * K);
  memcpy(values, h->F, sizeof(double) * lambda * K);
  memcpy(values + lambda * K, h->Fprev, sizeof(double) * lambda * K);
  sort_sample_indices(h, values, n2, h->sorted);
  free(values);
  const double path_factor = sqrt(h->cc * (2. - h->cc));
  const double dd = 1.0 + 2.0 * fmax(0.0, sqrt(((double)mu - 1.) / ((double)N + 1.)) - 1.0) + h->cc;
  const double chi_n = sqrt((double)N) * (1. - 1. / (4. * (double)N) + 1. / (21. * (double)N * (double)N));
  for (uint64_t i = 0; i < lambda; ++i) {
    double* C = h->C[CUR] + i * N * N;
    double* pc = h->P[CUR] + i * N;
    h->PS[CUR][i] *= (1. - h->cp);
    if ((uint64_t)h->sorted[i] >= n2 - mu) h->PS[CUR][i] += h->cp;
    h->PS[CUR][i] = fmin(h->PS[CUR][i], 1.0);
    for (uint64_t d = 0; d < N; ++d) pc[d] = (1. - h->cc) * pc[d];
    const double* parent = h->X[PAR] + h->parent_index[i] * N;
    for (uint64_t d = 0; d < N; ++d) pc[d] += path_factor / sqrt(C[d * N + d]) * (h->X[CUR][i * N + d] - parent[d]) / h->S[CUR][i];
    double len = 0.;
    for (uint64_t d = 0; d < N; ++d) len += pc[d] * pc[d];
    len = sqrt(len);
    for (uint64_t d = 0; d < N; ++d)
      for (uint64_t e = 0; e < N; ++e) C[d * N + e] = (1. - h->ccov) * C[d * N + e] + h->ccov * pc[d] * pc[e];
    if (h->PS[CUR][i] >= h->target)
      for (uint64_t d = 0; d < N; ++d)
        for (uint64_t e = 0; e < N; ++e) C[d * N + e] += h->ccov * path_factor * path_factor * C[d * N + e];
    h->S[CUR][i] *= exp(h->cc / dd * (len / chi_n - 1.0));
  }
  for (uint64_t i = 0; i < n2; ++i)
    if (h->sorted[i] >= (int)(n2 - mu)) {
      const uint64_t pidx = n2 - (uint64_t)h->sorted[i] - 1;
      if (i < lambda) copy_individual(h, PAR, pidx, CUR, i);
      else copy_individual(h, PAR, pidx, PREV, i - lambda);
    }
  /* updateStatistics */
  memcpy(h->prev_best, h->cur_best, sizeof(double) * K);
  memcpy(h->prev_best_x, h->cur_best_x, sizeof(double) * K * N);
  for (uint64_t k = 0; k < K; k++) h->cur_best[k] = -INFINITY;
  for (uint64_t i = 0; i < lambda; i++) { h->min_sd[i] = INFINITY; h->max_sd[i] = -INFINITY; }
  for (uint64_t i = 0; i < lambda; ++i)
    for (uint64_t k = 0; k < K; ++k)
      if (h->F[i * K + k] > h->cur_best[k]) {
        h->cur_best[k] = h->F[i * K + k];
        memcpy(h->cur_best_x + k * N, h->X[CUR] + i * N, sizeof(double) * N);
        h->val_diff[k] = h->cur_best[k] - h->prev_best[k];
        double l2 = 0.;
        for (uint64_t d = 0; d < N; ++d) l2 += pow(h->prev_best_x[k * N + d] - h->cur_best_x[k * N + d], 2.);
        h->var_diff[k] = sqrt(l2);
      }
  for (uint64_t k = 0; k < K; ++k)
    if (h->cur_best[k] > h->best_ever[k]) {
      h->best_ever[k] = h->cur_best[k];
      memcpy(h->best_ever_x + k * N, h->cur_best_x + k * N, sizeof(double) * N);
    }
  /* :452-461 indexes _currentSigma by the DIMENSION d (not by the sample i); restated as written, clamped to the array */
  for (uint64_t i = 0; i < lambda; ++i)
    for (uint64_t d = 0; d < N; ++d) {
      const double sdev = h->S[CUR][d < lambda ? d : lambda - 1] * sqrt(h->C[CUR][i * N * N + d * N + d]);
      if (sdev > h->max_sd[i]) h->max_sd[i] = sdev;
      if (sdev < h->min_sd[i]) h->min_sd[i] = sdev;
    }
  /* non dominated samples of the current generation :464-481, archive merge :483-519 */
  uint64_t* cand = malloc(sizeof(uint64_t) * lambda);
  uint64_t nc = 0;
  for (uint64_t i = 0; i < lambda; ++i) {
    int dominated = 0;
    for (uint64_t j = 0; j < lambda && !dominated; ++j)
      if (j != i) {
        uint64_t nd = 0;
        for (uint64_t k = 0; k < K; ++k)
          if (h->F[j * K + k] > h->F[i * K + k]) nd++;
        if (nd == K) dominated = 1;
      }
    if (!dominated) cand[nc++] = i;
  }
  h->nondom = nc;
  char* keep_c = malloc(nc + 1); char* keep_s = malloc(h->coll_n + 1);
  memset(keep_c, 1, nc + 1); memset(keep_s, 1, h->coll_n + 1);
  for (uint64_t a = 0; a < nc; ++a)
    for (size_t j = 0; j < h->coll_n; ++j) {
      uint64_t cdom = 0, sdom = 0;
      for (uint64_t k = 0; k < K; ++k) {
        if (h->F[cand[a] * K + k] > h->coll_f[j * K + k]) cdom++;
        if (h->F[cand[a] * K + k] < h->coll_f[j * K + k]) sdom++;
      }
      if (cdom == K) keep_s[j] = 0;
      if (sdom == K) keep_c[a] = 0;
    }
  size_t w = 0;
  for (size_t j = 0; j < h->coll_n; ++j)
    if (keep_s[j]) {
      if (w != j) { memmove(h->coll_x + w * N, h->coll_x + j * N, sizeof(double) * N); memmove(h->coll_f + w * K, h->coll_f + j * K, sizeof(double) * K); }
      w++;
    }
  for (uint64_t a = 0; a < nc; ++a)
    if (keep_c[a]) {
      if (w + 1 > h->coll_cap) {
        h->coll_cap = h->coll_cap ? 2 * h->coll_cap : 256;
        h->coll_x = realloc(h->coll_x, sizeof(double) * h->coll_cap * N); h->coll_f = realloc(h->coll_f, sizeof(double) * h->coll_cap * K);
      }
      memcpy(h->coll_x + w * N, h->X[CUR] + cand[a] * N, sizeof(double) * N);
      memcpy(h->coll_f + w * K, h->F + cand[a] * K, sizeof(double) * K);
      w++;
    }
  h->coll_n = w;
  free(cand); free(keep_c); free(keep_s);
  h->gen++;
  return 0;
}

int omocma_run_generation(omocma_t* h) { return omocma_ask(h) || omocma_eval(h) || omocma_tell(h); }

/* generated checkTermination (MOCMAES.config "Termination Criteria", optimizer.config, solver.config); h->gen is the generation
 * about to run, like _k->_currentGeneration when Experiment::run tests the criteria (experiment.cpp.base:57-100) */
int omocma_check_termination(omocma_t* h, int* finished, const char** reason) {
  h->reason[0] = 0;
  int fin = 0;
  const uint64_t K = h->K, lambda = h->lambda;
  double mx;
#define ADD(msg) do { fin = 1; strncat(h->reason, msg, sizeof(h->reason) - strlen(h->reason) - 1); } while (0)
  if (h->gen > 1) {
    mx = -INFINITY; for (uint64_t k = 0; k < K; k++) if (h->val_diff[k] > mx) mx = h->val_diff[k];
    if (fabs(mx) < h->tc_min_value_diff) ADD("Min Max Value Difference Threshold;");   /* the criterion reads the BASE class threshold */
    mx = -INFINITY; for (uint64_t k = 0; k < K; k++) if (h->var_diff[k] > mx) mx = h->var_diff[k];
    if (mx < h->tc_min_var_diff) ADD("Min Variable Difference Threshold;");
    mx = -INFINITY; for (uint64_t i = 0; i < lambda; i++) if (h->min_sd[i] > mx) mx = h->min_sd[i];
    if (mx <= h->tc_min_sd) ADD("Min Standard Deviation;");
    mx = INFINITY; for (uint64_t i = 0; i < lambda; i++) if (h->max_sd[i] < mx) mx = h->max_sd[i];
    if (mx >= h->tc_max_sd) ADD("Max Standard Deviation;");
  }
  if (h->tc_max_model_evaluations <= (double)h->model_evals) ADD("solver['Max Model Evaluations'];");
  if ((double)h->gen > h->tc_max_generations) ADD("solver['Max Generations'];");
#undef ADD
  *finished = fin;
  if (reason) *reason = h->reason;
  return 0;
}

int omocma_run(omocma_t* h, uint64_t max_generations, uint64_t* done) {
  uint64_t g = 0;
  for (; g < max_generations; g++) {
    int fin; const char* why;
    omocma_check_termination(h, &fin, &why);
    if (fin) break;
    if (omocma_run_generation(h)) return 1;
  }
  if (done) *done = g;
  return 0;
}

int omocma_get_array(omocma_t* h, const char* key, double* out, size_t cap, size_t* count) {
  const uint64_t N = h->N, lambda = h->lambda, mu = h->mu, K = h->K;
  const double* src = NULL; size_t n = 0;
  static const char* pop[3] = {"Current", "Previous", "Parent"};
  char name[96];
  for (int s = 0; s < 3 && !src; s++) {
    const uint64_t rows = s == PAR ? mu : lambda;
    snprintf(name, sizeof(name), "%s Sample Population", pop[s]); if (!strcmp(key, name)) { src = h->X[s]; n = rows * N; }
    snprintf(name, sizeof(name), "%s Sigma", pop[s]); if (!strcmp(key, name)) { src = h->S[s]; n = rows; }
    snprintf(name, sizeof(name), "%s Covariance Matrix", pop[s]); if (!strcmp(key, name)) { src = h->C[s]; n = rows * N * N; }
    snprintf(name, sizeof(name), "%s Evolution Paths", pop[s]); if (!strcmp(key, name)) { src = h->P[s]; n = rows * N; }
    snprintf(name, sizeof(name), "%s Success Probabilities", pop[s]); if (!strcmp(key, name)) { src = h->PS[s]; n = rows; }
  }
  if (!src) {
    if (!strcmp(key, "Current Values")) { src = h->F; n = lambda * K; }
    else if (!strcmp(key, "Previous Values")) { src = h->Fprev; n = lambda * K; }
    else if (!strcmp(key, "Best Ever Values")) { src = h->best_ever; n = K; }
    else if (!strcmp(key, "Current Best Values")) { src = h->cur_best; n = K; }
    else if (!strcmp(key, "Previous Best Values")) { src = h->prev_best; n = K; }
    else if (!strcmp(key, "Best Ever Variables Vector")) { src = h->best_ever_x; n = K * N; }
    else if (!strcmp(key, "Current Best Variables Vector")) { src = h->cur_best_x; n = K * N; }
    else if (!strcmp(key, "Current Best Value Differences")) { src = h->val_diff; n = K; }
    else if (!strcmp(key, "Current Best Variable Differences")) { src = h->var_diff; n = K; }
    else if (!strcmp(key, "Current Min Standard Deviations")) { src = h->min_sd; n = lambda; }
    else if (!strcmp(key, "Current Max Standard Deviations")) { src = h->max_sd; n = lambda; }
    else if (!strcmp(key, "Sample Collection")) { src = h->coll_x; n = h->coll_n * N; }
    else if (!strcmp(key, "Sample Value Collection")) { src = h->coll_f; n = h->coll_n * K; }
  }
  if (!src && (!strcmp(key, "Parent Index") || !strcmp(key, "Sorted Indices"))) {
    n = !strcmp(key, "Parent Index") ? lambda : 2 * lambda;
    if (count) *count = n;
    if (!out) return 0;
    if (cap < n) return failf(h, "get_array(%s): capacity %zu < %zu", key, cap, n);
    for (size_t i = 0; i < n; i++) out[i] = !strcmp(key, "Parent Index") ? (double)h->parent_index[i] : (double)h->sorted[i];
    return 0;
  }
  if (!src && n == 0 && strcmp(key, "Sample Collection") && strcmp(key, "Sample Value Collection")) return failf(h, "unknown array key '%s'", key);
  if (count) *count = n;
  if (!out) return 0;
  if (cap < n) return failf(h, "get_array(%s): capacity %zu < %zu", key, cap, n);
  if (n) memcpy(out, src, sizeof(double) * n);
  return 0;
}

int omocma_get_scalar(omocma_t* h, const char* key, double* out) {
  if (!strcmp(key, "Current Non Dominated Sample Count")) *out = (double)h->nondom;
  else if (!strcmp(key, "Infeasible Sample Count")) *out = (double)h->infeasible;
  else if (!strcmp(key, "Model Evaluation Count")) *out = (double)h->model_evals;
  else if (!strcmp(key, "Current Generation")) *out = (double)h->gen;
  else if (!strcmp(key, "Sample Collection Size")) *out = (double)h->coll_n;
  else if (!strcmp(key, "Population Size")) *out = (double)h->lambda;
  else if (!strcmp(key, "Mu Value")) *out = (double)h->mu;
  else if (!strcmp(key, "Evolution Path Adaption Strength")) *out = h->cc;
  else if (!strcmp(key, "Covariance Learning Rate")) *out = h->ccov;
  else if (!strcmp(key, "Termination Criteria/Min Value Difference Threshold")) *out = h->tc_min_value_diff;
  else if (!strcmp(key, "Termination Criteria/Min Variable Difference Threshold")) *out = h->tc_min_var_diff;
  else if (!strcmp(key, "Termination Criteria/Min Standard Deviation")) *out = h->tc_min_sd;
  else if (!strcmp(key, "Termination Criteria/Max Standard Deviation")) *out = h->tc_max_sd;
  else if (!strcmp(key, "Termination Criteria/Max Generations")) *out = h->tc_max_generations;
  else if (!strcmp(key, "Termination Criteria/Max Model Evaluations")) *out = h->tc_max_model_evaluations;
  else return failf(h, "unknown scalar key '%s'", key);
  return 0;
}

int omocma_set_scalar(omocma_t* h, const char* key, double v) {
  if (!strcmp(key, "Termination Criteria/Min Value Difference Threshold")) h->tc_min_value_diff = v;
  else if (!strcmp(key, "Termination Criteria/Min Variable Difference Threshold")) h->tc_min_var_diff = v;
  else if (!strcmp(key, "Termination Criteria/Min Standard Deviation")) h->tc_min_sd = v;
  else if (!strcmp(key, "Termination Criteria/Max Standard Deviation")) h->tc_max_sd = v;
  else if (!strcmp(key, "Termination Criteria/Max Generations")) h->tc_max_generations = v;
  else if (!strcmp(key, "Termination Criteria/Max Model Evaluations")) h->tc_max_model_evaluations = v;
  else return failf(h, "unknown scalar key '%s'", key);
  return 0;
}

uint64_t omocma_launch_count(const omocma_t* h) { (void)h; return 0; }
