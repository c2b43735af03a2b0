/*
 * odea.c — CPU ORACLE of the Differential Evolution generation loop. TEST INFRASTRUCTURE ONLY (same rules as okcma.c: only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may load it; the product never links or calls it).
 *
 * Plain-C restatement of /root/reference/source/modules/solver/optimizer/DEA/DEA.cpp.base, loop for loop (each function cites the
 * lines it follows; -O2 -ffp-contract=off), with ONE stated difference: the reference draws every random number from a single
 * sequential MT19937 stream; here the draws come from the counter-based Philox streams defined in include/kdea.h (indices,
 * crossover, fixInfeasible, initial population), which is what the device uses. PARITY of this file against the reference itself
 * is therefore pinned only through the reference's statistical thresholds (tests/statistical/optimizers/correctness/run-dea.py:
 * checkMin(e, 0.23246, ...), ported in tests/test_oracle_dea.py) — the reference ships no DEA trajectory to replay: "parity
 * unpinned" at the bit level for this solver.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/kdea.h"
#include "okcma.h"

typedef void (*odea_objective_fn)(void* user, const double* x, uint64_t n, double* f_out);

typedef struct odea {
  kdea_cfg cfg;
  uint64_t N, lambda, gen, model_evals, infeasible, best_idx;
  double *lower, *upper, *coef;
  double *X, *Xc, *F, *Fprev, *mean, *prev_mean, *best_ever, *cur_best, *maxdist;
  double best_ever_value, prev_best_ever_value, cur_best_value, prev_best_value, min_step;
  double tc_max_infeasible, tc_min_value, tc_min_step, tc_max_value, tc_min_value_diff, tc_max_generations, tc_max_model_evaluations;
  odea_objective_fn obj_fn; void* obj_user;
  int have_inj_f;
  char err[512], reason[512];
} odea_t;

static int failf(odea_t* h, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  if (h) vsnprintf(h->err, sizeof(h->err), fmt, ap);
  va_end(ap);
  return 1;
}

static double unit_open(uint32_t lo, uint32_t hi) {
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (double)(v >> 12) * 0x1.0p-52 + 0x1.0p-53;
}
/* include/kdea.h: key = { seed_lo, seed_hi ^ "DEA!" }, ctr = { base + (k >> 1), sample, attempt, generation } */
static double dea_uniform(uint64_t seed, uint32_t generation, uint32_t attempt, uint64_t sample, uint32_t base, uint32_t k) {
  uint32_t ctr[4] = {base + (k >> 1), (uint32_t)sample, attempt, generation};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x44454121u};
  uint32_t r[4];
  okcma_philox4x32_10(ctr, key, r);
  return (k & 1u) ? unit_open(r[2], r[3]) : unit_open(r[0], r[1]);
}
#define DEA_IDX 0u
#define DEA_CROSS (1u << 20)
#define DEA_FIX (1u << 21)

void odea_cfg_defaults(kdea_cfg* c) {
  memset(c, 0, sizeof(*c));
  c->abi_version = KDEA_ABI_VERSION;
  c->population_size = 200; c->crossover_rate = 0.9; c->mutation_rate = 0.5;
  c->mutation_rule = KDEA_MUTATION_FIXED; c->parent_selection_rule = KDEA_PARENT_RANDOM; c->accept_rule = KDEA_ACCEPT_GREEDY;
  c->fix_infeasible = 1;
}

static double* dcopy(const double* s, size_t n, double fill) {
  double* p = (double*)malloc(sizeof(double) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) p[i] = s ? s[i] : fill;
  return p;
}

void odea_destroy(odea_t* h) {
  if (!h) return;
  free(h->lower); free(h->upper); free(h->coef); free(h->X); free(h->Xc); free(h->F); free(h->Fprev); free(h->mean); free(h->prev_mean);
  free(h->best_ever); free(h->cur_best); free(h->maxdist); free(h);
}

/* ref: DEA.cpp.base:91-101 */
static void init_samples(odea_t* h) {
  const uint64_t N = h->N;
  for (uint64_t i = 0; i < h->lambda; ++i)
    for (uint64_t d = 0; d < N; ++d) {
      double width = h->upper[d] - h->lower[d];
      h->Xc[i * N + d] = h->lower[d] + width * dea_uniform(h->cfg.seed, 0, 0, i, DEA_CROSS, (uint32_t)d);
      h->X[i * N + d] = h->Xc[i * N + d];
    }
}

/* ref: DEA.cpp.base:15-64 */
int odea_create(const kdea_cfg* cfg, odea_t** out) {
  odea_t* h = (odea_t*)calloc(1, sizeof(odea_t));
  h->cfg = *cfg;
  const uint64_t N = cfg->n, L = cfg->population_size;
  h->N = N; h->lambda = L;
  if (!N || !cfg->lower_bound || !cfg->upper_bound) { odea_destroy(h); return 1; }
  if (cfg->mutation_rule != KDEA_MUTATION_FIXED || L < 4) { odea_destroy(h); return 1; }
  h->lower = dcopy(cfg->lower_bound, N, 0); h->upper = dcopy(cfg->upper_bound, N, 0);
  h->coef = dcopy(cfg->objective_coef, N, 0);
  if (!cfg->objective_coef) for (uint64_t d = 0; d < N; d++) h->coef[d] = pow(10.0, 6.0 * (double)d / (double)(N > 1 ? N - 1 : 1));
  for (uint64_t d = 0; d < N; d++)
    if (h->upper[d] < h->lower[d] || !isfinite(h->upper[d]) || !isfinite(h->lower[d])) { odea_destroy(h); return 1; }
  h->X = dcopy(NULL, L * N, 0); h->Xc = dcopy(NULL, L * N, 0);
  h->F = dcopy(NULL, L, -INFINITY); h->Fprev = dcopy(NULL, L, -INFINITY);
  h->mean = dcopy(NULL, N, 0); h->prev_mean = dcopy(NULL, N, 0); h->best_ever = dcopy(NULL, N, 0); h->cur_best = dcopy(NULL, N, 0);
  h->maxdist = dcopy(NULL, N, 0);
  h->prev_best_value = h->cur_best_value = h->prev_best_ever_value = h->best_ever_value = -INFINITY;
  h->min_step = INFINITY;
  h->tc_max_infeasible = 1e7; h->tc_min_value = -INFINITY; h->tc_min_step = -INFINITY; h->tc_max_value = INFINITY;
  h->tc_min_value_diff = -INFINITY; h->tc_max_generations = 1e10; h->tc_max_model_evaluations = 1e9;
  h->gen = 1; /* _currentGeneration during the first runGeneration */
  init_samples(h);
  for (uint64_t i = 0; i < L; ++i)
    for (uint64_t d = 0; d < N; ++d) h->mean[d] += h->X[i * N + d] / ((double)L);
  *out = h;
  return 0;
}

const char* odea_last_error(const odea_t* h) { return h ? h->err : "odea_create failed"; }
void odea_set_objective_callback(odea_t* h, odea_objective_fn fn, void* user) { h->obj_fn = fn; h->obj_user = user; }

/* ref: optimizer.cpp.base:5-14 */
static int is_feasible(const odea_t* h, const double* x) {
  for (uint64_t d = 0; d < h->N; d++) {
    if (!isfinite(x[d])) return 0;
    if (x[d] < h->lower[d]) return 0;
    if (x[d] > h->upper[d]) return 0;
  }
  return 1;
}

/* ref: DEA.cpp.base:123-186 (Mutation Rule "Fixed") */
static void mutate_single(odea_t* h, uint64_t i, uint32_t attempt) {
  const uint64_t N = h->N, L = h->lambda;
  uint32_t k = 0;
#define U() dea_uniform(h->cfg.seed, (uint32_t)h->gen, attempt, i, DEA_IDX, k++)
  uint64_t a, b, c = 0;
  do { a = (uint64_t)(U() * (double)L); } while (a == i);
  do { b = (uint64_t)(U() * (double)L); } while (b == i || b == a);
  const double* parent;
  if (h->cfg.parent_selection_rule == KDEA_PARENT_RANDOM) {
    do { c = (uint64_t)(U() * (double)L); } while (c == i || c == a || c == b);
    parent = h->X + c * N;
  } else {
    parent = h->X + h->best_idx * N;
  }
  const uint64_t rn = (uint64_t)(U() * (double)N);
#undef U
  for (uint64_t d = 0; d < N; ++d) {
    if ((dea_uniform(h->cfg.seed, (uint32_t)h->gen, attempt, i, DEA_CROSS, (uint32_t)d) < h->cfg.crossover_rate) || (d == rn))
      h->Xc[i * N + d] = parent[d] + h->cfg.mutation_rate * (h->X[a * N + d] - h->X[b * N + d]);
    else
      h->Xc[i * N + d] = h->X[i * N + d];
  }
}

/* ref: DEA.cpp.base:188-201 */
static void fix_infeasible(odea_t* h, uint64_t i, uint32_t attempt) {
  const uint64_t N = h->N;
  for (uint64_t d = 0; d < N; ++d) {
    double len = 0.0;
    if (h->Xc[i * N + d] < h->lower[d]) len = h->Xc[i * N + d] - h->lower[d];
    if (h->Xc[i * N + d] > h->upper[d]) len = h->Xc[i * N + d] - h->upper[d];
    h->Xc[i * N + d] = h->X[i * N + d] - len * dea_uniform(h->cfg.seed, (uint32_t)h->gen, attempt, i, DEA_FIX, (uint32_t)d);
  }
}

/* ref: DEA.cpp.base:103-121 */
int odea_ask(odea_t* h) {
  if (h->gen > 1)
    for (uint64_t i = 0; i < h->lambda; ++i) {
      int feasible = 1;
      uint32_t attempt = 0;
      do {
        mutate_single(h, i, attempt);
        if (h->cfg.fix_infeasible && !feasible) fix_infeasible(h, i, attempt);
        feasible = is_feasible(h, h->Xc + i * h->N);
        if (!feasible) h->infeasible++;
        attempt++;
        if (attempt > 100000) return failf(h, "sample %lu never becomes feasible", (unsigned long)i);
      } while (!feasible);
    }
  memcpy(h->Fprev, h->F, sizeof(double) * h->lambda);
  return 0;
}

/* ref: DEA.cpp.base:72-85 (one Sample per candidate), optimization.cpp.base:26-34 */
int odea_eval(odea_t* h) {
  h->model_evals += h->lambda;
  if (h->have_inj_f) { h->have_inj_f = 0; return 0; }
  if (h->obj_fn) {
    for (uint64_t i = 0; i < h->lambda; i++) h->obj_fn(h->obj_user, h->Xc + i * h->N, h->N, h->F + i);
  } else if (h->cfg.objective != KCMA_OBJ_EXTERNAL) {
    okcma_objective(h->cfg.objective, h->N, h->lambda, h->Xc, h->coef, h->F);
  } else {
    return failf(h, "objective is External: set a callback or inject F");
  }
  for (uint64_t i = 0; i < h->lambda; i++)
    if (!isfinite(h->F[i])) return failf(h, "Non finite value of function evaluation detected: %f\n", h->F[i]);
  return 0;
}

/* ref: DEA.cpp.base:203-282 */
int odea_tell(odea_t* h) {
  const uint64_t N = h->N, L = h->lambda;
  uint64_t best = 0;
  for (uint64_t i = 1; i < L; i++) if (h->F[i] > h->F[best]) best = i; /* std::max_element: the first maximum */
  h->best_idx = best;
  h->prev_best_ever_value = h->best_ever_value;
  h->prev_best_value = h->cur_best_value;
  h->cur_best_value = h->F[best];
  for (uint64_t d = 0; d < N; ++d) h->cur_best[d] = h->Xc[best * N + d];
  memcpy(h->prev_mean, h->mean, sizeof(double) * N);
  for (uint64_t d = 0; d < N; ++d) h->mean[d] = 0.0;
  if (h->cur_best_value > h->best_ever_value) memcpy(h->best_ever, h->cur_best, sizeof(double) * N);
  switch (h->cfg.accept_rule) {
    case KDEA_ACCEPT_BEST:
      if (h->cur_best_value > h->best_ever_value) {
        for (uint64_t d = 0; d < N; ++d) h->X[best * N + d] = h->Xc[best * N + d];
        h->best_ever_value = h->cur_best_value;
      }
      break;
    case KDEA_ACCEPT_GREEDY:
      for (uint64_t i = 0; i < L; ++i)
        if (h->F[i] > h->Fprev[i]) memcpy(h->X + i * N, h->Xc + i * N, sizeof(double) * N);
      if (h->cur_best_value > h->best_ever_value) h->best_ever_value = h->cur_best_value;
      break;
    case KDEA_ACCEPT_IMPROVED:
      for (uint64_t i = 0; i < L; ++i)
        if (h->F[i] > h->best_ever_value)
          for (uint64_t d = 0; d < N; ++d) h->X[i * N + d] = h->Xc[i * N + d];
      if (h->cur_best_value > h->best_ever_value) h->best_ever_value = h->cur_best_value;
      break;
    case KDEA_ACCEPT_ITERATIVE:
      for (uint64_t i = 0; i < L; ++i)
        if (h->F[i] > h->best_ever_value)
          for (uint64_t d = 0; d < N; ++d) {
            h->X[i * N + d] = h->Xc[i * N + d];
            h->best_ever_value = h->F[i];
          }
      break;
    default: return failf(h, "Accept Rule (%d) not recognized.\n", h->cfg.accept_rule);
  }
  for (uint64_t i = 0; i < L; ++i)
    for (uint64_t d = 0; d < N; ++d) h->mean[d] += h->X[i * N + d] / ((double)L);
  for (uint64_t d = 0; d < N; ++d) {
    double mx = -INFINITY, mn = +INFINITY;
    for (uint64_t i = 0; i < L; ++i) {
      if (h->X[i * N + d] > mx) mx = h->X[i * N + d];
      if (h->X[i * N + d] < mn) mn = h->X[i * N + d];
    }
    h->maxdist[d] = mx - mn;
  }
  /* ref :280-281: the result of std::min is discarded — _currentMinimumStepSize stays +Inf, "Min Step Size" never fires */
  h->min_step = INFINITY;
  h->gen++;
  return 0;
}

int odea_run_generation(odea_t* h) { return odea_ask(h) || odea_eval(h) || odea_tell(h); }

/* generated checkTermination: DEA.config:62-83, optimizer.config, solver.config (all evaluated, reasons concatenated) */
int odea_check_termination(odea_t* h, int* finished, const char** reason) {
  int fin = 0;
  h->reason[0] = 0;
  const uint64_t gen = h->gen;
  if ((double)h->infeasible > h->tc_max_infeasible) { strcat(h->reason, "DEA['Max Infeasible Resamplings'];"); fin = 1; }
  if (gen > 1 && (-h->best_ever_value < h->tc_min_value)) { strcat(h->reason, "DEA['Min Value'];"); fin = 1; }
  if (h->min_step < h->tc_min_step) { strcat(h->reason, "DEA['Min Step Size'];"); fin = 1; }
  if (gen > 1 && (+h->best_ever_value > h->tc_max_value)) { strcat(h->reason, "optimizer['Max Value'];"); fin = 1; }
  if (gen > 1 && (fabs(h->cur_best_value - h->prev_best_value) < h->tc_min_value_diff)) { strcat(h->reason, "optimizer['Min Value Difference Threshold'];"); fin = 1; }
  if (h->tc_max_model_evaluations <= (double)h->model_evals) { strcat(h->reason, "solver['Max Model Evaluations'];"); fin = 1; }
  if ((double)gen > h->tc_max_generations) { strcat(h->reason, "solver['Max Generations'];"); fin = 1; }
  *finished = fin;
  if (reason) *reason = h->reason;
  return 0;
}

int odea_run(odea_t* h, uint64_t max_generations, uint64_t* done) {
  uint64_t n = 0;
  int fin = 0;
  while (n < max_generations) {
    odea_check_termination(h, &fin, NULL);
    if (fin) break;
    if (odea_run_generation(h)) { if (done) *done = n; return 1; }
    n++;
  }
  if (done) *done = n;
  return 0;
}

int odea_inject_f(odea_t* h, const double* f, size_t count) {
  if (count != h->lambda) return failf(h, "inject F: expected %zu values", (size_t)h->lambda);
  memcpy(h->F, f, sizeof(double) * count);
  h->have_inj_f = 1;
  return 0;
}

static double* find_arr(odea_t* h, const char* key, size_t* n) {
  const size_t N = h->N, L = h->lambda;
#define A(K, P, C) if (!strcmp(key, K)) { *n = (C); return (P); }
  A("Sample Population", h->X, L * N) A("Candidate Population", h->Xc, L * N) A("Value Vector", h->F, L) A("Previous Value Vector", h->Fprev, L)
  A("Current Mean", h->mean, N) A("Previous Mean", h->prev_mean, N) A("Best Ever Variables", h->best_ever, N)
  A("Current Best Variables", h->cur_best, N) A("Max Distances", h->maxdist, N)
#undef A
  return NULL;
}
int odea_get_array(odea_t* h, const char* key, double* out, size_t cap, size_t* count) {
  size_t n = 0;
  double* p = find_arr(h, key, &n);
  if (!p) return failf(h, "unknown array key '%s'", key);
  if (count) *count = n;
  if (!out) return 0;
  if (cap < n) return failf(h, "buffer too small for '%s'", key);
  memcpy(out, p, sizeof(double) * n);
  return 0;
}
int odea_set_array(odea_t* h, const char* key, const double* in, size_t count) {
  size_t n = 0;
  double* p = find_arr(h, key, &n);
  if (!p || n != count) return failf(h, "bad array key / size '%s'", key);
  memcpy(p, in, sizeof(double) * n);
  return 0;
}
static double* find_sca(odea_t* h, const char* key) {
#define S(K, F) if (!strcmp(key, K)) return &h->F;
  S("Best Ever Value", best_ever_value) S("Current Best Value", cur_best_value) S("Previous Best Value", prev_best_value)
  S("Previous Best Ever Value", prev_best_ever_value) S("Current Minimum Step Size", min_step)
  S("Termination Criteria/Max Infeasible Resamplings", tc_max_infeasible) S("Termination Criteria/Min Value", tc_min_value)
  S("Termination Criteria/Min Step Size", tc_min_step) S("Termination Criteria/Max Value", tc_max_value)
  S("Termination Criteria/Min Value Difference Threshold", tc_min_value_diff) S("Termination Criteria/Max Generations", tc_max_generations)
  S("Termination Criteria/Max Model Evaluations", tc_max_model_evaluations)
#undef S
  return NULL;
}
int odea_get_scalar(odea_t* h, const char* key, double* out) {
  double* p = find_sca(h, key);
  if (p) { *out = *p; return 0; }
  if (!strcmp(key, "Best Sample Index")) { *out = (double)h->best_idx; return 0; }
  if (!strcmp(key, "Infeasible Sample Count")) { *out = (double)h->infeasible; return 0; }
  if (!strcmp(key, "Current Generation")) { *out = (double)(h->gen - 1); return 0; }
  if (!strcmp(key, "Model Evaluation Count")) { *out = (double)h->model_evals; return 0; }
  return failf(h, "unknown scalar key '%s'", key);
}
int odea_set_scalar(odea_t* h, const char* key, double v) {
  double* p = find_sca(h, key);
  if (p) { *p = v; return 0; }
  return failf(h, "unknown scalar key '%s'", key);
}
uint64_t odea_launch_count(const odea_t* h) { (void)h; return 0; }
