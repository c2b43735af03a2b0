/*
 * okcma.h — CPU ORACLE for the CMA-ES generation loop. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference algorithm
 *   /root/reference/source/modules/solver/optimizer/CMAES/CMAES.cpp.base
 * (each function cites the lines it follows), with the same loop nests and the same
 * floating-point evaluation order (compiled with -O2 -ffp-contract=off), plus
 * restatements of the two GSL 2.6 routines the path uses (GSL is an un-vendored meson
 * wrap, subprojects/gsl.wrap:1-6, absent from /root/reference):
 *   - gsl_rng_mt19937 + gsl_ran_gaussian (polar Box-Muller)   -> bit-exact (pinned by fixture)
 *   - gsl_eigen_symmv + sort ABS_ASC                           -> Householder+QL; pinned by residuals
 *
 * PARITY IS PINNED: tests/test_oracle_golden.py replays the reference's own saved
 * trajectory tests/python/plot/cmaes/gen00000000..100.json (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library. The product (libkcma.so) never links or calls it.
 */
#ifndef OKCMA_H
#define OKCMA_H
#include "../include/kcma.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct okcma okcma_t;

/* Host callbacks standing in for the reference's per-sample Conduit dispatch
 * (CMAES.cpp.base:205-224 "Evaluate", :349-369 "Evaluate Constraints"). */
typedef void (*okcma_objective_fn)(void* user, const double* x, uint64_t n, double* f_out);
typedef void (*okcma_constraints_fn)(void* user, const double* x, uint64_t n, double* g_out, uint64_t n_constraints);

void okcma_cfg_defaults(kcma_cfg* cfg);
int okcma_create(const kcma_cfg* cfg, okcma_t** out);
void okcma_destroy(okcma_t* h);
const char* okcma_last_error(const okcma_t* h);
const char* okcma_take_warnings(okcma_t* h);
void okcma_set_objective_callback(okcma_t* h, okcma_objective_fn fn, void* user);
void okcma_set_constraints_callback(okcma_t* h, okcma_constraints_fn fn, void* user);

int okcma_run_generation(okcma_t* h);
int okcma_ask(okcma_t* h);
int okcma_eval(okcma_t* h);
int okcma_tell(okcma_t* h);
int okcma_check_termination(okcma_t* h, int* finished, const char** reason);
int okcma_run(okcma_t* h, uint64_t max_generations, uint64_t* done);
int okcma_inject(okcma_t* h, int kind, const double* host, size_t count);

int okcma_get_array(okcma_t* h, const char* key, double* out, size_t capacity, size_t* count);
int okcma_set_array(okcma_t* h, const char* key, const double* in, size_t count);
int okcma_get_index_array(okcma_t* h, const char* key, uint64_t* out, size_t capacity, size_t* count);
int okcma_get_scalar(okcma_t* h, const char* key, double* out);
int okcma_set_scalar(okcma_t* h, const char* key, double value);

/* Stand-alone pieces (same semantics as the kcma_k_* entry points). */
void okcma_sort_index(const double* f, uint64_t n, uint64_t* index_out);
int okcma_eigen(uint64_t n, const double* c, double* eigenvalues, double* q);
void okcma_sample(uint64_t n, uint64_t rows, const double* z, const double* b, const double* d,
                  const double* mean, double sigma, double* y_out, double* x_out);
void okcma_rank_mu(uint64_t n, uint64_t rows, const double* t, const double* w, double* p_out);
void okcma_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void okcma_philox_normal(uint64_t seed, uint64_t generation, uint64_t row_begin, uint64_t rows, uint64_t n, double* z_out);
void okcma_objective(int objective, uint64_t n, uint64_t rows, const double* x, const double* coef, double* f_out);
void okcma_objective_gradient(int objective, uint64_t n, uint64_t rows, const double* x, const double* coef, double* g_out);
/* gsl_rng_mt19937 seeded like gsl_rng_set(seed), `count` draws of gsl_ran_gaussian(rng, 1.0). */
void okcma_mt19937_gaussian(uint64_t seed, uint64_t skip, uint64_t count, double* out);

#ifdef __cplusplus
}
#endif
#endif
