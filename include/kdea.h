/*
 * kdea.h — C ABI of the B200-native Differential Evolution generation loop (a sibling population solver on the
 * kernels of the CMA-ES path — Philox streams, batched device objectives, the batched host conduit). Replaces
 *   /root/reference/source/modules/solver/optimizer/DEA/DEA.cpp.base   (setInitialConfiguration :15-64, runGeneration :66-89,
 *   initSamples :91-101, prepareGeneration :103-121, mutateSingle :123-186, fixInfeasible :188-201, updateSolver :203-282)
 * behind "Solver": {"Type": "Optimizer/DEA"}. Same conventions as kcma.h: extern "C", plain pointers and sizes, caller-owned host
 * buffers that the library copies, non-zero return + kdea_last_error on failure, one caller thread per handle, no CPU fallback.
 *
 * Random numbers: the reference draws everything from ONE sequential MT19937 stream (_uniformGenerator). Here: counter-based
 * Philox4x32-10, key = { seed_lo, seed_hi ^ "DEA!" }, counter = { block, sample, attempt, generation }, draw k of a sub-stream =
 * half (k & 1) of block base + (k >> 1), with the sub-streams  indices (a, b, c, rn; base 0), crossover (one draw per dimension;
 * base 2^20), fixInfeasible (one draw per dimension; base 2^21), initial population (generation 0, base 2^20). The oracle
 * (oracle/odea.c) restates the same streams, so device and oracle agree bit for bit.
 */
#ifndef KDEA_H
#define KDEA_H
#include <stddef.h>
#include <stdint.h>

#include "kcma.h" /* objective ids (KCMA_OBJ_*), kcma_host_objective_fn */

#ifdef __cplusplus
extern "C" {
#endif

#define KDEA_ABI_VERSION 1u

enum { KDEA_MUTATION_FIXED = 0 };                       /* "Mutation Rule"; "Self Adaptive" (DEA.cpp.base:136-156) is not built */
enum { KDEA_PARENT_RANDOM = 0, KDEA_PARENT_BEST = 1 };  /* "Parent Selection Rule" :158-172 */
enum { KDEA_ACCEPT_BEST = 0, KDEA_ACCEPT_GREEDY = 1, KDEA_ACCEPT_IMPROVED = 2, KDEA_ACCEPT_ITERATIVE = 3 }; /* "Accept Rule" :222-267 */

typedef struct kdea kdea_t;

/* DEA.config "Configuration Settings" (:1-60) and "Module Defaults" (:150-213). */
typedef struct kdea_cfg {
  uint32_t abi_version, reserved0;
  uint64_t n;                /* variables */
  uint64_t population_size;  /* "Population Size" (default 200) */
  double crossover_rate;     /* "Crossover Rate" 0.9 */
  double mutation_rate;      /* "Mutation Rate" 0.5 */
  int32_t mutation_rule, parent_selection_rule, accept_rule, fix_infeasible;
  uint64_t seed;             /* "Random Seed" of the Uniform Generator */
  int32_t objective;         /* KCMA_OBJ_* (device objective) or KCMA_OBJ_EXTERNAL (host conduit / injection) */
  int32_t device;
  const double* lower_bound; /* n, finite and lower <= upper (:19-21): the initial population is uniform in the box (:91-101) */
  const double* upper_bound;
  const double* objective_coef; /* n or NULL */
} kdea_cfg;

void kdea_cfg_defaults(kdea_cfg* cfg);
int kdea_create(const kdea_cfg* cfg, kdea_t** out);
void kdea_destroy(kdea_t* h);
const char* kdea_last_error(const kdea_t* h);

/* runGeneration (:66-89) = prepareGeneration (mutation + rejection loop) + batched evaluation + updateSolver. */
int kdea_run_generation(kdea_t* h);
int kdea_ask(kdea_t* h);   /* prepareGeneration :103-121 (generation 1: the initial candidates) */
int kdea_eval(kdea_t* h);  /* the per-sample Conduit dispatch :72-85 as one batched evaluation */
int kdea_tell(kdea_t* h);  /* updateSolver :203-282 */
/* Batched host conduit: fn(user, X[rows x n], rows, n, F[rows]) once per generation (same type as the CMA-ES path). */
int kdea_set_host_objective(kdea_t* h, kcma_host_objective_fn fn, void* user);
/* Parity hook: F(x) of the current candidates from outside (population_size values). */
int kdea_inject_f(kdea_t* h, const double* f, size_t count);

/* generated checkTermination chain: DEA.config:62-83, optimizer.config, solver.config */
int kdea_check_termination(kdea_t* h, int* finished, const char** reason);
int kdea_run(kdea_t* h, uint64_t max_generations, uint64_t* done);

/* State by Korali key name ("Sample Population", "Candidate Population", "Value Vector", "Previous Value Vector", "Current Mean",
 * "Previous Mean", "Best Ever Variables", "Current Best Variables", "Max Distances"; scalars "Best Ever Value", "Current Best Value",
 * "Previous Best Value", "Previous Best Ever Value", "Best Sample Index", "Infeasible Sample Count", "Current Minimum Step Size",
 * "Current Generation", "Model Evaluation Count", "Termination Criteria/<name>"). */
int kdea_get_array(kdea_t* h, const char* key, double* out, size_t capacity, size_t* count);
int kdea_set_array(kdea_t* h, const char* key, const double* in, size_t count);
int kdea_get_scalar(kdea_t* h, const char* key, double* out);
int kdea_set_scalar(kdea_t* h, const char* key, double value);
uint64_t kdea_launch_count(const kdea_t* h);

#ifdef __cplusplus
}
#endif
#endif
