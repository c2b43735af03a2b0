/*
 * kmocma.h — C ABI of the B200-native multi-objective CMA-ES generation loop (a sibling population solver on the
 * kernels and conventions of the CMA-ES path — Philox streams, batched device objectives, the batched host conduit). Replaces
 *   /root/reference/source/modules/solver/optimizer/MOCMAES/MOCMAES.cpp.base   (setInitialConfiguration :10-144, runGeneration
 *   :146-175, prepareGeneration :177-189, sampleSingle :191-230, sortSampleIndices :232-342, updateDistribution :344-418,
 *   updateStatistics :420-520)
 * behind "Solver": {"Type": "Optimizer/MOCMAES"} with "Problem": {"Num Objectives": K >= 2}. Same conventions as kcma.h / kdea.h:
 * extern "C", plain pointers and sizes, caller-owned host buffers that the library copies, non-zero return + kmocma_last_error on
 * failure, one caller thread per handle, no CPU fallback.
 *
 * Random numbers: the reference draws the parent index from its _uniformGenerator and the offspring from
 * gsl_ran_multivariate_gaussian (x = parent + sigma * L z, L the lower Cholesky factor of the parent's covariance), both on
 * sequential MT19937 streams. Here: counter-based Philox4x32-10, key = { seed_lo, seed_hi ^ "MOCM" },
 *   parent index of offspring i : counter { 0, i, 0, generation }, first 52-bit uniform of the block;
 *   z of offspring i, attempt a  : counter { 2^20 + pair, i, a, generation } -> Box-Muller pair (z[2 pair], z[2 pair + 1]) as in kcma.h.
 * The oracle (oracle/omocma.c) restates the same streams.
 */
#ifndef KMOCMA_H
#define KMOCMA_H
#include <stddef.h>
#include <stdint.h>

#include "kcma.h"

#ifdef __cplusplus
extern "C" {
#endif

#define KMOCMA_ABI_VERSION 1u
#define KMOCMA_MAX_OBJECTIVES 8

/* built-in device models (examples/optimization/multiobjective/_model/model.py:5-38); External = host conduit / injection */
enum { KMOCMA_OBJ_EXTERNAL = 0, KMOCMA_OBJ_NEG_ROSENBROCK_AND_SPHERE = 1, KMOCMA_OBJ_NEG_ROSENBROCK_AND_TWO_SPHERES = 2 };

typedef struct kmocma kmocma_t;

/* MOCMAES.config "Configuration Settings" and "Module Defaults". */
typedef struct kmocma_cfg {
  uint32_t abi_version, reserved0;
  uint64_t n;                 /* variables (<= 158 on the device: one Cholesky factor per CTA in shared memory) */
  uint64_t num_objectives;    /* "Num Objectives" of the problem (>= 2, :20-21) */
  uint64_t population_size;   /* "Population Size" (0: ceil(4 + floor(3 ln n)), :24); <= 5120 on the device */
  uint64_t mu_value;          /* "Mu Value" (0: population / 2, :25) */
  double evolution_path_adaption_strength; /* < 0: 2 / (n + 2) (:133) */
  double covariance_learning_rate;         /* < 0: 2 / (n^2 + 6) (:134) */
  double target_success_rate;              /* 0.175 */
  double threshold_probability;            /* 0.44 (unused by the reference's loop) */
  double success_learning_rate;            /* 0.08 */
  uint64_t seed;
  int32_t objective;          /* KMOCMA_OBJ_* */
  int32_t device;
  const double* lower_bound;  /* n (may be -inf) */
  const double* upper_bound;  /* n (may be +inf) */
  const double* initial_value;   /* n; NaN = (lower + upper) / 2 (:70-75). Not read by the reference's loop beyond the defaults */
  const double* initial_stddev;  /* n; NaN = 0.3 (upper - lower) (:77-82) */
} kmocma_cfg;

/* batched host conduit: F[rows x num_objectives] from X[rows x n], once per generation (operation "Evaluate Multiple", :155-170) */
typedef void (*kmocma_host_objective_fn)(void* user, const double* x, uint64_t rows, uint64_t n, double* f_out, uint64_t num_objectives);

void kmocma_cfg_defaults(kmocma_cfg* cfg);
int kmocma_create(const kmocma_cfg* cfg, kmocma_t** out);
void kmocma_destroy(kmocma_t* h);
const char* kmocma_last_error(const kmocma_t* h);

int kmocma_run_generation(kmocma_t* h); /* runGeneration :146-175 */
int kmocma_ask(kmocma_t* h);            /* prepareGeneration :177-189 + sampleSingle :191-230 */
int kmocma_eval(kmocma_t* h);           /* the per-sample dispatch :152-170 as one batched evaluation */
int kmocma_tell(kmocma_t* h);           /* updateDistribution :344-418 + updateStatistics :420-520 */
int kmocma_set_host_objective(kmocma_t* h, kmocma_host_objective_fn fn, void* user);
int kmocma_inject_f(kmocma_t* h, const double* f, size_t count); /* population_size x num_objectives */

/* generated checkTermination chain: MOCMAES.config "Termination Criteria" + optimizer.config + solver.config */
int kmocma_check_termination(kmocma_t* h, int* finished, const char** reason);
int kmocma_run(kmocma_t* h, uint64_t max_generations, uint64_t* done);

/* State by Korali key name. Arrays: "Current Sample Population", "Previous Sample Population", "Parent Sample Population",
 * "Current Values", "Previous Values", "Current Sigma", "Parent Sigma", "Current Covariance Matrix", "Parent Covariance Matrix",
 * "Current Evolution Paths", "Parent Evolution Paths", "Current Success Probabilities", "Parent Success Probabilities",
 * "Parent Index", "Sorted Indices" (of the 2 lambda merged values), "Best Ever Values", "Current Best Values",
 * "Best Ever Variables Vector" (K x n), "Current Best Variables Vector", "Current Best Value Differences",
 * "Current Best Variable Differences", "Current Min Standard Deviations", "Current Max Standard Deviations",
 * "Sample Collection" (archive of non-dominated samples, rows x n), "Sample Value Collection" (rows x K).
 * Scalars: "Current Non Dominated Sample Count", "Infeasible Sample Count", "Model Evaluation Count", "Current Generation",
 * "Sample Collection Size", "Termination Criteria/<name>". */
int kmocma_get_array(kmocma_t* h, const char* key, double* out, size_t capacity, size_t* count);
int kmocma_get_scalar(kmocma_t* h, const char* key, double* out);
int kmocma_set_scalar(kmocma_t* h, const char* key, double value);
uint64_t kmocma_launch_count(const kmocma_t* h);

#ifdef __cplusplus
}
#endif
#endif
