"""``korali_b200.fCMAES`` — the ask/tell surface of the reference's float CMA-ES
(/root/reference/source/modules/solver/learner/deepSupervisor/optimizers/fCMAES.{hpp,cpp}: constructor :9-94, reset :96-119,
prepareGeneration :204-219, updateDistribution :248-337, checkTermination :480-502) on the B200 generation loop.

Same member names and meaning (``_initialMeans``, ``_samplePopulation``, ``prepareGeneration()``, ``updateDistribution(evaluations)`` ...);
arrays cross the boundary as float32 like the reference's ``std::vector<float>``. The state itself lives in a libkcma handle and the
generation loop computes in FP64 (``dtype`` of the path): a superset of the reference's float arithmetic, NOT a bit-level restatement of
it — fCMAES draws from ``std::default_random_engine`` / ``std::normal_distribution<float>`` and sums in float, here the draws are the
Philox stream of include/kcma.h and the sums are FP64 DMMA tiles. What is the same: the algorithm (fCMAES.cpp is CMAES.cpp.base without
the constraint path: Linear / Equal / Logarithmic weights, resample-until-feasible, rank-1 + rank-mu update, eigendecomposition every
generation), its defaults and its termination chain, which is evaluated here on the host from the handle's scalars.
"""
import math
import numpy as np
from . import _lib
from ._abi import INJ_F

_INF = float("inf")


class fCMAES:
    def __init__(self, nVars, populationSize=0, muSize=0, device=0):
        self._nVars = int(nVars)
        self._populationSize = int(populationSize) or int(math.ceil(4.0 + math.floor(3 * math.log(float(nVars)))))   # :61
        self._muValue = int(muSize) or self._populationSize // 2                                                      # :62 (muSize: see note)
        self._muType = "Linear"
        self._initialSigmaCumulationFactor = -1.0
        self._initialDampFactor = -1.0
        self._isSigmaBounded = False
        self._initialCumulativeCovariance = -1.0
        self._isDiagonal = False
        n = self._nVars
        self._initialMeans = np.full(n, np.nan, dtype=np.float32)
        self._initialStandardDeviations = np.full(n, np.nan, dtype=np.float32)
        self._lowerBounds = np.full(n, -np.inf, dtype=np.float32)
        self._upperBounds = np.full(n, np.inf, dtype=np.float32)
        self._minMeanUpdates = np.full(n, -np.inf, dtype=np.float32)
        self._maxGenerations = 10000000
        self._maxInfeasibleResamplings = 10000000
        self._maxConditionCovarianceMatrix = _INF
        self._minValue = -_INF
        self._maxValue = _INF
        self._minValueDifferenceThreshold = -_INF
        self._targetMaxStandardDeviation = -_INF
        self._seed = 0
        self._device = int(device)
        self._h = None
        self._currentGeneration = 1
        self._samplePopulation = np.zeros((self._populationSize, n), dtype=np.float32)

    # fCMAES::setSeed :468
    def setSeed(self, seed):
        self._seed = int(seed)

    # fCMAES::reset :96-119 (initMuWeights :121-159, initCovariance :161-191 run inside kcma_create)
    def reset(self):
        if self._h is not None:
            self._h.close()
        if not (np.all(np.isfinite(self._initialMeans)) and np.all(np.isfinite(self._initialStandardDeviations))):
            raise RuntimeError("fCMAES: _initialMeans and _initialStandardDeviations must be set before reset()")
        bounded = bool(np.any(np.isfinite(self._lowerBounds)) or np.any(np.isfinite(self._upperBounds)))
        kw = dict(n=self._nVars, population_size=self._populationSize, mu_value=self._muValue, mu_type=self._muType,
                  initial_sigma_cumulation_factor=float(self._initialSigmaCumulationFactor),
                  initial_damp_factor=float(self._initialDampFactor), is_sigma_bounded=int(bool(self._isSigmaBounded)),
                  initial_cumulative_covariance=float(self._initialCumulativeCovariance),
                  diagonal_covariance=int(bool(self._isDiagonal)), objective="External", keep_population=1, seed=self._seed,
                  initial_value=self._initialMeans.astype(np.float64), initial_stddev=self._initialStandardDeviations.astype(np.float64),
                  max_infeasible_resamplings=int(self._maxInfeasibleResamplings), device=self._device)
        if bounded:
            kw["lower_bound"] = self._lowerBounds.astype(np.float64)
            kw["upper_bound"] = self._upperBounds.astype(np.float64)
        self._h = _lib.Solver(**kw)
        self._h.set_scalar("Termination Criteria/Max Model Evaluations", 1e300)
        self._currentGeneration = 1
        self._pull()

    def _live(self):
        if self._h is None:
            raise RuntimeError("fCMAES: call reset() first")
        return self._h

    def _pull(self):
        h = self._h
        f32 = np.float32
        self._sigma = f32(h.scalar("Sigma"))
        self._currentMean = h.get("Current Mean").astype(f32)
        self._previousMean = h.get("Previous Mean").astype(f32)
        self._meanUpdate = h.get("Mean Update").astype(f32)
        self._bestEverValue = f32(h.scalar("Best Ever Value"))
        self._previousBestEverValue = f32(h.scalar("Previous Best Ever Value"))
        self._currentBestValue = f32(h.scalar("Current Best Value"))
        self._previousBestValue = f32(h.scalar("Previous Best Value"))
        self._bestEverVariables = h.get("Best Ever Variables").astype(f32)
        self._currentBestVariables = h.get("Current Best Variables").astype(f32)
        self._infeasibleSampleCount = int(h.scalar("Infeasible Sample Count"))
        self._minimumCovarianceEigenvalue = f32(h.scalar("Minimum Covariance Eigenvalue"))
        self._maximumCovarianceEigenvalue = f32(h.scalar("Maximum Covariance Eigenvalue"))
        self._currentMinStandardDeviation = f32(h.scalar("Current Min Standard Deviation"))
        self._currentMaxStandardDeviation = f32(h.scalar("Current Max Standard Deviation"))
        self._conjugateEvolutionPathL2Norm = f32(h.scalar("Conjugate Evolution Path L2 Norm"))

    # fCMAES::prepareGeneration :204-219 (updateEigensystem + sampleSingle + resample until feasible)
    def prepareGeneration(self):
        h = self._live()
        h.ask()
        self._samplePopulation = h.get("Sample Population").reshape(self._populationSize, self._nVars).astype(np.float32)
        self._infeasibleSampleCount = int(h.scalar("Infeasible Sample Count"))

    # fCMAES::updateDistribution :248-337
    def updateDistribution(self, evaluations):
        h = self._live()
        ev = np.ascontiguousarray(evaluations, dtype=np.float64).reshape(-1)
        if ev.size != self._populationSize:
            raise RuntimeError("fCMAES::updateDistribution: %d evaluations for a population of %d" % (ev.size, self._populationSize))
        h.inject(INJ_F, ev)
        h.eval()
        h.tell()
        self._valueVector = ev.astype(np.float32)
        self._sortingIndex = h.get_index("Sorting Index")
        self._pull()

    # fCMAES::checkTermination :480-502 (the caller advances _currentGeneration, as in the reference)
    def checkTermination(self):
        g = self._currentGeneration
        if g > 1:
            if self._maxInfeasibleResamplings > 0 and self._infeasibleSampleCount >= self._maxInfeasibleResamplings: return True
            if self._maximumCovarianceEigenvalue >= self._maxConditionCovarianceMatrix * self._minimumCovarianceEigenvalue: return True
            if -self._bestEverValue < self._minValue: return True
            if self._bestEverValue > self._maxValue: return True
            if abs(self._currentBestValue - self._previousBestValue) < self._minValueDifferenceThreshold: return True
            if self._currentMaxStandardDeviation <= self._targetMaxStandardDeviation: return True
            if not np.any(np.abs(self._meanUpdate) > self._minMeanUpdates): return True
        return g >= self._maxGenerations

    def printInfo(self):   # :448-466
        print("sigma=%f" % self._sigma)
        print("currentMean=%s" % self._currentMean)
        print("bestEverValue=%f currentBestValue=%f" % (self._bestEverValue, self._currentBestValue))

    def covarianceMatrix(self):
        return self._live().get("Covariance Matrix").reshape(self._nVars, self._nVars).astype(np.float32)

    def close(self):
        if self._h is not None:
            self._h.close()
            self._h = None
