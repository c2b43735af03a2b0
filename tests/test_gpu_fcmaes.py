"""korali_b200.fCMAES: the ask/tell surface of the reference's float CMA-ES (fCMAES.cpp:204-337) on the device generation loop.
The reference ships no test or caller of fCMAES (it is only compiled, deepSupervisor/optimizers/meson.build), so the checks are: the
facade IS the C-ABI handle underneath (bitwise the float32 image of an identical kcma handle driven with inject / tell), the
reference's defaults, bounds by resampling, termination chain, convergence on the sphere."""
import math
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import korali_b200  # noqa: E402
from korali_b200 import _lib  # noqa: E402
from korali_b200._abi import INJ_F  # noqa: E402


def _make(n=12, pop=0, seed=77, lo=None, hi=None):
    o = korali_b200.fCMAES(n, pop)
    o._initialMeans[:] = 1.5
    o._initialStandardDeviations[:] = 0.8
    if lo is not None:
        o._lowerBounds[:] = lo; o._upperBounds[:] = hi
    o.setSeed(seed)
    o.reset()
    return o


def test_defaults_follow_the_reference_constructor():
    o = korali_b200.fCMAES(10)
    assert o._populationSize == int(math.ceil(4.0 + math.floor(3 * math.log(10.0)))) and o._muValue == o._populationSize // 2
    assert o._muType == "Linear" and o._maxGenerations == 10000000 and o._samplePopulation.dtype == np.float32
    with pytest.raises(RuntimeError):
        o.reset()        # means / standard deviations are NaN until the caller sets them (:38-39)


def test_ask_tell_is_the_handle_underneath():
    o = _make(n=12, pop=32)
    s = _lib.Solver(n=12, population_size=32, mu_value=16, mu_type="Linear", objective="External", keep_population=1, seed=77,
                    initial_value=np.full(12, np.float32(1.5), dtype=np.float64), initial_stddev=np.full(12, np.float32(0.8), dtype=np.float64),
                    max_infeasible_resamplings=10000000)
    for g in range(15):
        o.prepareGeneration(); s.ask()
        x = s.get("Sample Population").reshape(32, 12)
        assert np.array_equal(o._samplePopulation, x.astype(np.float32))
        f = (-np.sum(o._samplePopulation.astype(np.float64) ** 2, axis=1)).astype(np.float32)      # the caller's float model
        o.updateDistribution(f)
        s.inject(INJ_F, f.astype(np.float64)); s.eval(); s.tell()
        assert np.array_equal(o._currentMean, s.get("Current Mean").astype(np.float32))
        assert o._sigma == np.float32(s.scalar("Sigma")) and o._bestEverValue == np.float32(s.scalar("Best Ever Value"))
        assert np.array_equal(o._sortingIndex, s.get_index("Sorting Index"))
        o._currentGeneration += 1
    assert not o.checkTermination()
    o.close(); s.close()


def test_converges_on_the_sphere_and_terminates():
    o = _make(n=8, pop=64, seed=5)
    o._maxValue = -1e-10
    gens = 0
    while not o.checkTermination() and gens < 2000:
        o.prepareGeneration()
        o.updateDistribution(-np.sum(o._samplePopulation.astype(np.float64) ** 2, axis=1))
        o._currentGeneration += 1; gens += 1
    assert o._bestEverValue > -1e-10 and gens < 2000
    assert np.abs(o._bestEverVariables).max() < 1e-4
    o.close()


def test_bounds_are_kept_by_resampling():
    o = _make(n=6, pop=48, seed=9, lo=1.0, hi=2.0)
    for _ in range(5):
        o.prepareGeneration()
        assert o._samplePopulation.min() >= 1.0 and o._samplePopulation.max() <= 2.0
        o.updateDistribution(-np.sum((o._samplePopulation - 1.2) ** 2, axis=1))
        o._currentGeneration += 1
    assert o._infeasibleSampleCount > 0
    o.close()
