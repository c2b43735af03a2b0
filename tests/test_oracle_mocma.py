"""CPU tests of the multi-objective CMA-ES oracle (oracle/omocma.c, the C restatement of MOCMAES.cpp.base on the Philox streams of
include/kmocma.h). The reference ships no MOCMAES trajectory or statistical test (only the configuration checks of
tests/unit/modules/solver/optimizers.cpp:2112-2180 and examples/optimization/multiobjective/run-mocmaes.py), so the oracle is anchored
by those: the configuration errors, the example's set-up and what its run must produce (a non-dominated archive whose ends approach
the optima of the two objectives), and the structure of a generation."""
import numpy as np
import pytest
from korali_b200._abi import KcmaError
from oracle import oracle as O

EXAMPLE = dict(n=4, num_objectives=2, population_size=32, mu_value=16, objective="NegRosenbrockAndSphere", lower_bound=-25.0,
               upper_bound=25.0, initial_stddev=3.0, seed=0xC0F33)     # run-mocmaes.py:20-37


def _dominated_pairs(f):
    return sum(1 for a in range(len(f)) for b in range(len(f)) if a != b and np.all(f[b] > f[a]))


def test_example_run_builds_a_pareto_front():
    o = O.OracleMOCMA(**EXAMPLE)
    for _ in range(300):
        o.run_generation()
    f = o.get("Sample Value Collection").reshape(-1, 2)
    x = o.get("Sample Collection").reshape(-1, 4)
    assert len(f) == o.scalar("Sample Collection Size") > 50 and _dominated_pairs(f) == 0
    be = o.get("Best Ever Values")
    assert be[1] > -1e-6 and be[0] > -0.5                      # sphere end at x = 0, Rosenbrock end approaching x = 1
    bx = o.get("Best Ever Variables Vector").reshape(2, 4)
    assert np.abs(bx[1]).max() < 1e-2
    # every archived point carries the values of the model at its parameters
    r1 = -np.sum(100 * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1 - x[:, :-1]) ** 2, axis=1)
    assert np.allclose(f[:, 0], r1, rtol=1e-12) and np.allclose(f[:, 1], -np.sum(x * x, axis=1), rtol=1e-12)
    assert o.scalar("Model Evaluation Count") == 300 * 32


def test_generation_structure():
    """Generation 1: every offspring descends from parent 0 (one non-dominated sample, :27, :196), the parents sit at the origin
    (:119) with C = diag(sd^2 / sigma^2) (:124); the mu best of offspring + previous offspring become the parents, best first."""
    o = O.OracleMOCMA(**EXAMPLE)
    o.ask()
    assert np.all(o.get("Parent Index") == 0)
    x = o.get("Current Sample Population").reshape(32, 4)
    assert np.all(np.abs(x) <= 25) and np.abs(x).max() > 0.5 and np.allclose(o.get("Current Sigma"), 3.0)
    o.eval(); o.tell()
    srt = o.get("Sorted Indices").astype(int)
    assert sorted(srt) == list(range(64))
    px = o.get("Parent Sample Population").reshape(16, 4)
    for i in range(32):                                          # previous values are -Inf: all parents come from the offspring
        if srt[i] >= 48:
            assert np.array_equal(px[63 - srt[i]], x[i])
    assert np.all(srt[32:] < 48)
    ps = o.get("Current Success Probabilities")
    assert np.allclose(ps[srt[:32] >= 48], 0.175 * 0.92 + 0.08) and np.allclose(ps[srt[:32] < 48], 0.175 * 0.92)
    # later generations draw their parents among min(mu, non-dominated) candidates
    for _ in range(5):
        o.run_generation()
    nd = o.scalar("Current Non Dominated Sample Count")
    o.ask()
    assert o.get("Parent Index").max() < min(16, nd)


def test_configuration_errors_of_the_reference_unit_test():
    with pytest.raises(KcmaError, match="multiple objectives"):
        O.OracleMOCMA(**dict(EXAMPLE, num_objectives=1, objective="External"))
    with pytest.raises(KcmaError, match="Mu Value"):
        O.OracleMOCMA(**dict(EXAMPLE, mu_value=33))
    with pytest.raises(KcmaError, match="Success Learning Rate"):
        O.OracleMOCMA(**dict(EXAMPLE, success_learning_rate=1.5))
    with pytest.raises(KcmaError, match="Target Success Rate"):
        O.OracleMOCMA(**dict(EXAMPLE, target_success_rate=0.0))
    with pytest.raises(KcmaError, match="cannot be inferred"):
        O.OracleMOCMA(n=2, num_objectives=2, objective="External", initial_stddev=1.0)      # no bounds, no initial value
    o = O.OracleMOCMA(n=10, num_objectives=2, objective="External", lower_bound=-1.0, upper_bound=1.0)
    assert o.population_size == 10 and o.mu_value == 5            # ceil(4 + floor(3 ln 10)), lambda / 2
    assert np.allclose(o.get("Parent Sigma"), 0.6)                 # 0.3 (upper - lower)


def test_termination_and_three_objectives():
    o = O.OracleMOCMA(**dict(EXAMPLE, num_objectives=3, objective="NegRosenbrockAndTwoSpheres"))
    o.set_scalar("Termination Criteria/Max Generations", 25)
    assert o.run(1000) == 25
    fin, why = o.check_termination()
    assert fin and why == "solver['Max Generations'];"
    f = o.get("Sample Value Collection").reshape(-1, 3)
    assert _dominated_pairs(f) == 0
    o = O.OracleMOCMA(**EXAMPLE)
    o.set_scalar("Termination Criteria/Min Value Difference Threshold", 1e-8)   # run-mocmaes.py:36 (read by "Min Max Value Difference Threshold")
    o.set_scalar("Termination Criteria/Min Variable Difference Threshold", 1e-8)
    done = o.run(100000)
    fin, why = o.check_termination()
    assert fin and done < 100000 and ("Value Difference" in why or "Variable Difference" in why)
