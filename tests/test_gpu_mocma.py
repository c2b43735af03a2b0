"""GPU tests of Optimizer/MOCMAES on the device (korali_b200/csrc/mocma.cu through include/kmocma.h) against the CPU oracle
(oracle/omocma.c). Given the same z the two are the same arithmetic (both built without FMA contraction); the device's log /
sincospi / exp differ from libm in the last bit, so continuous state is compared to 1e-10 over a free run and the discrete state
(parent indices, the order of the 2 lambda merged samples, the archive size) exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from korali_b200 import _mocma  # noqa: E402
from korali_b200._abi import KcmaError  # noqa: E402
from oracle import oracle as O  # noqa: E402

EXAMPLE = dict(n=4, num_objectives=2, population_size=32, mu_value=16, objective="NegRosenbrockAndSphere", lower_bound=-25.0,
               upper_bound=25.0, initial_stddev=3.0, seed=0xC0F33)     # examples/optimization/multiobjective/run-mocmaes.py:20-37


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("case", [EXAMPLE,
                                  dict(n=9, num_objectives=3, population_size=24, mu_value=24, objective="NegRosenbrockAndTwoSpheres",
                                       lower_bound=-3.0, upper_bound=4.0, seed=11),                      # mu == lambda: parent i -> offspring i
                                  dict(n=20, num_objectives=2, population_size=0, mu_value=0, objective="NegRosenbrockAndSphere",
                                       lower_bound=-2.0, upper_bound=2.0, initial_stddev=np.linspace(0.2, 1.0, 20), seed=5)])
def test_lockstep_against_oracle(case):
    s = _mocma.Solver(**case); o = O.OracleMOCMA(**case)
    assert (s.population_size, s.mu_value) == (o.population_size, o.mu_value)
    for g in range(25):
        s.ask(); o.ask()
        assert np.array_equal(s.get("Parent Index"), o.get("Parent Index")), g
        assert relerr(s.get("Current Sample Population"), o.get("Current Sample Population")) < 1e-10, g
        s.eval(); o.eval()
        assert relerr(s.get("Current Values"), o.get("Current Values")) < 1e-9, g
        s.tell(); o.tell()
        assert np.array_equal(s.get("Sorted Indices"), o.get("Sorted Indices")), g
        for k in ["Current Sigma", "Current Covariance Matrix", "Current Evolution Paths", "Current Success Probabilities", "Parent Sigma",
                  "Parent Sample Population", "Parent Covariance Matrix", "Parent Evolution Paths", "Parent Success Probabilities",
                  "Best Ever Values", "Current Best Values", "Best Ever Variables Vector", "Current Min Standard Deviations",
                  "Current Max Standard Deviations", "Current Best Variable Differences"]:
            e = relerr(s.get(k), o.get(k))
            assert e < 1e-9, (g, k, e)
        assert s.scalar("Current Non Dominated Sample Count") == o.scalar("Current Non Dominated Sample Count"), g
        assert s.scalar("Sample Collection Size") == o.scalar("Sample Collection Size"), g
    assert relerr(s.get("Sample Value Collection"), o.get("Sample Value Collection")) < 1e-9
    assert s.launch_count() > 0
    s.close(); o.close()


def test_given_the_same_values_the_update_is_bit_identical():
    """Injecting the oracle's population is not part of the ABI; injecting VALUES is: with identical F the ranking, the success
    probabilities and the selection of the parents are exact."""
    s = _mocma.Solver(**EXAMPLE); o = O.OracleMOCMA(**EXAMPLE)
    rng = np.random.default_rng(3)
    for g in range(6):
        s.ask(); o.ask()
        f = rng.standard_normal((32, 2))
        f[rng.integers(0, 32, 6)] = f[rng.integers(0, 32, 6)]      # equal rows: ties in the non-dominance levels and hypervolumes
        s.inject_f(f); o.inject_f(f)
        s.eval(); o.eval(); s.tell(); o.tell()
        assert np.array_equal(s.get("Sorted Indices"), o.get("Sorted Indices")), g
        assert np.array_equal(s.get("Current Success Probabilities"), o.get("Current Success Probabilities")), g
        assert np.array_equal(s.get("Best Ever Values"), o.get("Best Ever Values")), g
        assert s.scalar("Sample Collection Size") == o.scalar("Sample Collection Size"), g


def test_host_conduit_bounds_and_termination():
    calls = []
    def model(X):
        calls.append(X.shape)
        return np.stack([-np.sum((X - 1.0) ** 2, axis=1), -np.sum((X + 1.0) ** 2, axis=1)], axis=1)
    s = _mocma.Solver(n=3, num_objectives=2, population_size=16, objective="External", lower_bound=-2.0, upper_bound=2.0, seed=2)
    s.set_host_objective(model)
    s.set_scalar("Termination Criteria/Max Generations", 60)
    assert s.run(1000) == 60 and calls[0] == (16, 3) and len(calls) == 60
    fin, why = s.check_termination()
    assert fin and why == "solver['Max Generations'];"
    x = s.get("Sample Collection").reshape(-1, 3)
    f = s.get("Sample Value Collection").reshape(-1, 2)
    assert np.all(np.abs(x) <= 2.0) and len(f) > 10
    assert sum(1 for a in range(len(f)) for b in range(len(f)) if a != b and np.all(f[b] > f[a])) == 0
    # the Pareto set of two spheres centred at +1 and -1 is the segment between the centres: x_0 = x_1 = x_2 in [-1, 1]
    assert np.abs(x - x.mean(axis=1, keepdims=True)).max() < 0.6 and np.abs(x).max() < 1.6   # (the archive keeps early, rough points too)
    with pytest.raises(KcmaError, match="multiple objectives"):
        _mocma.Solver(n=3, num_objectives=1, objective="External", lower_bound=-2.0, upper_bound=2.0)
    with pytest.raises(KcmaError, match="non|Non"):
        t = _mocma.Solver(n=3, num_objectives=2, population_size=8, objective="External", lower_bound=-2.0, upper_bound=2.0)
        t.set_host_objective(lambda X: np.full((8, 2), np.nan))
        t.run_generation()


def _example_experiment(objective, gens=150):
    """examples/optimization/multiobjective/run-mocmaes.py:16-44 (silent, no result files)."""
    import korali_b200 as korali
    e = korali.Experiment()
    e["Random Seed"] = 0xC0F33
    e["Problem"]["Type"] = "Optimization"
    e["Problem"]["Objective Function"] = objective
    e["Problem"]["Num Objectives"] = 2
    for i in range(4):
        e["Variables"][i]["Name"] = "X" + str(i)
        e["Variables"][i]["Lower Bound"] = -25.0
        e["Variables"][i]["Upper Bound"] = +25.0
        e["Variables"][i]["Initial Standard Deviation"] = 3.0
    e["Solver"]["Type"] = "Optimizer/MOCMAES"
    e["Solver"]["Population Size"] = 32
    e["Solver"]["Mu Value"] = 16
    e["Solver"]["Termination Criteria"]["Min Value Difference Threshold"] = 1e-8
    e["Solver"]["Termination Criteria"]["Min Variable Difference Threshold"] = 1e-8
    e["Solver"]["Termination Criteria"]["Max Generations"] = gens
    e["Console Output"]["Verbosity"] = "Silent"
    e["File Output"]["Enabled"] = False
    return korali, e


def negative_rosenbrock_and_sphere(p):      # examples/optimization/multiobjective/_model/model.py:5-18
    x = p["Parameters"]
    dim = len(x)
    res_one = 0.
    for i in range(dim - 1):
        res_one += 100 * (x[i + 1] - x[i] ** 2) ** 2 + (1 - x[i]) ** 2
    res_two = 0.
    for i in range(dim):
        res_two += x[i] ** 2
    p["F(x)"] = [-res_one, -res_two]


def test_reference_example_through_the_korali_api():
    """The reference's example script runs unchanged (its Python model, per sample, operation "Evaluate Multiple") and gives the run
    of the device model 'RosenbrockAndSphere' on the same seed; e["Results"]["Pareto Optimal Samples"] as MOCMAES::finalize :539-552."""
    korali, e1 = _example_experiment(negative_rosenbrock_and_sphere)
    korali.Engine().run(e1)
    korali, e2 = _example_experiment("RosenbrockAndSphere")
    korali.Engine().run(e2)
    for e in (e1, e2):
        assert e["Current Generation"] == 150 and e["Solver"]["Model Evaluation Count"] == 150 * 32
        f = np.array(e["Results"]["Pareto Optimal Samples"]["F(x)"])          # (a koraliJson cursor is not a persistent sub-view)
        x = np.array(e["Results"]["Pareto Optimal Samples"]["Parameters"])
        assert f.shape[1] == 2 and x.shape == (len(f), 4) and len(f) == len(e["Solver"]["Sample Collection"]) > 10
        assert sum(1 for a in range(len(f)) for b in range(len(f)) if a != b and np.all(f[b] > f[a])) == 0
        assert np.allclose(f[:, 1], -np.sum(x * x, axis=1), rtol=1e-12)
    f1, f2 = np.array(e1["Results"]["Pareto Optimal Samples"]["F(x)"]), np.array(e2["Results"]["Pareto Optimal Samples"]["F(x)"])
    assert f1.shape == f2.shape and np.abs(f1 - f2).max() <= 1e-9 * np.abs(f1).max()      # Python floats vs the device model: same sums
    with pytest.raises(RuntimeError, match="multiple objectives"):
        korali, e3 = _example_experiment("RosenbrockAndSphere")
        e3["Problem"]["Num Objectives"] = 1
        korali.Engine().run(e3)
