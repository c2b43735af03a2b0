"""GPU parity of the Differential Evolution path (korali_b200/csrc/dea.cu through include/kdea.h) against the oracle
(oracle/odea.c = DEA.cpp.base restated on the same Philox streams): populations, candidates, values, indices and counters are
BIT-EXACT generation by generation, for every parent-selection / accept rule; plus the reference's own run-dea.py through the
Korali API and convergence on the built-in objectives."""
import numpy as np
import pytest
import torch
import korali_b200 as korali
from korali_b200 import _dea
from korali_models.models import evalmodel
from oracle import oracle as O

pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

ARR = ["Sample Population", "Candidate Population", "Value Vector", "Previous Value Vector", "Current Mean", "Previous Mean",
       "Best Ever Variables", "Current Best Variables", "Max Distances"]
SCA = ["Best Ever Value", "Current Best Value", "Previous Best Value", "Previous Best Ever Value", "Best Sample Index",
       "Infeasible Sample Count", "Current Generation", "Model Evaluation Count", "Current Minimum Step Size"]


def same_state(s, o, where):
    for k in ARR:
        assert np.array_equal(s.get(k), o.get(k)), (where, k)
    for k in SCA:
        assert s.scalar(k) == o.scalar(k), (where, k, s.scalar(k), o.scalar(k))


@pytest.mark.parametrize("parent", ["Random", "Best"])
@pytest.mark.parametrize("accept", ["Best", "Greedy", "Improved", "Iterative"])
@pytest.mark.parametrize("fix", [1, 0])
def test_lockstep_bit_exact_against_oracle(parent, accept, fix):
    kw = dict(n=7, population_size=40, objective="NegRosenbrock", lower_bound=-2.0, upper_bound=np.array([2.0, 2.0, 1.5, 2.0, 3.0, 2.0, 2.0]),
              seed=1234, parent_selection_rule=parent, accept_rule=accept, fix_infeasible=fix, crossover_rate=0.8, mutation_rate=0.6)
    s = _dea.Solver(**kw); o = O.OracleDEA(**kw)
    same_state(s, o, "create")
    for g in range(25):
        s.ask(); o.ask()
        assert np.array_equal(s.get("Candidate Population"), o.get("Candidate Population")), g
        s.eval(); o.eval()
        assert np.array_equal(s.get("Value Vector"), o.get("Value Vector")), g     # polynomial objective: same summation tree
        s.tell(); o.tell()
        same_state(s, o, g)
    assert s.scalar("Infeasible Sample Count") > 0      # the rejection loop (:107-118) and fixInfeasible were exercised
    s.close()


@pytest.mark.parametrize("n,lam,obj,cr", [(1, 10, "NegSphere", 0.9), (33, 257, "NegEllipsoid", 0.9), (100, 4096, "NegSphere", 0.3),
                                           (1000, 2048, "NegEllipsoid", 0.01)])
def test_shapes_bit_exact(n, lam, obj, cr):
    # (with many dimensions and a high crossover rate nearly every mutant leaves the box and the reference's rejection loop :107-118
    # spins for ever — DEA's own behaviour; the wide cases therefore cross over few dimensions)
    kw = dict(n=n, population_size=lam, objective=obj, lower_bound=-3.0, upper_bound=5.0, seed=5, crossover_rate=cr)
    s = _dea.Solver(**kw); o = O.OracleDEA(**kw)
    for g in range(4):
        s.run_generation(); o.run_generation()
        same_state(s, o, g)
    s.close()


def test_transcendental_objective_and_injection():
    kw = dict(n=20, population_size=64, objective="NegAckley", lower_bound=-4.0, upper_bound=4.0, seed=9)
    s = _dea.Solver(**kw); o = O.OracleDEA(**kw)
    for g in range(10):
        s.ask(); o.ask(); o.eval()
        s.eval()
        f = o.get("Value Vector")
        assert np.abs(s.get("Value Vector") - f).max() <= 1e-13 * np.abs(f).max()
        s.inject_f(f); s.set_scalar("Model Evaluation Count", s.scalar("Model Evaluation Count") - 64); s.eval()   # pin F: strict lockstep
        s.tell(); o.tell()
        same_state(s, o, g)
    s.close()


def test_host_conduit_matches_device_objective():
    kw = dict(n=6, population_size=32, lower_bound=-2.0, upper_bound=2.0, seed=77)
    a = _dea.Solver(objective="NegSumSq", **kw)
    b = _dea.Solver(objective="External", **kw)
    b.set_host_objective(lambda X: -np.sum(X * X, axis=1))
    for g in range(20):
        a.run_generation(); b.run_generation()
    assert abs(a.scalar("Best Ever Value") - b.scalar("Best Ever Value")) <= 1e-12 * abs(a.scalar("Best Ever Value")) + 1e-300
    a.close(); b.close()


def test_converges_and_terminates():
    # (of the rule combinations only Greedy + Best parent converges fast — Greedy compares a candidate with the PREVIOUS CANDIDATE's value,
    # :120 and :233, so with random parents the population drifts; the oracle behaves the same way)
    s = _dea.Solver(n=10, population_size=200, objective="NegSphere", lower_bound=-5.0, upper_bound=5.0, seed=1, parent_selection_rule="Best")
    s.set_scalar("Termination Criteria/Min Value", 1e-9)
    done = s.run(5000)
    fin, why = s.check_termination()
    assert fin and "DEA['Min Value']" in why and done < 5000
    assert -s.scalar("Best Ever Value") < 1e-9 and np.abs(s.get("Best Ever Variables")).max() < 1e-3
    s.close()


def model(sample):
    evalmodel(sample)


# tests/statistical/optimizers/correctness/run-dea.py, same experiment definitions / seeds / thresholds (console output silenced)
@pytest.mark.parametrize("parent,accept,tol", [("Random", "Greedy", 1e-4), ("Random", "Best", 1e-2), ("Random", "Improved", 1e-2),
                                                ("Random", "Iterative", 1e-2), ("Best", "Greedy", 1e-4), ("Best", "Iterative", 1e-2)])
def test_run_dea_through_the_korali_api(parent, accept, tol):
    e = korali.Experiment()
    e["Problem"]["Type"] = "Optimization"
    e["Problem"]["Objective Function"] = model
    e["Variables"][0]["Name"] = "X"
    e["Variables"][0]["Lower Bound"] = -10.0
    e["Variables"][0]["Upper Bound"] = +10.0
    e["Solver"]["Type"] = "Optimizer/DEA"
    e["Solver"]["Population Size"] = 10
    e["Solver"]["Termination Criteria"]["Max Generations"] = 100
    e["Solver"]["Parent Selection Rule"] = parent
    e["Solver"]["Accept Rule"] = accept
    e["Console Output"]["Verbosity"] = "Silent"
    e["File Output"]["Enabled"] = False
    e["Random Seed"] = 1337
    korali.Engine().run(e)
    assert np.isclose(0.23246, e["Solver"]["Best Ever Value"], atol=tol)
    assert e["Current Generation"] == 100 and e["Solver"]["Model Evaluation Count"] == 1000
    assert e["Results"]["Best Sample"]["F(x)"] == e["Solver"]["Best Ever Value"]
    # the same run on the oracle (Uniform Generator seed = Random Seed + 1): identical trajectory
    o = O.OracleDEA(n=1, population_size=10, objective="External", lower_bound=-10.0, upper_bound=10.0, seed=1338,
                    parent_selection_rule=parent, accept_rule=accept)
    def f(x):
        s = {"Parameters": [float(x[0])]}
        evalmodel(s)
        return s["F(x)"]
    o.set_objective(f)
    o.set_scalar("Termination Criteria/Max Generations", 100)
    o.run(1000)
    assert o.scalar("Best Ever Value") == e["Solver"]["Best Ever Value"]


def test_dea_device_objective_and_resume_through_the_korali_api(tmp_path):
    def make(gens, path):
        e = korali.Experiment()
        e["Problem"]["Type"] = "Optimization"
        e["Problem"]["Objective Function"] = "Rosenbrock"
        for i in range(5):
            e["Variables"][i]["Name"] = "X%d" % i
            e["Variables"][i]["Lower Bound"] = -3.0
            e["Variables"][i]["Upper Bound"] = 3.0
        e["Solver"]["Type"] = "Optimizer/DEA"
        e["Solver"]["Population Size"] = 64
        e["Solver"]["Termination Criteria"]["Max Generations"] = gens
        e["Console Output"]["Verbosity"] = "Silent"
        e["File Output"]["Path"] = str(tmp_path / path)
        e["File Output"]["Frequency"] = 20
        e["Random Seed"] = 31
        return e
    k = korali.Engine()
    full = make(40, "full"); k.run(full)
    part = make(20, "part"); k.run(part)
    r = korali.Experiment()
    r["Problem"]["Objective Function"] = "Rosenbrock"
    r.loadState(str(tmp_path / "part" / "latest"))
    r["Solver"]["Termination Criteria"]["Max Generations"] = 40
    r["File Output"]["Enabled"] = False
    k.run(r)
    assert r["Current Generation"] == 40
    assert r["Solver"]["Best Ever Value"] == full["Solver"]["Best Ever Value"]          # counter-based draws: resume is bitwise
    assert r["Solver"]["Sample Population"] == full["Solver"]["Sample Population"]
    with pytest.raises(RuntimeError, match="Self Adaptive"):
        e = make(5, "x"); e["Solver"]["Mutation Rule"] = "Self Adaptive"; k.run(e)
