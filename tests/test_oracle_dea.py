"""CPU tests of the Differential Evolution oracle (oracle/odea.c, the C restatement of DEA.cpp.base on the Philox streams of
include/kdea.h). The reference ships no DEA trajectory, so the oracle is anchored by the reference's own statistical thresholds:
tests/statistical/optimizers/correctness/run-dea.py (checkMin(e, 0.23246, tol) for every rule combination it runs)."""
import ctypes as C
import numpy as np
import pytest
from korali_b200._dea import KdeaCfg
from korali_models.models import evalmodel
from oracle import oracle as O


def model(x):
    s = {"Parameters": [float(v) for v in x]}
    evalmodel(s)
    return s["F(x)"]


# run-dea.py: (Parent Selection Rule, Accept Rule, tolerance of checkMin)
@pytest.mark.parametrize("parent,accept,tol", [("Random", "Greedy", 1e-4), ("Random", "Best", 1e-2), ("Random", "Improved", 1e-2),
                                                ("Random", "Iterative", 1e-2), ("Best", "Greedy", 1e-4), ("Best", "Iterative", 1e-2)])
def test_reference_thresholds_of_run_dea(parent, accept, tol):
    o = O.OracleDEA(n=1, population_size=10, objective="External", lower_bound=-10.0, upper_bound=10.0, seed=1337,
                    parent_selection_rule=parent, accept_rule=accept)
    o.set_objective(model)
    o.set_scalar("Termination Criteria/Max Generations", 100)
    assert o.run(1000) == 100
    assert np.isclose(0.23246, o.scalar("Best Ever Value"), atol=tol), o.scalar("Best Ever Value")
    assert o.scalar("Model Evaluation Count") == 1000
    fin, why = o.check_termination()
    assert fin and why == "solver['Max Generations'];"


def test_generation_structure_and_quirks():
    """Generation 1 evaluates the uniform initial population (:91-101, :105); Greedy accepts everything against -Inf (:231-235);
    candidates stay inside the box (rejection loop :107-118); 'Current Minimum Step Size' stays +Inf (:280-281 discards std::min)."""
    lo, up = np.array([-1.0, 0.0, 2.0]), np.array([1.0, 5.0, 2.5])
    o = O.OracleDEA(n=3, population_size=12, objective="NegSphere", lower_bound=lo, upper_bound=up, seed=7)
    x0 = o.get("Sample Population").reshape(12, 3)
    assert np.all(x0 >= lo) and np.all(x0 <= up) and np.array_equal(x0, o.get("Candidate Population").reshape(12, 3))
    assert np.allclose(o.get("Current Mean"), x0.mean(0), rtol=1e-14)
    o.run_generation()
    assert np.array_equal(o.get("Sample Population").reshape(12, 3), x0)
    assert o.scalar("Best Ever Value") == o.get("Value Vector").max() and o.scalar("Best Sample Index") == o.get("Value Vector").argmax()
    for _ in range(30):
        o.run_generation()
        xc = o.get("Candidate Population").reshape(12, 3)
        assert np.all(xc >= lo) and np.all(xc <= up)
        assert np.all(o.get("Max Distances") >= 0)
    assert o.scalar("Current Minimum Step Size") == np.inf
    assert o.scalar("Infeasible Sample Count") > 0            # mutants do leave this narrow box and are drawn again


def test_termination_criteria():
    o = O.OracleDEA(n=2, population_size=16, objective="NegSphere", lower_bound=-3.0, upper_bound=3.0, seed=3)
    o.set_scalar("Termination Criteria/Min Value", 1e-3)       # -bestEver < 1e-3
    done = o.run(10000)
    fin, why = o.check_termination()
    assert fin and "DEA['Min Value']" in why and done < 10000 and -o.scalar("Best Ever Value") < 1e-3
    o = O.OracleDEA(n=2, population_size=16, objective="NegSphere", lower_bound=-3.0, upper_bound=3.0, seed=3)
    o.set_scalar("Termination Criteria/Max Model Evaluations", 160)
    assert o.run(10000) == 10


def test_cfg_struct_layout_matches_c():
    import os, subprocess, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "kdea.h"
    int main(){ printf("%zu %zu %zu %zu %zu\n", sizeof(kdea_cfg), offsetof(kdea_cfg, mutation_rule), offsetof(kdea_cfg, seed),
      offsetof(kdea_cfg, device), offsetof(kdea_cfg, objective_coef)); return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    assert [int(x) for x in out] == [C.sizeof(KdeaCfg), KdeaCfg.mutation_rule.offset, KdeaCfg.seed.offset, KdeaCfg.device.offset,
                                     KdeaCfg.objective_coef.offset]
