"""Pins the CPU oracle (oracle/okcma.c) against the reference's own saved trajectory
tests/python/plot/cmaes/gen00000000..100.json (fixture: tests/golden/cmaes_plot_trajectory.npz)."""
import numpy as np
import pytest
from oracle import oracle as O
from conftest import relerr
from korali_b200._abi import INJ_BD, INJ_F, INJ_X, INJ_BDZ, INJ_GRAD, KcmaError

N, LAM, MU = 10, 32, 16


def make(golden, **kw):
    return O.Oracle(n=N, population_size=LAM, objective="NegSphere", seed=int(golden["Normal Generator Seed"][0]),
                    lower_bound=golden["Lower Bound"], upper_bound=golden["Upper Bound"],
                    initial_value=golden["Initial Value"], initial_stddev=golden["Initial Standard Deviation"], **kw)


def test_init_constants_bit_exact(golden):
    """initMuWeights / initCovariance (CMAES.cpp.base:233-313) reproduce the saved constants bit-for-bit."""
    o = make(golden)
    assert np.array_equal(o.get("Mu Weights"), golden["Mu Weights"][1])
    for k in ["Effective Mu", "Sigma Cumulation Factor", "Damp Factor", "Cumulative Covariance", "Chi Square Number", "Trace"]:
        assert o.scalar(k) == golden[k][1], k
    assert o.scalar("Sigma") == np.sqrt(22.5)
    assert np.array_equal(o.get("Axis Lengths"), np.ones(N))
    assert np.array_equal(o.get("Covariance Eigenvector Matrix").reshape(N, N), np.eye(N))


def test_mt19937_gaussian_stream_bit_exact(golden):
    """gsl_rng_mt19937 + gsl_ran_gaussian restatement: generation 1 samples are m + sigma*z bit-for-bit (F7)."""
    z = O.mt19937_gaussian(int(golden["Normal Generator Seed"][0]), LAM * N)
    x = golden["Initial Value"][None, :] + np.sqrt(22.5) * (1.0 * z.reshape(LAM, N))
    assert np.array_equal(x.ravel(), golden["Sample Population"][1])
    assert np.array_equal(z, golden["BDZ Matrix"][1])


def test_first_two_generations_free_running(golden):
    """Free-running oracle (own RNG stream): generation 1 fully bit-exact; generation 2 'BDZ Matrix' bit-exact when
    the eigenvectors GSL produced are injected (eigenvector signs are solver-specific)."""
    o = make(golden)
    o.run_generation()
    assert np.array_equal(o.get("Sample Population"), golden["Sample Population"][1])
    assert np.array_equal(o.get("Value Vector"), golden["Value Vector"][1]) or relerr(o.get("Value Vector"), golden["Value Vector"][1]) < 4e-16
    o.inject(INJ_F, golden["Value Vector"][1])  # python-summed F(x) of the fixture
    o2 = make(golden)
    o2.ask(); o2.inject(INJ_F, golden["Value Vector"][1]); o2.eval(); o2.tell()
    assert np.array_equal(o2.get_index("Sorting Index"), golden["Sorting Index"][1].astype(np.uint64))
    assert np.array_equal(o2.get("Current Mean"), golden["Current Mean"][1])
    assert relerr(o2.get("Covariance Matrix"), golden["Covariance Matrix"][1]) < 1e-15
    o2.inject(INJ_BD, np.concatenate([golden["Covariance Eigenvector Matrix"][2], golden["Axis Lengths"][2]]))
    o2.ask()
    assert np.array_equal(o2.get("BDZ Matrix"), golden["BDZ Matrix"][2])
    assert np.array_equal(o2.get("Sample Population"), golden["Sample Population"][2])


def load_state(o, golden, g):
    for k in ["Covariance Matrix", "Current Mean", "Previous Mean", "Evolution Path", "Conjugate Evolution Path",
              "Best Ever Variables"]:
        o.set(k, golden[k][g])
    for k in ["Sigma", "Best Ever Value", "Current Best Value", "Previous Best Value", "Previous Best Ever Value"]:
        o.set_scalar(k, golden[k][g])
    o.set_scalar("Current Generation", g)
    o.set_scalar("Model Evaluation Count", golden["Model Evaluation Count"][g])


def test_all_100_transitions(golden):
    """State(g-1) + the reference's {B, D, X, F}(g) -> State(g) for g = 1..100."""
    worst = {}
    for g in range(1, 101):
        o = make(golden)
        if g > 1:
            load_state(o, golden, g - 1)
        o.inject(INJ_BD, np.concatenate([golden["Covariance Eigenvector Matrix"][g], golden["Axis Lengths"][g]]))
        o.inject(INJ_X, golden["Sample Population"][g])
        o.inject(INJ_F, golden["Value Vector"][g])
        o.run_generation()
        assert np.array_equal(o.get_index("Sorting Index"), golden["Sorting Index"][g].astype(np.uint64)), g
        for k in ["Current Mean", "Previous Mean", "Mean Update", "Evolution Path"]:
            assert np.array_equal(o.get(k), golden[k][g]), (g, k)
        for k in ["Conjugate Evolution Path", "Covariance Matrix", "Best Ever Variables", "Current Best Variables"]:
            e = relerr(o.get(k), golden[k][g]); worst[k] = max(worst.get(k, 0), e)
            assert e < 2e-15, (g, k, e)
        for k in ["Sigma", "Conjugate Evolution Path L2 Norm", "Best Ever Value", "Current Best Value",
                  "Maximum Diagonal Covariance Matrix Element", "Minimum Diagonal Covariance Matrix Element",
                  "Current Min Standard Deviation", "Current Max Standard Deviation",
                  "Maximum Covariance Eigenvalue", "Minimum Covariance Eigenvalue"]:
            a, b = o.scalar(k), golden[k][g]
            assert abs(a - b) <= 2e-15 * abs(b), (g, k, a, b)
        assert o.scalar("Current Generation") == g
        assert o.scalar("Model Evaluation Count") == golden["Model Evaluation Count"][g]


def test_eigen_restatement_against_gsl_outputs(golden):
    """Householder+QL restatement of gsl_eigen_symmv: eigenvalues vs the fixture's 'Axis Lengths'^2, residuals,
    orthogonality, ascending order, and B D^2 B^T invariance (the only B-dependent quantity that must agree)."""
    for g in range(2, 101):
        c = golden["Covariance Matrix"][g - 1].reshape(N, N)
        w, q = O.eigen(c)
        assert np.all(np.diff(np.abs(w)) >= 0)
        assert relerr(w, golden["Axis Lengths"][g] ** 2) < 1e-13
        assert np.abs(q @ np.diag(w) @ q.T - c).max() < 1e-13 * np.abs(c).max()
        assert np.abs(q.T @ q - np.eye(N)).max() < 1e-13
        bref = golden["Covariance Eigenvector Matrix"][g].reshape(N, N)
        # same vectors up to sign
        dots = np.abs(np.sum(q * bref, axis=0))
        assert np.all(dots > 1 - 1e-9), (g, dots.min())


def test_sort_index_semantics():
    """sort_index (CMAES.cpp.base:940-950): descending; ascending index among equal values (defined tie-break)."""
    f = np.array([1.0, 3.0, 3.0, -np.inf, 2.0, 3.0, 0.0, -0.0])
    assert list(O.sort_index(f)) == [1, 2, 5, 4, 0, 6, 7, 3]
    rng = np.random.default_rng(0)
    f = rng.standard_normal(1000)
    assert np.array_equal(O.sort_index(f), np.argsort(-f, kind="stable").astype(np.uint64))
    f = rng.integers(0, 5, 1000).astype(np.float64)
    assert np.array_equal(O.sort_index(f), np.argsort(-f, kind="stable").astype(np.uint64))
    assert O.sort_index(np.array([])).size == 0


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normals_are_standard_normal():
    z = O.philox_normal(1337, 1, 0, 4096, 100)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 0.06
    # odd N: the last column uses the cosine half of its pair
    z2 = O.philox_normal(1337, 1, 0, 8, 7)
    z3 = O.philox_normal(1337, 1, 0, 8, 8)
    assert np.array_equal(z2, z3[:, :7])


# ---------------------------------------------------------------- Use Gradient Information ------------------
@pytest.mark.parametrize("obj", ["NegSphere", "NegSumSq", "NegEllipsoid", "NegRosenbrock", "NegAckley", "NegSphereSin2"])
def test_builtin_objective_gradients_match_finite_differences(obj):
    """The analytic dF/dx the built-in objectives hand to the gradient path (the reference takes them from the user model,
    examples/optimization/stochastic/_model/model.py:10-63) against central differences of the objective itself."""
    rng = np.random.default_rng(4)
    n = 7
    x = rng.standard_normal((5, n)) * 0.7 + 0.3
    coef = 10.0 ** (2.0 * np.arange(n) / (n - 1))
    g = O.objective_gradient(obj, x, coef)
    h = 1e-6
    for d in range(n):
        e = np.zeros(n); e[d] = h
        fd = (O.objective(obj, x + e, coef) - O.objective(obj, x - e, coef)) / (2 * h)
        assert np.abs(g[:, d] - fd).max() < 1e-6 * max(1.0, np.abs(fd).max()), (obj, d)
    if obj == "NegRosenbrock":   # the reference's own formula (model.py:23-34)
        ref = np.zeros_like(x)
        for i in range(n - 1):
            ref[:, i] += 2. * (1 - x[:, i]) + 200. * (x[:, i + 1] - x[:, i]**2) * 2 * x[:, i]
            ref[:, i + 1] -= 200. * (x[:, i + 1] - x[:, i]**2)
        assert np.abs(g - ref).max() < 1e-12 * np.abs(ref).max()


def test_gradient_step_of_the_mean_follows_the_reference_loop():
    """CMAES.cpp.base:603-621: weighted mean, then mean[d] += sum_i w_i * step / sqrt(N) * gradient[sorted_i][d]."""
    # (the reference's step is not normalised: Rosenbrock gradients of ~1e4 need a tiny step size or the mean runs away)
    case = dict(n=6, population_size=16, objective="NegRosenbrock", initial_value=0.3, initial_stddev=0.8, seed=5)
    a = O.Oracle(use_gradient_information=1, gradient_step_size=2e-5, **case)
    for g in range(6):
        a.ask(); a.eval()
        x = a.get("Sample Population").reshape(16, 6)
        grads = a.get("Gradients").reshape(16, 6)
        assert np.array_equal(grads, O.objective_gradient("NegRosenbrock", x))
        a.tell()
        idx = a.get_index("Sorting Index").astype(int)
        w = a.get("Mu Weights")
        want = np.zeros(6)
        for d in range(6):
            m = 0.0
            for i in range(len(w)):
                m += w[i] * x[idx[i], d]
            for i in range(len(w)):
                m += w[i] * float(np.float32(2e-5)) / np.sqrt(6.0) * grads[idx[i], d]   # the step size is a float in the reference (Q9)
            want[d] = m
        assert np.array_equal(a.get("Current Mean"), want), g
    plain = O.Oracle(**case)
    for g in range(6):
        plain.run_generation()
    assert not np.allclose(plain.get("Current Mean"), a.get("Current Mean"))   # the gradient step really changes the search
    with pytest.raises(KcmaError, match="Gradient Step Size must be larger than 0.0"):
        O.Oracle(use_gradient_information=1, gradient_step_size=0.0, **case)


def test_injected_gradients_replace_the_model_gradients():
    case = dict(n=4, population_size=8, objective="External", initial_value=1.0, initial_stddev=0.5, seed=2,
                use_gradient_information=1, gradient_step_size=0.1)
    o = O.Oracle(**case)
    o.ask()
    x = o.get("Sample Population").reshape(8, 4)
    o.inject(INJ_F, -0.5 * (x**2).sum(1))
    with pytest.raises(KcmaError, match="inject the gradients"):
        o.eval()
    o.inject(INJ_F, -0.5 * (x**2).sum(1)); o.inject(INJ_GRAD, (-x).ravel())
    o.eval(); o.tell()
    assert np.isfinite(o.get("Current Mean")).all()


# ---------------------------------------------------------------- discrete variables (Granularity) ----------
def test_discrete_variables_restatement():
    """CMAES.cpp.base:44-50, 515-544, 668, 730-734, 834-867: samples live on the grid of the discrete variables, the masking
    matrices follow their definition, the step-size update uses the masked path length, the search reaches the grid optimum."""
    n = 6
    gran = np.array([1.0, 1.0, 0.0, 0.0, 0.5, 0.0])
    o = O.Oracle(n=n, population_size=16, objective="NegSphere", initial_value=3.3, initial_stddev=2.0, seed=3, granularity=gran,
                 lower_bound=-20.0, upper_bound=20.0)
    o.set_scalar("Oracle/RNG Kind", 1)
    assert o.scalar("Number Of Discrete Mutations") == 0 and o.scalar("Number Masking Matrix Entries") == 0
    assert abs(o.scalar("Chi Square Number Discrete Mutations") - o.scalar("Chi Square Number")) < 1e-15
    for g in range(80):
        o.ask()
        x = o.get("Sample Population").reshape(16, n)
        disc = gran > 0
        assert np.array_equal(x[:, disc], np.round(x[:, disc] / gran[disc]) * gran[disc]), g        # discretize()
        ndm = int(o.scalar("Number Of Discrete Mutations"))
        dm = o.get("Discrete Mutations").reshape(16, n)
        assert not dm[ndm:].any()                                   # only the first samples are mutated ...
        mask = o.get("Masking Matrix")
        for i in range(max(ndm - 1, 0)):                            # ... in ONE masked dimension, by a multiple of its granularity
            nz = np.flatnonzero(dm[i])
            assert len(nz) <= 1 and all(mask[d] == 1.0 and abs(dm[i, d] / gran[d]) >= 1 and dm[i, d] / gran[d] == round(dm[i, d] / gran[d]) for d in nz)
        sigma_before, ps_before = o.scalar("Sigma"), None
        o.eval(); o.tell()
        c = o.get("Covariance Matrix").reshape(n, n)
        ps = o.get("Conjugate Evolution Path")
        # updateDiscreteMutationMatrix with the sigma of before updateSigma and the new C
        sd = sigma_before * np.sqrt(np.diag(c))
        cs = o.scalar("Sigma Cumulation Factor")
        msig = np.where(sd / np.sqrt(cs) < 0.2 * gran, 0.0, 1.0)
        mk = np.where(2.0 * sd < gran, 1.0, 0.0)
        assert np.array_equal(o.get("Masking Matrix Sigma"), msig) and np.array_equal(o.get("Masking Matrix"), mk), g
        entries = n + 1 - int((msig == 0).sum())
        chi = np.sqrt(entries) * (1. - 1. / (4. * entries) + 1. / (21. * entries * entries))
        assert abs(o.scalar("Chi Square Number Discrete Mutations") - chi) < 1e-15
        assert o.scalar("Number Of Discrete Mutations") == min(round(16 / 10.0 + mk.sum() + 1), np.floor(16 / 2.0) - 1)
        want_sigma = sigma_before * np.exp(cs / o.scalar("Damp Factor") * (np.sqrt((msig * ps * ps).sum()) / chi - 1.))
        # (escape-flat / sigma-bound corrections may apply on top: only check when they did not)
        if abs(o.scalar("Sigma") - want_sigma) > 1e-12 * want_sigma:
            assert o.scalar("Sigma") > want_sigma * 0.999
    xb = o.get("Best Ever Variables")
    assert np.all(xb[gran > 0] == 0.0) and np.abs(xb).max() < 1e-2 and o.scalar("Number Masking Matrix Entries") == 3
    with pytest.raises(KcmaError, match="Negative granularity"):
        O.Oracle(n=2, population_size=8, objective="NegSphere", initial_value=1.0, initial_stddev=1.0, granularity=np.array([1.0, -1.0]))
