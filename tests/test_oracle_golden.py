"""Pins the CPU oracle (oracle/okcma.c) against the reference's own saved trajectory
tests/python/plot/cmaes/gen00000000..100.json (fixture: tests/golden/cmaes_plot_trajectory.npz)."""
import numpy as np
import pytest
from oracle import oracle as O
from conftest import relerr
from korali_b200._abi import INJ_BD, INJ_F, INJ_X, INJ_BDZ

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
