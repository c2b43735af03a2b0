"""CPU tests of the secular-equation solver of the divide & conquer eigensolver (korali_b200/csrc/dc_inner.cuh).

The header is host/device code; here it is built with g++ (one lane instead of the 32 lanes of a warp) and checked against
numpy: the roots must interlace the poles, agree with the eigenvalues of diag(dl) + rho w w^T, and — the property the GPU
path relies on (Gu & Eisenstat) — the eigenvectors rebuilt from the accurate pole differences through the Loewner weights
must be orthonormal to FP64 precision even when roots sit within 1e-28 of a pole.
"""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def secular(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dcinner") / "libdcinner.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "helpers", "dc_inner_host.cpp")])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.dc_secular_host.argtypes = [C.c_int, dp, dp, C.c_double, dp, dp, C.POINTER(C.c_int)]

    def run(dl, w, rho):
        k = len(dl)
        dl = np.ascontiguousarray(dl, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64)
        lam = np.zeros(k); delta = np.zeros((k, k)); it = np.zeros(k, dtype=np.int32)
        lib.dc_secular_host(k, dl.ctypes.data_as(dp), w.ctypes.data_as(dp), rho, lam.ctypes.data_as(dp),
                            delta.ctypes.data_as(dp), it.ctypes.data_as(C.POINTER(C.c_int)))
        return lam, delta, it
    return run


def loewner_vectors(dl, w, delta):
    """Eigenvectors of diag(dl) + rho w_hat w_hat^T from the pole differences (dlaed3's formula)."""
    k = len(dl)
    what = np.zeros(k)
    for i in range(k):
        pr = delta[i, i]
        for j in range(k):
            if j != i:
                pr *= delta[j, i] / (dl[i] - dl[j])
        what[i] = np.copysign(np.sqrt(abs(pr)), w[i])
    u = what[None, :] / delta
    u /= np.linalg.norm(u, axis=1)[:, None]
    return u, what


CASES = {
    "random": lambda rng, k: (np.sort(rng.uniform(0, 10, k)), rng.standard_normal(k), 1.3),
    "clustered": lambda rng, k: (1.0 + 1e-7 * np.sort(rng.uniform(0, 1, k)), rng.standard_normal(k), 0.4),
    "tiny weights": lambda rng, k: (np.sort(rng.uniform(0, 10, k)), rng.standard_normal(k) * 10.0 ** rng.uniform(-14, 0, k), 1.7),
    "graded": lambda rng, k: (np.sort(10.0 ** rng.uniform(-6, 3, k)), rng.standard_normal(k), 2e-3),
    "large rho": lambda rng, k: (np.sort(rng.uniform(0, 1, k)), rng.standard_normal(k), 1e6),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("k", [1, 2, 3, 17, 64, 300])
def test_secular_roots_and_loewner_vectors(secular, case, k):
    rng = np.random.default_rng(k * 131 + len(case))
    dl, w, rho = CASES[case](rng, k)
    w = w / np.linalg.norm(w)
    lam, delta, it = secular(dl, w, rho)
    assert it.max() < 60
    # interlacing (strict on the pole side the root was expanded from, inclusive after rounding on the other)
    assert np.all(lam >= dl) and np.all(lam[:-1] <= dl[1:]) and lam[-1] <= dl[-1] + rho * (w @ w) * (1 + 1e-15)
    ref = np.linalg.eigvalsh(np.diag(dl) + rho * np.outer(w, w))
    scale = max(abs(dl).max(), rho)
    assert np.abs(lam - ref).max() <= 4e-14 * scale
    # the differences carry the sign pattern of interlacing: dl_i - lam_j < 0 for i <= j, > 0 for i > j
    i, j = np.meshgrid(np.arange(k), np.arange(k))
    assert np.all(delta[i <= j] < 0) and np.all(delta[i > j] > 0)
    u, what = loewner_vectors(dl, w, delta)
    assert np.abs(u @ u.T - np.eye(k)).max() <= 1e-13
    # residual of the modified problem, and the modified weights are a tiny relative perturbation of w
    a_hat = np.diag(dl) + rho * np.outer(what, what) / (what @ what) * (w @ w)
    assert np.abs(what / np.linalg.norm(what) - w).max() <= 1e-10
    assert np.abs(a_hat @ u.T - u.T * lam).max() <= 1e-13 * scale


def test_root_next_to_a_pole(secular):
    dl = np.array([1.0, 2.0, 3.0, 7.999998, 8.0 + 1e-9, 8.79])
    w = np.array([0.3, 0.4, 0.5, 1e-9, 0.5, 0.3]); w /= np.linalg.norm(w)
    lam, delta, it = secular(dl, w, 1.7)
    assert it.max() < 30
    assert 0 < -delta[3, 3] < 1e-16          # root 3 sits ~1e-18 above its pole; the difference is still resolved
    u, _ = loewner_vectors(dl, w, delta)
    assert np.abs(u @ u.T - np.eye(6)).max() < 1e-14
