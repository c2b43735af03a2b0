"""GPU parity tests of the tridiagonalisation-based eigensolver (korali_b200/csrc/tridiag.cu, dc.cu), stage by stage through the
C ABI (kcma_k_tridiag_stage) and end to end (kcma_k_eigen): replaces eigen() = gsl_eigen_symmv + sort, CMAES.cpp.base:896-938.
The reference's eigenvectors are unique only up to sign / rotation inside clusters, so the checks are the invariants SURVEY 8(c)
names: residual, orthonormality, eigenvalues against numpy (LAPACK), ascending order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from korali_b200 import _lib  # noqa: E402


def spd(n, seed, lo=0.0, hi=3.0):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.sort(10.0 ** rng.uniform(lo, hi, n))
    c = (q * lam) @ q.T
    return 0.5 * (c + c.T), lam


def q_from_reflectors(vr, tau):
    n = vr.shape[0]
    q = np.eye(n)
    for i in range(n - 2, -1, -1):          # Q = H_0 H_1 ... H_{n-2}
        v = vr[i]
        q -= tau[i] * np.outer(v, v @ q)
    return q


@pytest.mark.parametrize("n", [4, 5, 33, 100, 257, 600, 1000, 1001])
def test_sytrd_reduces_to_a_similar_tridiagonal(n):
    c, lam = spd(n, n)
    d, e, tau, vr = _lib.k_sytrd(c)
    t = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.abs(np.linalg.eigvalsh(t) - lam).max() <= 1e-13 * lam.max()
    # reflectors: unit leading entry at i+1, zero up to i, H_i orthogonal (tau = 2 / v.v or 0)
    for i in (0, n // 2, n - 3):
        if i < 0 or i > n - 2:
            continue
        assert np.all(vr[i, : i + 1] == 0) and vr[i, i + 1] == 1.0
        assert tau[i] == 0 or abs(tau[i] * (vr[i] @ vr[i]) - 2.0) < 1e-14
    if n <= 600:
        q = q_from_reflectors(vr, tau)
        assert np.abs(q @ t @ q.T - c).max() <= 1e-13 * np.abs(c).max()


def test_sytrd_global_memory_variant_matches_register_variant(monkeypatch):
    """Same algorithm, different summation order of the products A v (registers + shared-memory partials vs. one warp per column):
    the tridiagonals agree to round-off, entry by entry (the sign choices of the reflectors are the same)."""
    for n in (300, 700, 1300):                   # sytrd_reg_kernel<1,4>, <2,7>, <3,11>
        c, lam = spd(n, 7 + n)
        monkeypatch.delenv("KCMA_SYTRD_RESIDENT", raising=False)
        d0, e0, tau0, vr0 = _lib.k_sytrd(c)
        monkeypatch.setenv("KCMA_SYTRD_RESIDENT", "0")
        d1, e1, tau1, vr1 = _lib.k_sytrd(c)
        sc = lam.max()
        # the ENTRIES of T are not determined to working precision (round-off grows along the N-1 dependent steps), its spectrum is
        t0 = np.linalg.eigvalsh(np.diag(d0) + np.diag(e0, 1) + np.diag(e0, -1))
        t1 = np.linalg.eigvalsh(np.diag(d1) + np.diag(e1, 1) + np.diag(e1, -1))
        assert np.abs(t0 - t1).max() <= 1e-13 * sc and np.abs(t0 - lam).max() <= 1e-13 * sc, n
        assert np.abs(d0 - d1).max() <= 1e-7 * sc and np.abs(e0 - e1).max() <= 1e-7 * sc, n
        assert np.abs(tau0 - tau1).max() <= 1e-6 and np.abs(vr0 - vr1).max() <= 1e-6, n


def test_sytrd_large_n_streams_from_global_memory():
    n = 1800                                   # above the shared-memory-resident limit
    c, lam = spd(n, 3, 0.0, 2.0)
    d, e, tau, vr = _lib.k_sytrd(c)
    t_eigs = np.linalg.eigvalsh(np.diag(d) + np.diag(e, 1) + np.diag(e, -1))
    assert np.abs(t_eigs - lam).max() <= 1e-13 * lam.max()


def tri_cases(n, rng):
    yield "random", rng.standard_normal(n), rng.standard_normal(n - 1)
    yield "clustered", 1.0 + 1e-7 * rng.standard_normal(n), 1e-7 * rng.standard_normal(n - 1)
    yield "identity", np.full(n, 2.0), np.zeros(n - 1)
    yield "toeplitz", np.full(n, 2.0), np.full(n - 1, -1.0)
    yield "wilkinson", np.abs(np.arange(n) - n // 2).astype(float), np.ones(n - 1)
    yield "graded", 10.0 ** np.linspace(0, -8, n), 10.0 ** np.linspace(-1, -9, n - 1)
    if n >= 4:
        e = rng.standard_normal(n - 1); e[n // 3] = 0.0; e[n // 2] = 1e-300
        yield "split", rng.standard_normal(n), e


@pytest.mark.parametrize("n", [2, 3, 31, 32, 33, 64, 100, 257, 1000])
def test_divide_and_conquer_on_tridiagonals(n):
    rng = np.random.default_rng(n)
    for name, d, e in tri_cases(n, rng):
        t = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        lam, zt = _lib.k_stedc(d, e)
        sc = max(np.abs(t).max(), 1e-300)
        assert np.all(np.diff(lam) >= 0), name
        assert np.abs(lam - np.linalg.eigvalsh(t)).max() <= 2e-14 * sc * max(1, n / 100), name
        assert np.abs(zt @ zt.T - np.eye(n)).max() <= 1e-13, name
        assert np.abs(t @ zt.T - zt.T * lam).max() <= 1e-13 * sc, name


@pytest.mark.parametrize("n", [25, 64, 100, 257, 1000, 1001, 2000])
def test_eigen_tridiag_path_end_to_end(n):
    c, lam = spd(n, 11 + n)
    w, v = _lib.k_eigen(c)
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - lam).max() <= 1e-13 * lam.max()
    assert np.abs(v @ np.diag(w) @ v.T - c).max() <= 1e-12 * np.abs(c).max()
    assert np.abs(v.T @ v - np.eye(n)).max() <= 1e-12
    # sign convention shared with the oracle: the component of largest magnitude of every eigenvector is positive
    assert np.all(v[np.abs(v).argmax(axis=0), np.arange(n)] > 0)


def test_eigen_tridiag_cma_like_clustered_spectrum():
    n = 1000
    rng = np.random.default_rng(5)
    e = rng.standard_normal((n, 3 * n))
    c = 0.97 * np.eye(n) + 0.03 * (e @ e.T) / (3 * n)
    w, v = _lib.k_eigen(c)
    assert np.abs(w - np.linalg.eigvalsh(c)).max() <= 1e-13
    assert np.abs(c @ v - v * w).max() <= 1e-12 and np.abs(v.T @ v - np.eye(n)).max() <= 1e-12


def test_eigen_paths_agree(monkeypatch):
    c, lam = spd(300, 99)
    w0, v0 = _lib.k_eigen(c)
    monkeypatch.setenv("KCMA_EIGEN", "jacobi")
    w1, v1 = _lib.k_eigen(c)
    assert np.abs(w0 - w1).max() <= 1e-12 * lam.max()
    assert np.abs(np.abs(np.sum(v0 * v1, axis=0)) - 1.0).max() <= 1e-9      # same vectors (spectrum is simple)
