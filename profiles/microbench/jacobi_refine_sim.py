"""CPU feasibility study for the next round: can a GEMM-only iterative refinement
(Ogita & Aishima 2018, "Iterative refinement for symmetric eigenvalue decomposition", Algorithm 1: 4 N^3-GEMMs per step,
quadratically convergent) replace the TAIL sweeps of the blocked one-sided Jacobi? A Jacobi sweep at N = 1000 is a chain of 255
dependent steps (1.7 ms on the B200); a refinement step is four 1000^3 GEMMs on the DMMA pipe (~0.5 ms).

For a config-3-like covariance (clustered spectrum, a few generations in) the script runs the ring-ordered blocked Jacobi of
jacobi_orderings_sim.py from the previous generation's eigenvectors and, after every sweep, hands the current basis to the
refinement: how many refinement steps until the decomposition is at the accuracy the Jacobi alone ends with?

    python profiles/microbench/jacobi_refine_sim.py [N]
"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import jacobi_orderings_sim as J


def refine_step(A, X, normA):
    n = A.shape[0]
    R = np.eye(n) - X.T @ X
    S = X.T @ (A @ X)
    lam = np.diag(S) / (1.0 - np.diag(R))
    delta = 2.0 * (np.linalg.norm(S - np.diag(lam), 2) + normA * np.linalg.norm(R, 2))
    dl = lam[None, :] - lam[:, None]                    # lam_j - lam_i
    sep = np.abs(dl) > delta
    with np.errstate(divide="ignore", invalid="ignore"):
        E = np.where(sep, (S + lam[None, :] * R) / dl, 0.5 * R)
    return X + X @ E, lam, float(sep.sum() - 0) / (n * (n - 1))


def quality(A, X, lam):
    n = A.shape[0]
    return (np.abs((X * lam) @ X.T - A).max() / np.abs(A).max(), np.abs(X.T @ X - np.eye(n)).max())


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rng = np.random.default_rng(0)
    def gen_c(c, cmu=0.035, m=4 * n):
        z = rng.standard_normal((m, n)) @ np.linalg.cholesky(c).T
        return (1 - cmu) * c + cmu * (z.T @ z) / m
    c = np.eye(n)
    for _ in range(4):
        c_prev, c = c, gen_c(c)
    _, v_prev = np.linalg.eigh(c_prev)
    normA = np.linalg.norm(c, 2)
    ev = np.linalg.eigvalsh(c)
    print("N = %d, spectrum %.3g .. %.3g, smallest / median gap %.1e / %.1e" % (n, ev[0], ev[-1], np.diff(ev).min(), np.median(np.diff(ev))))
    G = (c @ v_prev).T.copy()
    tol2 = (4 * 2.2e-16 * np.sqrt(n)) ** 2
    for s in range(0, 12):
        if s:
            rot, big = J.sweep(G, 4, tol2, 1e-16, ordering="ring")
        V = np.linalg.solve(c, G.T)                     # the basis the kernel carries alongside (G = C V)
        nrm = np.sqrt(np.einsum("ij,ij->i", G, G))
        cosmax = np.abs((G / nrm[:, None]) @ (G / nrm[:, None]).T - np.eye(n)).max()
        X = V.copy()
        hist = []
        for it in range(8):
            X, lam, sepfrac = refine_step(c, X, normA)
            res, orth = quality(c, X, lam)
            hist.append("%.0e/%.0e(%.0f%%)" % (res, orth, 100 * sepfrac))
            if res < 1e-13 and orth < 1e-13:
                break
        print("after %2d Jacobi sweeps: max |cos| %.1e   refinement residual/orthonormality(separated pairs) per step: %s" % (s, cosmax, "  ".join(hist)), flush=True)
        if s and (rot == 0 or not big):
            print("Jacobi alone: done after %d sweeps" % s)
            break
