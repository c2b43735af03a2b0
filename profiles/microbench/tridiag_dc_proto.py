"""NumPy prototype of the round-2 eigensolver (tridiagonalisation + divide & conquer + WY back-transform), written to pin the
algebra of the CUDA kernels (korali_b200/csrc/tridiag.cu, dc.cu) before any GPU time is spent:

  1. sytrd_fused : Householder tridiagonalisation in the ONE-exchange-per-step formulation of the persistent kernel
                   (every CTA owns columns; per step it receives p = A v and the next column, forms w, the next reflector,
                   and in one pass over its columns applies the rank-2 update and the next matrix-vector product)
  2. dc_eig      : Cuppen / Gu-Eisenstat divide & conquer on (d, e): deflation, secular roots by a bracketed
                   Bunch-Nielsen-Sorensen iteration in shifted coordinates, Loewner weights, "deflation-oblivious" merge GEMM
  3. back_wy     : X^T = Z^T H_{n-3} ... H_0 panel by panel with compact-WY factors

    python profiles/microbench/tridiag_dc_proto.py
"""
import numpy as np

EPS = np.finfo(float).eps


# ------------------------------------------------------------------------------------------- 1. tridiagonalisation
def sytrd_fused(a):
    n = a.shape[0]
    a = a.copy()
    d = np.zeros(n); e = np.zeros(max(n - 1, 0)); tau = np.zeros(max(n - 1, 0))
    vr = np.zeros((n, n))                      # row i = reflector v_i (support i+1..n-1, v_i[i+1] = 1)
    v_old = np.zeros(n); w_old = np.zeros(n)   # rank-2 update of the previous step (zero at the start)
    p = np.zeros(n)
    col = a[:, 0].copy()                       # "published" column 0
    tau_old = 0.0
    for i in range(n):
        # ---- received: p (= A_{i-1} v_{i-1} on i..n-1) and col (= column i of the matrix before the step i-1 update)
        if i > 0:
            y = tau_old * p
            alpha = -0.5 * tau_old * np.dot(y[i:], v_old[i:])
            w_old = y + alpha * v_old
            w_old[:i] = 0.0
            col = col - v_old * w_old[i] - w_old * v_old[i]
        d[i] = col[i]
        if i == n - 1:
            break
        x = col[i + 1:]
        alph = x[0]
        xnorm = np.sqrt(np.dot(x[1:], x[1:]))
        v_new = np.zeros(n)
        if xnorm == 0.0:
            t = 0.0; e[i] = alph
            v_new[i + 1] = 1.0
        else:
            beta = -np.copysign(np.hypot(alph, xnorm), alph)
            t = (beta - alph) / beta
            v_new[i + 2:] = x[1:] / (alph - beta)
            v_new[i + 1] = 1.0
            e[i] = beta
        tau[i] = t; vr[i] = v_new
        # ---- fused pass over the columns c >= i+1 (each CTA: its own columns): update with (v_old, w_old), then p = A v_new
        r = slice(i + 1, n)
        a[r, r] -= np.outer(v_old[r], w_old[r]) + np.outer(w_old[r], v_old[r])
        p = np.zeros(n)
        p[r] = a[r, r].T @ v_new[r]
        col = np.zeros(n)
        col[r] = a[r, i + 1]                   # the owner of column i+1 publishes it (updated up to step i-1)
        v_old, tau_old = v_new, t
    return d, e, tau, vr


# ------------------------------------------------------------------------------------------- 2. divide & conquer
def secular_root(j, dl, w2, rho):
    """Root j of 1/rho + sum_i w2_i / (dl_i - lam) = 0 in (dl_j, dl_{j+1}) (last: (dl_{K-1}, dl_{K-1} + rho*sum w2)).
    Returns (origin index o, mu) with lam = dl[o] + mu, and delta_i = (dl_i - dl_o) - mu."""
    k = len(dl)
    rinv = 1.0 / rho
    last = j == k - 1
    if last:
        o = j
        lo, hi = 0.0, rho * w2.sum()
    else:
        gap = dl[j + 1] - dl[j]
        mid = 0.5 * gap
        dm = (dl - dl[j]) - mid
        fm = rinv + np.sum(w2 / dm)
        if fm >= 0.0:
            o = j; lo, hi = 0.0, mid
        else:
            o = j + 1; lo, hi = -mid, 0.0
    dd = dl - dl[o]
    # initial guess: the two (one) nearest poles exact, the rest frozen at the far end of the bracket
    far = hi if o == j else lo
    if last:
        dm = dd - 0.5 * hi
        rest = rinv + np.sum(w2[:j] / dm[:j])
        mu = hi if rest <= 0 else min(hi, max(w2[j] / rest, 0.0))     # rest + w2_j/(0-mu) = 0
        if not (lo < mu < hi):
            mu = 0.5 * hi
    else:
        dm = dd - far
        mask = np.ones(k, bool); mask[j] = mask[j + 1] = False
        c = rinv + np.sum(w2[mask] / dm[mask])
        # c + a/(d1-mu) + b/(d2-mu) = 0
        d1, d2 = dd[j], dd[j + 1]
        aa, bb = w2[j], w2[j + 1]
        qa = c; qb = -(c * (d1 + d2) + aa + bb); qc = c * d1 * d2 + aa * d2 + bb * d1
        mu = None
        if qa != 0.0:
            disc = qb * qb - 4 * qa * qc
            if disc >= 0:
                sq = np.sqrt(disc)
                q = -0.5 * (qb + np.copysign(sq, qb))
                for r in ((q / qa), (qc / q if q != 0 else np.inf)):
                    if lo < r < hi:
                        mu = r
        elif qb != 0.0:
            r = -qc / qb
            if lo < r < hi:
                mu = r
        if mu is None:
            mu = 0.5 * (lo + hi)
    for it in range(80):
        dlt = dd - mu
        terms = w2 / dlt
        if last:
            psi = terms.sum(); phi = 0.0
            dpsi = np.sum(terms / dlt); dphi = 0.0
        else:
            psi = terms[:j + 1].sum(); phi = terms[j + 1:].sum()
            dpsi = np.sum(terms[:j + 1] / dlt[:j + 1]); dphi = np.sum(terms[j + 1:] / dlt[j + 1:])
        f = rinv + psi + phi
        err = EPS * (8.0 * (np.abs(terms).sum()) + rinv + abs(mu) * (dpsi + dphi)) 
        if abs(f) <= err:
            break
        if f < 0: lo = max(lo, mu)
        else: hi = min(hi, mu)
        if hi - lo <= 2 * EPS * max(abs(lo), abs(hi)):
            mu = 0.5 * (lo + hi); break
        D1 = dlt[j]
        if last:
            S = dpsi * D1 * D1; s = psi - dpsi * D1
            c = rinv + s
            eta = D1 + S / c if c != 0 else np.inf     # c + S/(D1-eta) = 0
        else:
            D2 = dlt[j + 1]
            S = dpsi * D1 * D1; s = psi - dpsi * D1
            R = dphi * D2 * D2; r_ = phi - dphi * D2
            c = rinv + s + r_
            qa = c; qb = c * (D1 + D2) + S + R; qc = D1 * D2 * f
            disc = abs(qb * qb - 4 * qa * qc)
            if qa == 0.0:
                eta = qc / qb if qb != 0 else np.inf
            elif qb <= 0:
                eta = (qb - np.sqrt(disc)) / (2 * qa)
            else:
                eta = 2 * qc / (qb + np.sqrt(disc))
        new = mu + eta
        if not (lo < new < hi) or not np.isfinite(new):
            new = 0.5 * (lo + hi)
        mu = new
    return o, mu, dd - mu


def merge(d1, q1, d2, q2, beta, stats):
    n1, n2 = len(d1), len(d2); n = n1 + n2
    rho = abs(beta)
    d = np.concatenate([d1, d2])
    z = np.concatenate([q1[-1, :], np.sign(beta) * q2[0, :] if beta != 0 else q2[0, :]])
    q = np.zeros((n, n)); q[:n1, :n1] = q1; q[n1:, n1:] = q2
    z = z / np.sqrt(2.0); rho = 2.0 * rho
    order = np.argsort(d, kind="stable")
    tol = 8.0 * EPS * max(np.abs(d).max(), np.abs(z).max())
    defl = np.zeros(n, bool)
    if rho * np.abs(z).max() <= tol:
        defl[:] = True
    else:
        pj = -1
        for idx in order:
            if rho * abs(z[idx]) <= tol:
                defl[idx] = True
                continue
            if pj < 0:
                pj = idx; continue
            s, c = z[pj], z[idx]
            t = np.hypot(c, s)
            tt = d[idx] - d[pj]
            c /= t; s = -s / t
            if abs(tt * c * s) <= tol:
                z[idx] = t; z[pj] = 0.0
                qp, qn = q[:, pj].copy(), q[:, idx].copy()
                q[:, pj] = c * qp + s * qn
                q[:, idx] = -s * qp + c * qn
                tnew = d[pj] * c * c + d[idx] * s * s
                d[idx] = d[pj] * s * s + d[idx] * c * c
                d[pj] = tnew
                defl[pj] = True
                pj = idx
            else:
                pj = idx
    # the rotations can move d[pj] slightly: the non-deflated list must be strictly ascending -> re-sort the survivors
    nd = [i for i in np.argsort(d, kind="stable") if not defl[i]]
    k = len(nd)
    stats["deflated"] = stats.get("deflated", 0) + (n - k)
    ut = np.zeros((n, n))                       # row = new eigenvector in the basis of the old columns
    lam_all = np.zeros(n)
    if k > 0:
        dl = d[nd]; w = z[nd]; w2 = w * w
        lam = np.zeros(k); delta = np.zeros((k, k))
        for j in range(k):
            o, mu, dlt = secular_root(j, dl, w2, rho)
            lam[j] = dl[o] + mu
            delta[j] = dlt
        # Loewner weights (Gu-Eisenstat): what[i]^2 = prod_j (lam_j - dl_i) / prod_{j != i} (dl_j - dl_i)
        what = np.zeros(k)
        for i in range(k):
            pr = delta[i, i]
            for jj in range(k):
                if jj != i:
                    pr *= delta[jj, i] / (dl[i] - dl[jj])
            what[i] = np.copysign(np.sqrt(-pr), w[i])
        for j in range(k):
            u = what / delta[j]
            u /= np.linalg.norm(u)
            ut[j, nd] = u
        lam_all[:k] = lam
    dn = [i for i in range(n) if defl[i]]
    for m, i in enumerate(dn):
        ut[k + m, i] = 1.0
        lam_all[k + m] = d[i]
    perm = np.argsort(lam_all, kind="stable")
    ut = ut[perm]; lam_all = lam_all[perm]
    # deflation-oblivious merge GEMM at full size. Q is block diagonal unless a deflating rotation paired a column of Q1 with one
    # of Q2 (the kernel halves the k-range when a device flag says no such rotation happened)
    qn = q @ ut.T
    return lam_all, qn


def leaf_eig(d, e):
    t = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    w, v = np.linalg.eigh(t)
    return w, v


def dc_eig(d, e, leaf=32, stats=None):
    stats = {} if stats is None else stats
    n = len(d)
    if n <= leaf:
        return leaf_eig(d, e)
    m = (n // 2 + 1) // 2 * 2                   # even split point (16-byte aligned sub-blocks in the GEMMs)
    beta = e[m - 1]
    d1 = d[:m].copy(); d2 = d[m:].copy()
    d1[-1] -= abs(beta); d2[0] -= abs(beta)
    w1, q1 = dc_eig(d1, e[:m - 1], leaf, stats)
    w2, q2 = dc_eig(d2, e[m:], leaf, stats)
    return merge(w1, q1, w2, q2, beta, stats)


# ------------------------------------------------------------------------------------------- 3. back-transform
def back_wy(zt, vr, tau, nb=64):
    """rows of zt = eigenvectors of T; returns rows = eigenvectors of A: X^T = Z^T H_{n-3} ... H_0."""
    n = zt.shape[0]
    xt = zt.copy()
    nref = n - 1
    panels = [(i0, min(i0 + nb, nref)) for i0 in range(0, nref, nb)]
    for i0, i1 in reversed(panels):
        v = vr[i0:i1].T                         # n x nb, column p = v_{i0+p}
        g = v.T @ v
        kb = i1 - i0
        t = np.zeros((kb, kb))
        for p in range(kb):                     # dlarft forward columnwise
            t[p, p] = tau[i0 + p]
            if p:
                t[:p, p] = -tau[i0 + p] * (t[:p, :p] @ g[:p, p])
        w = xt @ v
        w2 = w @ t.T
        xt -= w2 @ v.T
    return xt


def eig_tridiag_dc(a, leaf=32, nb=64, stats=None):
    d, e, tau, vr = sytrd_fused(a)
    lam, z = dc_eig(d, e, leaf, stats)
    xt = back_wy(z.T.copy(), vr, tau, nb)
    return lam, xt.T, (d, e)


def check(name, a, **kw):
    n = a.shape[0]
    stats = {}
    lam, x, (d, e) = eig_tridiag_dc(a, stats=stats, **kw)
    t = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    ref = np.linalg.eigvalsh(a)
    sc = max(np.abs(a).max(), 1e-300)
    print("%-28s n=%4d  tri-eig %.1e  residual %.1e  orth %.1e  eig %.1e  deflated %d" % (
        name, n, np.abs(np.linalg.eigvalsh(t) - ref).max() / sc, np.abs(a @ x - x * lam).max() / sc,
        np.abs(x.T @ x - np.eye(n)).max(), np.abs(lam - ref).max() / sc, stats.get("deflated", 0)), flush=True)
    assert np.abs(a @ x - x * lam).max() <= 1e-12 * sc and np.abs(x.T @ x - np.eye(n)).max() <= 1e-12


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (3, 10, 33, 64, 100, 257):
        q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        lam = np.sort(10.0 ** rng.uniform(0, 3, n))
        a = (q * lam) @ q.T; a = 0.5 * (a + a.T)
        check("spread 1..1e3", a, leaf=8)
        ee = rng.standard_normal((n, n))
        check("clustered I+1e-6 E", np.eye(n) + 1e-6 * 0.5 * (ee + ee.T), leaf=8)
        check("cma-like I+0.03 P", np.eye(n) * 0.97 + 0.03 * (ee @ ee.T) / n, leaf=8)
        check("identity", np.eye(n) * 2.0, leaf=8)
        lam = np.sort(10.0 ** rng.uniform(-6, 0, n))
        a = (q * lam) @ q.T; a = 0.5 * (a + a.T)
        check("cond 1e6", a, leaf=8)
        lam = np.repeat(np.arange(1, n // 4 + 2), 4)[:n].astype(float)
        a = (q * lam) @ q.T; a = 0.5 * (a + a.T)
        check("4-fold multiple", a, leaf=8)
        check("wilkinson-like tri", np.diag(np.abs(np.arange(n) - n // 2).astype(float)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1), leaf=4)
        check("diag + tiny", np.diag(np.arange(1.0, n + 1)) + 1e-14 * 0.5 * (ee + ee.T), leaf=4)
