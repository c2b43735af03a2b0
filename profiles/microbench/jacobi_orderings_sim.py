"""CPU simulation (numpy) of the blocked one-sided Jacobi used by korali_b200/csrc/eigen.cu: how many sweeps do different
block sizes / orderings / inner strategies need on the kind of matrix config 3 produces early on (clustered spectrum,
cold or warm start)? Evidence for DESIGN.md section 10 - not part of the product.

    python profiles/microbench/jacobi_orderings_sim.py [N]
"""
import sys
import numpy as np


def rr_pairs(nb, step):
    m = nb - 1
    out = []
    for k in range(nb // 2):
        a, b = (m, step) if k == 0 else ((step + k) % m, (step - k + m) % m)
        out.append((min(a, b), max(a, b)))
    return out


def ring_schedule(blocks):
    """Tournament in which one block of every pair STAYS with its CTA for long runs of steps.
    blocks: list whose length is a power of two. First half = anchors A, second half = movers B: in step s anchor A[i] meets
    B[(i + s) mod n/2] (the movers travel round a ring, one neighbour hop per step, the anchors never move: n/2 steps); then the
    tournaments inside A and inside B run side by side with the same construction (anchors of A keep their CTA again).
    n - 1 steps of n/2 disjoint pairs, every pair exactly once; an anchor changes only log2(n) times per sweep."""
    n = len(blocks)
    if n == 2:
        return [[(blocks[0], blocks[1])]]
    h = n // 2
    A, B = blocks[:h], blocks[h:]
    steps = [[(A[i], B[(i + s) % h]) for i in range(h)] for s in range(h)]
    for x, y in zip(ring_schedule(A), ring_schedule(B)):
        steps.append(x + y)
    return steps


def xor_schedule(nb, order="descending"):
    """Step s pairs block i with block i XOR s (nb = power of two): a 1-factorisation for s = 1 .. nb-1 in any order.
    descending: s = nb-1 .. 1 (partners differing in the top bit first, neighbours i ^ 1 last); ascending: the reverse;
    gray: s runs through the top-bit-first levels like the ring order, inside a level in Gray-code order."""
    ss = list(range(nb - 1, 0, -1)) if order == "descending" else list(range(1, nb))
    if order == "gray":
        ss, h = [], nb // 2
        while h >= 1:
            ss += [h + (j ^ (j >> 1)) for j in range(h)]
            h //= 2
    return [[(i, i ^ s) for i in range(nb) if i < (i ^ s)] for s in ss]


def oddeven_schedule(nb):
    """Odd-even transposition: the blocks sit on a line; even steps pair positions (0,1)(2,3).., odd steps (1,2)(3,4).. and
    the two blocks of a pair swap places after meeting. nb steps; a step of the odd phase leaves the two end blocks idle."""
    pos = list(range(nb))
    steps = []
    for s in range(nb):
        st = []
        for a in range(s % 2, nb - 1, 2):
            st.append((pos[a], pos[a + 1]))
            pos[a], pos[a + 1] = pos[a + 1], pos[a]
        steps.append(st)
    return steps


def check_schedule(steps, nb):
    seen = set()
    for st in steps:
        used = [b for pr in st for b in pr]
        assert len(used) == nb and len(set(used)) == nb            # every block exactly once per step
        for a, b in st:
            key = (min(a, b), max(a, b))
            assert key not in seen
            seen.add(key)
    assert len(seen) == nb * (nb - 1) // 2 and len(steps) == nb - 1    # every pair exactly once per sweep
    stay = sum(1 for s in range(1, len(steps)) for k in range(nb // 2) if steps[s][k][0] == steps[s - 1][k][0])
    return stay / ((len(steps) - 1) * (nb // 2))


INNER_SUBBLOCK = False   # set by callers: order of the cross pairs inside a step (see sweep)


def rotate_pairs(gam, R, pairs, tol2, big_thr):
    """gam: [P, r, r] Gram matrices, R: [P, r, r]; pairs: list of disjoint (p, q). One round, vectorised over P."""
    rot = 0
    big = False
    for p, q in pairs:
        a, b, g = gam[:, p, p], gam[:, q, q], gam[:, p, q]
        on = g * g > tol2 * a * b
        big |= bool(np.any(g * g > big_thr * a * b))
        rot += int(on.sum())
        d = b - a
        g2 = 2 * g
        h = np.sqrt(d * d + g2 * g2)
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.where(on, np.sign(np.where(d == 0, 1.0, d)) * g2 / (np.abs(d) + h), 0.0)
        t = np.nan_to_num(t)
        c = 1.0 / np.sqrt(1 + t * t)
        s = c * t
        c, s = c[:, None], s[:, None]
        for M in (gam, R):                      # rows
            x, y = M[:, p, :].copy(), M[:, q, :].copy()
            M[:, p, :] = c * x - s * y
            M[:, q, :] = s * x + c * y
        x, y = gam[:, :, p].copy(), gam[:, :, q].copy()      # columns of gamma
        gam[:, :, p] = c * x - s * y
        gam[:, :, q] = s * x + c * y
    return rot, big


def sweep(G, b, tol2, big_thr, inner_passes=1, sort_rows=False, ordering="round-robin"):
    n = G.shape[0]
    nb = n // b
    rot_total, big_any = 0, False
    ring = ring_schedule(list(range(nb))) if ordering.startswith("ring") else None
    if ordering.startswith("xor-"):
        ring = xor_schedule(nb, ordering[4:])
    if ordering == "odd-even":
        ring = oddeven_schedule(nb)
    if ordering == "ring-reversed":      # innermost sub-tournaments first, the A x B phase of the top level last
        ring = ring[::-1]
    if ordering.startswith("ring+local"):   # "ring+local<m>x<r>": after the sweep, r more passes of the sub-tournaments inside
        m, r = (int(x) for x in ordering[len("ring+local"):].split("x"))   # groups of m consecutive blocks (the last m - 1 steps)
        ring = ring + ring[-(m - 1):] * r
    for step in range(len(ring) if ring else nb - 1):
        prs = [(min(a, c), max(a, c)) for a, c in ring[step]] if ring else rr_pairs(nb, step)
        if step == 0:
            first_prs = prs
        idx = np.array([list(range(I * b, I * b + b)) + list(range(J * b, J * b + b)) for I, J in prs])   # [P, 2b]
        rows = G[idx]                                                                                     # [P, 2b, n]
        gam = np.einsum("pik,pjk->pij", rows, rows)
        R = np.tile(np.eye(2 * b), (len(prs), 1, 1))
        for _ in range(inner_passes):
            if step == 0 or (ring and step >= nb - 1 and prs == first_prs):    # all pairs of the 2b rows
                for r in range(2 * b - 1):
                    ro, bg = rotate_pairs(gam, R, rr_pairs(2 * b, r), tol2, big_thr)
                    rot_total += ro; big_any |= bg
            elif INNER_SUBBLOCK and b % 2 == 0:   # cross pairs as two local phases of half-block problems: (A0,B0) | (A1,B1), then (A0,B1) | (A1,B0)
                hb = b // 2
                for ph in range(2):
                    for r in range(hb):
                        prs_in = [(k, b + ph * hb + (k + r) % hb) for k in range(hb)] + [(hb + k, b + (1 - ph) * hb + (k + r) % hb) for k in range(hb)]
                        ro, bg = rotate_pairs(gam, R, prs_in, tol2, big_thr)
                        rot_total += ro; big_any |= bg
            else:            # cross pairs only
                for r in range(b):
                    ro, bg = rotate_pairs(gam, R, [(k, b + (k + r) % b) for k in range(b)], tol2, big_thr)
                    rot_total += ro; big_any |= bg
        G[idx] = np.einsum("pij,pjk->pik", R, rows)
    if sort_rows:   # de Rijk: keep the rows ordered by decreasing norm
        G[:] = G[np.argsort(-np.einsum("ij,ij->i", G, G), kind="stable")]
    return rot_total, big_any


def solve(C, V0, b, inner_passes=1, sort_rows=False, big_thr=1e-16, max_sweeps=40, ordering="round-robin"):
    n = C.shape[0]
    G = (C @ V0).T.copy()
    tol2 = (4 * 2.2e-16 * np.sqrt(n)) ** 2
    for s in range(1, max_sweeps + 1):
        o = ordering
        if ordering == "ring-alternating":   # forward on odd sweeps, reversed on even ones
            o = "ring" if s % 2 else "ring-reversed"
        rot, big = sweep(G, b, tol2, big_thr, inner_passes, sort_rows, o)
        if rot == 0 or not big:
            break
    lam = np.sqrt(np.einsum("ij,ij->i", G, G))
    off = np.abs((G / lam[:, None]) @ (G / lam[:, None]).T - np.eye(n)).max()
    return s, off


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rng = np.random.default_rng(0)
    # config-3-like early covariance: C = (1 - cmu) I + cmu * sample covariance of mu_eff-ish samples, a few generations deep
    def gen_c(c, cmu=0.035, m=4 * n):
        z = rng.standard_normal((m, n)) @ np.linalg.cholesky(c).T
        return (1 - cmu) * c + cmu * (z.T @ z) / m
    c = np.eye(n)
    for _ in range(4):
        c_prev, c = c, gen_c(c)
    w_prev, v_prev = np.linalg.eigh(c_prev)
    print("N = %d, spectrum spread of C: %.3g .. %.3g" % (n, *np.linalg.eigvalsh(c)[[0, -1]]))
    for nb in (8, 64, 256):
        print("  ring schedule, %3d blocks: valid 1-factorisation, the first block of a CTA's pair stays in %.1f %% of the step transitions"
              % (nb, 100 * check_schedule(ring_schedule(list(range(nb))), nb)))
    for name, kw in [("4-row blocks, one cross pass (the kernels)", dict(b=4)),
                     ("4-row blocks, ring schedule (one block of each pair stays put)", dict(b=4, ordering="ring")),
                     ("4-row blocks, two cross passes per step", dict(b=4, inner_passes=2)),
                     ("8-row blocks, one cross pass", dict(b=8)),
                     ("16-row blocks, one cross pass (large-N kernel)", dict(b=16)),
                     ("4-row blocks + rows re-sorted by norm after each sweep", dict(b=4, sort_rows=True)),
                     ("4-row blocks, last-sweep criterion 1e-20 (old)", dict(b=4, big_thr=1e-20)),
                     ("4-row blocks, last-sweep criterion 1e-14", dict(b=4, big_thr=1e-14))]:
        cold = solve(c, np.eye(n), **kw)
        warm = solve(c, v_prev, **kw)
        print("  %-58s cold start: %2d sweeps (max |cos| left %.1e)   warm start: %2d sweeps (%.1e)" % (name, cold[0], cold[1], warm[0], warm[1]))
