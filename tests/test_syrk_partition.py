"""Host-side statement of the stream-K partition of the rank-mu product (korali_b200/csrc/gemm_tma.cu: syrk_sk_tma_kernel and
launch_syrk_sk_tma): the weighted (tile, k-step) space of the lower-triangular tiles is cut into one contiguous span per CTA; part p of
a tile goes to slab p. The invariants the device code relies on, checked here for the shapes of every config: every k-step of every
tile is covered exactly once, the parts of a tile are numbered 0, 1, ... without gaps beyond the bound the launcher zeroes, no
(tile, part) slot is written twice, and the segment table of a CTA fits. (The arithmetic itself is checked on the GPU against the
oracle: tests/test_gpu_parity.py::test_rank_mu_matches_oracle, tests/test_gpu_fullsize.py.)"""
import pytest

TB, W_FULL, W_DIAG, MAX_SEG = 128, 8, 5, 64


def partition(n, k_rows, sms, max_splits=16, k_dev=None):
    """k_rows: the host's bound on the row count (sizes the grid); k_dev: the count the kernel reads from device memory (<= k_rows)."""
    nt = (n + TB - 1) // TB
    nk = (k_rows + 15) // 16
    sw = W_FULL * (nt * (nt - 1) // 2) + W_DIAG * nt
    total = nk * sw
    wmax = W_FULL if nt > 1 else W_DIAG
    g = min(sms, max(1, sw * nk // 64))
    g = min(g, max(1, (max_splits - 1) * sw // wmax))
    parts_bound = min(max_splits, (wmax * g + sw - 1) // sw + 1)
    if nt * (nt + 1) // 2 // g + 3 > MAX_SEG:
        return None                                   # the launcher falls back to the split-K kernel
    if k_dev is not None:                              # a rank only knows on the device how many selected samples it owns
        nk = (k_dev + 15) // 16
        total = nk * sw
    bound = lambda c: c * total // g
    cover, max_seg = {}, 0
    for c in range(g):
        b0, b1 = bound(c), bound(c + 1)
        if b1 <= b0:
            continue
        bi = bj = cum = 0
        w = lambda: W_DIAG if bi == bj else W_FULL
        while bi < nt and cum + w() * nk <= b0:
            cum += w() * nk
            bi, bj = (bi + 1, 0) if bj == bi else (bi, bj + 1)
        pos, ns = b0, 0
        while pos < b1 and bi < nt:
            end = cum + w() * nk
            k_lo = (pos - cum) // w()
            k_hi = nk if b1 >= end else (b1 - cum) // w()
            if k_hi > k_lo:
                cf = cum * g // total
                while bound(cf + 1) <= cum:
                    cf += 1
                while bound(cf) > cum:
                    cf -= 1
                part = c - cf
                assert 0 <= part < parts_bound, (n, k_rows, sms, part, parts_bound)
                assert (bi, bj, part) not in cover
                cover[(bi, bj, part)] = (k_lo, k_hi)
                ns += 1
            cum = end
            bi, bj = (bi + 1, 0) if bj == bi else (bi, bj + 1)
            pos = cum
        max_seg = max(max_seg, ns)
    for bi in range(nt):
        for bj in range(bi + 1):
            pos = 0
            for lo, hi in sorted(v for k, v in cover.items() if k[:2] == (bi, bj)):
                assert lo == pos, (n, k_rows, sms, bi, bj)
                pos = hi
            assert pos == nk, (n, k_rows, sms, bi, bj, pos, nk)
    assert max_seg <= MAX_SEG
    return g, parts_bound, max_seg


@pytest.mark.parametrize("n,k_rows", [(1000, 32768), (1000, 4096), (1000, 17), (100, 2048), (10, 16), (10, 5), (4096, 65536), (4096, 1 << 19),
                                      (257, 1000), (128, 33), (129, 64), (2000, 777), (64, 0), (1184, 8192)])
@pytest.mark.parametrize("sms", [148, 132, 7, 1])
def test_every_k_step_of_every_tile_exactly_once(n, k_rows, sms):
    partition(n, k_rows, sms)


def test_config3_balance():
    g, parts, segs = partition(1000, 32768, 148)
    assert (g, parts, segs) == (148, 6, 2)     # 148 CTAs, at most 6 parts per tile (the slabs the launcher zeroes), 2 segments per CTA


@pytest.mark.parametrize("n", [10, 100, 257, 1000, 4096])
def test_device_row_count_below_the_host_bound(n):
    """Multi-GPU: the grid is sized for the host's bound (mu, or the shard size), the partition is computed on the device from the
    number of selected samples the rank really owns."""
    for k_host in (100, 8192, 1 << 17):
        for k_dev in (0, 1, 17, k_host // 8, k_host // 2, k_host - 1):
            partition(n, k_host, 148, k_dev=k_dev)

