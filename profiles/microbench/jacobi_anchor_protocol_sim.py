"""CPU emulation of the DATAFLOW PROTOCOL of the experimental resident-anchor variant of jacobi_pipe_kernel
(ORDER = 2, KCMA_JACOBI_ORDER=anchor; eigen.cu). Not a numerical simulation: blocks carry opaque values, a "rotation" of
the pair (I, J) in step s replaces both values by a hash of (old I, old J, s), or leaves them alone when the step rotates
nothing (rot == 0). What is checked, under random interleavings of the CTAs and of their G and V groups:

  * the protocol never deadlocks (flags per block, epoch counting, the kept block's flag not published while it is resident);
  * every value a CTA combines is the CURRENT value of that block (no stale copy in shared memory or in global memory);
  * after every sweep global memory holds the current value of every block (the release / write-back rules);
  * how many block loads / stores / flag waits the variant saves.

The state machine below mirrors the kernel line by line: keep_prev / keep_next from ring_pair, g_dirty, v_valid, v_dirty.

    python profiles/microbench/jacobi_anchor_protocol_sim.py [nb] [sweeps] [seed]
"""
import random
import sys


def ring_pair(np_, step, k):
    base = 0
    while True:
        h = np_ >> 1
        if step < h:
            return base + k, base + h + ((k + step) & (h - 1))
        step -= h
        if k >= (h >> 1):
            base += h
            k -= h >> 1
        np_ = h


def mix(a, b, s, which):
    return hash((a, b, s, which)) & 0xFFFFFFFFFFFF


class Sim:
    def __init__(self, nb, seed, p_rot0):
        self.nb, self.rng, self.p_rot0 = nb, random.Random(seed), p_rot0
        self.G = [("g", i) for i in range(nb)]          # global memory, per block
        self.V = [("v", i) for i in range(nb)]
        self.readyG = [0] * nb
        self.readyV = [0] * nb
        self.truthG = list(self.G)                      # what a plain (non-resident) execution would hold
        self.truthV = list(self.V)
        self.stats = dict(g_loads=0, g_stores=0, v_loads=0, v_stores=0, flag_waits=0, steps=0)
        self.rot_of = {}                                # (sweep, step, cta) -> rot (decided once, shared by the G and V group)

    def rot(self, sweep, step, k):
        key = (sweep, step, k)
        if key not in self.rot_of:
            self.rot_of[key] = 0 if self.rng.random() < self.p_rot0 else 1
        return self.rot_of[key]

    def g_group(self, k, sweeps, g_head):
        nb = self.nb
        epoch, smem, dirty = 0, None, False             # smem: value of Gs rows 0-3
        for sweep in range(sweeps):
            for step in range(nb - 1):
                I, J = ring_pair(nb, step, k)
                keep_prev = step > 0 and ring_pair(nb, step - 1, k)[0] == I
                keep_next = step + 2 < nb and ring_pair(nb, step + 1, k)[0] == I
                if not keep_prev:
                    dirty = False
                for blk in ([J] if keep_prev else [I, J]):          # flag wait
                    self.stats["flag_waits"] += 1
                    while self.readyG[blk] < epoch:
                        yield
                if keep_prev:
                    a = smem                                        # Gram operands of rows 0-3 from shared memory
                else:
                    a = self.G[I]; smem = a; self.stats["g_loads"] += 1
                b = self.G[J]; self.stats["g_loads"] += 1
                assert a == self.truthG[I] and b == self.truthG[J], ("stale G", sweep, step, k)
                rot = self.rot(sweep, step, k)
                yield
                if rot:
                    na, nb_ = mix(a, b, step, 0), mix(a, b, step, 1)
                    self.truthG[I], self.truthG[J] = na, nb_
                    if keep_next:
                        smem = na                                   # in place in shared memory
                    else:
                        self.G[I] = na; self.stats["g_stores"] += 1
                    self.G[J] = nb_; self.stats["g_stores"] += 1
                    dirty = keep_next
                elif (not keep_next) and dirty:
                    self.G[I] = smem; self.stats["g_stores"] += 1  # write-back of a kept block released in a rotation-free step
                yield
                if not keep_next:
                    self.readyG[I] = epoch + 1
                self.readyG[J] = epoch + 1
                g_head[k] = epoch + 1
                epoch += 1
                self.stats["steps"] += 1
            yield ("sweep", sweep)

    def v_group(self, k, sweeps, g_head):
        nb = self.nb
        epoch, smem, valid, dirty = 0, None, False, False
        for sweep in range(sweeps):
            for step in range(nb - 1):
                I, J = ring_pair(nb, step, k)
                keep_prev = step > 0 and ring_pair(nb, step - 1, k)[0] == I
                keep_next = step + 2 < nb and ring_pair(nb, step + 1, k)[0] == I
                if not keep_prev:
                    valid = dirty = False
                while g_head[k] <= epoch:                           # R of this step posted by the G group
                    yield
                rot = self.rot(sweep, step, k)
                for blk in ([J] if keep_prev else [I, J]):
                    while self.readyV[blk] < epoch:
                        yield
                if rot:
                    if valid:
                        a = smem
                    else:
                        a = self.V[I]; self.stats["v_loads"] += 1
                    b = self.V[J]; self.stats["v_loads"] += 1
                    assert a == self.truthV[I] and b == self.truthV[J], ("stale V", sweep, step, k)
                    yield
                    na, nb_ = mix(a, b, step, 2), mix(a, b, step, 3)
                    self.truthV[I], self.truthV[J] = na, nb_
                    if keep_next:
                        smem = na
                    else:
                        self.V[I] = na; self.stats["v_stores"] += 1
                    self.V[J] = nb_; self.stats["v_stores"] += 1
                    valid = dirty = keep_next
                elif (not keep_next) and dirty:
                    self.V[I] = smem; self.stats["v_stores"] += 1
                yield
                if not keep_next:
                    self.readyV[I] = epoch + 1
                self.readyV[J] = epoch + 1
                epoch += 1
            yield ("sweep", sweep)


def run(nb, sweeps, seed, p_rot0):
    sim = Sim(nb, seed, p_rot0)
    ncta = nb // 2
    g_head = [0] * ncta                                             # steps whose R the G group of a CTA has posted
    live = []
    for k in range(ncta):
        live.append([sim.g_group(k, sweeps, g_head), None])
        live.append([sim.v_group(k, sweeps, g_head), None])
    spins = 0
    while live:
        waiting = [it for it in live if it[1] is not None]
        if len(waiting) == len(live):                               # grid.sync between sweeps: everyone has arrived
            assert len({it[1] for it in live}) == 1
            assert sim.G == sim.truthG and sim.V == sim.truthV, ("global memory not current after sweep", live[0][1])
            for it in live:
                it[1] = None
        it = sim.rng.choice(live)
        if it[1] is not None:
            continue
        before = (sum(sim.readyG), sum(sim.readyV), sum(g_head))
        try:
            ev = next(it[0])
        except StopIteration:
            live.remove(it)
            continue
        if ev is not None:
            it[1] = ev[1]
        spins = spins + 1 if (sum(sim.readyG), sum(sim.readyV), sum(g_head)) == before and ev is None else 0
        assert spins < 20000 * len(live) + 100000, "deadlock"
    return sim


if __name__ == "__main__":
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    for p_rot0 in (0.0, 0.3, 0.9, 1.0):
        sim = run(nb, sweeps, seed, p_rot0)
        st = sim.stats
        pair_steps = sweeps * (nb - 1) * (nb // 2)
        print("nb=%d sweeps=%d P(rot=0)=%.1f: ok, no deadlock, no stale block; per pair-step: G loads %.2f (2.00 without), G stores %.2f, "
              "V loads %.2f, V stores %.2f, G flag waits %.2f (2.00)" % (nb, sweeps, p_rot0, st["g_loads"] / pair_steps, st["g_stores"] / pair_steps,
              st["v_loads"] / pair_steps, st["v_stores"] / pair_steps, st["flag_waits"] / pair_steps))
