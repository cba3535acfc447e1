"""Protocol-level simulation of the mbarrier pipelines that have not run on a B200 yet: the tcgen05 attention backward
(csrc/attention_bwd_tc5.cu: dQ and dK/dV kernels) and the cluster-fused decode GEMM with the in-kernel RMSNorm operand
(csrc/gemm_decode_fused.cu, NORM_IN).  Each warp role is transcribed as a coroutine that performs the kernel's waits, arrivals,
TMA issues, tcgen05.mma issues / commits and shared-memory / TMEM reads and writes in the kernel's order; a random scheduler
interleaves the roles and completes the asynchronous operations (TMA copies, MMAs in issue order) at random later points.

What it checks, over many random schedules:
  * liveness -- every role terminates (a lost arrival or a wrong phase parity shows up as a deadlock);
  * every MMA reads the operand tiles it is meant to read, at issue time AND at completion time (a tile refilled too early, or
    not yet filled, is a label mismatch);
  * every TMEM / shared buffer read by the softmax warps holds the tile it is meant to hold.

mbarrier semantics modelled: arrival count, expect_tx / complete_tx byte accounting, phase parity with try_wait(parity)
(a fresh barrier passes a wait on parity 1).  tcgen05.commit arrives when all previously issued MMAs have completed.
TEST INFRASTRUCTURE ONLY -- it pins the synchronisation design, not the data path.
"""
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.done = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.done += 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier was initialised for"
        self.pending -= 1
        self._check()

    def expect_tx(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        assert self.tx >= 0
        self._check()

    def ready(self, parity):
        return (self.done & 1) != parity


class Sim:
    def __init__(self, seed):
        self.rng = random.Random(seed)
        self.agents, self.async_ops, self.mma_queue = [], [], []
        self.mem = {}                                    # buffer name -> label of its current content

    def spawn(self, name, gen):
        self.agents.append([name, gen, None])            # [name, coroutine, pending wait predicate]

    # asynchronous engines ---------------------------------------------------------------------------------------------
    def tma(self, buf, label, bar, nbytes):
        self.async_ops.append(lambda: (self.mem.__setitem__(buf, label), bar.complete_tx(nbytes)))

    def mma(self, reads, write=None, chain=()):
        """reads: [(buffer, expected label)], write: (buffer, label).  Shared-memory operands are checked now and again at
        completion; ``chain`` = the accumulator this MMA adds to (TMEM): the tensor core executes MMAs in issue order, so that
        dependency is checked at completion only."""
        for b, lab in reads:
            assert self.mem.get(b) == lab, f"MMA issued while {b} holds {self.mem.get(b)}, wanted {lab}"
        self.mma_queue.append(("mma", list(reads) + list(chain), write))

    def commit(self, bar):
        self.mma_queue.append(("commit", bar, None))

    def _step_mma(self):
        kind, a, b = self.mma_queue.pop(0)
        if kind == "commit":
            a.arrive()
            return
        for buf, lab in a:
            assert self.mem.get(buf) == lab, f"MMA completed after {buf} was refilled with {self.mem.get(buf)}, wanted {lab}"
        if b is not None:
            self.mem[b[0]] = b[1]

    def run(self, max_steps=200000):
        for _ in range(max_steps):
            choices = []
            for ag in self.agents:
                if ag[2] is None or ag[2]():
                    choices.append(("agent", ag))
            if self.async_ops:
                choices.append(("async", None))
            if self.mma_queue:
                choices.append(("mma", None))
            if not choices:
                if all(ag[1] is None for ag in self.agents):
                    return
                blocked = [ag[0] for ag in self.agents if ag[1] is not None]
                raise AssertionError(f"deadlock: {blocked} blocked with no asynchronous work outstanding")
            kind, ag = self.rng.choice(choices)
            if kind == "async":
                self.async_ops.pop(self.rng.randrange(len(self.async_ops)))()
            elif kind == "mma":
                self._step_mma()
            else:
                try:
                    ag[2] = next(ag[1])                  # the coroutine yields its next wait predicate (or None to just yield)
                except StopIteration:
                    ag[1], ag[2] = None, (lambda: False)
            self.agents = [a for a in self.agents if a[1] is not None] + [a for a in self.agents if a[1] is None]
            if all(a[1] is None for a in self.agents) and not self.async_ops and not self.mma_queue:
                return
        raise AssertionError("simulation did not finish")


def wait(bar, parity):
    return lambda: bar.ready(parity)


# ====================================================================================================================== dQ kernel
def sim_dq(nt, seed):
    s = Sim(seed)
    qdo, kv_full, kv_empty = MBar(1), [MBar(1), MBar(1)], [MBar(1), MBar(1)]            # K/V ring of two stages
    s_full, s_free, ds_full, ds_free, o_done = MBar(1), MBar(4), MBar(4), MBar(1), MBar(1)
    s.mem.update({"Q": None, "dO": None, "K0": None, "V0": None, "K1": None, "V1": None, "dS": None, "S": None, "dP": None,
                  "dQ": ("dQ", -1)})

    def producer():
        qdo.expect_tx(4)
        for b in ("Q", "Q", "dO", "dO"):
            s.tma(b, b, qdo, 1)
        for j in range(nt):
            st = j & 1
            yield wait(kv_empty[st], ((j >> 1) & 1) ^ 1)
            kv_full[st].expect_tx(4)
            s.tma(f"K{st}", ("K", j), kv_full[st], 2)
            s.tma(f"V{st}", ("V", j), kv_full[st], 2)
            yield None

    def mma():
        yield wait(qdo, 0)
        for j in range(nt):
            st = j & 1
            yield wait(kv_full[st], (j >> 1) & 1)
            if j > 0:
                yield wait(s_free, (j - 1) & 1)
            s.mma([("Q", "Q"), (f"K{st}", ("K", j))], ("S", ("S", j)))
            s.mma([("dO", "dO"), (f"V{st}", ("V", j))], ("dP", ("dP", j)))
            s.commit(s_full)
            yield wait(ds_full, j & 1)
            s.mma([("dS", ("dS", j)), (f"K{st}", ("K", j))], ("dQ", ("dQ", j)), chain=[("dQ", ("dQ", j - 1))])
            s.commit(kv_empty[st])
            s.commit(ds_free)
            yield None
        s.commit(o_done)

    def softmax(w):
        def g():
            for j in range(nt):
                yield wait(s_full, j & 1)
                assert s.mem["S"] == ("S", j) and s.mem["dP"] == ("dP", j), (s.mem["S"], s.mem["dP"], j)
                yield None
                s_free.arrive()
                if j > 0:
                    yield wait(ds_free, (j - 1) & 1)
                s.mem["dS"] = ("dS", j) if w == 3 else s.mem["dS"]      # the tile is complete once the last warp has written
                yield None
                ds_full.arrive()
            yield wait(o_done, 0)
            assert s.mem["dQ"] == ("dQ", nt - 1)
        return g()

    s.spawn("tma", producer())
    s.spawn("mma", mma())
    for w in range(4):
        s.spawn(f"softmax{w}", softmax(w))
    s.run()


# ====================================================================================================================== dK/dV kernel
def sim_dkv(n_iter, seed):
    s = Sim(seed)
    kv_bar, qd_full, qd_empty = MBar(1), [MBar(1), MBar(1)], [MBar(1), MBar(1)]          # Q/dO ring of two stages
    s_full, s_free, pd_full, pd_free, done = MBar(1), MBar(4), MBar(4), MBar(1), MBar(1)
    s.mem.update({"K": None, "V": None, "Q0": None, "dO0": None, "Q1": None, "dO1": None, "PT": None, "dST": None, "ST": None, "dPT": None,
                  "dV": ("dV", -1), "dK": ("dK", -1)})

    def producer():
        kv_bar.expect_tx(2)
        s.tma("K", "K", kv_bar, 1)
        s.tma("V", "V", kv_bar, 1)
        for n in range(n_iter):
            st = n & 1
            yield wait(qd_empty[st], ((n >> 1) & 1) ^ 1)
            qd_full[st].expect_tx(2)
            s.tma(f"Q{st}", ("Q", n), qd_full[st], 1)
            s.tma(f"dO{st}", ("dO", n), qd_full[st], 1)
            yield None

    def mma():
        yield wait(kv_bar, 0)
        for n in range(n_iter):
            st = n & 1
            yield wait(qd_full[st], (n >> 1) & 1)
            if n > 0:
                yield wait(s_free, (n - 1) & 1)
            s.mma([("K", "K"), (f"Q{st}", ("Q", n))], ("ST", ("ST", n)))
            s.mma([("V", "V"), (f"dO{st}", ("dO", n))], ("dPT", ("dPT", n)))
            s.commit(s_full)
            yield wait(pd_full, n & 1)
            s.mma([("PT", ("PT", n)), (f"dO{st}", ("dO", n))], ("dV", ("dV", n)), chain=[("dV", ("dV", n - 1))])
            s.mma([("dST", ("dST", n)), (f"Q{st}", ("Q", n))], ("dK", ("dK", n)), chain=[("dK", ("dK", n - 1))])
            s.commit(qd_empty[st])
            s.commit(pd_free)
            yield None
        s.commit(done)

    def softmax(w):
        def g():
            for n in range(n_iter):
                yield wait(s_full, n & 1)
                assert s.mem["ST"] == ("ST", n) and s.mem["dPT"] == ("dPT", n)
                yield None
                s_free.arrive()
                if n > 0:
                    yield wait(pd_free, (n - 1) & 1)
                if w == 3:
                    s.mem["PT"], s.mem["dST"] = ("PT", n), ("dST", n)
                yield None
                pd_full.arrive()
            yield wait(done, 0)
            assert s.mem["dV"] == ("dV", n_iter - 1) and s.mem["dK"] == ("dK", n_iter - 1)
        return g()

    s.spawn("tma", producer())
    s.spawn("mma", mma())
    for w in range(4):
        s.spawn(f"softmax{w}", softmax(w))
    s.run()


# ====================================================================================================================== fused decode GEMM, NORM_IN
def sim_fused_norm_in(nkb, stages, seed):
    s = Sim(seed)
    full = [MBar(5) for _ in range(stages)]               # weight tile (expect_tx arrival) + four B-producer warps
    empty = [MBar(1) for _ in range(stages)]
    acc = MBar(1)
    for st in range(stages):
        s.mem[f"A{st}"], s.mem[f"B{st}"] = None, None
    s.mem["acc"] = ("acc", -1)

    def producer():
        npre = min(nkb, stages)
        for i in range(npre):
            full[i].expect_tx(1)
            s.tma(f"A{i}", ("A", i), full[i], 1)
        yield None                                        # pdl_wait
        for i in range(npre, nkb):
            st = i % stages
            yield wait(empty[st], ((i // stages) & 1) ^ 1)
            full[st].expect_tx(1)
            s.tma(f"A{st}", ("A", i), full[st], 1)
            yield None

    def mma():
        for i in range(nkb):
            st = i % stages
            yield wait(full[st], (i // stages) & 1)
            s.mma([(f"A{st}", ("A", i)), (f"B{st}", ("B", i))], ("acc", ("acc", i)), chain=[("acc", ("acc", i - 1))])
            s.commit(empty[st])
            yield None
        s.commit(acc)

    def bwarp(w):
        def g():
            yield None                                    # pdl_wait, rstd staging
            for i in range(nkb):
                st = i % stages
                yield wait(empty[st], ((i // stages) & 1) ^ 1)
                if w == 3:
                    s.mem[f"B{st}"] = ("B", i)
                yield None
                full[st].arrive()
            yield wait(acc, 0)
            assert s.mem["acc"] == ("acc", nkb - 1)
        return g()

    s.spawn("tma", producer())
    s.spawn("mma", mma())
    for w in range(4):
        s.spawn(f"bwarp{w}", bwarp(w))
    s.run()


@pytest.mark.parametrize("nt", [1, 2, 3, 6])
def test_dq_pipeline(nt):
    for seed in range(60):
        sim_dq(nt, seed)


@pytest.mark.parametrize("n_iter", [1, 2, 5, 8])
def test_dkv_pipeline(n_iter):
    for seed in range(60):
        sim_dkv(n_iter, seed)


@pytest.mark.parametrize("nkb,stages", [(1, 3), (3, 3), (7, 3), (12, 2), (11, 4)])
def test_fused_decode_norm_in_pipeline(nkb, stages):
    for seed in range(60):
        sim_fused_norm_in(nkb, stages, seed)


def test_the_simulator_catches_a_missing_wait():
    """Self-check: a TMA producer that does NOT wait for the K/V slot to be released (kv_empty) refills K under MMAs that still
    read it; the simulation must report that on some schedule.  (In this one-tile-in-flight design several other waits -- s_free,
    ds_free -- are implied transitively by kv_empty; they are kept in the kernels as the pipeline is meant to be deepened.)"""
    def broken(seed):
        s = Sim(seed)
        kv_full, s_full, ds_full = MBar(1), MBar(1), MBar(4)
        s.mem.update({"K": None, "dS": None, "S": None, "dQ": ("dQ", -1)})

        def producer():
            for j in range(4):                               # no kv_empty wait
                kv_full.expect_tx(1)
                s.tma("K", ("K", j), kv_full, 1)
                yield wait(kv_full, j & 1)                   # (only waits for its own copy to land)

        def mma():
            for j in range(4):
                yield wait(kv_full, j & 1)
                s.mma([("K", ("K", j))], ("S", ("S", j)))
                s.commit(s_full)
                yield wait(ds_full, j & 1)
                s.mma([("dS", ("dS", j)), ("K", ("K", j))], ("dQ", ("dQ", j)), chain=[("dQ", ("dQ", j - 1))])
                yield None

        def softmax(w):
            def g():
                for j in range(4):
                    yield wait(s_full, j & 1)
                    if w == 3:
                        s.mem["dS"] = ("dS", j)
                    yield None
                    ds_full.arrive()
            return g()

        s.spawn("tma", producer())
        s.spawn("mma", mma())
        for w in range(4):
            s.spawn(f"sm{w}", softmax(w))
        s.run()

    caught = 0
    for seed in range(200):
        try:
            broken(seed)
        except AssertionError:
            caught += 1
    assert caught > 0
