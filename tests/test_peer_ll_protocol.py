"""Protocol check of the low-latency tensor-parallel all-reduce (chatts_b200/csrc/allreduce_ll.cu) on CPU.

The kernel's correctness rests on a wire protocol, not on arithmetic: 16-byte units {d0, epoch, d1, epoch} whose 8-byte halves may
land separately, pollers that accept a unit only when both epoch words match, two buffer sets that alternate between consecutive
calls, and the claim that finishing call n+1 licenses overwriting the units of call n.  This file restates the kernel's three phases
with the SAME unit-index formulas and runs W ranks x (C x T) CTAs as coroutines under random schedules: every store half and every
poll attempt is a scheduling point, a rank's call n+1 starts only when all its CTAs of call n have finished (stream order), ranks
drift freely otherwise.  Checked: every rank ends every call with the h the rank-ordered fp32 sum defines (bit for bit), identical
h / norm_out on all ranks, no deadlock.  The mutations at the bottom show the simulation can fail: an epoch that does not advance, or a
poller that checks one epoch word, break it.

No GPU, no library call: the model dtype is emulated by float16 (numpy has no bfloat16; the rounding points are what matters)."""
import numpy as np
import pytest

KMAXC = 8


def rnd(x):
    return np.float32(np.float16(x))


class Rank:
    def __init__(self, W, Tmax, h):
        hw = h // W
        # per buffer set: units as rows [d0, f0, d1, f1] (float64 holds fp32 payloads and epochs exactly)
        self.rs = [np.zeros((W * Tmax * hw // 2, 4)) for _ in range(2)]
        self.ag = [np.zeros((W * Tmax * hw // 4, 4)) for _ in range(2)]
        self.sq = [np.zeros((W * Tmax * KMAXC, 4)) for _ in range(2)]
        self.epoch_state = 0


def st_ll(arr, unit, d0, d1, epoch, torn=True):
    """A vector store whose two 8-byte halves are delivered separately (any order is a schedule away: we yield in between)."""
    arr[unit, 0], arr[unit, 1] = d0, epoch
    if torn:
        yield
    arr[unit, 2], arr[unit, 3] = d1, epoch
    yield


def ll_ok(u, epoch, both=True):
    return (u[1] == epoch and u[3] == epoch) if both else (u[1] == epoch)


def cta(ranks, me, which, c, t, C, W, Tmax, h, S, local_part, resid, norm_w, eps, out_h, out_n, both=True):
    """One CTA (c, t) of rank `me`, transcribed from peer_allreduce_ll_kernel (same index arithmetic, thread loops flattened)."""
    hw, R = h // W, ranks[me]
    wc = hw // C
    epoch = R.epoch_state + 1
    # phase A
    upo = wc // 2
    for u in range(W * upo):
        j, k = u // upo, (u % upo) * 2
        col = j * hw + c * wc + k
        a0, a1 = np.float32(local_part[0, t, col]), np.float32(local_part[0, t, col + 1])
        for s in range(1, S):
            a0, a1 = np.float32(a0 + local_part[s, t, col]), np.float32(a1 + local_part[s, t, col + 1])
        unit = ((me * Tmax + t) * hw + c * wc + k) >> 1
        yield from st_ll(ranks[j].rs[which], unit, a0, a1, epoch)
    # phase B
    ss_parts = []
    for q in range(wc // 4):
        k = q * 4
        while True:
            us = []
            for r in range(W):
                unit = ((r * Tmax + t) * hw + c * wc + k) >> 1
                us.append((R.rs[which][unit].copy(), R.rs[which][unit + 1].copy()))
            if all(ll_ok(a, epoch, both) and ll_ok(b, epoch, both) for a, b in us):
                break
            yield "poll"
        acc = [np.float32(0)] * 4
        for r in range(W):
            a, b = us[r]
            acc = [np.float32(acc[0] + np.float32(a[0])), np.float32(acc[1] + np.float32(a[2])), np.float32(acc[2] + np.float32(b[0])),
                   np.float32(acc[3] + np.float32(b[2]))]
        col = me * hw + c * wc + k
        hv = [rnd(np.float32(resid[t, col + i]) + rnd(acc[i])) for i in range(4)]
        ss_parts.append(np.float32(sum(np.float32(x * x) for x in hv)))
        unit = ((me * Tmax + t) * hw + c * wc + k) >> 2
        for j in range(W):
            # payload words carry two model-dtype values each; the pair is kept as a tuple index into a side table
            yield from st_ll(ranks[j].ag[which], unit, _pack(hv[0], hv[1]), _pack(hv[2], hv[3]), epoch)
    ssq = np.float32(0)
    for x in ss_parts:
        ssq = np.float32(ssq + x)
    for j in range(W):
        yield from st_ll(ranks[j].sq[which], (me * Tmax + t) * KMAXC + c, ssq, 0.0, epoch)
    # phase C
    sq = []
    for i in range(W * C):
        o, cc = i // C, i % C
        while True:
            u = R.sq[which][(o * Tmax + t) * KMAXC + cc].copy()
            if ll_ok(u, epoch, both):
                break
            yield "poll"
        sq.append(np.float32(u[0]))
    tot = np.float32(0)
    for x in sq:
        tot = np.float32(tot + x)
    inv = np.float32(1.0) / np.sqrt(np.float32(tot / np.float32(h) + np.float32(eps)))
    upo = wc // 4
    for u in range(W * upo):
        o, k = u // upo, (u % upo) * 4
        unit = ((o * Tmax + t) * hw + c * wc + k) >> 2
        while True:
            v = R.ag[which][unit].copy()
            if ll_ok(v, epoch, both):
                break
            yield "poll"
        x = list(_unpack(v[0])) + list(_unpack(v[2]))
        off = o * hw + c * wc + k
        for i in range(4):
            out_h[t, off + i] = x[i]
            out_n[t, off + i] = rnd(np.float32(norm_w[off + i]) * rnd(np.float32(x[i]) * inv))
    yield "done"


def _pack(a, b):
    """Two float16 values in one 32-bit payload word (kept exact inside a float64 cell)."""
    return float(int(np.float16(a).view(np.uint16)) | (int(np.float16(b).view(np.uint16)) << 16))


def _unpack(w):
    w = int(w)
    return (np.float32(np.uint16(w & 0xFFFF).view(np.float16)), np.float32(np.uint16(w >> 16).view(np.float16)))


def expected(parts, resid, norm_w, eps):
    """parts [W][S][T][h] -> h_new, norm (rank-ordered fp32 sum; the statistic order is checked through cross-rank identity)."""
    W, S, T, h = parts.shape
    acc = np.zeros((T, h), np.float32)
    for r in range(W):
        loc = parts[r, 0].astype(np.float32)
        for s in range(1, S):
            loc = (loc + parts[r, s]).astype(np.float32)
        acc = (acc + loc).astype(np.float32)
    hn = (resid.astype(np.float32) + acc.astype(np.float16).astype(np.float32)).astype(np.float16).astype(np.float32)
    return hn


def simulate(W, C, T, h, S, calls, seed, alternate=True, both=True, max_steps=2_000_000, stuck_epoch=False):
    rng = np.random.default_rng(seed)
    Tmax = T + 1
    ranks = [Rank(W, Tmax, h) for _ in range(W)]
    resid = [rng.standard_normal((T, h)).astype(np.float16).astype(np.float32) for _ in range(W)]
    resid = [resid[0].copy() for _ in range(W)]                       # h is identical on all ranks when a call starts
    norm_w = rng.standard_normal(h).astype(np.float16).astype(np.float32)
    all_parts = [rng.standard_normal((W, S, T, h)).astype(np.float32) for _ in range(calls)]
    call_of = [0] * W                                                  # the call each rank is executing
    live = {r: None for r in range(W)}
    outs = {}

    def start(r):
        n = call_of[r]
        which = (n % 2) if alternate else 0
        oh, on = np.zeros((T, h), np.float32), np.zeros((T, h), np.float32)
        outs[(r, n)] = (oh, on)
        live[r] = [cta(ranks, r, which, c, t, C, W, Tmax, h, S, all_parts[n][r], resid[r], norm_w, 1e-6, oh, on, both)
                   for t in range(T) for c in range(C)]

    for r in range(W):
        start(r)
    steps = idle = 0
    while any(live[r] is not None for r in range(W)):
        steps += 1
        if steps > max_steps:
            return "deadlock", None
        r = int(rng.integers(W))
        if live[r] is None:
            continue
        g = live[r][int(rng.integers(len(live[r])))]
        try:
            ev = next(g)
        except StopIteration:
            ev = "done"
        if ev == "poll":
            idle += 1
            if idle > 200_000:
                return "deadlock", None
        else:
            idle = 0
        if ev == "done":
            live[r].remove(g)
            if not live[r]:                                            # kernel complete on this rank: publish epoch, next call
                ranks[r].epoch_state += 0 if stuck_epoch else 1
                resid[r] = outs[(r, call_of[r])][0].copy()             # resid_out aliases resid_in (st.h)
                call_of[r] += 1
                if call_of[r] < calls:
                    start(r)
                else:
                    live[r] = None
    # verify
    ref = resid[0] * 0
    cur = None
    for n in range(calls):
        base = outs[(0, n)]
        for r in range(W):
            oh, on = outs[(r, n)]
            if not (np.array_equal(oh, base[0]) and np.array_equal(on, base[1])):
                return "ranks differ", n
    # replay the expected chain
    rng2 = np.random.default_rng(seed)
    r0 = rng2.standard_normal((T, h)).astype(np.float16).astype(np.float32)
    cur = r0
    for n in range(calls):
        cur = expected(all_parts[n], cur, norm_w, 1e-6)
        if not np.array_equal(cur, outs[(0, n)][0]):
            return "wrong sum", n
        rstd = 1.0 / np.sqrt((cur.astype(np.float64) ** 2).mean(-1, keepdims=True) + 1e-6)
        want = norm_w * (cur * rstd)
        if not np.allclose(outs[(0, n)][1], want, rtol=3e-3, atol=3e-3):
            return "wrong norm", n
    return "ok", steps


@pytest.mark.parametrize("W,C,T,h,S", [(2, 2, 2, 32, 2), (4, 1, 3, 32, 1), (4, 2, 2, 64, 3), (8, 1, 1, 64, 2)])
def test_two_shot_ll_protocol_under_random_schedules(W, C, T, h, S):
    for seed in range(4):
        verdict, info = simulate(W, C, T, h, S, calls=5, seed=seed)
        assert verdict == "ok", (verdict, info, seed)


def test_a_single_buffer_set_is_safe_too():
    """Stronger than the kernel needs: a rank can only finish call n after every owner has consumed the call's reduce-scatter units
    (its phase C waits for all owners' phase B), and it only writes all-gather units of call n+1 after every peer has started call
    n+1 -- so even without alternating the two sets nothing is overwritten early.  (The model keeps alternating; this pins that a
    repeated set, e.g. two o_proj-type calls in a row, is not a hazard.)"""
    for seed in range(6):
        verdict, info = simulate(2, 2, 2, 32, 1, calls=6, seed=seed, alternate=False)
        assert verdict == "ok", (verdict, info, seed)


def test_simulation_detects_a_stuck_epoch():
    """If the epoch did not advance between calls, units of the previous call would be accepted as this call's."""
    bad = 0
    for seed in range(8):
        verdict, _ = simulate(2, 2, 2, 32, 1, calls=4, seed=seed, alternate=False, stuck_epoch=True, max_steps=400_000)
        bad += verdict != "ok"
    assert bad > 0


def test_simulation_detects_a_one_word_epoch_check():
    """Accepting a unit on its first epoch word reads the second half of a torn store too early."""
    bad = 0
    for seed in range(12):
        verdict, _ = simulate(2, 2, 2, 32, 1, calls=4, seed=seed, both=False)
        bad += verdict != "ok"
    assert bad > 0


# ---------------------------------------------------------------------------------------------------------------------------
# The same protocol at (tile, token) granularity inside the decode GEMM (gemm_decode_fused.cu, tensor-parallel tail): the owner of
# a 128-column tile is one rank, all-gather units are 8-byte halves of the 16-byte layout, the statistic is one 8-byte unit per
# (token, tile).  Calls of the stand-alone kernel and of the fused tail alternate on the SAME regions and epoch counter.
# ---------------------------------------------------------------------------------------------------------------------------
TILE = 8            # stands for the kernel's 128 (the index arithmetic only needs hw % TILE == 0 and TILE % 4 == 0)


def cta_fused(ranks, me, which, tile, split, S, W, Tmax, n, local_part, resid, out_h, out_ssq, ssq_cells):
    """One CTA of the cluster of `tile` on rank `me`: tokens split, split+S, ...; pair lanes as in the kernel."""
    hw, R = n // W, ranks[me]
    f0 = tile * TILE
    owner = f0 // hw
    epoch = R.epoch_state + 1
    T = local_part.shape[0]
    toks = range(split, T, S)
    for t in toks:                                                       # pass A
        for ft in range(0, TILE, 2):
            col = f0 - owner * hw + ft
            unit = ((me * Tmax + t) * hw + col) >> 1
            yield from st_ll(ranks[owner].rs[which], unit, np.float32(local_part[t, f0 + ft]), np.float32(local_part[t, f0 + ft + 1]), epoch)
    if owner == me:                                                      # pass B
        for t in toks:
            ss = np.float32(0)
            for ft in range(0, TILE, 2):
                col = f0 - owner * hw + ft
                while True:
                    us = [R.rs[which][((r * Tmax + t) * hw + col) >> 1].copy() for r in range(W)]
                    if all(ll_ok(u, epoch) for u in us):
                        break
                    yield "poll"
                a0 = a1 = np.float32(0)
                for u in us:
                    a0, a1 = np.float32(a0 + np.float32(u[0])), np.float32(a1 + np.float32(u[2]))
                h0 = rnd(np.float32(resid[t, f0 + ft]) + rnd(a0))
                h1 = rnd(np.float32(resid[t, f0 + ft + 1]) + rnd(a1))
                ss = np.float32(ss + np.float32(h0 * h0 + h1 * h1))
                unit = (me * Tmax + t) * hw + col
                for j in range(W):                                       # 8-byte half of the 16-byte all-gather unit
                    cell = ranks[j].ag[which][unit >> 2]
                    half = (unit >> 1) & 1
                    cell[2 * half], cell[2 * half + 1] = _pack(h0, h1), epoch
                    yield
            for j in range(W):
                ssq_cells[j][which][t, tile] = (float(ss), epoch)
                yield
    for t in toks:                                                       # pass C
        for ft in range(0, TILE, 2):
            col = f0 - owner * hw + ft
            unit = (owner * Tmax + t) * hw + col
            half = (unit >> 1) & 1
            while True:
                cell = R.ag[which][unit >> 2]
                if cell[2 * half + 1] == epoch:
                    w = cell[2 * half]
                    break
                yield "poll"
            out_h[t, f0 + ft], out_h[t, f0 + ft + 1] = _unpack(w)
        while ssq_cells[me][which][t, tile][1] != epoch:
            yield "poll"
        out_ssq[t, tile] = ssq_cells[me][which][t, tile][0]
    yield "done"


def simulate_mixed(W, T, n, S, kinds, seed):
    """kinds: sequence of 'll' (stand-alone kernel, C = 1) and 'fused' calls over the same regions / epochs."""
    rng = np.random.default_rng(seed)
    Tmax, tiles = T + 1, n // TILE
    ranks = [Rank(W, Tmax, n) for _ in range(W)]
    ssq_cells = [[np.zeros((Tmax, tiles), dtype=object) for _ in range(2)] for _ in range(W)]
    for r in range(W):
        for b in range(2):
            for i in range(Tmax):
                for j in range(tiles):
                    ssq_cells[r][b][i, j] = (0.0, 0)
    r0 = rng.standard_normal((T, n)).astype(np.float16).astype(np.float32)
    resid = [r0.copy() for _ in range(W)]
    norm_w = np.ones(n, np.float32)
    parts = [rng.standard_normal((W, 1, T, n)).astype(np.float32) for _ in kinds]
    call_of, live, outs = [0] * W, {}, {}

    def start(r):
        i = call_of[r]
        which = i % 2
        oh, aux = np.zeros((T, n), np.float32), np.zeros((T, max(n, tiles)), np.float32)
        outs[(r, i)] = (oh, aux)
        if kinds[i] == "ll":
            live[r] = [cta(ranks, r, which, 0, t, 1, W, Tmax, n, 1, parts[i][r], resid[r], norm_w, 1e-6, oh, aux) for t in range(T)]
        else:
            live[r] = [cta_fused(ranks, r, which, tile, sp, S, W, Tmax, n, parts[i][r, 0], resid[r], oh, aux, ssq_cells)
                       for tile in range(tiles) for sp in range(S)]

    for r in range(W):
        start(r)
    idle = 0
    while any(live[r] is not None for r in range(W)):
        r = int(rng.integers(W))
        if live[r] is None:
            continue
        g = live[r][int(rng.integers(len(live[r])))]
        try:
            ev = next(g)
        except StopIteration:
            ev = "done"
        idle = idle + 1 if ev == "poll" else 0
        if idle > 300_000:
            return "deadlock", None
        if ev == "done":
            live[r].remove(g)
            if not live[r]:
                ranks[r].epoch_state += 1
                resid[r] = outs[(r, call_of[r])][0].copy()
                call_of[r] += 1
                if call_of[r] < len(kinds):
                    start(r)
                else:
                    live[r] = None
    cur = r0
    for i, kind in enumerate(kinds):
        cur = expected(parts[i], cur, norm_w, 1e-6)
        for r in range(W):
            if not np.array_equal(outs[(r, i)][0], cur):
                return "wrong sum", (i, r)
            if not np.array_equal(outs[(r, i)][1], outs[(0, i)][1]):
                return "ranks differ", (i, r)
        if kind == "fused":                                             # per-tile statistic = sum of squares of the tile's h
            want = (cur.reshape(T, tiles, TILE).astype(np.float64) ** 2).sum(-1)
            if not np.allclose(outs[(0, i)][1][:, :tiles], want, rtol=1e-5):
                return "wrong statistic", i
    return "ok", None


@pytest.mark.parametrize("W,T,n,S", [(2, 3, 32, 2), (4, 2, 64, 1), (2, 5, 48, 3)])
def test_fused_gemm_tail_and_stand_alone_kernel_share_regions(W, T, n, S):
    for seed in range(3):
        verdict, info = simulate_mixed(W, T, n, S, ["fused", "fused", "ll", "fused", "ll", "ll", "fused"], seed)
        assert verdict == "ok", (verdict, info, seed)
