"""Host index logic (chatts_b200/layout.py) against the slow oracle statement -- bit-exact."""
import numpy as np
import pytest

from chatts_b200 import layout as L
from oracle import merge as om

TS = 990


def _rand_case(rng, B, S, left_pad):
    ids = rng.integers(0, 900, size=(B, S))
    am = np.ones((B, S), dtype=np.int64)
    n_series = 0
    for b in range(B):
        pad = int(rng.integers(0, S // 2))
        if left_pad:
            am[b, :pad] = 0
        else:
            am[b, S - pad:] = 0
        real = np.nonzero(am[b])[0]
        k = int(rng.integers(0, 4))
        cand = real[:-1]
        if k and len(cand) > 2 * k:
            pos = np.sort(rng.choice(cand[::2], size=k, replace=False))
            for p in pos:
                ids[b, p] = TS
                ids[b, p + 1] = TS + 1
            n_series += k
        ids[b, am[b] == 0] = 999
    pc = rng.integers(0, 9, size=n_series)
    return ids, am, pc


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("left_pad", [True, False])
@pytest.mark.parametrize("mode", ["insert", "overwrite"])
def test_hf_layout_matches_oracle(seed, left_pad, mode):
    rng = np.random.default_rng(seed)
    ids, am, pc = _rand_case(rng, B=int(rng.integers(1, 5)), S=40, left_pad=left_pad)
    lay = L.hf_layout(ids, am, pc, TS, mode)
    ref = om.hf_layout(ids, am, pc.tolist(), TS, mode)
    flat = [e for sample in ref for e in sample]
    assert lay.total == len(flat)
    assert lay.cu_seqlens.tolist() == np.cumsum([0] + [len(s) for s in ref]).tolist()
    # token rows: same id, same source column; patch rows: same global row at the same position
    b_of = np.repeat(np.arange(len(ref)), [len(s) for s in ref])
    for t, (kind, idx) in enumerate(flat):
        if kind == "tok":
            assert lay.ids[t] == ids[b_of[t], idx] and lay.src_col[t] == idx
        else:
            assert lay.ids[t] == -1 and lay.row_map[idx] == t
    pos_ref = np.concatenate([np.arange(len(s)) for s in ref]) if flat else np.zeros(0)
    assert lay.positions.tolist() == pos_ref.tolist()


def test_hf_layout_count_mismatch_raises():
    ids = np.array([[1, TS, TS + 1, 2]])
    with pytest.raises(AssertionError):
        L.hf_layout(ids, None, [3, 4], TS)
    with pytest.raises(AssertionError):
        om.hf_layout(ids, np.ones_like(ids), [3, 4], TS)


def test_hf_layout_overwrite_mode_replaces_the_pair():
    """mode "overwrite": <ts><ts/> give up their two slots, the P rows take their place (P - 2 new positions per series)."""
    ids = np.array([[5, TS, TS + 1, 6, TS, TS + 1]])
    lay = L.hf_layout(ids, None, [3, 1], TS, "overwrite")
    assert lay.ids.tolist() == [5, -1, -1, -1, 6, -1] and lay.row_map.tolist() == [1, 2, 3, 5]
    assert lay.src_col.tolist() == [0, -1, -1, -1, 3, -1] and lay.total == 6 + (3 - 2) + (1 - 2)
    ins = L.hf_layout(ids, None, [3, 1], TS, "insert")
    assert ins.total == 6 + 3 + 1 and ins.ids.tolist() == [5, TS, -1, -1, -1, TS + 1, 6, TS, -1, TS + 1]
    with pytest.raises(ValueError):
        L.hf_layout(ids, None, [3, 1], TS, "replace")


def test_hf_layout_empty_series_and_no_series():
    ids = np.array([[5, TS, TS + 1, 6], [7, 8, 9, 10]])
    lay = L.hf_layout(ids, None, [0], TS)
    assert lay.total == 8 and lay.row_map.shape == (0,)
    lay = L.hf_layout(ids[1:], None, [], TS)
    assert lay.ids.tolist() == [7, 8, 9, 10]


def test_vllm_layout_and_expand():
    toks = [3, TS, TS + 1, 4, TS, TS + 1, 5]
    exp = L.expand_prompt_vllm(toks, [[11, TS, 12], [TS]], [3, 2], TS)
    assert exp == om.vllm_expand_prompt(toks, [[11, TS, 12], [TS]], [3, 2], TS)
    assert exp == [3, 11, TS, 12, TS, TS, 4, TS, TS, 5]
    lay = L.vllm_layout(exp, 5, TS)
    assert lay.row_map.tolist() == [2, 4, 5, 7, 8]
    assert (lay.ids[lay.row_map] == -1).all()
    with pytest.raises(ValueError):
        L.vllm_layout(exp, 4, TS)
