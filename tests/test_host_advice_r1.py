"""Regression tests for the round-1 advisor findings, all on CPU through the C-ABI test double:
completion text / token ids end at the row's own stop token, stop ids are a union, streaming decodes incrementally; KV pages
never leak when an allocation fails; sampling defaults follow HF (temperature 1.0, fresh seed per unseeded call); an over-long
series / sequence raises instead of indexing past a table; one invalid request does not take its prefill group down."""
import numpy as np
import pytest
import torch

from tests.test_host_model import _build, _series


class _Utf8Tokenizer:
    """ids < 256 are UTF-8 bytes; 300 = <|im_end|> (a special token that an HF tokenizer would render as text unless skipped)."""
    pad_token_id, eos_token_id = 301, 300

    def decode(self, ids, skip_special_tokens=False):
        out = bytes(int(t) for t in ids if int(t) < 256).decode("utf-8", errors="replace")
        if not skip_special_tokens:
            out += "".join("<|im_end|>" for t in ids if int(t) == 300)
        return out


def test_cut_at_stop_and_union_of_stop_ids():
    from chatts_b200 import ChatTSConfig
    from chatts_b200.vllm_compat import cut_at_stop, decode_text, eos_ids
    cfg = ChatTSConfig.tiny()
    stop = eos_ids(cfg, [7, 8], [cfg.eos_token_id, 5])
    assert stop == sorted({cfg.eos_token_id, 5, 7, 8})                 # request stop ids ADD to the model's EOS
    toks, fin = cut_at_stop([65, 66, 8, 999, 999], stop)
    assert toks == [65, 66] and fin == "stop"                          # the stop token and the pad fill after it are dropped
    toks, fin = cut_at_stop([65, 66, 67], stop)
    assert toks == [65, 66, 67] and fin == "length"
    assert cut_at_stop([65, 8, 66], stop, ignore_eos=True) == ([65, 8, 66], "length")
    assert decode_text(_Utf8Tokenizer(), [72, 105, 300]) == "Hi"       # skip_special_tokens=True reaches the tokenizer


def test_incremental_decoder_holds_back_partial_utf8():
    from chatts_b200.vllm_compat import IncrementalDecoder
    dec = IncrementalDecoder(_Utf8Tokenizer())
    data = "时序ok".encode("utf-8")
    pieces = [dec.push([b]) for b in data]
    assert "�" not in "".join(pieces)
    assert "".join(pieces) + dec.flush() == "时序ok"
    assert pieces[0] == "" and pieces[1] == "" and pieces[2] == "时"     # a character appears when its last byte arrives


def test_llm_outputs_end_at_each_rows_own_eos(cabi_double, monkeypatch):
    from chatts_b200 import vllm_compat
    cfg, sd, model, proc = _build(cabi_double)
    llm = vllm_compat.LLM(model=model)
    S_holder = {}

    def fake_generate(**kw):
        ids = kw["input_ids"]
        S_holder["eos"] = kw["eos_token_id"]
        new = torch.tensor([[72, 105, cfg.eos_token_id, cfg.pad_token_id, cfg.pad_token_id], [72, 101, 108, 108, 111]])
        return torch.cat([ids, new], 1)

    monkeypatch.setattr(model, "generate", fake_generate)
    outs = llm.generate(["a", "b"], vllm_compat.SamplingParams(max_tokens=5, stop_token_ids=[7]))
    assert set(S_holder["eos"]) == {cfg.eos_token_id, 7}
    a, b = outs[0].outputs[0], outs[1].outputs[0]
    assert a.text == "Hi" and a.token_ids == [72, 105] and a.finish_reason == "stop"
    assert b.text == "Hello" and len(b.token_ids) == 5 and b.finish_reason == "length"


def test_alloc_pages_takes_nothing_when_it_fails(cabi_double):
    cfg, sd, model, proc = _build(cabi_double)
    free0 = len(model.pool.free)
    with pytest.raises(ValueError):
        model._alloc_pages(np.array([10, model.max_seq_len + model.page_size * 2]), 0)     # second row over-long
    assert len(model.pool.free) == free0
    with pytest.raises(RuntimeError):
        model._alloc_pages(np.array([model.max_seq_len] * 8), 0) and model._alloc_pages(np.array([model.max_seq_len] * 8), 0)
    # whatever the first call of the line above took is all that is gone; a failing call itself takes nothing
    held_now = free0 - len(model.pool.free)
    assert held_now in (0, 8 * model.max_pages)


def test_sequences_beyond_the_rotary_table_are_rejected(cabi_double):
    cfg, sd, model, proc = _build(cabi_double)
    assert model.n_pos >= model.max_pages * model.page_size or model.n_pos == cfg.max_position_embeddings
    enc = proc(text=["x" * 20], timeseries=[], return_tensors="pt")
    with pytest.raises(ValueError):
        model.generate(**enc, max_new_tokens=model.n_pos, ignore_eos=True)


def test_do_sample_without_temperature_samples_and_unseeded_calls_differ(cabi_double):
    cfg, sd, model, proc = _build(cabi_double)
    enc = proc(text=["A <ts><ts/> ?"], timeseries=[_series()[0]], return_tensors="pt")
    S = enc["input_ids"].shape[1]
    greedy = model.generate(**enc, max_new_tokens=12, ignore_eos=True)[0, S:].tolist()
    runs = [model.generate(**enc, max_new_tokens=12, ignore_eos=True, do_sample=True)[0, S:].tolist() for _ in range(4)]
    assert any(r != greedy for r in runs)                               # temperature defaults to 1.0: not silently greedy
    assert len({tuple(r) for r in runs}) > 1                            # unseeded calls draw fresh seeds
    a = model.generate(**enc, max_new_tokens=12, ignore_eos=True, do_sample=True, seed=5)[0, S:].tolist()
    b = model.generate(**enc, max_new_tokens=12, ignore_eos=True, do_sample=True, seed=5)[0, S:].tolist()
    assert a == b                                                       # a seed still reproduces a generation


def test_series_longer_than_the_position_table_raises_index_error(cabi_double):
    cfg, sd, model, proc = _build(cabi_double)
    n = cfg.ts["max_sequence_length"] + 16
    enc = proc(text=["A <ts><ts/> ?"], timeseries=[np.sin(np.arange(n) / 7.0)], return_tensors="pt")
    with pytest.raises(IndexError):
        model.forward(enc["input_ids"], enc["attention_mask"], enc["timeseries"])


def test_engine_fails_only_the_invalid_request(cabi_double):
    from chatts_b200.engine import ContinuousEngine
    cfg, sd, model, proc = _build(cabi_double)
    good = proc(text=["plain prompt"], timeseries=[], return_tensors="pt")
    want = model.generate(**good, max_new_tokens=4, ignore_eos=True)[0, good["input_ids"].shape[1]:].tolist()
    pages0 = len(model.pool.free)
    eng = ContinuousEngine(model, slots=2, steps_per_round=2, max_prefill_batch=2)
    eng.add_request(good["input_ids"][0], good["timeseries"], max_new_tokens=4, ignore_eos=True)
    eng.add_request(torch.arange(10) % 200, None, max_new_tokens=model.max_seq_len * 2, ignore_eos=True)     # can never fit
    eng.add_request(good["input_ids"][0], good["timeseries"], max_new_tokens=4, ignore_eos=True)
    done = eng.run()
    assert [r.rid for r in done] == [0, 1, 2]
    assert done[0].tokens == want and done[2].tokens == want and done[0].error is None
    assert isinstance(done[1].error, ValueError) and done[1].tokens == []
    eng.close()
    assert len(model.pool.free) == pages0
