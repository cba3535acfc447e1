"""Host logic of the drop-in surfaces on CPU: the package driven through the torch test double of the C-ABI
(tests/cabi_double.py), checked against the oracle.  Covers what the GPU tests cannot see without a device: prompt
expansion / merge order (chatts_vllm.py:538-574), left padding, paged-KV bookkeeping across generate() calls, split-K
workspace plumbing, sampling arguments and the vLLM request shape (demo/demo_vllm.py:18-63).
"""
import numpy as np
import pytest
import torch

from chatts_b200 import ChatTSConfig, ChatTSProcessor, SimpleTokenizer
from chatts_b200.weights import synthetic_state_dict
from oracle import decoder as od
from oracle import merge as om
from oracle import ts_encoder as ote

DT = torch.bfloat16


def _series():
    x = np.arange(256)
    a = np.sin(x / 10) * 5.0
    a[100:] -= 10.0
    return a, (x * 0.05)[:100]


def _build(cabi_double, split=1, qwen3=False, **kw):
    from chatts_b200.model import ChatTSForCausalLM

    cabi_double.split = split
    cfg = ChatTSConfig.tiny()
    if qwen3:
        cfg.qk_norm, cfg.attention_bias = True, False
    sd = synthetic_state_dict(cfg, seed=1234, device="cpu", dtype=DT, std=0.05)
    model = ChatTSForCausalLM(cfg, sd, device="cpu", dtype=DT, max_batch=8, max_seq_len=512, page_size=16,
                              use_cuda_graph=False, **kw)
    proc = ChatTSProcessor(SimpleTokenizer(cfg.ts_token_start_index, cfg.pad_token_id, cfg.eos_token_id), cfg)
    return cfg, sd, model, proc


def _oracle_embeds(cfg, sd, enc):
    ts_w = {k[len("ts_encoder."):]: v for k, v in sd.items() if k.startswith("ts_encoder.")}
    feats, pc = ote.forward(enc["timeseries"].to(DT), cfg.ts, ts_w)
    return om.hf_merge(enc["input_ids"], enc["attention_mask"], sd["model.embed_tokens.weight"], feats, pc.tolist(),
                       cfg.ts_token_start_index)


PROMPTS = ["A <ts><ts/> and B <ts><ts/> ?", "Only text, no series, but a longer prompt to left-pad the other one"]


@pytest.mark.parametrize("split,qwen3", [(1, False), (3, False), (2, True)])
def test_forward_matches_oracle_through_double(cabi_double, split, qwen3):
    cfg, sd, model, proc = _build(cabi_double, split, qwen3)
    enc = proc(text=PROMPTS, timeseries=list(_series()), padding=True, return_tensors="pt")
    out = model.forward(enc["input_ids"], enc["attention_mask"], enc["timeseries"], logits_to_keep=0).logits
    for b, e in enumerate(_oracle_embeds(cfg, sd, enc)):
        ref = od.logits(od.forward_hidden(e, sd, cfg.to_dict(), od.State(cfg.num_hidden_layers)), sd).float()
        assert out[b].shape == ref.shape                      # expanded length = text + patch rows of this sample
        err = float((out[b].float() - ref).abs().max() / ref.abs().max())
        assert err < 2.5e-2, err                               # two bf16 evaluations with different accumulation order


def test_generate_bookkeeping_and_teacher_forced_tokens(cabi_double):
    cfg, sd, model, proc = _build(cabi_double, split=2)
    enc = proc(text=PROMPTS, timeseries=list(_series()), padding=True, return_tensors="pt")
    S = enc["input_ids"].shape[1]
    new = 12
    ids = model.generate(**enc, max_new_tokens=new, ignore_eos=True)
    assert ids.shape == (2, S + new) and torch.equal(ids[:, :S], enc["input_ids"])
    assert len(model.pool.free) == model.pool.num_pages        # every KV page returned
    # teacher-forced: feed OUR tokens to the oracle; wherever we differ from its argmax the two logits must be a near tie
    for b, e in enumerate(_oracle_embeds(cfg, sd, enc)):
        st = od.State(cfg.num_hidden_layers)
        lg = od.logits(od.forward_hidden(e, sd, cfg.to_dict(), st)[-1:], sd)[0].float()
        exact = 0
        for t in range(new):
            tok = int(ids[b, S + t])
            top = float(lg.max())
            assert top - float(lg[tok]) <= 3e-2 * max(1.0, abs(top)), (b, t)
            exact += int(int(lg.argmax()) == tok)
            nxt = sd["model.embed_tokens.weight"][tok][None, :]
            lg = od.logits(od.forward_hidden(nxt, sd, cfg.to_dict(), st), sd)[0].float()
        assert exact >= new - 2
    # a second call reuses the pool and gives the same answer (no state leaks between calls)
    again = model.generate(**enc, max_new_tokens=new, ignore_eos=True)
    assert torch.equal(again, ids)


def test_generate_arguments(cabi_double):
    cfg, sd, model, proc = _build(cabi_double)
    enc = proc(text=PROMPTS, timeseries=list(_series()), padding=True, return_tensors="pt")
    S = enc["input_ids"].shape[1]
    greedy = model.generate(**enc, max_new_tokens=6, ignore_eos=True)
    # max_length counts the prompt (inference_tsmllm_deepspeed.py:101); synced_gpus is accepted and ignored
    ml = model.generate(**enc, max_length=S + 6, ignore_eos=True, synced_gpus=False)
    assert torch.equal(ml, greedy)
    # EOS stop: declare the 3rd generated token of sample 0 to be EOS -> that row is padded after it
    eos = int(greedy[0, S + 2])
    stopped = model.generate(**enc, max_new_tokens=6, eos_token_id=eos)
    row = stopped[0, S:].tolist()
    first = row.index(eos)
    assert first <= 2 and all(t == cfg.pad_token_id or t == eos for t in row[first + 1:])
    # sampling: seeded -> reproducible; temperature/top_p accepted (inference_tsmllm_deepspeed.py:103-105)
    s1 = model.generate(**enc, max_new_tokens=5, do_sample=True, temperature=0.7, top_p=0.9, ignore_eos=True, seed=7)
    s2 = model.generate(**enc, max_new_tokens=5, do_sample=True, temperature=0.7, top_p=0.9, ignore_eos=True, seed=7)
    assert s1.shape == (2, S + 5) and torch.equal(s1, s2)
    assert len(model.pool.free) == model.pool.num_pages


def test_placeholder_mismatch_raises(cabi_double):
    cfg, sd, model, proc = _build(cabi_double)
    with pytest.raises(AssertionError):                        # encoding_utils.py:58,68
        proc(text=["one <ts><ts/> only"], timeseries=list(_series()), padding=True, return_tensors="pt")


def test_vllm_request_shape(cabi_double):
    from chatts_b200.vllm_compat import LLM, SamplingParams

    cfg, sd, model, proc = _build(cabi_double)
    a, b = _series()
    llm = LLM(model=model)
    reqs = [{"prompt": "two: <ts><ts/> <ts><ts/>", "multi_modal_data": {"timeseries": [a, b.tolist()]}},
            {"prompt": "none"}]
    outs = llm.generate(reqs, SamplingParams(max_tokens=6, ignore_eos=True))
    assert [len(o.outputs[0].token_ids) for o in outs] == [6, 6]
    assert all(isinstance(o.outputs[0].text, str) for o in outs)
    with pytest.raises(TypeError):                             # chatts_vllm.py:277-279
        llm.generate([{"prompt": "x <ts><ts/>", "multi_modal_data": {"timeseries": ["not a series"]}}],
                     SamplingParams(max_tokens=2))
    # plain strings = text-only prompts through the same call (llm_utils.py:127), keyword form of demo_vllm.py:59
    plain = llm.generate(["none", "another"], sampling_params=SamplingParams(max_tokens=6, ignore_eos=True), use_tqdm=False)
    assert plain[0].prompt == "none" and plain[0].outputs[0].token_ids == outs[1].outputs[0].token_ids
    assert len(llm.generate("none", SamplingParams(max_tokens=3, ignore_eos=True))) == 1


def test_vllm_sampling_params_n_stop_topk(cabi_double):
    """The SamplingParams fields the reference's workers set (llm_utils.py:153): n completions per request, stop strings,
    stop_token_ids, top_k / top_p / temperature."""
    from chatts_b200.vllm_compat import LLM, SamplingParams

    cfg, sd, model, proc = _build(cabi_double)
    a, _ = _series()
    llm = LLM(model=model)
    reqs = [{"prompt": "s: <ts><ts/>", "multi_modal_data": {"timeseries": [a]}}, {"prompt": "plain"}]
    outs = llm.generate(reqs, SamplingParams(max_tokens=5, temperature=0.7, top_p=0.9, top_k=50, n=3, seed=4, ignore_eos=True))
    assert len(outs) == 2 and all(len(o.outputs) == 3 for o in outs)
    assert all(len(c.token_ids) == 5 for o in outs for c in o.outputs)
    assert len({tuple(c.token_ids) for c in outs[0].outputs}) > 1           # independent draws per completion
    again = llm.generate(reqs, SamplingParams(max_tokens=5, temperature=0.7, top_p=0.9, top_k=50, n=3, seed=4, ignore_eos=True))
    assert [[c.token_ids for c in o.outputs] for o in again] == [[c.token_ids for c in o.outputs] for o in outs]
    g = llm.generate(reqs[1:], SamplingParams(max_tokens=8, ignore_eos=True))[0].outputs[0]
    if len(g.text) >= 3:                                                       # cut BEFORE the first stop string
        stop = g.text[2:3]
        cut = llm.generate(reqs[1:], SamplingParams(max_tokens=8, ignore_eos=True, stop=[stop]))[0].outputs[0]
        assert cut.text == g.text[: g.text.find(stop)] and stop not in cut.text


@pytest.mark.parametrize("split,qwen3", [(1, False), (3, True)])
def test_native_step_executor_generates_the_same_tokens(cabi_double, split, qwen3):
    """use_native_step=True routes the decode step through ctx.decoder_step (cts_decoder_step: one C call per step): same
    tokens as the per-kernel orchestration, same paged-KV bookkeeping."""
    cfg, sd, model, proc = _build(cabi_double, split, qwen3)
    enc = proc(text=PROMPTS, timeseries=list(_series()), padding=True, return_tensors="pt")
    ref = model.generate(**enc, max_new_tokens=9, ignore_eos=True)
    cfg2, sd2, native, _ = _build(cabi_double, split, qwen3, use_native_step=True)
    out = native.generate(**enc, max_new_tokens=9, ignore_eos=True)
    assert torch.equal(out, ref) and len(native.pool.free) == native.pool.num_pages


def test_from_pretrained_reads_generation_config_defaults(cabi_double, tmp_path):
    """HF's generate() falls back to the checkpoint's generation_config.json when the caller passes no sampling arguments
    (README.md:102): from_pretrained picks those defaults up; explicit arguments still win."""
    import json
    from safetensors.torch import save_file
    from chatts_b200.model import ChatTSForCausalLM

    cfg, sd, model, proc = _build(cabi_double)
    d = tmp_path / "ckpt"
    d.mkdir()
    json.dump({**cfg.to_dict(), "architectures": ["Qwen2TSForCausalLM"]}, open(d / "config.json", "w"))
    save_file({k: v.contiguous() for k, v in sd.items()}, str(d / "model.safetensors"))
    json.dump({"do_sample": True, "temperature": 0.7, "top_p": 0.8, "top_k": 20, "eos_token_id": [998, 997]}, open(d / "generation_config.json", "w"))
    m = ChatTSForCausalLM.from_pretrained(str(d), device="cpu", torch_dtype=DT, max_batch=4, max_seq_len=256, page_size=16, use_cuda_graph=False)
    assert m.generation_defaults["temperature"] == 0.7 and m.generation_defaults["top_k"] == 20
    enc = proc(text=["hello"], timeseries=[], return_tensors="pt")
    a = m.generate(**enc, max_new_tokens=8, seed=1, ignore_eos=True)
    b = m.generate(**enc, max_new_tokens=8, seed=2, ignore_eos=True)
    g1 = m.generate(**enc, max_new_tokens=8, do_sample=False, ignore_eos=True)
    g2 = model.generate(**enc, max_new_tokens=8, ignore_eos=True)
    assert not torch.equal(a, b)                       # the checkpoint's sampling defaults are in force
    assert torch.equal(g1, g2)                         # explicit greedy request == the model without a generation config


@pytest.mark.parametrize("split,qwen3", [(1, False), (3, True)])
def test_fused_decode_gemms_generate_the_same_tokens(cabi_double, split, qwen3):
    """use_fused_decode=True: QKV / o_proj / gate_up / down_proj through ctx.gemm_decode_fused (cluster-reduced split-K + fused
    tail) and plain RMSNorms between them -- same tokens, same KV cache contents as the nine-stage layer."""
    cfg, sd, model, proc = _build(cabi_double, split, qwen3)
    enc = proc(text=PROMPTS, timeseries=list(_series()), padding=True, return_tensors="pt")
    ref = model.generate(**enc, max_new_tokens=9, ignore_eos=True)
    for level in (1, 2):                       # 7 stages (plain RMSNorm launches kept) and 5 stages (RMSNorm inside the projections)
        cfg2, sd2, fused, _ = _build(cabi_double, split, qwen3, use_fused_decode=level)
        out = fused.generate(**enc, max_new_tokens=9, ignore_eos=True)
        assert torch.equal(out, ref) and len(fused.pool.free) == fused.pool.num_pages, level
        if level == 1:
            assert torch.equal(fused.kv, model.kv)
        else:                                  # the row statistic is summed tile by tile: at most an ulp of bf16 in the cache
            assert float((fused.kv.float() - model.kv.float()).abs().max()) <= 2 ** -6 * float(model.kv.float().abs().max())


def test_async_engine_streams_cumulative_text(cabi_double):
    """AsyncLLMEngine.generate as chatts/utils/vllm_stream_qa.py:52-59 drives it: an async generator whose outputs[0].text grows
    token by token and ends with exactly what the blocking call returns."""
    import asyncio

    from chatts_b200.vllm_compat import LLM, AsyncEngineArgs, AsyncLLMEngine, SamplingParams

    cfg, sd, model, proc = _build(cabi_double)
    a, _ = _series()
    llm = LLM(model=model)
    eng = AsyncLLMEngine.from_engine_args(AsyncEngineArgs(model=None, max_model_len=512, limit_mm_per_prompt={"timeseries": 15}), llm=llm)
    req = {"prompt": "one: <ts><ts/> ?", "multi_modal_data": {"timeseries": [a]}}
    sp = SamplingParams(max_tokens=7, ignore_eos=True)
    want = llm.generate([req], sp)[0].outputs[0]

    async def run():
        seen = []
        async for out in eng.generate(req, sp, request_id=1.0):
            seen.append((out.outputs[0].text, list(out.outputs[0].token_ids)))
        return seen

    seen = asyncio.run(run())
    assert len(seen) == 8                                    # 7 partial outputs + the final one
    for (t0, k0), (t1, k1) in zip(seen[:-2], seen[1:-1]):
        assert t1.startswith(t0) and k1[: len(k0)] == k0 and len(k1) == len(k0) + 1
    assert seen[-1] == (want.text, want.token_ids) and seen[-2][1] == want.token_ids
    # a plain string prompt, and an error inside the worker reaches the consumer
    assert len(asyncio.run(_collect(eng.generate("plain text", SamplingParams(max_tokens=2, ignore_eos=True))))) == 3
    with pytest.raises(TypeError):
        asyncio.run(_collect(eng.generate({"prompt": "x <ts><ts/>", "multi_modal_data": {"timeseries": ["bad"]}}, sp)))


async def _collect(agen):
    return [x async for x in agen]

