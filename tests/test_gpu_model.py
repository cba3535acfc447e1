"""End-to-end parity of the B200 path (processor -> TS encoder -> merge -> decoder -> logits / generate) against
the CPU oracle on identical seeded weights and inputs (BASELINE.json configs[0]-style, at a size the oracle
finishes in seconds).  Index placement is bit-exact; logits within the north_star tolerance."""
import numpy as np
import pytest
import torch

from oracle import decoder as od
from oracle import merge as om
from oracle import ts_encoder as ote
from tests.gpu_util import parity_gate, record, rel_err

pytestmark = pytest.mark.gpu
DT = torch.bfloat16


def _setup(seed=0, **cfg_kw):
    from chatts_b200 import ChatTSConfig, ChatTSProcessor, SimpleTokenizer
    from chatts_b200.model import ChatTSForCausalLM
    from chatts_b200.weights import synthetic_state_dict
    cfg = ChatTSConfig.tiny(**cfg_kw)
    sd = synthetic_state_dict(cfg, seed=1234 + seed, device="cpu", dtype=DT, std=0.05)
    model = ChatTSForCausalLM(cfg, sd, dtype=DT, max_batch=8, max_seq_len=512, page_size=16)
    tok = SimpleTokenizer(cfg.ts_token_start_index, cfg.pad_token_id, cfg.eos_token_id)
    proc = ChatTSProcessor(tok, cfg)
    return cfg, sd, model, proc


def _oracle_cfg(cfg):
    d = cfg.to_dict()
    return d


def _oracle_embeds(cfg, sd, enc):
    """Per-sample merged input embeddings by the oracle (TS encoder restatement + HF-layout merge)."""
    ts_w = {k[len("ts_encoder."):]: v for k, v in sd.items() if k.startswith("ts_encoder.")}
    x = enc["timeseries"].to(DT)
    if x.shape[0] > 0:
        feats, pc = ote.forward(x, cfg.ts, ts_w)
        pc = pc.tolist()
    else:
        feats, pc = torch.zeros(0, cfg.hidden_size, dtype=DT), []
    return om.hf_merge(enc["input_ids"], enc["attention_mask"], sd["model.embed_tokens.weight"], feats, pc,
                       cfg.ts_token_start_index), pc


def _demo_series():
    x = np.arange(256)
    ts1 = np.sin(x / 10) * 5.0
    ts1[100:] -= 10.0
    ts2 = x * 0.05
    ts2[103] += 10.0
    return ts1, ts2


def test_forward_logits_all_positions_match_oracle():
    cfg, sd, model, proc = _setup()
    ts1, ts2 = _demo_series()
    rng = np.random.default_rng(0)
    prompts = ["I have 2 series. TS1: <ts><ts/>; TS2: <ts><ts/>. Compare them.", "Short <ts><ts/> one", "no series here at all"]
    series = [ts1, ts2, rng.normal(size=77)]
    enc = proc(text=prompts, timeseries=series, padding=True, return_tensors="pt")
    out = model.forward(enc["input_ids"], enc["attention_mask"], enc["timeseries"], logits_to_keep=0).logits
    embeds, pc = _oracle_embeds(cfg, sd, enc)
    assert pc == [16, 16, 5]
    worst = 0.0
    for b, e in enumerate(embeds):
        st = od.State(cfg.num_hidden_layers)
        ref = od.logits(od.forward_hidden(e, sd, _oracle_cfg(cfg), st), sd)
        assert out[b].shape == ref.shape                       # merged length = S_text + sum(P): index placement
        worst = max(worst, rel_err(out[b], ref))
    record("model_forward_all_positions", err=worst)
    assert worst < 1.85e-2      # bf16 vs bf16 with different accumulation order, every position (measured on a B200: 1.23e-2)


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def test_next_token_logits_comparative_gate(dt):
    """BASELINE.json configs[0] in both 16-bit dtypes (fp16 is what the reference runs): the north_star metric
    max|d|/max|ref| on the next-token logits against the oracle in the same dtype AND against the fp32 oracle, gated comparatively
    (tests/gpu_util.py:parity_gate -- the same-dtype CPU oracle is itself 1.1e-2 (bf16) / 1.5e-3 (fp16) from fp32)."""
    from chatts_b200 import ChatTSConfig, ChatTSProcessor, SimpleTokenizer
    from chatts_b200.model import ChatTSForCausalLM
    from chatts_b200.weights import synthetic_state_dict
    cfg = ChatTSConfig.tiny()
    sd = synthetic_state_dict(cfg, seed=1235, device="cpu", dtype=dt, std=0.05)
    model = ChatTSForCausalLM(cfg, sd, dtype=dt, max_batch=4, max_seq_len=512, page_size=16)
    proc = ChatTSProcessor(SimpleTokenizer(cfg.ts_token_start_index, cfg.pad_token_id, cfg.eos_token_id), cfg)
    ts1, _ = _demo_series()
    # configs[0]: one 256-point sine series + a 64-token prompt
    enc = proc(text=["Describe <ts><ts/> please, in detail, with numbers and dates: " + "x" * 20], timeseries=[ts1], return_tensors="pt")
    lg = model.forward(enc["input_ids"], enc["attention_mask"], enc["timeseries"]).logits[:, 0]
    refs = []
    for fp32 in (False, True):
        w = {k: v.float() for k, v in sd.items()} if fp32 else sd
        ts_w = {k[len("ts_encoder."):]: v for k, v in w.items() if k.startswith("ts_encoder.")}
        x = enc["timeseries"].to(dt)
        feats, pc = ote.forward(x.float() if fp32 else x, cfg.ts, ts_w)
        emb = om.hf_merge(enc["input_ids"], enc["attention_mask"], w["model.embed_tokens.weight"], feats, pc.tolist(),
                          cfg.ts_token_start_index)[0]
        refs.append(od.logits(od.forward_hidden(emb, w, _oracle_cfg(cfg), od.State(cfg.num_hidden_layers))[-1:], w))
    # fixed bounds: 1.5 x the B200 measurements (bf16 1.0e-2; fp16 ~1.3e-3 against fp32, same order against the fp16 oracle)
    parity_gate("model_next_token_logits", lg, refs[0], refs[1], dt, 1.57e-2 if dt == torch.bfloat16 else 1.6e-3)          # measured: 1.05e-2 / 1.04e-3


@pytest.mark.parametrize("use_graph", [True, False])
def test_generate_greedy_matches_oracle(use_graph):
    cfg, sd, model, proc = _setup(seed=2)
    model.use_cuda_graph = use_graph
    ts1, ts2 = _demo_series()
    prompts = ["A <ts><ts/> and B <ts><ts/> ?", "Only text, no series, but a longer prompt to left-pad the other one"]
    enc = proc(text=prompts, timeseries=[ts1, ts2[:100]], padding=True, return_tensors="pt")
    new = 40
    ids = model.generate(**enc, max_new_tokens=new, ignore_eos=True)
    S = enc["input_ids"].shape[1]
    assert ids.shape == (2, S + new)
    assert torch.equal(ids[:, :S], enc["input_ids"])          # README.md:103: the first S columns are the original ids
    embeds, _ = _oracle_embeds(cfg, sd, enc)
    # Teacher-forced check against the oracle: feed the oracle the tokens the B200 path produced and require, at every
    # step, that the produced token is the oracle's argmax or within bf16 noise of it (random weights give near-ties;
    # a free-running comparison would diverge at the first one).  Also report how long the free-running sequences agree.
    agree, worst_gap, exact = [], 0.0, 0
    for b, e in enumerate(embeds):
        got = ids[b, S:].tolist()
        ref_toks, _ = od.greedy_generate(e, sd, _oracle_cfg(cfg), new)
        agree.append(next((i for i, (a, r) in enumerate(zip(got, ref_toks)) if a != r), new))
        st = od.State(cfg.num_hidden_layers)
        lg = od.logits(od.forward_hidden(e, sd, _oracle_cfg(cfg), st)[-1:], sd)[0].float()
        for tok in got:
            gap = float((lg.max() - lg[tok]) / lg.abs().max())
            worst_gap = max(worst_gap, gap)
            exact += int(int(lg.argmax()) == tok)
            lg = od.logits(od.forward_hidden(sd["model.embed_tokens.weight"][tok][None, :], sd, _oracle_cfg(cfg), st), sd)[0].float()
    record("generate_greedy_agreement", use_graph=use_graph, free_running_first_divergence=str(agree), steps=new,
           teacher_forced_exact=exact, teacher_forced_total=2 * new, worst_gap_rel=worst_gap)
    assert worst_gap < 1e-2            # every produced token is (within bf16 noise) the oracle's choice (measured on a B200: 6.0e-3)
    assert exact >= int(0.9 * 2 * new)


def test_generate_is_deterministic_and_pages_are_released():
    cfg, sd, model, proc = _setup(seed=3)
    ts1, _ = _demo_series()
    enc = proc(text=["x <ts><ts/> y"], timeseries=[ts1], return_tensors="pt")
    free0 = len(model.pool.free)
    a = model.generate(**enc, max_new_tokens=12, ignore_eos=True)
    b = model.generate(**enc, max_new_tokens=12, ignore_eos=True)
    assert torch.equal(a, b)
    assert len(model.pool.free) == free0
    # sampling path runs and respects the length contract
    s = model.generate(**enc, max_new_tokens=5, do_sample=True, temperature=0.7, top_p=0.9, ignore_eos=True, seed=1)
    assert s.shape[1] == enc["input_ids"].shape[1] + 5


def test_vllm_surface_request_shape():
    from chatts_b200.vllm_compat import LLM, SamplingParams
    cfg, sd, model, proc = _setup(seed=4)
    llm = LLM(model=model)
    ts1, ts2 = _demo_series()
    reqs = [{"prompt": "two: <ts><ts/> <ts><ts/>", "multi_modal_data": {"timeseries": [ts1, ts2.tolist()]}},
            {"prompt": "none"}]
    outs = llm.generate(reqs, SamplingParams(max_tokens=6, ignore_eos=True))
    assert len(outs) == 2 and all(len(o.outputs[0].token_ids) == 6 for o in outs)
    with pytest.raises(TypeError):
        llm.generate([{"prompt": "<ts><ts/>", "multi_modal_data": {"timeseries": ["bad"]}}])
    with pytest.raises(AssertionError):
        llm.generate([{"prompt": "<ts><ts/> <ts><ts/>", "multi_modal_data": {"timeseries": [ts1]}}])


def test_full_size_layer_shapes_property():
    """ChatTS-14B layer shapes (hidden 5120, 40/8 heads of 128, inter 13824) with 1 layer and a small vocab:
    decode through the CUDA graph must reproduce the prefill logits of the same tokens (KV-cache consistency)."""
    from chatts_b200 import ChatTSConfig
    from chatts_b200.model import ChatTSForCausalLM
    cfg = ChatTSConfig(num_hidden_layers=1, vocab_size=2048, ts_token_start_index=2000, eos_token_id=2047, pad_token_id=2046,
                       max_position_embeddings=4096)
    model = ChatTSForCausalLM.from_synthetic(cfg, seed=7, max_batch=4, max_seq_len=256)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 1900, (2, 70), generator=g)
    new = 6
    out = model.generate(input_ids=ids, max_new_tokens=new, ignore_eos=True)
    # teacher-forced check: prefill over prompt + generated tokens must predict the same next tokens
    full = out[:, : 70 + new - 1]
    lg = model.forward(full, None, None, logits_to_keep=0).logits
    for b in range(2):
        pred = lg[b][69:].float().argmax(-1).cpu().tolist()
        assert pred == out[b, 70:].tolist()
