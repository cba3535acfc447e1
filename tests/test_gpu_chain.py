"""Persistent decode-chain kernel (one launch per layer between attention calls) against the oracle and against the
multi-kernel decode path."""
import numpy as np
import pytest
import torch

from oracle import decoder as od
from oracle import merge as om
from oracle import ts_encoder as ote
from tests.gpu_util import record

pytestmark = pytest.mark.gpu
DT = torch.bfloat16


def _mk(seed, use_chain, batch=8, **kw):
    from chatts_b200 import ChatTSConfig, ChatTSProcessor, SimpleTokenizer
    from chatts_b200.model import ChatTSForCausalLM
    from chatts_b200.weights import synthetic_state_dict
    cfg = ChatTSConfig.tiny(**kw)
    sd = synthetic_state_dict(cfg, seed=500 + seed, device="cpu", dtype=DT, std=0.05)
    model = ChatTSForCausalLM(cfg, sd, dtype=DT, max_batch=batch, max_seq_len=512, page_size=16, use_chain=use_chain)
    proc = ChatTSProcessor(SimpleTokenizer(cfg.ts_token_start_index, cfg.pad_token_id, cfg.eos_token_id), cfg)
    return cfg, sd, model, proc


def _teacher_forced_gap(cfg, sd, enc, ids, samples):
    ts_w = {k[len("ts_encoder."):]: v for k, v in sd.items() if k.startswith("ts_encoder.")}
    feats, pc = ote.forward(enc["timeseries"].to(DT), cfg.ts, ts_w)
    embeds = om.hf_merge(enc["input_ids"], enc["attention_mask"], sd["model.embed_tokens.weight"], feats, pc.tolist(),
                         cfg.ts_token_start_index)
    S = enc["input_ids"].shape[1]
    worst, exact, total = 0.0, 0, 0
    for b in samples:
        st = od.State(cfg.num_hidden_layers)
        lg = od.logits(od.forward_hidden(embeds[b], sd, cfg.to_dict(), st)[-1:], sd)[0].float()
        for tok in ids[b, S:].tolist():
            worst = max(worst, float((lg.max() - lg[tok]) / lg.abs().max()))
            exact += int(int(lg.argmax()) == tok)
            total += 1
            lg = od.logits(od.forward_hidden(sd["model.embed_tokens.weight"][tok][None], sd, cfg.to_dict(), st), sd)[0].float()
    return worst, exact, total


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("qwen3", [False, True])
def test_chain_decode_matches_oracle(graph, qwen3):
    kw = dict(qk_norm=True, attention_bias=False) if qwen3 else {}
    cfg, sd, model, proc = _mk(1, True, **kw)
    model.use_cuda_graph = graph
    assert model._chain_ok(2)
    x = np.arange(200)
    enc = proc(text=["chain <ts><ts/> test", "second prompt without series but longer"], timeseries=[np.sin(x / 8.0) * 4], padding=True,
               return_tensors="pt")
    ids = model.generate(**enc, max_new_tokens=24, ignore_eos=True)
    worst, exact, total = _teacher_forced_gap(cfg, sd, enc, ids, [0, 1])
    record("chain_decode", graph=graph, qwen3=qwen3, worst_gap_rel=worst, exact=exact, total=total)
    assert worst < 2e-2 and exact >= int(0.9 * total)


def test_chain_batch32_equals_multikernel_path():
    """T = 32 (BN = 32 tiles): chain path and multi-kernel path run the same arithmetic -> (nearly) identical tokens."""
    rng = np.random.default_rng(0)
    prompts = [f"p{b} " + "x" * int(rng.integers(3, 30)) for b in range(32)]
    outs = []
    for use_chain in (False, True):
        cfg, sd, model, proc = _mk(2, use_chain, batch=32)
        enc = proc(text=prompts, timeseries=[], padding=True, return_tensors="pt")
        outs.append(model.generate(**enc, max_new_tokens=12, ignore_eos=True))
    same = float((outs[0] == outs[1]).float().mean())
    record("chain_vs_multikernel_b32", identical_fraction=same)
    assert same > 0.97


def test_chain_full_size_layer_shapes():
    """ChatTS-14B layer shapes through the chain kernel: decode must reproduce the prefill predictions (KV consistency)."""
    from chatts_b200 import ChatTSConfig
    from chatts_b200.model import ChatTSForCausalLM
    cfg = ChatTSConfig(num_hidden_layers=2, vocab_size=2048, ts_token_start_index=2000, eos_token_id=2047, pad_token_id=2046,
                       max_position_embeddings=4096)
    model = ChatTSForCausalLM.from_synthetic(cfg, seed=7, max_batch=4, max_seq_len=256, use_chain=True)
    assert model._chain_ok(2)
    ids = torch.randint(0, 1900, (2, 70), generator=torch.Generator().manual_seed(0))
    new = 6
    out = model.generate(input_ids=ids, max_new_tokens=new, ignore_eos=True)
    full = out[:, : 70 + new - 1]
    lg = model.forward(full, None, None, logits_to_keep=0).logits
    for b in range(2):
        pl = lg[b][69:].float()
        for i, tok in enumerate(out[b, 70:].tolist()):
            gap = float((pl[i].max() - pl[i][tok]) / pl[i].abs().max())
            assert gap < 2e-2, (b, i, gap)
