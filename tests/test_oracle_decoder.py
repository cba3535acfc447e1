"""Pin oracle/decoder.py against the installed transformers Qwen2ForCausalLM (the third-party code the
reference calls, README.md:88 / chatts_vllm.py:483-488) -- prefill and KV-cached decode, fp32 and bf16."""
import pytest
import torch

from oracle import decoder as od

CFG = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
           num_key_value_heads=2, head_dim=64, vocab_size=320, rms_norm_eps=1e-6, rope_theta=1e6,
           max_position_embeddings=512)


def _hf_model(dtype):
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(0)
    c = Qwen2Config(hidden_size=CFG["hidden_size"], intermediate_size=CFG["intermediate_size"],
                    num_hidden_layers=CFG["num_hidden_layers"], num_attention_heads=CFG["num_attention_heads"],
                    num_key_value_heads=CFG["num_key_value_heads"], vocab_size=CFG["vocab_size"],
                    rms_norm_eps=CFG["rms_norm_eps"], rope_theta=CFG["rope_theta"],
                    max_position_embeddings=CFG["max_position_embeddings"], tie_word_embeddings=False,
                    attn_implementation="eager")
    m = Qwen2ForCausalLM(c).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
            if "layernorm" in n or n.endswith("norm.weight"):
                p.uniform_(0.5, 1.5)
    return m.to(dtype)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_prefill_and_decode_match_transformers(dtype, tol):
    m = _hf_model(dtype)
    w = {k: v.detach() for k, v in m.state_dict().items()}
    torch.manual_seed(1)
    T = 37
    x = (torch.randn(T, CFG["hidden_size"]) * 0.5).to(dtype)
    st = od.State(CFG["num_hidden_layers"])
    h = od.forward_hidden(x, w, CFG, st)
    lg = od.logits(h, w)
    with torch.no_grad():
        out = m(inputs_embeds=x[None], use_cache=True)
    ref = out.logits[0]
    scale = ref.float().abs().max()
    assert (lg.float() - ref.float()).abs().max() / scale < tol
    # three cached decode steps
    past = out.past_key_values
    for step in range(3):
        e = (torch.randn(1, CFG["hidden_size"]) * 0.5).to(dtype)
        h = od.forward_hidden(e, w, CFG, st)
        lg = od.logits(h, w)
        with torch.no_grad():
            o2 = m(inputs_embeds=e[None], past_key_values=past, use_cache=True)
        past = o2.past_key_values
        assert (lg.float() - o2.logits[0].float()).abs().max() / scale < tol, step


def test_greedy_generate_matches_transformers_fp32():
    m = _hf_model(torch.float32)
    w = {k: v.detach() for k, v in m.state_dict().items()}
    torch.manual_seed(2)
    x = torch.randn(11, CFG["hidden_size"]) * 0.5
    toks, _ = od.greedy_generate(x, w, CFG, 6)
    with torch.no_grad():
        ids = m.generate(inputs_embeds=x[None], max_new_tokens=6, do_sample=False)
    assert ids[0].tolist() == toks


def test_qwen3_variant_matches_transformers():
    """ChatTS-8B uses the Qwen3 decoder (chatts_vllm.py:633-668): per-head q/k RMSNorm, no qkv bias."""
    from transformers import Qwen3Config, Qwen3ForCausalLM
    torch.manual_seed(0)
    c = Qwen3Config(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                    head_dim=64, vocab_size=320, rms_norm_eps=1e-6, rope_theta=1e6, max_position_embeddings=512,
                    tie_word_embeddings=False, attention_bias=False, attn_implementation="eager")
    m = Qwen3ForCausalLM(c).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:
                p.uniform_(0.5, 1.5)
    w = {k: v.detach() for k, v in m.state_dict().items()}
    cfg = dict(CFG)
    x = torch.randn(23, 256) * 0.5
    st = od.State(2)
    lg = od.logits(od.forward_hidden(x, w, cfg, st), w)
    with torch.no_grad():
        out = m(inputs_embeds=x[None], use_cache=True)
    scale = out.logits.abs().max()
    assert (lg - out.logits[0]).abs().max() / scale < 2e-5
    e = torch.randn(1, 256) * 0.5
    lg = od.logits(od.forward_hidden(e, w, cfg, st), w)
    with torch.no_grad():
        o2 = m(inputs_embeds=e[None], past_key_values=out.past_key_values, use_cache=True)
    assert (lg - o2.logits[0]).abs().max() / scale < 2e-5
