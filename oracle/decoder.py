"""Oracle: Qwen2 decoder forward (prefill + KV-cached decode).  TEST INFRASTRUCTURE ONLY.

The reference calls this arithmetic through third-party code (chatts/vllm/chatts_vllm.py:483-488,595-598;
README.md:88): ``transformers==4.52.4`` (requirements.txt:7; 5.5.0 installed here)
``models/qwen2/modeling_qwen2.py``.  This file restates that published algorithm, citing the installed
file's lines, in the same dtype discipline (every nn.Linear / elementwise result rounded to the model
dtype; fp32 RMSNorm statistics; fp32 softmax; cos/sin computed in fp32 then cast).
PARITY UNPINNED against the reference repo itself (it holds no test/golden for the decoder);
pinned against the installed transformers Qwen2ForCausalLM in tests/test_oracle_decoder.py.

Weights: plain dict with HF names (``model.layers.{i}.self_attn.q_proj.weight`` ...).
One sample at a time (no padding): x [T, H].
"""

import torch
import torch.nn.functional as F


def rope_tables(cfg, n_pos, dtype):
    """modeling_qwen2.py:88-113: inv_freq fp32, freqs = pos*inv_freq, cat, cos/sin fp32 -> dtype."""
    d = int(cfg.get("head_dim") or cfg["hidden_size"] // cfg["num_attention_heads"])
    base = float(cfg.get("rope_theta", 10000.0))
    inv_freq = 1.0 / (base ** (torch.arange(0, d, 2, dtype=torch.int64).to(torch.float32) / d))
    pos = torch.arange(n_pos, dtype=torch.float32)
    freqs = (inv_freq[:, None] @ pos[None, :]).transpose(0, 1)          # [n_pos, d/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rms_norm(x, w, eps):
    """modeling_qwen2.py:258-263."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q, k, cos, sin):
    """modeling_qwen2.py:116-146.  q [T, nh, d], k [T, nkv, d], cos/sin [T, d]."""
    c, s = cos[:, None, :], sin[:, None, :]
    return (q * c) + (_rot_half(q) * s), (k * c) + (_rot_half(k) * s)


def attention(q, k_all, v_all, n_rep, q_pos0):
    """modeling_qwen2.py:161-184 (eager): causal GQA, scale 1/sqrt(d), fp32 softmax cast back to dtype.
    q [T, nh, d]; k_all/v_all [S, nkv, d] (S = q_pos0 + T)."""
    T, nh, d = q.shape
    S = k_all.shape[0]
    kk = k_all.repeat_interleave(n_rep, dim=1)      # [S, nh, d]
    vv = v_all.repeat_interleave(n_rep, dim=1)
    w = torch.einsum("thd,shd->hts", q, kk) * (d ** -0.5)
    qpos = torch.arange(T)[:, None] + q_pos0
    mask = torch.arange(S)[None, :] > qpos
    w = w.masked_fill(mask[None], float("-inf"))
    w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.einsum("hts,shd->thd", w, vv)
    return o.reshape(T, nh * d)


class State:
    """Per-sample KV cache: k[l], v[l] : [S, nkv, d]."""

    def __init__(self, n_layers):
        self.k = [None] * n_layers
        self.v = [None] * n_layers
        self.len = 0


def forward_hidden(x, w, cfg, state):
    """modeling_qwen2.py:269-310 x n_layers, then final norm.  x [T,H] -> [T,H]; appends to ``state``."""
    nh, nkv = int(cfg["num_attention_heads"]), int(cfg["num_key_value_heads"])
    d = int(cfg.get("head_dim") or cfg["hidden_size"] // nh)
    eps = float(cfg.get("rms_norm_eps", 1e-6))
    T = x.shape[0]
    p0 = state.len
    cos, sin = rope_tables(cfg, p0 + T, x.dtype)
    cos, sin = cos[p0:], sin[p0:]
    h = x
    for l in range(int(cfg["num_hidden_layers"])):
        pre = f"model.layers.{l}."
        r = h
        a = rms_norm(h, w[pre + "input_layernorm.weight"], eps)
        q = F.linear(a, w[pre + "self_attn.q_proj.weight"], w.get(pre + "self_attn.q_proj.bias")).view(T, nh, d)
        k = F.linear(a, w[pre + "self_attn.k_proj.weight"], w.get(pre + "self_attn.k_proj.bias")).view(T, nkv, d)
        v = F.linear(a, w[pre + "self_attn.v_proj.weight"], w.get(pre + "self_attn.v_proj.bias")).view(T, nkv, d)
        if (pre + "self_attn.q_norm.weight") in w:
            # Qwen3 (ChatTS-8B, chatts_vllm.py:633-668): per-head RMSNorm of q and k over head_dim before RoPE
            # (transformers models/qwen3/modeling_qwen3.py Qwen3Attention.forward)
            q = rms_norm(q, w[pre + "self_attn.q_norm.weight"], eps)
            k = rms_norm(k, w[pre + "self_attn.k_norm.weight"], eps)
        q, k = apply_rope(q, k, cos, sin)
        state.k[l] = k if state.k[l] is None else torch.cat([state.k[l], k], 0)
        state.v[l] = v if state.v[l] is None else torch.cat([state.v[l], v], 0)
        o = attention(q, state.k[l], state.v[l], nh // nkv, p0)
        h = r + F.linear(o, w[pre + "self_attn.o_proj.weight"])
        r = h
        a = rms_norm(h, w[pre + "post_attention_layernorm.weight"], eps)
        g = F.silu(F.linear(a, w[pre + "mlp.gate_proj.weight"])) * F.linear(a, w[pre + "mlp.up_proj.weight"])
        h = r + F.linear(g, w[pre + "mlp.down_proj.weight"])
    state.len = p0 + T
    return rms_norm(h, w["model.norm.weight"], eps)


def logits(hidden, w):
    lm = w["lm_head.weight"] if "lm_head.weight" in w else w["model.embed_tokens.weight"]   # tied (chatts_vllm.py:619-623)
    return F.linear(hidden, lm)


def greedy_generate(x_prompt, w, cfg, max_new_tokens):
    """HF generate(do_sample=False) restated for one sample: returns (new token ids, logits per step)."""
    st = State(int(cfg["num_hidden_layers"]))
    h = forward_hidden(x_prompt, w, cfg, st)
    out, lg_all = [], []
    lg = logits(h[-1:], w)
    for _ in range(max_new_tokens):
        lg_all.append(lg[0].clone())
        tok = int(torch.argmax(lg[0].float()))
        out.append(tok)
        e = w["model.embed_tokens.weight"][tok][None, :]
        h = forward_hidden(e, w, cfg, st)
        lg = logits(h, w)
    return out, torch.stack(lg_all)
