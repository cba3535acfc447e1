"""Oracle: one LoRA fine-tune step of the decoder.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference repo holds no training code (README.md:216-218 points to the external
ChatTS-Training project; demo/demo_lora.ipynb cells 3-4 only LOAD an adapter with peft, which is not installed here).
This file restates the published algorithms the external recipe is made of, on top of the decoder oracle:

  * peft ``lora.Linear.forward``:  y = base(x) + lora_B(lora_A(x)) * (lora_alpha / r); A ~ kaiming_uniform(a=sqrt(5)),
    B = 0 at init; adapters on q/k/v/o/gate/up/down_proj, base weights frozen.
  * transformers ``ForCausalLMLoss``: logits upcast to fp32, shift by one, cross entropy with ignore_index = -100,
    mean over the counted positions (the prompt part of a record carries -100: only the ``output`` text is learnt,
    chatts/align/uts_template_qa.py:127-131 gives the record shape {input, output, timeseries}).
  * torch.optim.AdamW on the adapter tensors (fp32 master copies).

Gradients come from torch.autograd over the same arithmetic as oracle/decoder.py (every Linear result rounded to the
model dtype).  The time-series encoder and the embeddings are frozen: the step starts from the merged input
embeddings (oracle/merge.py), exactly like the decoder oracle.
"""
import math

import torch
import torch.nn.functional as F

from . import decoder as od

TARGETS = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


def _module_of(proj):
    return "self_attn" if proj in ("q_proj", "k_proj", "v_proj", "o_proj") else "mlp"


def adapter_shapes(cfg, r, targets=TARGETS):
    """name -> ((r, in), (out, r)) with peft's names: ...layers.{l}.{self_attn|mlp}.{proj}.lora_{A,B}.weight."""
    H, I = int(cfg["hidden_size"]), int(cfg["intermediate_size"])
    nh, nkv = int(cfg["num_attention_heads"]), int(cfg["num_key_value_heads"])
    d = int(cfg.get("head_dim") or H // nh)
    io = {"q_proj": (H, nh * d), "k_proj": (H, nkv * d), "v_proj": (H, nkv * d), "o_proj": (nh * d, H),
          "gate_proj": (H, I), "up_proj": (H, I), "down_proj": (I, H)}
    out = {}
    for l in range(int(cfg["num_hidden_layers"])):
        for p in targets:
            base = f"model.layers.{l}.{_module_of(p)}.{p}"
            out[base] = ((r, io[p][0]), (io[p][1], r))
    return out


def init_adapters(cfg, r, seed=0, targets=TARGETS, b_std=0.0):
    """peft init (A: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)); B: zeros).  ``b_std`` > 0 gives B a small
    normal init instead so that a single test step exercises every gradient path (with B = 0, dA is exactly 0)."""
    g = torch.Generator().manual_seed(seed)
    ad = {}
    for base, (sa, sb) in adapter_shapes(cfg, r, targets).items():
        bound = 1.0 / math.sqrt(sa[1])
        ad[base + ".lora_A.weight"] = (torch.rand(sa, generator=g) * 2 - 1) * bound
        ad[base + ".lora_B.weight"] = torch.randn(sb, generator=g) * b_std if b_std > 0 else torch.zeros(sb)
    return ad


def _lin(x, w, name, ad, scaling, bias=None):
    """peft lora.Linear.forward in the model dtype."""
    y = F.linear(x, w[name + ".weight"], bias)
    a = ad.get(name + ".lora_A.weight")
    if a is not None:
        b = ad[name + ".lora_B.weight"]
        y = y + F.linear(F.linear(x, a.to(x.dtype)), b.to(x.dtype)) * scaling
    return y


def forward_hidden(x, w, ad, scaling, cfg):
    """oracle/decoder.forward_hidden (modeling_qwen2.py:269-310 x n_layers) with LoRA linears, no KV state.  x [T,H]."""
    nh, nkv = int(cfg["num_attention_heads"]), int(cfg["num_key_value_heads"])
    d = int(cfg.get("head_dim") or cfg["hidden_size"] // nh)
    eps = float(cfg.get("rms_norm_eps", 1e-6))
    T = x.shape[0]
    cos, sin = od.rope_tables(cfg, T, x.dtype)
    h = x
    for l in range(int(cfg["num_hidden_layers"])):
        pre = f"model.layers.{l}."
        r = h
        a = od.rms_norm(h, w[pre + "input_layernorm.weight"], eps)
        q = _lin(a, w, pre + "self_attn.q_proj", ad, scaling, w.get(pre + "self_attn.q_proj.bias")).view(T, nh, d)
        k = _lin(a, w, pre + "self_attn.k_proj", ad, scaling, w.get(pre + "self_attn.k_proj.bias")).view(T, nkv, d)
        v = _lin(a, w, pre + "self_attn.v_proj", ad, scaling, w.get(pre + "self_attn.v_proj.bias")).view(T, nkv, d)
        if (pre + "self_attn.q_norm.weight") in w:
            q = od.rms_norm(q, w[pre + "self_attn.q_norm.weight"], eps)
            k = od.rms_norm(k, w[pre + "self_attn.k_norm.weight"], eps)
        q, k = od.apply_rope(q, k, cos, sin)
        o = od.attention(q, k, v, nh // nkv, 0)
        h = r + _lin(o, w, pre + "self_attn.o_proj", ad, scaling)
        r = h
        a = od.rms_norm(h, w[pre + "post_attention_layernorm.weight"], eps)
        g = F.silu(_lin(a, w, pre + "mlp.gate_proj", ad, scaling)) * _lin(a, w, pre + "mlp.up_proj", ad, scaling)
        h = r + _lin(g, w, pre + "mlp.down_proj", ad, scaling)
    return od.rms_norm(h, w["model.norm.weight"], eps)


def causal_lm_loss(embeds, labels, w, ad, scaling, cfg):
    """transformers ForCausalLMLoss over a list of samples: embeds[i] [T_i,H], labels[i] int64[T_i] (-100 = ignore).
    Returns (mean loss over counted positions, number of counted positions)."""
    total, count = 0.0, 0
    for x, y in zip(embeds, labels):
        hid = forward_hidden(x, w, ad, scaling, cfg)
        lg = od.logits(hid, w).float()
        tgt = y[1:]
        keep = tgt != -100
        n = int(keep.sum())
        if n:
            total = total + F.cross_entropy(lg[:-1][keep], tgt[keep], reduction="sum")
            count += n
    return total / max(count, 1), count


def grads(embeds, labels, w, ad, scaling, cfg):
    """(loss, {adapter name: grad fp32}) by autograd; adapters are the only leaves."""
    leaves = {k: v.clone().float().requires_grad_(True) for k, v in ad.items()}
    loss, _ = causal_lm_loss(embeds, labels, w, leaves, scaling, cfg)
    names = list(leaves)
    gs = torch.autograd.grad(loss, [leaves[n] for n in names], allow_unused=True)
    return float(loss.detach()), {n: (torch.zeros_like(leaves[n]) if g is None else g.detach()) for n, g in zip(names, gs)}


def adamw_update(p, g, m, v, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """torch.optim.AdamW (decoupled decay, bias-corrected), restated for one tensor; ``step`` counts from 1.
    Returns (p, m, v) updated copies."""
    b1, b2 = betas
    p = p * (1.0 - lr * weight_decay)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mhat = m / (1 - b1 ** step)
    vhat = v / (1 - b2 ** step)
    return p - lr * mhat / (vhat.sqrt() + eps), m, v
