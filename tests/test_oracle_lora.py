"""Pin oracle/lora.py (row A9): loss and adapter gradients against the installed transformers Qwen3ForCausalLM with
LoRA modules wrapped around its Linear layers (peft's published forward; peft itself is not installed) and
transformers' own label-shifted loss; the AdamW restatement against torch.optim.AdamW."""
import torch
import torch.nn as nn

from oracle import lora as ol

CFG = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
           head_dim=64, vocab_size=320, rms_norm_eps=1e-6, rope_theta=1e6, max_position_embeddings=512)


class _LoRALinear(nn.Module):
    """peft lora.Linear.forward: base(x) + lora_B(lora_A(x)) * scaling."""

    def __init__(self, base, a, b, scaling):
        super().__init__()
        self.base, self.scaling = base, scaling
        self.lora_A, self.lora_B = nn.Parameter(a.clone()), nn.Parameter(b.clone())
        for p in base.parameters():
            p.requires_grad_(False)

    def forward(self, x):
        return self.base(x) + nn.functional.linear(nn.functional.linear(x, self.lora_A), self.lora_B) * self.scaling


def _hf(qwen3):
    torch.manual_seed(0)
    kw = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
              vocab_size=320, rms_norm_eps=1e-6, rope_theta=1e6, max_position_embeddings=512, tie_word_embeddings=False,
              attn_implementation="eager")
    if qwen3:
        from transformers import Qwen3Config, Qwen3ForCausalLM
        m = Qwen3ForCausalLM(Qwen3Config(head_dim=64, attention_bias=False, **kw))
    else:
        from transformers import Qwen2Config, Qwen2ForCausalLM
        m = Qwen2ForCausalLM(Qwen2Config(**kw))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:
                p.uniform_(0.5, 1.5)
            elif n.endswith("bias"):
                p.normal_(0, 0.05)
    return m.eval()


def _check(qwen3):
    m = _hf(qwen3)
    w = {k: v.detach().clone() for k, v in m.state_dict().items()}
    r, alpha = 4, 8.0
    ad = ol.init_adapters(CFG, r, seed=3, b_std=0.05)
    for p in m.parameters():
        p.requires_grad_(False)
    wrapped = {}
    for l, layer in enumerate(m.model.layers):
        for proj in ol.TARGETS:
            mod = layer.self_attn if proj in ("q_proj", "k_proj", "v_proj", "o_proj") else layer.mlp
            base = f"model.layers.{l}.{'self_attn' if mod is layer.self_attn else 'mlp'}.{proj}"
            wl = _LoRALinear(getattr(mod, proj), ad[base + ".lora_A.weight"], ad[base + ".lora_B.weight"], alpha / r)
            setattr(mod, proj, wl)
            wrapped[base] = wl
    torch.manual_seed(5)
    lens = [19, 31]
    embeds = [torch.randn(n, 256) * 0.5 for n in lens]
    labels = []
    for n in lens:
        y = torch.randint(0, 320, (n,))
        y[: n // 2] = -100                                  # the "input" part of a record is not learnt
        labels.append(y)
    # transformers: one padded batch (right padding, masked), its own shifted loss (mean over counted positions)
    T = max(lens)
    xb = torch.zeros(2, T, 256)
    yb = torch.full((2, T), -100, dtype=torch.long)
    am = torch.zeros(2, T, dtype=torch.long)
    for i, n in enumerate(lens):
        xb[i, :n], yb[i, :n], am[i, :n] = embeds[i], labels[i], 1
    out = m(inputs_embeds=xb, attention_mask=am, labels=yb)
    out.loss.backward()
    loss, g = ol.grads(embeds, labels, w, ad, alpha / r, CFG)
    assert abs(loss - float(out.loss)) < 1e-4 * max(1.0, abs(loss))        # fp32 on a multi-threaded host: reduction order varies with load
    worst = 0.0
    for base, wl in wrapped.items():
        for nm, par in (("lora_A", wl.lora_A), ("lora_B", wl.lora_B)):
            ref = par.grad
            got = g[f"{base}.{nm}.weight"]
            assert ref.abs().max() > 0, base                 # every path carries gradient (B != 0)
            worst = max(worst, float((got - ref).abs().max() / ref.abs().max()))
    assert worst < 2e-4, worst


def test_lora_grads_match_transformers_qwen3():
    _check(True)


def test_lora_grads_match_transformers_qwen2():
    _check(False)


def test_adamw_restated():
    torch.manual_seed(0)
    p0, steps = torch.randn(7, 5), 4
    gs = [torch.randn(7, 5) for _ in range(steps)]
    par = nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([par], lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for s, g in enumerate(gs, 1):
        par.grad = g.clone()
        opt.step()
        p, m, v = ol.adamw_update(p, g, m, v, s, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
        assert torch.allclose(p, par.detach(), atol=1e-6, rtol=1e-5)


def test_zero_b_means_zero_grad_for_a():
    """peft's default init (B = 0): the first step moves only B."""
    m = _hf(True)
    w = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ad = ol.init_adapters(CFG, 4, seed=1)
    x = [torch.randn(12, 256) * 0.5]
    y = [torch.randint(0, 320, (12,))]
    _, g = ol.grads(x, y, w, ad, 2.0, CFG)
    assert all(float(v.abs().max()) == 0 for k, v in g.items() if "lora_A" in k)
    assert any(float(v.abs().max()) > 0 for k, v in g.items() if "lora_B" in k)
