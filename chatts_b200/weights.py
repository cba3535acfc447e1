"""Weights: HF checkpoint names, a seeded synthetic factory (no checkpoint exists offline),
safetensors loading, and the tensor-parallel shard plan.

Names follow the checkpoint / the reference's mapper (chatts/vllm/chatts_vllm.py:454-470,612-625):
  model.embed_tokens.weight, model.layers.{i}.self_attn.{q,k,v}_proj.{weight,bias}, ...o_proj.weight,
  model.layers.{i}.mlp.{gate,up,down}_proj.weight, ...{input,post_attention}_layernorm.weight,
  model.norm.weight, lm_head.weight (absent when tied, :619-623),
  ts_encoder.mlp.{0,2,..}.{weight,bias}, ts_encoder.position_embedding.weight.
"""
import glob
import json
import os

import torch


def ts_encoder_shapes(cfg):
    ts = cfg.ts
    shapes = {}
    in_size = cfg.ts_input_size()
    for li in range(int(ts["num_layers"])):
        shapes[f"ts_encoder.mlp.{2 * li}.weight"] = (ts["hidden_size"], in_size)
        shapes[f"ts_encoder.mlp.{2 * li}.bias"] = (ts["hidden_size"],)
        in_size = ts["hidden_size"]
    if ts.get("use_position_embedding", False):
        shapes["ts_encoder.position_embedding.weight"] = (ts["max_sequence_length"] + 1, ts.get("embedding_dim", 16))
    return shapes


def decoder_shapes(cfg, layers=None):
    H, I, d = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim
    nh, nkv = cfg.num_attention_heads, cfg.num_key_value_heads
    shapes = {"model.embed_tokens.weight": (cfg.vocab_size, H), "model.norm.weight": (H,)}
    if not cfg.tie_word_embeddings:
        shapes["lm_head.weight"] = (cfg.vocab_size, H)
    for l in range(cfg.num_hidden_layers if layers is None else layers):
        p = f"model.layers.{l}."
        shapes[p + "self_attn.q_proj.weight"] = (nh * d, H)
        shapes[p + "self_attn.k_proj.weight"] = (nkv * d, H)
        shapes[p + "self_attn.v_proj.weight"] = (nkv * d, H)
        if getattr(cfg, "attention_bias", True):
            shapes[p + "self_attn.q_proj.bias"] = (nh * d,)
            shapes[p + "self_attn.k_proj.bias"] = (nkv * d,)
            shapes[p + "self_attn.v_proj.bias"] = (nkv * d,)
        if getattr(cfg, "qk_norm", False):
            shapes[p + "self_attn.q_norm.weight"] = (d,)
            shapes[p + "self_attn.k_norm.weight"] = (d,)
        shapes[p + "self_attn.o_proj.weight"] = (H, nh * d)
        shapes[p + "mlp.gate_proj.weight"] = (I, H)
        shapes[p + "mlp.up_proj.weight"] = (I, H)
        shapes[p + "mlp.down_proj.weight"] = (H, I)
        shapes[p + "input_layernorm.weight"] = (H,)
        shapes[p + "post_attention_layernorm.weight"] = (H,)
    return shapes


def all_shapes(cfg):
    s = decoder_shapes(cfg)
    s.update(ts_encoder_shapes(cfg))
    return s


def synthetic_state_dict(cfg, seed=1234, device="cpu", dtype=torch.bfloat16, std=0.02, names=None):
    """Seeded random weights at the config's shapes: N(0, std^2) for linears / embeddings /
    biases, U(0.5, 1.5) for norm weights (ones would hide a missing multiply).  Generated in fp32 on `device`
    with a torch.Generator, then cast -- so a CPU call gives the oracle and the GPU model identical values."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = {}
    for name, shape in all_shapes(cfg).items():
        if names is not None and name not in names:
            # still advance the generator identically? No: names-filtered dicts are only used for benchmarks.
            continue
        if name.endswith(("layernorm.weight", "q_norm.weight", "k_norm.weight")) or name == "model.norm.weight":
            t = torch.rand(shape, generator=g, device=device, dtype=torch.float32) + 0.5
        else:
            t = torch.randn(shape, generator=g, device=device, dtype=torch.float32) * std
        out[name] = t.to(dtype)
    return out


def load_checkpoint(path, device="cpu", dtype=None):
    """Read an HF checkpoint directory (config.json + *.safetensors [+ index]) into a name->tensor dict."""
    try:
        from safetensors import safe_open
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("safetensors is required to load a checkpoint") from e
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    sd = {}
    for f in files:
        with safe_open(f, framework="pt", device=str(device)) as sf:
            for k in sf.keys():
                t = sf.get_tensor(k)
                sd[k] = t.to(dtype) if dtype is not None and t.is_floating_point() else t
    return sd


# --------------------------------------------------------------------------------------------------
# tensor parallel shard plan (Megatron style): column split of QKV / gate / up / lm_head,
# row split of o_proj / down_proj, kv heads divided across ranks (nkv % tp == 0), everything else replicated.
# --------------------------------------------------------------------------------------------------
def shard_range(total, rank, size):
    assert total % size == 0, f"{total} not divisible by tensor-parallel size {size}"
    per = total // size
    return rank * per, (rank + 1) * per


def shard_tensor(name, t, cfg, rank, size):
    if size == 1:
        return t
    d = cfg.head_dim
    if name.endswith(("q_proj.weight", "q_proj.bias")):
        a, b = shard_range(cfg.num_attention_heads, rank, size)
        return t[a * d:b * d]
    if name.endswith(("k_proj.weight", "k_proj.bias", "v_proj.weight", "v_proj.bias")):
        a, b = shard_range(cfg.num_key_value_heads, rank, size)
        return t[a * d:b * d]
    if name.endswith("o_proj.weight"):
        a, b = shard_range(cfg.num_attention_heads, rank, size)
        return t[:, a * d:b * d]
    if name.endswith(("gate_proj.weight", "up_proj.weight")):
        a, b = shard_range(cfg.intermediate_size, rank, size)
        return t[a:b]
    if name.endswith("down_proj.weight"):
        a, b = shard_range(cfg.intermediate_size, rank, size)
        return t[:, a:b]
    if name == "lm_head.weight":
        a, b = shard_range(cfg.vocab_size, rank, size)
        return t[a:b]
    return t
