"""Small host-side units (no GPU): config parsing, weight shapes, page pool, split heuristics are GPU-side."""
import json

import pytest
import torch

from chatts_b200 import ChatTSConfig
from chatts_b200.weights import all_shapes, decoder_shapes, shard_tensor, synthetic_state_dict, ts_encoder_shapes


def test_14b_layer_parameter_count_matches_survey():
    """275 268 608 parameters per Qwen2.5-14B decoder layer; TS encoder ~106 M."""
    cfg = ChatTSConfig.chatts_14b()
    per_layer = sum(int(torch.Size(s).numel()) for n, s in decoder_shapes(cfg, layers=1).items() if n.startswith("model.layers.0."))
    assert per_layer == 275_268_608
    ts = sum(int(torch.Size(s).numel()) for s in ts_encoder_shapes(cfg).values())
    assert 105_000_000 < ts < 107_000_000
    assert cfg.ts_input_size() == 16 + 16 * 16 and cfg.ts_mode() == 1
    assert cfg.ts_token_end_index == cfg.ts_token_start_index + 1          # chatts_vllm.py:441


def test_config_from_checkpoint_style_json(tmp_path):
    d = dict(architectures=["Qwen2TSForCausalLM"], model_type="chatts", hidden_size=512, intermediate_size=1024, num_hidden_layers=3,
             num_attention_heads=8, num_key_value_heads=2, vocab_size=2048, rms_norm_eps=1e-6, rope_theta=1000000.0,
             max_position_embeddings=4096, torch_dtype="float16", tie_word_embeddings=False, eos_token_id=[7, 9],
             ts=dict(patch_size=16, num_layers=5, hidden_size=512, num_features=2, max_sequence_length=2048,
                     use_position_embedding=True, embedding_dim=16),
             ts_token_start_index=2000, unknown_field=123)
    (tmp_path / "config.json").write_text(json.dumps(d))
    cfg = ChatTSConfig.from_json(str(tmp_path))
    assert cfg.head_dim == 64 and cfg.eos_token_id == 7 and cfg.torch_dtype == "float16"
    assert cfg.ts["max_sequence_length"] == 2048 and not cfg.qk_norm and cfg.attention_bias
    d3 = dict(d, architectures=["Qwen3TSForCausalLM"], head_dim=128)
    cfg3 = ChatTSConfig.from_dict(d3)
    assert cfg3.qk_norm and not cfg3.attention_bias and cfg3.head_dim == 128       # chatts_vllm.py:633-668
    sd_names = set(all_shapes(cfg3))
    assert "model.layers.0.self_attn.q_norm.weight" in sd_names and "model.layers.0.self_attn.q_proj.bias" not in sd_names


def test_synthetic_weights_are_seeded_and_dtype_cast():
    cfg = ChatTSConfig.tiny()
    a = synthetic_state_dict(cfg, seed=5, device="cpu", dtype=torch.bfloat16)
    b = synthetic_state_dict(cfg, seed=5, device="cpu", dtype=torch.bfloat16)
    c = synthetic_state_dict(cfg, seed=6, device="cpu", dtype=torch.bfloat16)
    assert all(torch.equal(a[k], b[k]) for k in a) and any(not torch.equal(a[k], c[k]) for k in a)
    assert a["model.norm.weight"].min() >= 0.5 and a["model.layers.0.mlp.down_proj.weight"].dtype == torch.bfloat16


def test_tp8_shard_shapes_of_the_14b_plan():
    """configs[3]: TP=8 -> 5 q heads + 1 kv head per GPU, 1728 intermediate columns (27 x 64: interleavable), 19008 vocab rows."""
    cfg = ChatTSConfig.chatts_14b()
    shapes = decoder_shapes(cfg, layers=1)
    meta = {n: torch.empty(s, device="meta") for n, s in shapes.items()}
    p = "model.layers.0."
    sh = lambda n: tuple(shard_tensor(n, meta[n], cfg, 3, 8).shape)
    assert sh(p + "self_attn.q_proj.weight") == (5 * 128, 5120) and sh(p + "self_attn.k_proj.weight") == (128, 5120)
    assert sh(p + "self_attn.o_proj.weight") == (5120, 640) and sh(p + "mlp.gate_proj.weight") == (1728, 5120)
    assert sh(p + "mlp.down_proj.weight") == (5120, 1728) and 1728 % 64 == 0
    assert sh("lm_head.weight") == (19008, 5120) and sh("model.norm.weight") == (5120,)


def test_page_pool():
    from chatts_b200.model import PagePool
    pool = PagePool(8)
    a = pool.alloc(3)
    b = pool.alloc(5)
    assert sorted(a + b) == list(range(8))
    with pytest.raises(RuntimeError):
        pool.alloc(1)
    pool.release(a)
    assert sorted(pool.alloc(3)) == sorted(a)


def test_gptq_int4_dequantisation_roundtrip(tmp_path):
    """GPTQ-Int4 checkpoints (README.md:262-263) are dequantised at load time: pack -> unpack reproduces the 4-bit grid, for
    both zero conventions, with and without g_idx, and a packed checkpoint loads through load_checkpoint + dequantize_gptq."""
    import torch
    from safetensors.torch import save_file
    from chatts_b200.weights import dequantize_gptq, dequantize_gptq_linear, load_checkpoint, pack_gptq_linear

    g = torch.Generator().manual_seed(0)
    w = torch.randn(48, 256, generator=g) * 0.05                       # [out, in]
    for zo in (1, 0):
        qw, qz, sc, gi = pack_gptq_linear(w, group_size=128, zero_offset=zo)
        assert qw.shape == (32, 48) and qz.shape == (2, 6) and sc.shape == (2, 48) and qw.dtype == torch.int32
        d1 = dequantize_gptq_linear(qw, qz, sc, gi, 128, zo, torch.float32)
        d2 = dequantize_gptq_linear(qw, qz, sc, None, 128, zo, torch.float32)
        assert torch.equal(d1, d2) and d1.shape == w.shape
        step = sc.float().t().repeat_interleave(128, 1)                  # quantisation step of every element
        assert float(((d1 - w).abs() / step).max()) <= 0.5 + 2e-2       # round-to-nearest on the 4-bit grid (fp16 scales)
    qw, qz, sc, gi = pack_gptq_linear(w, 128, 1)
    save_file({"model.layers.0.mlp.up_proj.qweight": qw, "model.layers.0.mlp.up_proj.qzeros": qz, "model.layers.0.mlp.up_proj.scales": sc,
               "model.layers.0.mlp.up_proj.g_idx": gi, "model.norm.weight": torch.ones(4)}, str(tmp_path / "model.safetensors"))
    sd = dequantize_gptq(load_checkpoint(str(tmp_path)), {"bits": 4, "group_size": 128}, dtype=torch.bfloat16)
    assert set(sd) == {"model.layers.0.mlp.up_proj.weight", "model.norm.weight"}
    assert sd["model.layers.0.mlp.up_proj.weight"].dtype == torch.bfloat16 and sd["model.layers.0.mlp.up_proj.weight"].shape == (48, 256)
    import pytest
    with pytest.raises(ValueError):
        dequantize_gptq({"a.qweight": qw, "a.qzeros": qz, "a.scales": sc}, {"bits": 8})
