"""N>1 host logic on CPU: the tensor-parallel shard plan recombines to the unsharded result, exercised with two
real processes over gloo (world_size 2): column-split GEMMs are concatenated, row-split GEMMs are all-reduced."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chatts_b200 import ChatTSConfig
from chatts_b200.weights import shard_tensor, synthetic_state_dict


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = ChatTSConfig.tiny()
    sd = synthetic_state_dict(cfg, seed=3, device="cpu", dtype=torch.float32)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, cfg.hidden_size, generator=g)
    p = "model.layers.0."
    sh = lambda n: shard_tensor(p + n, sd[p + n], cfg, rank, world)
    # attention block: column-split q/k/v (whole heads per rank), row-split o_proj, all-reduce
    q = x @ sh("self_attn.q_proj.weight").T + sh("self_attn.q_proj.bias")
    full_q = x @ sd[p + "self_attn.q_proj.weight"].T + sd[p + "self_attn.q_proj.bias"]
    per = cfg.num_attention_heads // world * cfg.head_dim
    assert torch.allclose(q, full_q[:, rank * per:(rank + 1) * per], atol=1e-5)
    o_part = q @ sh("self_attn.o_proj.weight").T
    dist.all_reduce(o_part)
    assert torch.allclose(o_part, full_q @ sd[p + "self_attn.o_proj.weight"].T, atol=1e-4)
    # MLP block
    gate, up = x @ sh("mlp.gate_proj.weight").T, x @ sh("mlp.up_proj.weight").T
    act = torch.nn.functional.silu(gate) * up
    d_part = act @ sh("mlp.down_proj.weight").T
    dist.all_reduce(d_part)
    full = (torch.nn.functional.silu(x @ sd[p + "mlp.gate_proj.weight"].T) * (x @ sd[p + "mlp.up_proj.weight"].T)) @ sd[p + "mlp.down_proj.weight"].T
    assert torch.allclose(d_part, full, atol=1e-4)
    # vocab-parallel lm_head + all-gather
    lg = x @ shard_tensor("lm_head.weight", sd["lm_head.weight"], cfg, rank, world).T
    parts = [torch.empty_like(lg) for _ in range(world)]
    dist.all_gather(parts, lg)
    assert torch.allclose(torch.cat(parts, -1), x @ sd["lm_head.weight"].T, atol=1e-4)
    # kv heads are divided, never replicated
    assert sh("self_attn.k_proj.weight").shape[0] == cfg.num_key_value_heads // world * cfg.head_dim
    ret[rank] = True
    dist.destroy_process_group()


def test_shard_plan_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_shard_divisibility_is_checked():
    cfg = ChatTSConfig.tiny()
    sd = synthetic_state_dict(cfg, seed=3, device="cpu", dtype=torch.float32)
    with pytest.raises(AssertionError):
        shard_tensor("model.layers.0.self_attn.k_proj.weight", sd["model.layers.0.self_attn.k_proj.weight"], cfg, 0, 4)
