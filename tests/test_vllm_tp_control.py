"""Control plane of vllm_compat.LLM(tensor_parallel_size=k) when it has to SPAWN its ranks (demo/demo_vllm.py:30 passes the argument
without torchrun): a real worker process over gloo mirrors every generate() call the driver broadcasts and stops on request.  The
data plane (shards, peer-memory all-reduce) is covered on GPUs by tools/tp_check.py and bench.py's tp_parity gate."""
import json
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_spawned_rank_mirrors_generate_calls_and_stops(tmp_path):
    from chatts_b200 import vllm_compat as vc
    from tests.tp_worker_stub import make_recorder
    log = str(tmp_path / "calls.jsonl")
    port = vc._free_port()
    ctx = mp.get_context("spawn")
    pr = ctx.Process(target=vc._tp_worker, args=(1, 2, port, "gloo", make_recorder, {"path": log}), daemon=True)
    pr.start()
    os.environ.update(RANK="0", LOCAL_RANK="0", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=0, world_size=2)
    try:
        sp = vc.SamplingParams(max_tokens=7, temperature=0.5)
        dist.broadcast_object_list([("generate", [{"prompt": "a <ts><ts/>", "multi_modal_data": {"timeseries": [[1.0, 2.0]]}}], sp)], src=0)
        dist.broadcast_object_list([("generate", [{"prompt": "b"}, {"prompt": "c"}], vc.SamplingParams(max_tokens=3))], src=0)
        dist.broadcast_object_list([("stop",)], src=0)
        pr.join(timeout=60)
        assert pr.exitcode == 0
    finally:
        dist.destroy_process_group()
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
            os.environ.pop(k, None)
    calls = [json.loads(l) for l in open(log)]
    assert calls == [{"prompts": ["a <ts><ts/>"], "max_tokens": 7, "temperature": 0.5}, {"prompts": ["b", "c"], "max_tokens": 3, "temperature": 0.0}]


def test_tensor_parallel_size_mismatch_and_unspawnable_model_raise(cabi_double):
    import pytest
    from chatts_b200 import vllm_compat as vc
    from tests.test_host_model import _build
    cfg, sd, model, proc = _build(cabi_double)
    with pytest.raises(ValueError):                      # a constructed model cannot be shipped to spawned ranks
        vc.LLM(model=model, tensor_parallel_size=2)
    assert vc.LLM(model=model, tensor_parallel_size=1).model is model
