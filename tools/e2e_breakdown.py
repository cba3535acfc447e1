#!/usr/bin/env python
"""Where the `e2e` second goes: bench.py's public call -- model.generate(**pinned_host_tensors, max_new_tokens=K, sync_every=1) on the
metric batch -- with a device synchronisation and a wall-clock mark around each stage of generate() (input preparation incl. the
patch-count kernels and the one host sync, page allocation, prefill incl. the TS encoder, decode-state set-up incl. graph capture on
first use, the K decode steps with their per-step D2H).  The synchronisations make the stages add up; the untimed call next to it is the
number bench.py reports.  One JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from chatts_b200 import ChatTSConfig
    from chatts_b200.model import ChatTSForCausalLM
    batch = int(os.environ.get("E2E_BATCH", "32"))
    new = int(os.environ.get("E2E_STEPS", "20"))
    cfg = ChatTSConfig.chatts_14b()
    if os.environ.get("E2E_LAYERS"):
        cfg.num_hidden_layers = int(os.environ["E2E_LAYERS"])
    model = ChatTSForCausalLM.from_synthetic(cfg, seed=1234, max_batch=32, max_seq_len=1024, page_size=64)
    enc = {k: v.pin_memory() for k, v in bench.make_batch(cfg, batch, seed=1).items()}
    model.generate(**enc, max_new_tokens=4, ignore_eos=True, sync_every=1)
    torch.cuda.synchronize()

    def plain():
        t0 = time.perf_counter()
        model.generate(**enc, max_new_tokens=new, ignore_eos=True, sync_every=1)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    untimed = [plain() for _ in range(3)]
    acc = {}

    def wrap(name):
        fn = getattr(model, name)

        def timed(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize()
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
            return r
        setattr(model, name, timed)

    for n in ("_prepare_inputs", "_alloc_pages", "_prefill", "_decode_state", "_decode_step"):
        wrap(n)
    ts_enc = model.ts_encoder
    if ts_enc is not None:
        f0 = ts_enc.encode

        def enc_timed(*a, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = f0(*a, **k)
            torch.cuda.synchronize()
            acc["ts_encoder.encode (inside _prefill)"] = acc.get("ts_encoder.encode (inside _prefill)", 0.0) + time.perf_counter() - t0
            return r
        ts_enc.encode = enc_timed
    t0 = time.perf_counter()
    model.generate(**enc, max_new_tokens=new, ignore_eos=True, sync_every=1)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    staged = sum(v for k, v in acc.items() if not k.startswith("ts_encoder"))
    print(json.dumps({"batch": batch, "new_tokens": new, "layers": cfg.num_hidden_layers, "untimed_call_s": untimed,
                      "tokens_per_s_untimed": batch * new / min(untimed), "staged_call_s": total,
                      "stages_s": {k: round(v, 5) for k, v in acc.items()}, "host_glue_s": round(total - staged, 5),
                      "prefill_positions": batch * 576,
                      "prefill_tflops": (batch * 576 * 2 * 13.2e9 * cfg.num_hidden_layers / 48) / max(acc.get("_prefill", 1e-9), 1e-9) / 1e12}))


if __name__ == "__main__":
    main()
