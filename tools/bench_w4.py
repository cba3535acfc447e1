#!/usr/bin/env python
"""W4A16 side line of the decode benchmark (README.md:52,262-263: ChatTS-14B-GPTQ-Int4): the same workload as bench.py -- prompts of
8 series x 256 points, greedy decode at b = 1 / 8 / 32 -- on a ChatTS-14B whose seven projections per layer are 4-bit (synthetic codes /
scales / zero points at the real shapes, group size 128; embeddings, norms, lm_head and the TS encoder stay bf16), against the same
model decoding through its dequantised bf16 copy.  ms/step from CUDA events over graph replays; one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    from chatts_b200 import ChatTSConfig
    from chatts_b200.model import ChatTSForCausalLM
    steps = int(os.environ.get("W4_STEPS", "32"))
    cfg = ChatTSConfig.chatts_14b()
    if os.environ.get("W4_LAYERS"):
        cfg.num_hidden_layers = int(os.environ["W4_LAYERS"])
    model = ChatTSForCausalLM.from_synthetic(cfg, seed=1234, max_batch=32, max_seq_len=1024, page_size=64)
    model.quantize_w4_synthetic(group_size=128)
    w4 = model.w4
    hbm, _ = bench.peaks()
    out = {"steps": steps, "layers": cfg.num_hidden_layers, "group_size": 128, "by_batch": {}}

    def run(batch):
        enc = bench.make_batch(cfg, batch)
        ids_cpu, am_cpu, counts, lay = model._prepare_inputs(enc["input_ids"], enc["attention_mask"], enc["timeseries"])
        pts, held = model._alloc_pages(lay.lens, steps + 16)
        try:
            logits = model._prefill(lay, counts, enc["timeseries"], pts)
            st = model._decode_state(batch, steps + 16)
            lens32 = torch.from_numpy(lay.lens.astype(np.int32))
            st.page_table.copy_(torch.from_numpy(pts)); st.positions.copy_(lens32 - 1); st.seq_lens.copy_(lens32); st.step_ptr.zero_()
            model.ctx.greedy_advance(logits, batch, st.out_tokens, st.step_ptr, st.cur_ids, st.positions, st.seq_lens, st.slot_map, st.page_table, model.page_size)
            for _ in range(4):
                model._decode_step(st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model._decode_step(st)
            e1.record()
            torch.cuda.synchronize()
            toks = st.out_tokens[:, : int(st.step_ptr[0])].cpu().numpy().copy()
            return e0.elapsed_time(e1) / steps, toks
        finally:
            model.pool.release(held)

    for b in (1, 8, 32):
        model.w4, model._steps = w4, {}
        ms4, t4 = run(b)
        model.w4, model._steps = None, {}
        ms16, t16 = run(b)
        n = min(t4.shape[1], t16.shape[1])
        agree = [int(next((i for i in range(n) if t4[r, i] != t16[r, i]), n)) for r in range(b)]
        # bytes a step streams: 4-bit codes + scales/zeros of the projections, bf16 for the rest (norms, lm_head), KV cache
        L = cfg.num_hidden_layers
        per_layer = (cfg.hidden_size * (cfg.num_attention_heads + 2 * cfg.num_key_value_heads) * cfg.head_dim + cfg.hidden_size * cfg.num_attention_heads * cfg.head_dim + 3 * cfg.hidden_size * cfg.intermediate_size)
        w4_bytes = L * per_layer * (0.5 + 3.0 / 128) + 2 * cfg.hidden_size * cfg.vocab_size
        kv = b * 600 * L * 2 * cfg.num_key_value_heads * cfg.head_dim * 2
        out["by_batch"][str(b)] = {"w4_ms_per_step": ms4, "bf16_ms_per_step": ms16, "speedup": ms16 / ms4, "w4_tokens_per_s": b / (ms4 / 1e3),
                                   "w4_whole_step_hbm_frac": (w4_bytes + kv) / (ms4 / 1e3) / 1e9 / hbm, "min_greedy_agreement_of_%d" % n: int(min(agree))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
