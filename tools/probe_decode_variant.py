#!/usr/bin/env python
"""Numeric guard for a decode variant that is NOT bit-identical to the default path (CTS_DECODE_FUSED=2: the RMSNorm statistic is summed in
another order).  Run by bench.py in a child process: two 4-layer models of the ChatTS-14B shape on the same weights -- the default path
and the variant -- are prefilled on the same batch and then decoded TEACHER-FORCED (both advance with the default path's greedy token),
and the next-token logits are compared at every step.  Prints one JSON line {"max_rel": ..., "finite": ..., "steps": ...}.

    python tools/probe_decode_variant.py --level 2 [--batch 32] [--steps 8] [--layers 4]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def compare(level, batches, steps, layers, cfg=None):
    from bench import make_batch
    from chatts_b200 import ChatTSConfig
    from chatts_b200.model import ChatTSForCausalLM
    cfg = cfg or ChatTSConfig.chatts_14b()
    cfg.num_hidden_layers = layers
    kw = dict(seed=1234, max_batch=max(batches), max_seq_len=1024, page_size=64, use_cuda_graph=False)
    models = [ChatTSForCausalLM.from_synthetic(cfg, use_fused_decode=0, **kw), ChatTSForCausalLM.from_synthetic(cfg, use_fused_decode=level, **kw)]
    worst, finite, n = 0.0, True, 0

    def rel(b, a):
        return float((b.float() - a.float()).abs().max() / a.float().abs().max().clamp_min(1e-6))

    for b in batches:
        enc = make_batch(cfg, b)
        runs = []
        for m in models:
            _, _, counts, lay = m._prepare_inputs(enc["input_ids"], enc["attention_mask"], enc["timeseries"])
            pts, held = m._alloc_pages(lay.lens, steps + 8)
            logits = m._prefill(lay, counts, enc["timeseries"], pts)
            st = m._decode_state(b, steps + 8)
            lens32 = torch.from_numpy(lay.lens.astype(np.int32))
            st.page_table.copy_(torch.from_numpy(pts))
            st.positions.copy_(lens32 - 1)
            st.seq_lens.copy_(lens32)
            st.step_ptr.zero_()
            runs.append((m, st, logits, held))
        try:
            la, lb = runs[0][2], runs[1][2]
            for _ in range(steps + 1):
                worst = max(worst, rel(lb, la))
                finite = finite and bool(torch.isfinite(lb.float()).all())
                n += 1
                forced = la.clone()                                  # both models continue with the DEFAULT path's greedy token
                for m, st, _, _ in runs:
                    m.ctx.greedy_advance(forced, b, st.out_tokens, st.step_ptr, st.cur_ids, st.positions, st.seq_lens, st.slot_map, st.page_table,
                                         m.page_size)
                for m, st, _, _ in runs:
                    m._decode_step(st, sample=False)
                la, lb = runs[0][1].full_logits, runs[1][1].full_logits
            worst = max(worst, rel(lb, la))
            finite = finite and bool(torch.isfinite(lb.float()).all())
        finally:
            for m, _, _, held in runs:
                m.pool.release(held)
    return {"level": level, "max_rel": worst, "finite": finite, "steps": n, "batches": list(batches)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--layers", type=int, default=4)
    a = ap.parse_args()
    print(json.dumps(compare(a.level, sorted({1, a.batch}), a.steps, a.layers)), flush=True)


if __name__ == "__main__":
    main()
