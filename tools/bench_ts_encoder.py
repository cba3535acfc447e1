#!/usr/bin/env python
"""TS encoder alone at the ChatTS-14B shapes (5 x 5120-wide layers, 212 MB of weights): the metric prompt (8 series x 256 points =
128 patch rows) and the benchmark batch, fused single-launch kernel vs the multi-launch path, CUDA-graph replay with the L2
flushed before every timed replay (the same measurement as bench.py's `ts_encoder` block).

    python tools/bench_ts_encoder.py [--batches 1 2 32]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 2, 32])
    ap.add_argument("--trace", action="store_true")
    a = ap.parse_args()
    from chatts_b200 import ChatTSConfig
    from chatts_b200.ts_encoder import TimeSeriesEmbedding
    from chatts_b200.weights import ts_encoder_shapes
    cfg = ChatTSConfig.chatts_14b()
    g = torch.Generator(device="cuda").manual_seed(5)
    w = {k: (torch.randn(s, generator=g, device="cuda") * 0.02).to(torch.bfloat16) for k, s in ts_encoder_shapes(cfg).items()}
    tse = TimeSeriesEmbedding(cfg.ts, w, dtype=torch.bfloat16)
    hbm_peak, _ = bench.peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for nb in a.batches:
        x = bench.make_batch(cfg, nb, seed=2)["timeseries"].to("cuda", torch.bfloat16)
        counts = tse.patch_counts(x)
        host = torch.stack([counts[1], counts[2]]).cpu()
        hc = (host[0], host[1])
        res = {}
        for fused in (True, False):
            tse.use_fused = fused
            l0 = tse.ctx.launches
            feats, pc = tse.encode(x, counts=counts, host_counts=hc)
            n_launch = tse.ctx.launches - l0
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                feats, pc = tse.encode(x, counts=counts, host_counts=hc)
            tot, reps = 0.0, 10
            for it in range(reps + 2):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gr.replay()
                e1.record()
                torch.cuda.synchronize()
                if it >= 2:
                    tot += e0.elapsed_time(e1)
            us = tot * 1e3 / reps
            rows = int(feats.shape[0])
            H, in0, nl = tse.hidden_size, tse.input_size, tse.num_layers
            alg = 2 * (in0 * H + (nl - 1) * H * H + nl * H) + x.numel() * 2 + rows * in0 * 2 * 2 + rows * H * 2 * (2 * nl - 1)
            res["fused" if fused else "multi_launch"] = {"us": us, "launches": n_launch, "rows": rows, "achieved_gbs": alg / (us * 1e-6) / 1e9,
                                                          "hbm_frac": alg / (us * 1e-6) / 1e9 / hbm_peak, "sum": float(feats.float().abs().sum())}
            del gr
        out[f"b{nb}"] = res
    print(json.dumps(out))
    if a.trace:
        # per-phase timeline of ONE fused launch (csrc/trace.cuh marks: 1 wait released, 4 layer start (after a grid barrier), 5 last
        # load requested, 6 accumulator complete, 7 cluster synchronised, 8 tail done, 9 second cluster sync, 3 exit)
        import numpy as np
        x = bench.make_batch(cfg, 1, seed=2)["timeseries"].to("cuda", torch.bfloat16)
        counts = tse.patch_counts(x)
        host = torch.stack([counts[1], counts[2]]).cpu()
        tse.use_fused = True
        tse.encode(x, counts=counts, host_counts=(host[0], host[1]))
        torch.cuda.synchronize()
        tse.ctx.trace_begin(1 << 18)
        tse.encode(x, counts=counts, host_counts=(host[0], host[1]))
        rec = tse.ctx.trace_end()
        np.save(os.path.join(ROOT, "gpurun_out", "trace_ts_fused.npy"), rec)
        tag = rec[:, 0].astype(np.uint64)
        ph = ((tag >> np.uint64(52)) & np.uint64(0xf)).astype(int)
        t = rec[:, 1].astype(np.int64)
        t0 = t.min()
        gz = int((tag[0] >> np.uint64(32)) & np.uint64(0xff))
        print(f"fused launch traced: {rec.shape[0]} records, split {gz}, span {(t.max() - t0) / 1e3:.1f} us")
        for code in (0, 1, 4, 5, 6, 7, 8, 9, 3):
            tt = np.sort(t[ph == code]) - t0
            if tt.size == 0:
                continue
            # the marks of one phase come in bursts (one per layer): show first / median / last of every burst
            order = np.argsort(tt)
            n_b = max(1, round(tt.size / max(1, (ph == 0).sum())))
            chunks = np.array_split(tt, n_b) if tt.size >= n_b else [tt]
            print(f"  phase {code}: " + " | ".join(f"{c[0] / 1e3:6.1f}..{np.median(c) / 1e3:6.1f}..{c[-1] / 1e3:6.1f}" for c in chunks))


if __name__ == "__main__":
    main()
