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


if __name__ == "__main__":
    main()
