#!/usr/bin/env python
"""cts_gemm_w4 alone at the ChatTS-14B projection shapes: packed GB/s per launch (distinct weights per launch, >> L2) for several
split factors, next to the bf16 decode GEMM on the dequantised weight.  `--once` runs a few launches only (for ncu)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chatts_b200 import _cabi  # noqa: E402


def main():
    once = "--once" in sys.argv
    c = _cabi.get_context()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1)
    res = {}
    shapes = {"gate_up": (27648, 5120), "down": (5120, 13824), "qkv": (7168, 5120), "o": (5120, 5120)}
    T = 32
    for name, (n, k) in shapes.items():
        L = 2 if once else max(2, int(600e6 // (n * k // 2)))          # enough distinct weights to defeat the L2
        qw = [torch.randint(0, 256, (n, k // 2), generator=g, device=dev, dtype=torch.uint8) for _ in range(L)]
        sc = (torch.rand(n, k // 128, generator=g, device=dev) * 0.01 + 0.005).to(torch.bfloat16)
        zp = torch.randint(1, 17, (n, k // 128), generator=g, device=dev, dtype=torch.uint8)
        x = (torch.randn(T, k, generator=g, device=dev) * 0.5).to(torch.bfloat16)
        wd = [torch.randn(n, k, generator=g, device=dev).to(torch.bfloat16) for _ in range(min(L, 4))]
        out = {}
        for split in ([c.gemm_w4_suggest_split(n, k)] if once else sorted({1, 2, 3, 4, 5, 7, c.gemm_w4_suggest_split(n, k)})):
            if split > k // 64 // 2:
                continue
            blocks = -(-(k // 64) // split) + 1
            if blocks * 64 // 128 + 2 > 44:
                continue
            ws = torch.empty(split * T * n, device=dev, dtype=torch.float32)
            us = bench._event_timer(lambda i: c.gemm_w4(x, qw[i % L], sc, zp, 128, ws, split, t=T), 2 * L if not once else 2)
            byts = n * k / 2 + n * (k // 128) * 3
            out[f"w4_split{split}"] = {"us": round(us, 2), "packed_gbs": round(byts / (us * 1e-6) / 1e9, 1)}
        if not once:
            sp = c.suggest_split(n if name != "gate_up" else n // 2, k, T, name == "gate_up")
            ws = torch.empty(sp * T * n, device=dev, dtype=torch.float32)
            us = bench._event_timer(lambda i: c.gemm(x, wd[i % len(wd)], ws, epilogue=3, split_k=sp, t=T), 8)
            out[f"bf16_split{sp}"] = {"us": round(us, 2), "gbs": round(n * k * 2 / (us * 1e-6) / 1e9, 1)}
        res[name] = out
        del qw, wd
        torch.cuda.empty_cache()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
