#!/bin/bash
# Round-2 GPU visit 7 (one GPU): per-phase timeline of the fused TS encoder; prefill attention after the K/V ring (tests, timing against
# the kernels installed on the box); repetition-penalty kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python tools/bench_ts_encoder.py --batches 1 --trace > gpurun_out/r2v7_ts_trace.log 2>&1; echo "rc=$?"; tail -n 14 gpurun_out/r2v7_ts_trace.log
timeout 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_zz_b_sampling.py tests/test_gpu_zz_c_train.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2v7_tests.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/r2v7_tests.log
timeout 600 python - > gpurun_out/r2v7_attn.log 2>&1 <<'PY'
import json, math, torch, sys
sys.path.insert(0, '.')
import bench
from chatts_b200 import _cabi
c = _cabi.get_context()
nh, nkv, d = 40, 8, 128
for (B, S) in ((32, 576), (8, 2464), (4, 4096)):
    T = B * S
    g = torch.Generator(device="cuda").manual_seed(7)
    q = (torch.randn(T, nh * d, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    k = (torch.randn(T, nkv * d, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    v = (torch.randn(T, nkv * d, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    o = torch.empty(T, nh * d, device="cuda", dtype=torch.bfloat16)
    cu = torch.arange(0, T + 1, S, dtype=torch.int32, device="cuda")
    sc = 1.0 / math.sqrt(d)
    us = bench._event_timer(lambda i: c.attn_prefill(q, k, v, cu, B, S, nh, nkv, d, sc, o), 8)
    flops = B * 4.0 * (S * S / 2.0) * d * nh
    print(json.dumps({"shape": [B, S], "ours_us": us, "ours_tflops": flops / (us * 1e-6) / 1e12,
                      "installed": bench.installed_attention(q, k, v, B, S, nh, nkv, d, sc, flops, bench._event_timer)}))
PY
echo "rc=$?"; cat gpurun_out/r2v7_attn.log | cut -c1-900
