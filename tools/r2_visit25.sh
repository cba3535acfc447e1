#!/bin/bash
# Round-2 visit 25 (one B200): W4A16 mma kernel with 128-K pipeline stages on one ring (tests, GEMM sweep, model side line).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -n 3
W4_TC5=0 timeout 900 python tools/bench_w4_gemm.py > gpurun_out/r2v25_w4_gemm.json 2> gpurun_out/r2v25_w4_gemm.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2v25_w4_gemm.json'))
for name, o in d.items():
    print(name, {k: ({kk: v[kk] for kk in ('suggested', 'best', 'best_us', 'suggested_us', 'best_packed_gbs', 'speedup_vs_bf16') if kk in v} if k.startswith('mma') else v['us']) for k, v in o.items()})
PY
tail -n 3 gpurun_out/r2v25_w4_gemm.err
timeout 900 python tools/bench_w4.py > gpurun_out/r2v25_w4_bench.json 2> gpurun_out/r2v25_w4_bench.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2v25_w4_bench.json; tail -n 3 gpurun_out/r2v25_w4_bench.err
