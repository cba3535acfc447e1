#!/bin/bash
# Round-2 visit 14 (one B200): the trimmed W4A16 mma kernel again (tests, GEMM sweep, model side line) and the default bench line with the
# new BASELINE configs[3] side block (second model instance, batch 8, 30 series x 512 points).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -n 3
W4_TC5=0 timeout 900 python tools/bench_w4_gemm.py > gpurun_out/r2v14_w4_gemm.json 2> gpurun_out/r2v14_w4_gemm.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2v14_w4_gemm.json'))
for name, o in d.items():
    print(name, {k: ({kk: v[kk] for kk in ('suggested', 'best', 'best_us', 'suggested_us', 'best_packed_gbs', 'speedup_vs_bf16') if kk in v} if k.startswith('mma') else v) for k, v in o.items()})
PY
tail -n 3 gpurun_out/r2v14_w4_gemm.err
timeout 900 python tools/bench_w4.py > gpurun_out/r2v14_w4_bench.json 2> gpurun_out/r2v14_w4_bench.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2v14_w4_bench.json; tail -n 3 gpurun_out/r2v14_w4_bench.err
echo "##### bench (with config4)"
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v14_bench.json 2> gpurun_out/r2v14_bench.err ) 2>&1 | grep real; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2v14_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print({k: d.get(k) for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'config4', d.get('config4'))
PY
tail -n 3 gpurun_out/r2v14_bench.err
