#!/bin/bash
# Round-2 visit 24 (one B200): the whole GPU suite, smoke, the default bench line and the attention micro-benchmark at the current HEAD.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x ) > gpurun_out/r2v24_pytest.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/r2v24_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2v24_smoke.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/r2v24_smoke.log | cut -c1-600
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v24_bench.json 2> gpurun_out/r2v24_bench.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2v24_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print({k: d.get(k) for k in ('value', 'ms_per_step')}, {b: round(v['ms_per_step'], 3) for b, v in d['by_batch'].items()}, 'e2e', d['e2e']['value'], 'attn', d['attention']['prefill']['achieved_tflops'], d['attention']['prefill']['us_per_launch'], 'config4', (d.get('config4') or {}).get('tokens_per_s'), 'roof', d['roofline']['frac'], d['roofline']['whole_step']['frac'])
PY
tail -n 3 gpurun_out/r2v24_bench.err
timeout 600 python tools/bench_attention.py > gpurun_out/r2v24_attention.json 2> gpurun_out/r2v24_attention.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2v24_attention.json
