#!/bin/bash
# Round-2 visit 39 (one B200): the default bench line at the final HEAD.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v39_bench.json 2> gpurun_out/r2v39_bench.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2v39_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print({k: d.get(k) for k in ('value', 'ms_per_step')}, {b: round(v['ms_per_step'], 3) for b, v in d['by_batch'].items()}, 'e2e', d['e2e']['value'], 'attn', d['attention']['prefill']['achieved_tflops'], 'config4', (d.get('config4') or {}).get('tokens_per_s'), 'w4 b1', d['w4a16']['by_batch']['1']['speedup'], 'roof', d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['clocks'])
PY
tail -n 2 gpurun_out/r2v39_bench.err
