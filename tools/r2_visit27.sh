#!/bin/bash
# Round-2 visit 27 (one B200): the default bench line with the w4a16 side block; the W4 test file.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v27_bench.json 2> gpurun_out/r2v27_bench.err ) 2>&1 | grep real; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2v27_bench.json'):
    if l.startswith('{'):
        d = json.loads(l); print({k: d.get(k) for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['value'], 'config4', (d.get('config4') or {}).get('tokens_per_s'), 'w4a16', json.dumps(d.get('w4a16'))[:1200])
PY
tail -n 3 gpurun_out/r2v27_bench.err
timeout 600 python -m pytest tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -n 2
