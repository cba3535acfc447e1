#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; : > gpurun_out/sweep2.txt
run() { echo "## $*" >> gpurun_out/sweep2.txt; env "$@" timeout 400 python bench.py --steps 48 --warmup 3 --no-cpu-baseline --sweep-only --only-batch 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({b:round(v['ms_per_step'],3) for b,v in d['by_batch'].items()})
" >> gpurun_out/sweep2.txt; }
run CTS_ATTN_SPLITS=1
run CTS_ATTN_SPLITS=2
run CTS_ATTN_SPLITS=3
run CTS_ATTN_SPLITS=5
cat gpurun_out/sweep2.txt
