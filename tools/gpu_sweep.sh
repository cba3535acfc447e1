#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; : > gpurun_out/sweep3.txt
run() { echo "## $*" >> gpurun_out/sweep3.txt; env "$@" timeout 400 python bench.py --steps 48 --warmup 3 --no-cpu-baseline --sweep-only 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({b:round(v['ms_per_step'],3) for b,v in d['by_batch'].items()})
" >> gpurun_out/sweep3.txt; }
run CTS_NORM_CLUSTER=8
run CTS_NORM_CLUSTER=4
run CTS_NORM_CLUSTER=2
run CTS_NORM_CLUSTER=1
cat gpurun_out/sweep3.txt
