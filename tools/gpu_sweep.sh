#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; : > gpurun_out/sweep.txt
run() { echo "## $*" >> gpurun_out/sweep.txt; env "$@" timeout 400 python bench.py --steps 48 --warmup 3 --no-cpu-baseline --sweep-only 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({b:round(v['ms_per_step'],3) for b,v in d['by_batch'].items()}, 'gu_frac', d['roofline'] and round(d['roofline']['frac'],3))
" >> gpurun_out/sweep.txt; }
run CTS_DECODE_SMEM_KB=75
run CTS_DECODE_SMEM_KB=75 CTS_SPLITS=7,11,2,11
run CTS_DECODE_SMEM_KB=75 CTS_SPLITS=7,10,2,10
run CTS_DECODE_SMEM_KB=75 CTS_SPLITS=5,7,1,7
run CTS_DECODE_SMEM_KB=56 CTS_SPLITS=10,14,2,14
run CTS_DECODE_SMEM_KB=56 CTS_SPLITS=7,11,2,11
run CTS_DECODE_SMEM_KB=62 CTS_SPLITS=7,11,2,11
run CTS_DECODE_SMEM_KB=75 CTS_SPLITS=2,3,2,3
cat gpurun_out/sweep.txt
