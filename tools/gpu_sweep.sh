#!/bin/bash
# tuning sweep: decode step time vs GEMM knobs (each line: knobs + by_batch ms)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out; : > gpurun_out/sweep.txt
run() { echo "## $*" >> gpurun_out/sweep.txt; env "$@" timeout 400 python bench.py --steps 48 --warmup 3 --no-cpu-baseline --sweep-only 2>>gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({b:round(v['ms_per_step'],3) for b,v in d['by_batch'].items()}, 'gu_frac', d['roofline'] and round(d['roofline']['frac'],3), d['clocks'])
" >> gpurun_out/sweep.txt; }
run CTS_L2_PREFETCH_MB=0
run CTS_L2_PREFETCH_MB=0
run CTS_L2_PREFETCH_MB=8
run CTS_L2_PREFETCH_MB=24
run CTS_L2_PREFETCH_MB=0 CTS_DECODE_SMEM_KB=70
run CTS_L2_PREFETCH_MB=0 CTS_DECODE_SMEM_KB=200
cat gpurun_out/sweep.txt
