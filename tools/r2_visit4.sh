#!/bin/bash
# Round-2 GPU visit 4 (one GPU): regression suite after the latency fixes (batched partial loads, vector SwiGLU tail, attention
# page preload / pre-wait KV prefetch / batched split merge, next-GEMM L2 prefetch, step barrier), then A/B of the prefetch budget
# and new timelines.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
echo "##### a. GPU suite"
timeout 900 python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2v4_suite.log 2>&1; echo "rc=$?"; tail -n 15 gpurun_out/r2v4_suite.log
echo "##### b. decode A/B, full model"
ab() { echo "## $*"; env "$@" timeout 500 python bench.py --steps 48 --warmup 3 --no-cpu-baseline --sweep-only --no-probe 2>>gpurun_out/r2v4_ab.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({'launches_per_step': d.get('launches_per_step'), 'by_batch': {b: (round(v['ms_per_step'],3), v['tokens_sha1']) for b, v in d['by_batch'].items()}, 'gu_us': round(d['roofline']['us_per_launch'],2), 'attn_dec_us': round(d['attention']['decode']['us_per_launch'],2)})
"; }
ab CTS_NEXT_PREFETCH_MB=48
ab CTS_NEXT_PREFETCH_MB=0
ab CTS_NEXT_PREFETCH_MB=24
ab CTS_NEXT_PREFETCH_MB=80
echo "##### c. timelines"
for b in 32 1; do
  TRACE_TAG=_v4 timeout 300 python tools/trace_decode_step.py --batch $b --rows 30 > gpurun_out/r2v4_trace_b$b.log 2>&1; echo "rc=$?"; head -2 gpurun_out/r2v4_trace_b$b.log; tail -8 gpurun_out/r2v4_trace_b$b.log
done
tail -n 5 gpurun_out/r2v4_ab.err
