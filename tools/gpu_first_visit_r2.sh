#!/bin/bash
# One box acquisition for the whole backlog of round 1 (everything that was written after the GPU budget ran out):
#   1. the regression suite of the validated path (tools/gpu_checks.sh: 130 parity tests, smoke, a short bench)
#   2. the pending training / sampling / executor / tcgen05-backward tests with --runxfail, bench_lora, ncu captures
#      (tools/gpu_train_checks.sh)
#   3. A/B lines for the opt-in variants that are meant to become defaults once green
# Logs under gpurun_out/.     gpurun --timeout 2400 -- 'bash tools/gpu_first_visit_r2.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "##### 1. regression"
RUN_BENCH=1 BENCH_STEPS=32 bash tools/gpu_checks.sh 2>&1 | tail -n 60
echo "##### 2. training path"
RUN_NCU=${RUN_NCU:-1} bash tools/gpu_train_checks.sh 2>&1 | tail -n 120
echo "##### 3. opt-in variants (A/B)"
ab() { echo "## $*"; env "$@" timeout 400 python tools/bench_lora.py --steps 4 --warmup 3 --layers 8 2>>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({'ms_per_step': round(d['ms_per_step'],2), 'tok_s': round(d['value']), 'frac': round(d['roofline']['frac'],3), 'loss': round(d['loss'],4)})
"; }
ab CTS_BASE=1
ab CTS_ATTN_BWD_TC5=1
ab CTS_WGRAD_MMA=1
ab CTS_ATTN_BWD_TC5=1 CTS_WGRAD_MMA=1
echo "## decode: native step executor / sampling kernel (tokens/s, b=32)"
for v in "CTS_BASE=1" "CTS_NATIVE_STEP=1" "CTS_DECODE_FUSED=1" "CTS_DECODE_FUSED=2"; do
  echo "## $v"; env $v timeout 400 python bench.py --steps 32 --warmup 3 --no-cpu-baseline --sweep-only 2>>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({'ms_per_step': round(d['ms_per_step'],3), 'launches_per_step': d.get('launches_per_step'), 'by_batch': {b: round(v['ms_per_step'],3) for b, v in d['by_batch'].items()}})
"; done
ls -la gpurun_out | tail -n 40
