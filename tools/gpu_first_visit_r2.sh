#!/bin/bash
# One box acquisition for the whole backlog of round 1 (everything that was written after the GPU budget ran out):
#   1. the pending tests with --runxfail (training / sampling / executor / tcgen05 backward / cluster-fused decode GEMMs / low-latency
#      all-reduce on one GPU), then bench_lora                                   (tools/gpu_train_checks.sh, ncu deferred to step 4)
#   2. A/B lines for the opt-in variants that are meant to become defaults once green
#   3. the regression suite of the validated path (tools/gpu_checks.sh: parity tests, smoke, a short bench)
#   4. ncu captures of the training kernels and of the fused decode GEMMs
# Most informative first: a timeout or a dead box costs the tail, not the head.
# Logs under gpurun_out/.     gpurun --timeout 2400 -- 'bash tools/gpu_first_visit_r2.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "##### 1. pending tests + config-5 bench"
RUN_NCU=0 bash tools/gpu_train_checks.sh 2>&1 | tail -n 120
echo "##### 2. opt-in variants (A/B)"
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
echo "##### 3. regression"
RUN_BENCH=1 BENCH_STEPS=32 bash tools/gpu_checks.sh 2>&1 | tail -n 60
echo "##### 4. ncu"
if grep -q "passed" gpurun_out/test_train.log 2>/dev/null && ! grep -q "failed" gpurun_out/test_train.log; then
  LORA="python tools/bench_lora.py --steps 1 --warmup 3 --layers 2"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_lora.csv $LORA > gpurun_out/ncu_launch_lora.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:attn_bwd_dq|attn_bwd_dkv|lora_wgrad|ce_loss_grad' -s 40 -c 8 \
     -o gpurun_out/prof_train_kernels -f $LORA > gpurun_out/ncu_train_kernels.log 2>&1
fi
if grep -q "passed" gpurun_out/test_zz_e_fused_decode.log 2>/dev/null && ! grep -q "failed" gpurun_out/test_zz_e_fused_decode.log; then
  CTS_DECODE_FUSED=2 PROFILE_TAG=fused2 LIST_SKIP=250 LIST_COUNT=260 bash tools/gpu_profile.sh 2>&1 | tail -n 20
fi
ls -la gpurun_out | tail -n 40
