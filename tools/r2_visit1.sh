#!/bin/bash
# Round-2 GPU visit 1 (one GPU): (a) reproduce the one hardware failure of round 1 (cts_decoder_step, Qwen2 + bias, eager) with
# CTS_DEBUG_SYNC=1; (b) A/B of every decode variant on the FULL 48-layer model at b = 1 / 8 / 32; (c) ncu launch lists of the
# default and fused-level-2 steps.  Logs under gpurun_out/r2v1_*.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "##### a. native step executor"
CTS_DEBUG_SYNC=1 timeout 300 python -m pytest tests/test_gpu_zz_a_native_step.py -q -m gpu --runxfail --no-header -p no:cacheprovider -x \
   > gpurun_out/r2v1_native.log 2>&1; echo "rc=$?"; tail -n 40 gpurun_out/r2v1_native.log
echo "##### b. decode variants, full model"
ab() { echo "## $*"; env "$@" timeout 500 python bench.py --steps 48 --warmup 3 --no-cpu-baseline --sweep-only --no-probe 2>>gpurun_out/r2v1_ab.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({'launches_per_step': d.get('launches_per_step'), 'by_batch': {b: (round(v['ms_per_step'],3), v['tokens_sha1']) for b, v in d['by_batch'].items()}, 'gu_us': round(d['roofline']['us_per_launch'],2), 'fused_gu': d['roofline'].get('fused_variant'), 'attn_dec_us': round(d['attention']['decode']['us_per_launch'],2)})
"; }
ab CTS_BASE=1
ab CTS_DECODE_FUSED=1
ab CTS_DECODE_FUSED=2
ab CTS_NATIVE_STEP=1
ab CTS_DECODE_CHAIN=1
echo "##### c. launch lists (eager step, b=32)"
KREG='regex:gemm_tn|gemm_decode_fused|attn_|reduce_|qkv_rope|embed_gather|greedy|sample_|rmsnorm|ts_|peer_'
for v in 0 2; do
  if [ $v = 0 ]; then SK=900; CN=440; else SK=700; CN=250; fi
  CTS_DECODE_FUSED=$v timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s $SK -c $CN --csv \
     --log-file gpurun_out/r2v1_launches_fused$v.csv python bench.py --steps 2 --warmup 3 --batch 32 --only-batch --no-cpu-baseline --no-graph --sweep-only --no-probe \
     > gpurun_out/r2v1_ncu_fused$v.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2v1_launches_fused$v.csv
done
tail -n 20 gpurun_out/r2v1_ab.err
