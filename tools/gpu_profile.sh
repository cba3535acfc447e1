#!/bin/bash
# ncu evidence for the bench workload (B200_PROFILING.md recipe).  One GPU only.  Keep counts small: ncu costs ~0.3 s/launch.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
KREG='regex:gemm_tn|gemm_decode_fused|attn_|reduce_|qkv_rope|embed_gather|greedy|sample_|rmsnorm|ts_|peer_'
B=${PROFILE_BATCH:-32}${PROFILE_TAG:+_$PROFILE_TAG}     # PROFILE_TAG names a variant run, e.g. CTS_DECODE_FUSED=2 PROFILE_TAG=fused2
BENCH="python bench.py --steps 2 --warmup 3 --batch ${PROFILE_BATCH:-32} --only-batch --no-cpu-baseline --no-graph --sweep-only"
echo "=== launch list (one eager decode step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s ${LIST_SKIP:-440} -c ${LIST_COUNT:-450} --csv \
   --log-file gpurun_out/launches_b$B.csv $BENCH > gpurun_out/ncu_launch_b$B.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_b$B.csv
echo "=== full set: decode GEMMs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s ${GEMM_SKIP:-260} -c 5 \
   -o gpurun_out/prof_gemm_b$B -f $BENCH > gpurun_out/ncu_gemm_b$B.log 2>&1
echo "rc=$?"
if [ -n "${CTS_DECODE_FUSED:-}" ]; then
  echo "=== full set: cluster-fused decode GEMMs"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_decode_fused -s ${FUSED_SKIP:-120} -c 6 \
     -o gpurun_out/prof_gemm_fused_b$B -f $BENCH > gpurun_out/ncu_gemm_fused_b$B.log 2>&1
  echo "rc=$?"
fi
echo "=== full set: prefill GEMMs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s 12 -c 4 \
   -o gpurun_out/prof_gemm_prefill_b$B -f $BENCH > gpurun_out/ncu_gemm_prefill_b$B.log 2>&1
echo "rc=$?"
echo "=== full set: decode attention"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 10 -c 2 \
   -o gpurun_out/prof_attn_b$B -f $BENCH > gpurun_out/ncu_attn_b$B.log 2>&1
echo "rc=$?"
echo "=== full set: prefill attention (tcgen05)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill -s 4 -c 2 \
   -o gpurun_out/prof_attn_prefill_b$B -f $BENCH > gpurun_out/ncu_attn_prefill_b$B.log 2>&1
echo "rc=$?"
ls -la gpurun_out/ | grep -E "ncu-rep|csv"
