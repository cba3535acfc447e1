#!/bin/bash
# ncu evidence for one decode step of the bench workload (B200_PROFILING.md recipe).  One GPU only.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
KREG='regex:gemm_tn|attn_|reduce_|qkv_rope|embed_gather|greedy|step_inc|ts_|peer_'
B=${PROFILE_BATCH:-32}
BENCH="python bench.py --steps 2 --warmup 3 --batch $B --only-batch --no-cpu-baseline --no-graph"
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c ${NCU_COUNT:-4500} --csv \
   --log-file gpurun_out/launches_b$B.csv $BENCH > gpurun_out/ncu_launch_b$B.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/ncu_launch_b$B.log; wc -l gpurun_out/launches_b$B.csv
if [ "${FULL:-1}" = "1" ]; then
echo "=== full set: decode GEMMs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tn -s ${GEMM_SKIP:-700} -c 6 \
   -o gpurun_out/prof_gemm_b$B -f $BENCH > gpurun_out/ncu_gemm_b$B.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/ncu_gemm_b$B.log
echo "=== full set: decode attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 60 -c 2 \
   -o gpurun_out/prof_attn_b$B -f $BENCH > gpurun_out/ncu_attn_b$B.log 2>&1
echo "rc=$?"; tail -n 2 gpurun_out/ncu_attn_b$B.log
fi
ls -la gpurun_out/ | head -30
