#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
KREG='regex:gemm_tn|attn_|reduce_|qkv_rope|embed_gather|greedy|ts_|peer_'
BENCH="python bench.py --steps 2 --warmup 3 --batch 32 --only-batch --no-cpu-baseline --no-graph --sweep-only"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 560 --csv --log-file gpurun_out/launches_prefill_b32.csv $BENCH > gpurun_out/ncu_launch_prefill.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_prefill_b32.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill -s 4 -c 2 -o gpurun_out/prof_attn_prefill_b32 -f $BENCH > gpurun_out/ncu_attn_prefill.log 2>&1
echo "rc=$?"
