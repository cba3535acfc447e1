#!/bin/bash
# Round-2 GPU visit 3 (2 GPUs): overlapped timelines of one decode step (tools/trace_decode_step.py) on 1 GPU (b = 32, 1) and under
# TP2 with the low-latency all-reduce (b = 32, 1); plus a loop over the one test that failed once in round 1 without CTS_DEBUG_SYNC.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for b in 32 1; do
  timeout 300 python tools/trace_decode_step.py --batch $b > gpurun_out/r2v3_trace_b$b.log 2>&1; echo "rc=$?"; head -3 gpurun_out/r2v3_trace_b$b.log; tail -12 gpurun_out/r2v3_trace_b$b.log
done
for b in 32 1; do
  CTS_PEER_LL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
     tools/trace_decode_step.py --batch $b > gpurun_out/r2v3_trace_tp2_b$b.log 2>&1; echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*" gpurun_out/r2v3_trace_tp2_b$b.log | head -3; tail -14 gpurun_out/r2v3_trace_tp2_b$b.log
done
echo "##### native step, 12 runs without CTS_DEBUG_SYNC"
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  timeout 120 python -m pytest "tests/test_gpu_zz_a_native_step.py::test_native_step_is_bit_identical" -q -m gpu --no-header -p no:cacheprovider -x 2>&1 | tail -n 1
done
