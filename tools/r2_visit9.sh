#!/bin/bash
# Round-2 GPU visit 9 (one GPU): (a) why the fused TS encoder gets a 3-way split (occupancy numbers, forced splits); (b) first run of the
# W4A16 decode GEMM: parity tests, then decode ms/step of a synthetic W4 ChatTS-14B at b = 1 / 8 / 32 against the bf16 model.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
echo "##### a. fused TS encoder occupancy"
CTS_TS_FUSED_DEBUG=1 timeout 300 python tools/bench_ts_encoder.py --batches 1 2>&1 | grep -E "ts_fused|fused" | head -14 | cut -c1-400
for s in 4 5 6 7; do echo "## forced split $s"; CTS_TS_FUSED_SPLIT=$s timeout 120 python tools/bench_ts_encoder.py --batches 1 2>&1 | tail -n 2 | cut -c1-330; done
echo "##### b. W4A16"
timeout 600 python -m pytest tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2v9_w4_tests.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/r2v9_w4_tests.log
timeout 900 python tools/bench_w4.py > gpurun_out/r2v9_w4_bench.json 2> gpurun_out/r2v9_w4_bench.err; echo "rc=$?"; cat gpurun_out/r2v9_w4_bench.json | cut -c1-1500; tail -n 5 gpurun_out/r2v9_w4_bench.err
