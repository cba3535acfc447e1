#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
echo "##### W4A16 after the conflict-free dequant mapping"
timeout 600 python -m pytest tests/test_gpu_w4.py tests/test_gpu_ts_encoder.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2v10a_tests.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/r2v10a_tests.log
timeout 900 python tools/bench_w4.py > gpurun_out/r2v10a_w4_bench.json 2> gpurun_out/r2v10a_w4_bench.err; echo "rc=$?"; cat gpurun_out/r2v10a_w4_bench.json | cut -c1-1500; tail -n 3 gpurun_out/r2v10a_w4_bench.err
echo "##### fused TS encoder, automatic split"
CTS_TS_FUSED_DEBUG=1 timeout 300 python tools/bench_ts_encoder.py --batches 1 2 2>&1 | grep -E "chosen split|\"b1\"" | head -4 | cut -c1-700
