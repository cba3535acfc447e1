#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ts_encoder.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/r2v8_ts_tests.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/r2v8_ts_tests.log
timeout 300 python tools/bench_ts_encoder.py --batches 1 2 --trace > gpurun_out/r2v8_ts_trace.log 2>&1; echo "rc=$?"; tail -n 14 gpurun_out/r2v8_ts_trace.log | cut -c1-1200
