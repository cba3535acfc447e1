#!/bin/bash
# Round-2 visit 31 (one B200): the one-launch TS encoder with three DSMEM items in flight per thread and batched TMEM loads in the
# epilogue: parity tests, timing at b = 1 / 2 (forced splits for comparison), phase timeline.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ts_encoder.py tests/test_gpu_zz_a_native_step.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -n 2
timeout 300 python tools/bench_ts_encoder.py > gpurun_out/r2v31_ts.json 2> gpurun_out/r2v31_ts.err; echo "rc=$?"; tail -n 12 gpurun_out/r2v31_ts.json | cut -c1-1200; tail -n 2 gpurun_out/r2v31_ts.err
for s in 4 6 8; do echo "## forced split $s"; CTS_TS_FUSED_SPLIT=$s timeout 120 python tools/bench_ts_encoder.py --batches 1 2>&1 | grep -E '^\{' | cut -c1-330; done
