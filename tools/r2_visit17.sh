#!/bin/bash
# Round-2 visit 17 (one B200): ncu --set full of the W4A16 mma kernel at t = 1 (scale-after-accumulate path) + the failing model test's output.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider -k "model_decodes" 2>&1 | grep -E "assert|Error|passed|failed|agree|err" | head -20
W4_ONCE_T=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_w4_mma_kernel -s 2 -c 2 -o gpurun_out/r2v17_prof_w4_mma_t1 -f python tools/bench_w4_gemm.py --once > gpurun_out/r2v17_ncu.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2v17_ncu.log
