#!/bin/bash
# Round-2 visit 23 (one B200): prefill attention with two softmax warps per row block: parity tests, the kernel alone.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_zz_c_train.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider -x 2>&1 | tail -n 3
timeout 600 python tools/bench_attention.py > gpurun_out/r2v23_attention.json 2> gpurun_out/r2v23_attention.err; echo "rc=$?"; cat gpurun_out/r2v23_attention.json | cut -c1-2500; tail -n 3 gpurun_out/r2v23_attention.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_tc5_kernel -s 1 -c 1 -o gpurun_out/r2v23_prof_attn_prefill -f python tools/bench_attention.py --once > gpurun_out/r2v23_ncu.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/r2v23_ncu.log
