#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -n 2
timeout 600 python tools/bench_w4_gemm.py > gpurun_out/r2v11_w4_gemm.json 2> gpurun_out/r2v11_w4_gemm.err; echo "rc=$?"; cat gpurun_out/r2v11_w4_gemm.json | cut -c1-2500; tail -n 3 gpurun_out/r2v11_w4_gemm.err
timeout 900 python tools/bench_w4.py > gpurun_out/r2v11_w4_bench.json 2> gpurun_out/r2v11_w4_bench.err; echo "rc=$?"; cat gpurun_out/r2v11_w4_bench.json | cut -c1-1500; tail -n 3 gpurun_out/r2v11_w4_bench.err
