#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python tools/bench_w4_gemm.py > gpurun_out/r2v10d_w4_gemm.json 2> gpurun_out/r2v10d_w4_gemm.err; echo "rc=$?"; cat gpurun_out/r2v10d_w4_gemm.json | cut -c1-2500; tail -n 3 gpurun_out/r2v10d_w4_gemm.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_w4_kernel -s 2 -c 2 -o gpurun_out/r2v10d_prof_w4 -f python tools/bench_w4_gemm.py --once > gpurun_out/r2v10d_ncu.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2v10d_ncu.log
