#!/bin/bash
# Round-2 visit 12 (one B200): the whole GPU suite on the rebuilt HEAD, smoke, the default bench line, the e2e stage breakdown and the
# W4 GEMM / model side lines of the 3-stage / 8-slot kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
echo "##### GPU suite"
( time timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x ) > gpurun_out/r2v12_pytest.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/r2v12_pytest.log
echo "##### smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2v12_smoke.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/r2v12_smoke.log
echo "##### bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2v12_bench.json 2> gpurun_out/r2v12_bench.err; echo "rc=$?"; cut -c1-1800 gpurun_out/r2v12_bench.json; tail -n 3 gpurun_out/r2v12_bench.err
echo "##### e2e breakdown"
timeout 600 python tools/e2e_breakdown.py > gpurun_out/r2v12_e2e.json 2> gpurun_out/r2v12_e2e.err; echo "rc=$?"; cat gpurun_out/r2v12_e2e.json; tail -n 3 gpurun_out/r2v12_e2e.err
echo "##### W4"
timeout 600 python tools/bench_w4_gemm.py > gpurun_out/r2v12_w4_gemm.json 2> gpurun_out/r2v12_w4_gemm.err; echo "rc=$?"; cut -c1-2500 gpurun_out/r2v12_w4_gemm.json; tail -n 3 gpurun_out/r2v12_w4_gemm.err
timeout 900 python tools/bench_w4.py > gpurun_out/r2v12_w4_bench.json 2> gpurun_out/r2v12_w4_bench.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2v12_w4_bench.json; tail -n 3 gpurun_out/r2v12_w4_bench.err
