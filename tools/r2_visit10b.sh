#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
echo "##### W4: racecheck of the one case that was not bit-identical, then the tests, then the bench"
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest "tests/test_gpu_w4.py::test_w4_partials_bit_identical_to_the_dense_gemm" -q -m gpu --no-header -p no:cacheprovider -k "5120-13824-128-1-11-dtype1 or 1024-2560" > gpurun_out/r2v10b_racecheck.log 2>&1; echo "rc=$?"; grep -E "RACECHECK|Race|hazard|ERROR SUMMARY|passed|failed" gpurun_out/r2v10b_racecheck.log | head -20
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -n 2; done
timeout 900 python tools/bench_w4.py > gpurun_out/r2v10b_w4_bench.json 2> gpurun_out/r2v10b_w4_bench.err; echo "rc=$?"; cat gpurun_out/r2v10b_w4_bench.json | cut -c1-1500; tail -n 3 gpurun_out/r2v10b_w4_bench.err
