#!/bin/bash
# One GPU-box visit: parity tests (each file in its own process so a trapped kernel cannot take the rest down),
# smoke, a short bench.  Everything is logged under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
STATUS=0
for f in gemm elementwise attention ts_encoder model configs; do
  echo "=== tests/test_gpu_$f.py"
  timeout ${TEST_TIMEOUT:-420} python -m pytest tests/test_gpu_$f.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  rc=$?
  tail -n 6 gpurun_out/test_$f.log
  echo "rc=$rc"
  [ $rc -ne 0 ] && STATUS=1
done
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/smoke.log
if [ "${RUN_BENCH:-1}" = "1" ]; then
  echo "=== bench"
  timeout ${BENCH_TIMEOUT:-900} python bench.py --steps ${BENCH_STEPS:-32} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"
  tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
fi
exit $STATUS
