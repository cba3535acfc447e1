#!/bin/bash
# Round-2 GPU visit 6 (one GPU): first execution of the fused TS encoder (csrc/ts_encoder_fused.cu) -- its tests, then timing against
# the multi-launch path at the 14B shapes, then a full default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
echo "##### a. TS encoder tests (fused kernel on the default path)"
timeout 600 python -m pytest tests/test_gpu_ts_encoder.py -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/r2v6_ts_tests.log 2>&1; echo "rc=$?"; tail -n 25 gpurun_out/r2v6_ts_tests.log
CTS_TS_FUSED_COOP=1 timeout 600 python -m pytest tests/test_gpu_ts_encoder.py -q -m gpu --no-header -p no:cacheprovider -x -k fused > gpurun_out/r2v6_ts_tests_coop.log 2>&1; echo "coop rc=$?"; tail -n 6 gpurun_out/r2v6_ts_tests_coop.log
timeout 600 python -m pytest tests/test_gpu_zz_a_native_step.py tests/test_gpu_configs.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2v6_other_tests.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/r2v6_other_tests.log
echo "##### b. TS encoder timing, 14B shapes"
timeout 600 python tools/bench_ts_encoder.py > gpurun_out/r2v6_ts_bench.json 2> gpurun_out/r2v6_ts_bench.err; echo "rc=$?"; cat gpurun_out/r2v6_ts_bench.json; tail -n 5 gpurun_out/r2v6_ts_bench.err
CTS_TS_FUSED_COOP=1 timeout 600 python tools/bench_ts_encoder.py --batches 1 > gpurun_out/r2v6_ts_bench_coop.json 2>> gpurun_out/r2v6_ts_bench.err; echo "coop rc=$?"; cat gpurun_out/r2v6_ts_bench_coop.json
echo "##### c. default bench line"
timeout 900 python bench.py --steps 32 --warmup 3 > gpurun_out/r2v6_bench.json 2> gpurun_out/r2v6_bench.err; echo "rc=$?"; python -c "
import json
for l in open('gpurun_out/r2v6_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k: d[k] for k in ('value','ms_per_step','launches_per_step')}, d['by_batch'], 'e2e', d['e2e']['value'], 'ts', {k:(round(v['us'],1), round(v['hbm_frac'],3), v['launches']) for k,v in d['ts_encoder']['cases'].items()}, 'whole_step', d['roofline']['whole_step']['frac'])
"; tail -n 5 gpurun_out/r2v6_bench.err
