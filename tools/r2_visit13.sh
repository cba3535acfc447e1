#!/bin/bash
# Round-2 visit 13 (one B200): the register-operand W4A16 kernel (csrc/gemm_w4_mma.cu) -- parity tests at every shape, the GEMM sweep
# (splits x token counts, against the bf16 GEMM and the tcgen05 W4 kernel), the model side line, one ncu --set full capture.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -n 4
timeout 900 python tools/bench_w4_gemm.py > gpurun_out/r2v13_w4_gemm.json 2> gpurun_out/r2v13_w4_gemm.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2v13_w4_gemm.json'))
for name, o in d.items():
    print(name, {k: ({kk: v[kk] for kk in ('suggested', 'best', 'best_us', 'suggested_us', 'best_packed_gbs', 'speedup_vs_bf16') if kk in v} if k.startswith('mma') else v) for k, v in o.items()})
PY
tail -n 3 gpurun_out/r2v13_w4_gemm.err
timeout 900 python tools/bench_w4.py > gpurun_out/r2v13_w4_bench.json 2> gpurun_out/r2v13_w4_bench.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2v13_w4_bench.json; tail -n 3 gpurun_out/r2v13_w4_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_w4_mma_kernel -s 2 -c 2 -o gpurun_out/r2v13_prof_w4_mma -f python tools/bench_w4_gemm.py --once > gpurun_out/r2v13_ncu.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2v13_ncu.log
