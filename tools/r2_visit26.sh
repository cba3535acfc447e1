#!/bin/bash
# Round-2 visit 26 (one B200): W4 mma kernel, CTAs per SM A/B with the 128-K stages (2 x 5 stages against 3 x 3), + ncu at t = 1.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for c in 2 3; do
  CTS_W4M_CTAS=$c W4_TC5=0 timeout 600 python tools/bench_w4_gemm.py > gpurun_out/r2v26_w4_gemm_ctas$c.json 2> gpurun_out/r2v26_w4_gemm_ctas$c.err; echo "ctas=$c rc=$?"
  python - <<PY
import json
d = json.load(open('gpurun_out/r2v26_w4_gemm_ctas$c.json'))
for name, o in d.items():
    print(name, {k: (v['best'], v['best_us']) for k, v in o.items() if k.startswith('mma')})
PY
done
W4_ONCE_T=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_w4_mma_kernel -s 2 -c 1 -o gpurun_out/r2v26_prof_w4_mma_t1 -f python tools/bench_w4_gemm.py --once > gpurun_out/r2v26_ncu.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/r2v26_ncu.log
timeout 300 python -m pytest tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider -k "gptq_checkpoint_directory" 2>&1 | tail -n 2
