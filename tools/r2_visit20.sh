#!/bin/bash
# Round-2 visit 20 (one B200): single-pass tcgen05 prefill attention (lazy rescale of O in TMEM): parity tests (attention, training LSE
# users, model), the kernel alone against the installed kernels, one ncu --set full capture.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_zz_c_train.py tests/test_gpu_zz_d_attn_bwd_tc5.py tests/test_gpu_model.py tests/test_gpu_configs.py tests/test_gpu_w4.py -q -m gpu --no-header -p no:cacheprovider -x 2>&1 | tail -n 4
timeout 600 python tools/bench_attention.py > gpurun_out/r2v20_attention.json 2> gpurun_out/r2v20_attention.err; echo "rc=$?"; cat gpurun_out/r2v20_attention.json | cut -c1-2500; tail -n 3 gpurun_out/r2v20_attention.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_tc5_kernel -s 1 -c 1 -o gpurun_out/r2v20_prof_attn_prefill -f python tools/bench_attention.py --once > gpurun_out/r2v20_ncu.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2v20_ncu.log
