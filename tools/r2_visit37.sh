#!/bin/bash
# Round-2 visit 37 (one B200): warp-per-row RMSNorm for prefill-sized T: parity tests (elementwise, configs, model, training), e2e stages.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_configs.py tests/test_gpu_model.py tests/test_gpu_zz_c_train.py tests/test_gpu_zz_a_native_step.py -q -m gpu --no-header -p no:cacheprovider -x 2>&1 | tail -n 3
timeout 600 python tools/e2e_breakdown.py > gpurun_out/r2v37_e2e.json 2> gpurun_out/r2v37_e2e.err; echo "rc=$?"; cat gpurun_out/r2v37_e2e.json; tail -n 2 gpurun_out/r2v37_e2e.err
