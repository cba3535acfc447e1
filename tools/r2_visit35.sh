#!/bin/bash
# Round-2 visit 35 (one B200): the GPU suite + smoke on the final HEAD (model.py's step allocation changed for the TP prefill exchange).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x ) > gpurun_out/r2v35_pytest.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/r2v35_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2v35_smoke.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/r2v35_smoke.log | cut -c1-300
