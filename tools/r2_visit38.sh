#!/bin/bash
# Round-2 visit 38 (one B200): the whole GPU suite + smoke on the final HEAD (elementwise.cu gained the rows RMSNorm kernel).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x ) > gpurun_out/r2v38_pytest.log 2>&1; echo "rc=$?"; tail -n 5 gpurun_out/r2v38_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2v38_smoke.log 2>&1; echo "rc=$?"; tail -n 1 gpurun_out/r2v38_smoke.log | cut -c1-300
