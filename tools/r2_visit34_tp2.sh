#!/bin/bash
# Round-2 visit 34 (two B200): tools/tp_check.py with a prefill whose token count is not a multiple of the world size (padded token shards
# of the reduce-scatter / all-gather exchange), all three exchanges.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for mode in rs_ag fp32 16bit; do
  CTS_TP_PREFILL_EXCHANGE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 100)) tools/tp_check.py 2>&1 | grep -E "tp_check|Error|error" | head -6; echo "$mode rc=${PIPESTATUS[0]}"
done
