#!/bin/bash
# Round-2 visit 32/33 (N B200): the tensor-parallel prefill's row-parallel exchange -- rs_ag (default: fp32 reduce-scatter + 16-bit
# all-gather of the result), fp32 (one fp32 all-reduce), 16bit (all-reduce in the model dtype) -- bench.py line at N (e2e, config4 prefill
# seconds, parity gate).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${TP:-2}
for mode in ${MODES:-rs_ag fp32 16bit}; do
  export CTS_TP_PREFILL_EXCHANGE=$mode
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 100)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2v32_bench_tp${N}_$mode.json 2> gpurun_out/r2v32_bench_tp${N}_$mode.err; echo "$mode rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/r2v32_bench_tp${N}_$mode.json'):
    if l.startswith('{'):
        d=json.loads(l); c4=d.get('config4') or {}
        print('$mode', {k: d.get(k) for k in ('value','ms_per_step')}, 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('seconds'), 'parity', {k:(d.get('tp_parity') or {}).get(k) for k in ('prefill_logits_max_rel','decode_logits_max_rel','greedy_agreement_of_8','pass')}, 'config4 prefill_s', c4.get('prefill_s'), 'e2e', (c4.get('e2e') or {}).get('value'))
PY
  grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2v32_bench_tp${N}_$mode.err | tail -n 2
done
