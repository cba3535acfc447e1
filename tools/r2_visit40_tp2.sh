#!/bin/bash
# Round-2 visit 18 (two B200): bench.py under tensor parallelism 2 with the BASELINE configs[3] side block (second TP model instance).
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${TP:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 100)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2v40_bench_tp$N.json 2> gpurun_out/r2v40_bench_tp$N.err ) 2>&1 | grep real; echo "rc=$?"
python - <<PY
import json
for l in open('gpurun_out/r2v40_bench_tp$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k: d.get(k) for k in ('value','ms_per_step','launches_per_step')}, {b:round(v['ms_per_step'],3) for b,v in d['by_batch'].items()}, 'e2e', (d.get('e2e') or {}).get('value'), 'tp_parity', d.get('tp_parity'), 'config4', d.get('config4'))
PY
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2v40_bench_tp$N.err | tail -n 6
