#!/bin/bash
# Round-2 multi-GPU visit (N GPUs, default 4): BASELINE config 5 under data parallelism (bench_lora --gpus N) and the tensor-parallel
# decode line of bench.py at N (tp_parity gate, e2e, NVLink bytes), plus a TP timeline from rank 0.
#   gpurun --gpus 4 --timeout 1500 -- 'bash tools/r2_visit10_multi.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${TP:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${RUN_LORA:-1}" = "1" ]; then
  echo "##### config 5: LoRA step, $N x DP"
  timeout 700 $TR --master-port $((29300 + RANDOM % 100)) tools/bench_lora.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r2v10_lora_$N.json 2> gpurun_out/r2v10_lora_$N.err; echo "rc=$?"
  python -c "
import json
for l in open('gpurun_out/r2v10_lora_$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k: d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, 'frac', d['roofline']['frac'], 'e2e', d.get('e2e',{}).get('value'), {k:v for k,v in d.items() if 'allreduce' in k or 'dp' in k})
"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2v10_lora_$N.err | tail -n 4
fi
echo "##### decode, TP$N"
timeout 700 $TR --master-port $((29400 + RANDOM % 100)) bench.py --gpus $N --steps 32 --warmup 3 > gpurun_out/r2v10_bench_tp$N.json 2> gpurun_out/r2v10_bench_tp$N.err; echo "rc=$?"
python -c "
import json
for l in open('gpurun_out/r2v10_bench_tp$N.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k: d.get(k) for k in ('value','ms_per_step','launches_per_step')}, {b:round(v['ms_per_step'],3) for b,v in d['by_batch'].items()}, 'e2e', (d.get('e2e') or {}).get('value'), 'tp_parity', d.get('tp_parity'), 'nvlink', d.get('nvlink_bytes_per_step_per_rank'))
"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2v10_bench_tp$N.err | tail -n 6
echo "##### timeline, TP$N, b=32 and b=1"
for b in 32 1; do
  TRACE_TAG=_v10 timeout 300 $TR --master-port $((29500 + RANDOM % 100)) tools/trace_decode_step.py --batch $b --rows 30 > gpurun_out/r2v10_trace_tp${N}_b$b.log 2>&1; echo "rc=$?"
  grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/r2v10_trace_tp${N}_b$b.log | grep -E "untraced|weight streams|->" | head -12
done
