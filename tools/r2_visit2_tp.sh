#!/bin/bash
# Round-2 GPU visit 2 (N GPUs, default 2): the tensor-parallel variants that had only run with streams as ranks on one GPU --
# correctness against the 1-GPU path first (tools/tp_check.py), then ms/step at b = 1 / 8 / 32 on the full 48-layer model.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/r2_visit2_tp.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${TP:-2}; OUT=gpurun_out/r2v2_tp${N}.txt; : > $OUT
nvidia-smi topo -m > gpurun_out/r2v2_topo.txt 2>&1
for v in "CTS_BASELINE=1" "CTS_PEER_LL=1" "CTS_PEER_LL=1 CTS_DECODE_FUSED=1" "CTS_PEER_LL=1 CTS_DECODE_FUSED=2"; do
  echo "## tp_check $v" >> $OUT
  env $v timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) \
      tools/tp_check.py 2>>gpurun_out/r2v2_tp.err | tail -n 6 >> $OUT
  echo "rc=${PIPESTATUS[0]}" >> $OUT
done
run() {
  echo "## $*" >> $OUT
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
      bench.py --gpus $N --steps 48 --warmup 3 --no-cpu-baseline --sweep-only 2>>gpurun_out/r2v2_tp.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({'launches': d.get('launches_per_step'), 'by_batch': {b:(round(v['ms_per_step'],3), v['tokens_sha1']) for b,v in d['by_batch'].items()}})
" >> $OUT
}
run CTS_BASELINE=1
run CTS_PEER_LL=1
run CTS_PEER_LL=1 CTS_DECODE_FUSED=1
run CTS_PEER_LL=1 CTS_DECODE_FUSED=2
cat $OUT
tail -n 30 gpurun_out/r2v2_tp.err
