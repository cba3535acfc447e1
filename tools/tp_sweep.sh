#!/bin/bash
# Tensor-parallel decode sweep on N GPUs of one box (default 2): split-K factors, flash-decode splits, decode-GEMM smem budget and
# RMSNorm cluster size were tuned for ONE GPU; under TP the per-rank GEMMs are 1/N the size and the step is latency-bound, so
# the best values differ.   gpurun --gpus 2 --timeout 1500 -- 'bash tools/tp_sweep.sh'      (charged N x the box time)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=${TP:-2}; OUT=gpurun_out/tp${N}_sweep.txt; : > $OUT
run() {
  echo "## $*" >> $OUT
  env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
      bench.py --gpus $N --steps 48 --warmup 3 --no-cpu-baseline --sweep-only 2>>gpurun_out/tp_sweep.err | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({b:round(v['ms_per_step'],3) for b,v in d['by_batch'].items()})
" >> $OUT
}
for v in "CTS_BASELINE=1" "CTS_PEER_LL=1" "CTS_PEER_LL=1 CTS_DECODE_FUSED=1" "CTS_PEER_LL=1 CTS_DECODE_FUSED=2"; do          # correctness first: TP vs the 1-GPU path on the same weights
  echo "## tp_check $v" >> $OUT
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29400 + RANDOM % 200)) \
      tools/tp_check.py 2>>gpurun_out/tp_sweep.err | tail -n 6 >> $OUT
done
run CTS_BASELINE=1
run CTS_DECODE_FUSED=1
run CTS_PEER_LL=1                                   # two-shot all-reduce with in-band flags (csrc/allreduce_ll.cu)
run CTS_PEER_LL=1 CTS_DECODE_FUSED=1
run CTS_PEER_LL=1 CTS_DECODE_FUSED=2                  # GEMM + all-reduce in one kernel, norms folded into the next projection
for s in "1,1,1,1" "2,1,2,1" "3,2,3,2" "5,3,4,3"; do run CTS_SPLITS=$s; done          # qkv,o,gu,d (decode-sized T only)
for a in 1 2 4; do run CTS_ATTN_SPLITS=$a; done
for k in 50 75 100; do run CTS_DECODE_SMEM_KB=$k; done
for c in 1 2 4 8; do run CTS_NORM_CLUSTER=$c; done
cat $OUT
