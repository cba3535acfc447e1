#!/bin/bash
# Round-2 GPU visit 5 (one GPU): the re-gated parity tests (bf16 + fp16, comparative gate), smoke, then BASELINE config 5 -- the
# ChatTS-8B LoRA step -- with the opt-in training kernels A/B'd at full depth, and the ncu evidence of the training kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
echo "##### a. parity tests + smoke"
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_model.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2v5_parity.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/r2v5_parity.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v5_smoke.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2v5_smoke.log
cp gpurun_out/parity_metrics.jsonl gpurun_out/r2v5_parity_metrics.jsonl 2>/dev/null
echo "##### b. config 5 (ChatTS-8B LoRA step), full depth, variants"
ab() { tag=$1; shift; echo "## $tag: $*"; env "$@" timeout 600 python tools/bench_lora.py --steps 4 --warmup 3 > gpurun_out/r2v5_lora_$tag.json 2>> gpurun_out/r2v5_lora.err; python -c "
import sys,json
for l in open('gpurun_out/r2v5_lora_$tag.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({'ms_per_step': round(d['ms_per_step'],2), 'positions_per_s': round(d['value']), 'frac': round(d['roofline']['frac'],3), 'loss': round(d.get('loss',0),4), 'e2e': d.get('e2e',{}).get('value')})
"; }
ab base CTS_BASE=1
ab tc5 CTS_ATTN_BWD_TC5=1
ab wgrad CTS_WGRAD_MMA=1
ab both CTS_ATTN_BWD_TC5=1 CTS_WGRAD_MMA=1
echo "##### c. ncu: launch list of one training step (2 layers) + full captures"
LORA="python tools/bench_lora.py --steps 1 --warmup 3 --layers 2"
CTS_ATTN_BWD_TC5=1 CTS_WGRAD_MMA=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2v5_launches_lora.csv $LORA > gpurun_out/r2v5_ncu_launch_lora.log 2>&1; echo "rc=$?"; wc -l gpurun_out/r2v5_launches_lora.csv
CTS_ATTN_BWD_TC5=1 CTS_WGRAD_MMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:attn_bwd_dq|attn_bwd_dkv|lora_wgrad|ce_loss_grad' -s 40 -c 8 \
   -o gpurun_out/r2v5_prof_train_kernels -f $LORA > gpurun_out/r2v5_ncu_train_kernels.log 2>&1; echo "rc=$?"
tail -n 8 gpurun_out/r2v5_lora.err
ls -la gpurun_out | grep r2v5
