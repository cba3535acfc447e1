#!/bin/bash
# Round-2 visit 28 (one B200): compute-sanitizer (memcheck, racecheck, synccheck) over the kernels written in this session -- the
# persistent single-pass prefill attention and the W4A16 mma kernel -- and the LoRA step (config 5) re-measured with the new attention forward.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
SEL_A='tests/test_gpu_attention.py -k "prefill and (lens1 or growing or many_items or fp16) and not 64-"'
SEL_W='tests/test_gpu_w4.py -k "mma_partials and (256-512-128-5-2 or 384-1024 or 528-1536 or 1024-2560 or 128-256-64)"'
for tool in memcheck racecheck synccheck; do
  extra=""; [ "$tool" = racecheck ] && extra="--racecheck-report analysis"
  eval timeout 900 compute-sanitizer --tool $tool $extra python -m pytest $SEL_A -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2v28_${tool}_attention.log 2>&1; echo "$tool attention rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2v28_${tool}_attention.log | tail -n 3
  eval timeout 900 compute-sanitizer --tool $tool $extra python -m pytest $SEL_W -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2v28_${tool}_w4.log 2>&1; echo "$tool w4 rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2v28_${tool}_w4.log | tail -n 3
done
timeout 600 python tools/bench_lora.py --steps 4 --warmup 3 > gpurun_out/r2v28_lora_1.json 2> gpurun_out/r2v28_lora_1.err; echo "rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2v28_lora_1.json'):
    if l.startswith('{'):
        d = json.loads(l); print({k: d.get(k) for k in ('value', 'ms_per_step')}, 'frac', d['roofline']['frac'], 'e2e', d.get('e2e', {}).get('value'))
PY
tail -n 2 gpurun_out/r2v28_lora_1.err
