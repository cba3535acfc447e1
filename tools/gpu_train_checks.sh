#!/bin/bash
# First B200 visit for the LoRA fine-tune step (row A9): the training-kernel parity tests with their xfail markers ignored
# (--runxfail: real pass/fail per kernel), then -- only if they pass -- the config-5 bench at reduced and full depth and the
# ncu evidence (launch list of one training step + full captures of the attention-backward and weight-gradient kernels).
# Everything is bounded by `timeout`; logs land in gpurun_out/.   One GPU.
#   gpurun --timeout 1500 -- 'bash tools/gpu_train_checks.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
STATUS=0
for f in train_kernels zz_a_native_step zz_b_sampling zz_d_attn_bwd_tc5 zz_e_fused_decode zz_f_peer_ll; do          # each file in its own process: a trapped kernel cannot take the rest down
  echo "=== tests/test_gpu_$f.py (--runxfail)"
  timeout 300 python -m pytest tests/test_gpu_$f.py -q -m gpu --runxfail --no-header -p no:cacheprovider -rfE --junitxml gpurun_out/junit_$f.xml > gpurun_out/test_$f.log 2>&1
  echo "rc=$?"; tail -n 12 gpurun_out/test_$f.log
done
# a file that failed above: rerun its first failing case alone with CTS_DEBUG_SYNC=1 so the log names the faulting entry point
for f in zz_a_native_step zz_b_sampling zz_d_attn_bwd_tc5 zz_e_fused_decode zz_f_peer_ll; do
  if grep -q "^FAILED\|^ERROR" gpurun_out/test_$f.log 2>/dev/null; then
    first=$(grep -m1 "^FAILED\|^ERROR" gpurun_out/test_$f.log | awk '{print $2}')
    echo "=== debug rerun of $first"
    CTS_DEBUG_SYNC=1 timeout 200 python -m pytest "$first" -q -m gpu --runxfail --no-header -p no:cacheprovider -x 2>&1 | tail -n 25 > gpurun_out/debug_$f.log
    tail -n 12 gpurun_out/debug_$f.log
  fi
done
echo "=== tests/test_gpu_zz_c_train.py (--runxfail)"
timeout ${TEST_TIMEOUT:-600} python -m pytest tests/test_gpu_zz_c_train.py -q -m gpu --runxfail --no-header -p no:cacheprovider -rfE --junitxml gpurun_out/junit_zz_c_train.xml \
   > gpurun_out/test_train.log 2>&1
rc=$?
tail -n 40 gpurun_out/test_train.log
echo "rc=$rc"
if [ $rc -ne 0 ] && [ "${FORCE_BENCH:-0}" != "1" ]; then
  echo "training kernels not green: skipping bench/profiles (FORCE_BENCH=1 overrides)"
  exit $rc
fi
echo "=== bench_lora (4 layers, smoke)"
timeout 600 python tools/bench_lora.py --steps 2 --warmup 3 --layers 4 > gpurun_out/bench_lora_4layers.json 2> gpurun_out/bench_lora_4layers.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_lora_4layers.json; tail -n 5 gpurun_out/bench_lora_4layers.err
echo "=== bench_lora (ChatTS-8B, full depth)"
timeout 900 python tools/bench_lora.py --steps ${LORA_STEPS:-6} --warmup 3 > gpurun_out/bench_lora.json 2> gpurun_out/bench_lora.err
echo "rc=$?"; tail -c 3000 gpurun_out/bench_lora.json; tail -n 5 gpurun_out/bench_lora.err
if [ "${RUN_NCU:-1}" = "1" ]; then
  LORA="python tools/bench_lora.py --steps 1 --warmup 3 --layers 2"
  echo "=== launch list (one training step, 2 layers)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_lora.csv $LORA \
     > gpurun_out/ncu_launch_lora.log 2>&1
  echo "rc=$?"; wc -l gpurun_out/launches_lora.csv
  echo "=== full set: attention backward, LoRA weight gradient, cross entropy"
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:attn_bwd_dq|attn_bwd_dkv|lora_wgrad|ce_loss_grad' -s 40 -c 8 \
     -o gpurun_out/prof_train_kernels -f $LORA > gpurun_out/ncu_train_kernels.log 2>&1
  echo "rc=$?"
fi
ls -la gpurun_out/ | grep -E "ncu-rep|csv|json"
