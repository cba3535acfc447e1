#!/bin/bash
# Round-2 visit 36 (one B200): ncu launch list (gpu__time_duration, --clock-control none) of the PREFILL of the bench batch at the final
# HEAD (the first 560 launches of an eager bench run = TS encode + 48 layers of prefill), for the e2e analysis.
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
KREG='regex:gemm_tn|attn_|reduce_|qkv_rope|embed_gather|greedy|ts_|peer_|swiglu'
BENCH="python bench.py --steps 2 --warmup 3 --batch 32 --only-batch --no-cpu-baseline --no-graph --sweep-only"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 560 --csv --log-file gpurun_out/r2v36_launches_prefill_b32.csv $BENCH > gpurun_out/r2v36_ncu_launch_prefill.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r2v36_launches_prefill_b32.csv
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2v36_launches_prefill_b32.csv')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 2:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(',', ''))
    except ValueError: continue
    name = r[kn].split('(')[0].replace('void <unnamed>::', '')[:70]
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{v[1] / 1e6:9.2f} ms {v[0]:5d} {100 * v[1] / tot:5.1f}%  {k}")
PY
