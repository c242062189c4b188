#!/bin/bash
# A/B of the tuning knobs in one process (same table, same batches) + the correctness subset + launch list.
#   tools/gpu_ab.sh <tag> "<ab spec>" [ncu-regex skip count]
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-ab}; AB=${2:-"claim_pf=0;seg_vpl=1;apply_pf=0;lookup_pf=1"}
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "pool_backward or hot_fid or hash_filter or sharded_direct or smoke or bench_shape or lookup or tower" > $O/${T}_tests.log 2>&1; tail -3 $O/${T}_tests.log
timeout 900 python bench.py --no-cpu-baseline --no-extras --steps 20 --repeats 5 --ab "$AB" > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -c 300 $O/${T}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("e2e", d["e2e"] and (round(d["e2e"]["ms_per_step"], 4), round(d["e2e"]["dense_tower_ms"], 4)))
print("ms/step", round(d["ms_per_step"], 4), "fwd", round(r["launch_ms"], 4), "bwd", round(r["backward"]["ms"], 4), "parity", d["parity_check"]["ok"], d["parity_check"]["hot_max_rel_err"])
for k, v in (d.get("ab") or {}).items():
  print("  ab", k, round(v["ms_per_step"], 4), v["regions_ms"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/${T}_ncu_l.log 2>&1
python - <<PY
import csv, re
rows = []
for line in csv.reader(open("gpurun_out/${T}_launches.csv")):
  if len(line) > 10 and line[0].isdigit():
    rows.append((re.sub(r"\(.*", "", line[4]).replace("void mono::", "").replace("mono::", ""), float(line[-1]) / 1000))
idx = [i for i, (k, _) in enumerate(rows) if k.startswith("lookup_kernel") or k.startswith("lookup_tma")]
a, b = idx[1], idx[2]
tot = 0
for k, v in rows[a:b]:
  print(f"{v:8.1f} us  {k[:60]}")
  tot += v
print(f"{tot:8.1f} us  total of {b - a} launches")
PY
if [ -n "$3" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$3" -s ${4:-6} -c ${5:-8} -o $O/${T}_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/${T}_ncu_f.log 2>&1; tail -1 $O/${T}_ncu_f.log
fi
