#!/bin/bash
# per-kernel times of the sharded step on ONE GPU (every 'remote' store is local)
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-s}; X=${2:-direct}
timeout 600 python bench.py --sharded --exchange $X --no-cpu-baseline --no-e2e --no-extras --steps 20 --repeats 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -c 300 $O/${T}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
print("sharded $X: ms/step", round(d["ms_per_step"], 4), "parity", d["parity_check"]["ok"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv python bench.py --sharded --exchange $X --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/${T}_ncu_l.log 2>&1
python - <<PY
import csv, re
rows = []
for line in csv.reader(open("gpurun_out/${T}_launches.csv")):
  if len(line) > 10 and line[0].isdigit():
    rows.append((re.sub(r"\(.*", "", line[4]).replace("void mono::", "").replace("mono::", ""), float(line[-1]) / 1000))
idx = [i for i, (k, _) in enumerate(rows) if k.startswith("fid_claim")]
a, b = idx[-2], idx[-1]
tot = 0
for k, v in rows[a:b]:
  print(f"{v:8.1f} us  {k[:60]}")
  tot += v
print(f"{tot:8.1f} us  total of {b - a} launches")
PY
