#!/bin/bash
# lean perf iteration: bench (no CPU arm / e2e / extras) + launch list + optional full ncu of the backward kernels
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-p}
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "pool_backward or hot_fid or hash_filter or sharded_direct or smoke" > $O/${T}_tests.log 2>&1; tail -3 $O/${T}_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-extras --steps 20 --repeats 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -c 300 $O/${T}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms/step", round(d["ms_per_step"], 4), "fwd", round(r["launch_ms"], 4), "bwd", round(r["backward"]["ms"], 4), "parity", d["parity_check"]["ok"], d["parity_check"]["hot_max_rel_err"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/${T}_ncu_l.log 2>&1
python - <<PY
import csv, re
rows = []
for line in csv.reader(open("gpurun_out/${T}_launches.csv")):
  if len(line) > 10 and line[0].isdigit():
    rows.append((re.sub(r"\(.*", "", line[4]).replace("void mono::", "").replace("mono::", ""), float(line[-1]) / 1000))
# one forward + one backward: from the first lookup kernel after the warm-up to the next one
idx = [i for i, (k, _) in enumerate(rows) if k.startswith("lookup_kernel") or k.startswith("lookup_tma")]
a, b = idx[1], idx[2]
tot = 0
for k, v in rows[a:b]:
  print(f"{v:8.1f} us  {k[:60]}")
  tot += v
print(f"{tot:8.1f} us  total of {b - a} launches")
PY
if [ -n "$2" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${3:-6} -c ${4:-8} -o $O/${T}_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/${T}_ncu_f.log 2>&1; tail -1 $O/${T}_ncu_f.log
fi
for V in 0 1; do
  MONO_LOOKUP_TMA=$V timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 10 --repeats 2 > $O/${T}_fwd_tma$V.json 2> $O/${T}_fwd_tma$V.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_fwd_tma$V.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("lookup_tma=$V: step", round(d["ms_per_step"], 4), "fwd zipf ms", round(r["launch_ms"], 4), "frac", round(r["frac"], 3), {k: (round(v["ms"], 4), round(v.get("frac", 0), 3)) for k, v in (r.get("extras") or {}).items()})
PY
done
