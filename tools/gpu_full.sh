#!/bin/bash
# full 1-GPU validation: whole GPU suite, the default bench line (with e2e), the C4 streaming workload (small + full)
cd "$(dirname "$0")/.."
O=gpurun_out; T=${1:-full}
mkdir -p $O
timeout 1500 python -m pytest tests -q -x -m gpu > $O/${T}_tests.log 2>&1; tail -4 $O/${T}_tests.log
timeout 900 python bench.py --no-cpu-baseline --steps 20 --repeats 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -c 400 $O/${T}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms/step", round(d["ms_per_step"], 4), "fwd", round(r["launch_ms"], 4), "frac", round(r["frac"], 3), "bwd", round(r["backward"]["ms"], 4), "parity", d["parity_check"]["ok"])
e = d["e2e"]
print("e2e ms", round(e["ms_per_step"], 4), "G/s", round(e["value"] / 1e9, 3), "tower", round(e["dense_tower_ms"], 4), "phases", e.get("device_phases"))
print("extras", {k: (round(v["ms"], 4), round(v.get("frac", 0), 3)) for k, v in (r.get("extras") or {}).items()})
PY
timeout 600 python bench.py --workload c4 --keys 20000000 --steps 16 --repeats 3 > $O/${T}_c4_small.json 2> $O/${T}_c4_small.err; tail -c 300 $O/${T}_c4_small.err; cut -c1-1500 $O/${T}_c4_small.json
timeout 900 python bench.py --workload c4 --steps 16 --repeats 3 > $O/${T}_c4.json 2> $O/${T}_c4.err; tail -c 300 $O/${T}_c4.err; cut -c1-1800 $O/${T}_c4.json
