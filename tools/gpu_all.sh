#!/bin/bash
# One GPU-box session of round 2: tests, benches, A/Bs and ncu captures; everything lands in gpurun_out/r2_*.
# Every step has its own timeout so that a hang in one does not eat the call.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r2_smi.txt 2>&1
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2_tests.log 2>&1; tail -12 $O/r2_tests.log
if ! tail -3 $O/r2_tests.log | grep -q " passed"; then
  echo "== tests (no -x, to see every failure)"; timeout 1500 python -m pytest tests -m gpu -q > $O/r2_tests_all.log 2>&1; tail -40 $O/r2_tests_all.log | cut -c1-220
fi
echo "== bench N=1"; timeout 900 python bench.py > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; tail -c 300 $O/r2_bench_n1.err; head -c 600 $O/r2_bench_n1.json; echo
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; head -c 500 $O/r2_bench_ref.json; echo
echo "== forward A/B: TMA-staged lookup"; MONO_LOOKUP_TMA=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-parity --steps 10 --repeats 3 > $O/r2_bench_tma.json 2> $O/r2_bench_tma.err; tail -c 300 $O/r2_bench_tma.err
python - <<'PY'
import json
for f in ("r2_bench_n1", "r2_bench_tma"):
  try:
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f, "ms/step", round(d["ms_per_step"], 4), "fwd ms", round(r["launch_ms"], 4), "frac", round(r["frac"], 3), "bwd ms", round(r["backward"]["ms"], 4),
          {k: round(v["ms"], 4) for k, v in (r.get("extras") or {}).items()})
  except Exception as e:
    print(f, "unreadable:", e)
PY
echo "== workload c3 (Criteo-shaped: 26 slots, dim 16)"; timeout 900 python bench.py --workload c3 --no-extras --steps 20 --repeats 3 > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err; tail -c 300 $O/r2_bench_c3.err; head -c 300 $O/r2_bench_c3.json; echo
echo "== workload c5 (dim sweep)"; timeout 1200 python bench.py --workload c5 > $O/r2_bench_c5.json 2> $O/r2_bench_c5.err; tail -c 300 $O/r2_bench_c5.err
python - <<'PY'
import json
try:
  d = json.loads(open("gpurun_out/r2_bench_c5.json").read().strip().splitlines()[-1])
  for r in d["rows"]:
    print("c5 dim", r["dim"], r["fids"], "lookup ms", round(r["lookup_ms"], 4), "frac", round(r["lookup_frac"], 3), "lookup+update ms", round(r["lookup_update_ms"], 4), "frac", round(r["lookup_update_frac"], 3))
except Exception as e:
  print("c5 unreadable:", e)
PY
for X in peer direct; do
  echo "== sharded step on one GPU, exchange=$X"
  timeout 600 python bench.py --sharded --exchange $X --no-cpu-baseline --no-e2e --no-extras --steps 10 --repeats 3 > $O/r2_bench_sharded_$X.json 2> $O/r2_bench_sharded_$X.err
  tail -c 400 $O/r2_bench_sharded_$X.err; head -c 400 $O/r2_bench_sharded_$X.json; echo
done
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_direct.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/r2_ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r2_launches_sharded_direct.csv python bench.py --sharded --exchange direct --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/r2_ncu_l2.log 2>&1
echo "== ncu full (one step of the single-GPU path)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fid_claim|claim_miss|radix|runs_|seg_reduce|tree_level|long_finish|lookup_kernel|lookup_tma|upsert_fin" -s 22 -c 17 -o $O/r2_step_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/r2_ncu_full.log 2>&1; tail -2 $O/r2_ncu_full.log
MONO_LOOKUP_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lookup_tma" -s 8 -c 2 -o $O/r2_tma_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/r2_ncu_tma.log 2>&1; tail -2 $O/r2_ncu_tma.log
echo "== done"
