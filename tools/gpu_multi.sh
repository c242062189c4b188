#!/bin/bash
# multi-GPU session: distributed parity tests + bench at N GPUs for the peer (host-driven) and direct (device-driven) exchange
cd "$(dirname "$0")/.."
N=${1:-2}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/r2_smi_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_distributed.py -q -v > $O/r2_tests_2gpu.log 2>&1; tail -12 $O/r2_tests_2gpu.log | cut -c1-200
fi
P=29511
for X in ${2:-peer direct}; do
  P=$((P+1))
  echo "== bench N=$N exchange=$X"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --exchange $X --steps 20 --repeats 3 ${3:-} > $O/r2_bench_n${N}_$X.json 2> $O/r2_bench_n${N}_$X.err
  tail -c 500 $O/r2_bench_n${N}_$X.err
  python - <<PY
import json
try:
  d = json.loads(open("$O/r2_bench_n${N}_$X.json").read().strip().splitlines()[-1])
  print("N=$N $X: ms/step", round(d["ms_per_step"], 4), "G lookups/s", round(d["value"] / 1e9, 3), "parity", d.get("parity_check"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
  print("unreadable:", e)
PY
done
if [ -n "$4" ]; then   # C4 streaming insert+evict through the sharded step: $4 = resident keys per GPU
  echo "== c4 N=$N keys/GPU=$4"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --workload c4 --gpus $N --keys $4 --steps 16 --repeats 2 > $O/r2_c4_n${N}.json 2> $O/r2_c4_n${N}.err
  tail -c 400 $O/r2_c4_n${N}.err; cut -c1-1600 $O/r2_c4_n${N}.json
fi
