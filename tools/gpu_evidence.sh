#!/bin/bash
# evidence session: optimizer / stream / tower tests, launch lists (direct + sharded at world 1), one full ncu capture of a step
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "optimizer or further or group_adagrad or tower or never_misses or two_host or smoke" > $O/ev_tests.log 2>&1; tail -3 $O/ev_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2b_launches_direct.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/ev_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r2b_launches_sharded_direct.csv python bench.py --sharded --exchange direct --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/ev_l2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fid_claim|claim_miss|radix|runs_|seg_reduce|tree_level|long_finish|lookup_kernel|upsert_fin" -s 22 -c 17 -o $O/r2b_step_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --no-extras --repeats 1 > $O/ev_full.log 2>&1; tail -2 $O/ev_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tower_grad" -s 2 -c 1 -o $O/r2b_tower_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extras --repeats 1 > $O/ev_tower.log 2>&1; tail -2 $O/ev_tower.log
ls -la $O/r2b_*
