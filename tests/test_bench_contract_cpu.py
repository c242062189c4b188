"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver
reads, the synthetic batches have the documented shape, and the algorithmic-byte formulas are SURVEY §8(d)'s."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_json_line():
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--keys", "50000", "--cpu-batch", "2048"], capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["metric"] == "fid_lookups_per_sec" and d["unit"] == "lookups/s"
  assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1
  assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
  assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
  assert d["e2e"] == {"value": d["value"], "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
  assert "workload" in d["config"] and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_quietly():
  env = dict(os.environ, RANK="1", WORLD_SIZE="2")
  out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
  assert out.returncode == 0 and out.stdout.strip() == ""


def test_batches_and_byte_formulas():
  import bench
  b = bench.make_batches(2, 1000, 5000, seed=3)
  assert len(b) == 2 and b[0].shape == (1000 * bench.SLOTS,) and b[0].dtype == np.int64
  slots = (b[0] >> 48) & 0x7FFF
  assert slots.reshape(-1, bench.SLOTS)[0].tolist() == list(range(1, bench.SLOTS + 1))      # sample-major, slot 1..S
  assert ((b[0] & ((1 << 48) - 1)) < 5000).all()
  assert not np.array_equal(b[0], b[1])
  again = bench.make_batches(2, 1000, 5000, seed=3)
  assert np.array_equal(b[0], again[0])                                                     # seeded
  M, U, D = 2_097_152, 500_000, 32
  assert bench.fwd_bytes(M, U, D) == 8 * M + U * (32 + 4 * D) + 4 * D * M
  assert bench.bwd_bytes(M, U, D) == 4 * D * M + U * (32 + 16 * D + 8)
  r = bench.make_batches(1, 4000, 5000, seed=1, remote=(0.875, 2, 0))[0]
  assert 0.82 < float((r % 2 == 1).mean()) < 0.93                                           # --remote-frac knob
