"""GPU debug: does the pipelined claim kernel insert duplicate keys?  (table size vs distinct FIDs, per step)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monolith_b200 import MultiHashTable, _lib, entry
from tests.helpers import table
lib = _lib.load()
dev = torch.device("cuda", 0)
for mode in (0, 1):
  lib.mono_set_option(b"claim_pipeline", mode)
  rng = np.random.default_rng(7)
  t = MultiHashTable({"t": table([(32, "adagrad", {})], [0.05], capacity=256, init=entry.RandomUniformInitializer(-0.1, 0.1), init_seed=5)}, device=dev)
  seen = set()
  for step in range(4):
    n = 200000
    ids = rng.integers(0, 50000, n)
    ids[rng.random(n) < 0.3] = rng.integers(0, 40, int((rng.random(n) < 0.3).sum()))[:1].repeat(1)[0]
    fids = (np.int64(7) << 48) | ids.astype(np.int64)
    seen.update(fids.tolist())
    pg = rng.standard_normal((n, 32)).astype(np.float32)
    t.pool_backward("t", torch.from_numpy(fids).to(dev), torch.from_numpy(pg).to(dev), None, "sum", req_time=10 + step)
    ks = np.concatenate([ids_.cpu().numpy() for ids_, _ in t.export("t", chunk=1 << 16)])
    print("mode", mode, "step", step, "size", t.size("t"), "distinct", len(seen), "exported", ks.size, "unique exported", np.unique(ks).size, flush=True)
