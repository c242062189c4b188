"""GPU debug: does a claim-kernel variant ever insert duplicate keys?  (table size vs distinct FIDs)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monolith_b200 import MultiHashTable, _lib, entry
from tests.helpers import table
lib = _lib.load()
dev = torch.device("cuda", 0)
for pipe, cg in ((0, 0), (0, 1), (1, 0), (1, 1)):
  lib.mono_set_option(b"claim_pipeline", pipe)
  lib.mono_set_option(b"claim_cg", cg)
  dups, t_ms = 0, []
  for trial in range(12):
    rng = np.random.default_rng(100 + trial)
    t = MultiHashTable({"t": table([(32, "adagrad", {})], [0.05], capacity=1 << 21, init=entry.RandomUniformInitializer(-0.1, 0.1), init_seed=5)}, device=dev)
    n = 1 << 20
    ids = rng.integers(0, 400000, n)
    hot = rng.random(n) < 0.3
    ids[hot] = rng.integers(0, 40, int(hot.sum()))
    fids = (np.int64(7) << 48) | ids.astype(np.int64)
    f = torch.from_numpy(fids).to(dev)
    pg = torch.randn(n, 32, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t.pool_backward("t", f, pg, None, "sum", req_time=10)
    torch.cuda.synchronize()
    t_ms.append(1e3 * (time.perf_counter() - t0))
    dups += t.size("t") - np.unique(fids).size
    t.close()
  print("pipeline", pipe, "cg", cg, "duplicates over 12 first-steps:", dups, "ms", round(float(np.median(t_ms)), 3), flush=True)
