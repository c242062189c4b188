"""GPU debug: extreme same-key contention in the claim kernel (ONE FID = 30 % of the batch): duplicates? torn reads?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from monolith_b200 import MultiHashTable, _lib, entry
from tests.helpers import table
lib = _lib.load()
dev = torch.device("cuda", 0)
for confirm, cg in ((0, 0), (1, 0), (0, 1)):
  lib.mono_set_option(b"claim_pipeline", confirm)   # = confirm knob
  lib.mono_set_option(b"claim_cg", cg)
  dups = []
  for trial in range(40):
    rng = np.random.default_rng(7 + trial)
    t = MultiHashTable({"t": table([(32, "adagrad", {})], [0.05], capacity=256, init=entry.RandomUniformInitializer(-0.1, 0.1), init_seed=5)}, device=dev)
    n = 200000
    ids = rng.integers(0, 50000, n)
    ids[rng.random(n) < 0.3] = int(rng.integers(0, 40))
    fids = (np.int64(7) << 48) | ids.astype(np.int64)
    for step in range(2):
      t.pool_backward("t", torch.from_numpy(fids).to(dev), torch.randn(n, 32, device=dev), None, "sum", req_time=10 + step)
    dups.append(t.size("t") - np.unique(fids).size)
    t.close()
  print("confirm", confirm, "cg", cg, "duplicates in 40 trials:", sum(dups), "trials with dups:", sum(1 for d in dups if d), "torn reads seen:", lib.mono_get_option(b"claim_torn"), flush=True)
