"""World-size-2 gloo test of the FID-hash sharded exchange (monolith_b200/distributed_ps.py) on CPU.

The exchange logic (dedup/bucket layout, count / FID / row / grad all-to-alls, split sizes, offsets)
is the product code; the five compute steps are served by the CPU oracle so that the test runs
without a GPU.  Expected results come from ONE global oracle table fed the union of both ranks'
batches — the sharded run must pool the same rows and leave the same per-key state
(ref protocol: native_training/distributed_ps_test.py:787-882, shard membership = fid mod N).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _configs():
  from tests.helpers import table
  return {"a": table([(8, "adagrad", {})], [0.1]), "b": table([(1, "ftrl", {}), (4, "sgd", {})], [0.05, 0.2])}


class OracleBackend:

  def __init__(self, tbl):
    self.t = tbl

  def reorder(self, fids_list, num_shards, dims):
    from tests import orc
    out, ss, sl, _, offs = orc.reorder_by_indices([f.numpy() for f in fids_list], num_shards, dims)
    return torch.from_numpy(out), ss.tolist(), sl.tolist(), torch.from_numpy(offs)

  def fused_lookup(self, ids, slot_sizes, num_shards):
    return torch.from_numpy(self.t.fused_lookup(ids.numpy(), slot_sizes, num_shards)[0])

  def fused_apply(self, ids, slot_sizes, grads, num_shards, req_time):
    self.t.fused_apply_gradient(ids.numpy(), slot_sizes, grads.numpy(), num_shards, req_time=req_time)

  def gather_pool(self, rows, offs, dim, row_offsets, pooling, out):
    from tests import orc
    ro = None if row_offsets is None else row_offsets.numpy()
    return torch.from_numpy(orc.gather_pool(rows.numpy(), offs.numpy(), dim, ro, pooling))

  def gather_pool_grad_into(self, grad_buf, pooled_grad, offs, dim, row_offsets, pooling):
    from tests import orc
    ro = None if row_offsets is None else row_offsets.numpy()
    grad_buf += torch.from_numpy(orc.gather_pool_grad(pooled_grad.numpy(), offs.numpy(), dim, grad_buf.numel(), ro, pooling))

  def zeros(self, n, like):
    return torch.zeros(n, dtype=torch.float32)


def _batch(rank, step):
  rng = np.random.default_rng(100 * step + rank)
  fa = (np.int64(1) << 48) | rng.integers(0, 60, 90)
  lens = rng.integers(0, 4, 40)
  offs_b = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
  fb = (np.int64(2) << 48) | rng.integers(0, 25, int(offs_b[-1]))
  ga = rng.standard_normal((90, 8)).astype(np.float32)
  gb = rng.standard_normal((40, 5)).astype(np.float32)
  return fa, fb, offs_b, ga, gb


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from monolith_b200.distributed_ps import PartitionedHashTable
  from tests import orc
  cfg = _configs()
  shard = orc.OracleMultiHashTable(cfg)
  pht = PartitionedHashTable(None, world, rank, backend=OracleBackend(shard), dims=shard.dims, names=shard.names)
  res = []
  for step in range(3):
    fa, fb, offs_b, ga, gb = _batch(rank, step)
    pooled, ctx = pht.lookup({"a": torch.from_numpy(fa), "b": torch.from_numpy(fb)},
                             row_offsets={"b": torch.from_numpy(offs_b)}, pooling={"b": "mean"})
    res.append((pooled["a"].numpy().copy(), pooled["b"].numpy().copy()))
    pht.apply_gradients(ctx, {"a": torch.from_numpy(ga), "b": torch.from_numpy(gb)}, req_time=10 + step)
  state = {n: (shard.keys(n), shard.lookup_entry(n, shard.keys(n))) for n in shard.names}
  q.put((rank, res, state))
  dist.barrier()
  dist.destroy_process_group()


def test_sharded_exchange_matches_single_table():
  from tests import orc
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + (os.getpid() % 400)
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  got = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0

  # expected: one global table; within a step both ranks look up BEFORE either applies (sync step)
  glob = orc.OracleMultiHashTable(_configs())
  for step in range(3):
    batches = [_batch(r, step) for r in range(world)]
    for r, (fa, fb, offs_b, ga, gb) in enumerate(batches):
      np.testing.assert_array_equal(got[r][1][step][0], glob.lookup_pool("a", fa, None, "sum"))
      np.testing.assert_array_equal(got[r][1][step][1], glob.lookup_pool("b", fb, offs_b, "mean"))
    # owners apply the requesters' grads in requester order: rank 0's unique rows, then rank 1's
    for r, (fa, fb, offs_b, ga, gb) in enumerate(batches):
      for name, f, g, ro, pool, D in (("a", fa, ga, None, "sum", 8), ("b", fb, gb, offs_b, "mean", 5)):
        u, inv = orc.dedup(f)
        ug = orc.gather_pool_grad(g, inv * D, D, u.size * D, ro, pool).reshape(-1, D)
        glob.apply_gradients({name: (u, ug)}, req_time=10 + step)
  for name in glob.names:
    keys = glob.keys(name)
    for r in range(world):
      mine = keys[(keys.view(np.uint64) % np.uint64(world)) == r]
      k_r, e_r = got[r][2][name]
      np.testing.assert_array_equal(k_r, mine)  # shard membership = fid mod N
      np.testing.assert_array_equal(e_r.view(np.uint32), glob.lookup_entry(name, mine).view(np.uint32))


# ---- host-side logic of the NVLink peer-window exchange (ShardedStep) ---------------------------------
@pytest.mark.parametrize("N", [1, 2, 3, 8])
def test_exchange_plan_layouts_are_all_to_all_v(N):
  """exchange_plan(): simulate the three transfers of the peer-window step with numpy arrays as windows and
  check that every received list is the compact, requester-major concatenation an all-to-all-v would give,
  and that rows / gradients line up with the FIDs they belong to."""
  from monolith_b200.distributed_ps import exchange_plan
  rng = np.random.default_rng(N)
  cnt = rng.integers(0, 7, (N, N)).astype(np.int64)
  cnt[rng.integers(0, N), rng.integers(0, N)] = 0
  plans = [exchange_plan(cnt, r) for r in range(N)]
  # requester r's bucketed unique list: items tagged (r, owner, i)
  uniq = [[(r, o, i) for o in range(N) for i in range(cnt[r, o])] for r in range(N)]
  for r in range(N):
    for o in range(N):
      assert uniq[r][plans[r]["bucket_at"][o]:plans[r]["bucket_at"][o] + cnt[r, o]] == [(r, o, i) for i in range(cnt[r, o])]
  # 1. FIDs: requester r puts bucket o at seg_at[o] of owner o's ids_in
  ids_in = [[None] * int(cnt[:, o].sum()) for o in range(N)]
  for r in range(N):
    for o in range(N):
      for i in range(cnt[r, o]):
        dst = plans[r]["seg_at"][o] + i
        assert ids_in[o][dst] is None                         # no two requesters collide
        ids_in[o][dst] = uniq[r][plans[r]["bucket_at"][o] + i]
  for o in range(N):
    assert ids_in[o] == [(r, o, i) for r in range(N) for i in range(cnt[r, o])]     # compact, requester-major
    assert plans[o]["recv"].tolist() == cnt[:, o].tolist()
    assert plans[o]["recv_at"].tolist() == np.concatenate([[0], np.cumsum(cnt[:, o])])[:N].tolist()
  # 2. rows: owner o stores the row of its k-th received id into requester r's rows_in at rows_dst[r] + i
  rows_in = [[None] * len(uniq[r]) for r in range(N)]
  for o in range(N):
    k = 0
    for r in range(N):
      for i in range(cnt[r, o]):
        rows_in[r][plans[o]["rows_dst"][r] + i] = ("row of", ids_in[o][k])
        k += 1
  for r in range(N):
    assert rows_in[r] == [("row of", u) for u in uniq[r]]     # row j of rows_in belongs to unique FID j
  # 3. gradients: requester r stores the gradient of its j-th unique FID at seg_at[o] + i of owner o's grads_in
  grads_in = [[None] * len(ids_in[o]) for o in range(N)]
  for r in range(N):
    for o in range(N):
      for i in range(cnt[r, o]):
        grads_in[o][plans[r]["seg_at"][o] + i] = ("grad of", uniq[r][plans[r]["bucket_at"][o] + i])
  for o in range(N):
    assert grads_in[o] == [("grad of", u) for u in ids_in[o]]  # aligned with ids_in: fused_apply sees (id, grad) pairs


def _hostcounts_worker(rank, world, port, q):
  try:
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from monolith_b200.distributed_ps import HostCounts
    hx = HostCounts(world, rank, 3)
    seen = []
    for e in range(200):
      row = [e * 10 + rank, rank, e]
      if rank == 1 and e % 17 == 0:
        import time
        time.sleep(0.002)            # a slow rank: the fast one must wait, never read a stale row
      seen.append(hx.exchange(row).tolist())
    q.put((rank, seen))
    dist.barrier()
    dist.destroy_process_group()
  except Exception:
    import traceback
    q.put((rank, "ERROR: " + traceback.format_exc()))
    raise


def test_host_counts_exchange_two_ranks():
  """HostCounts (the /dev/shm count matrix that replaces the count all-to-all): 200 back-to-back exchanges on 2
  ranks, one of them randomly slow; every rank sees every rank's row of the same epoch."""
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29600 + (os.getpid() % 300)
  procs = [ctx.Process(target=_hostcounts_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  got = dict(q.get(timeout=120) for _ in range(world))
  for p in procs:
    p.join(timeout=60)
  for r in range(world):
    assert not isinstance(got[r], str), got[r]
    assert got[r] == [[[e * 10 + rr, rr, e] for rr in range(world)] for e in range(200)]
