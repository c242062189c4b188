"""A model of the lock-free cuckoo insert / lookup protocol of csrc/common.cuh (cuckoo_insert) and csrc/rowops.cuh
(probe_lane + probe_lane_confirm_miss), explored over MANY interleavings on the CPU.

Why: the protocol's claims — (1) concurrent inserters never leave a key in the table twice and never lose one, (2) a
reader on another stream never misses a key that is resident for the whole lookup — are about races that a GPU stress test
hits once in 1e9 steps.  (It paid for itself at once: the first version of the in-flight check read the move counter before
the in-flight count and did not bump the counter when a carried entry was put back; schedule 79787 of the random search
showed a reader declaring a resident key absent.)  Here every atomic memory operation of every actor (128-bit CAS / load on a slot, atomic add on a
counter) is one scheduling point and the schedules are enumerated: all schedules with at most two pre-emptions, plus
random ones.  The actors below restate the device code step by step (same order of atomics, same decisions); the table is
tiny (3-4 buckets of 2 slots) so that displacements, marks, the take-back path and the exchange fallback all happen.

This is a model of the ALGORITHM (test infrastructure, like oracle/): it does not execute the CUDA code.  The CUDA code is
exercised by tests/test_gpu_parity.py (`test_lookup_never_misses_during_inserts_on_another_stream`, the streaming and
growth tests)."""
import itertools
import random
import zlib

import pytest

SLOTS = 2          # slots per bucket in the model (4 on the device)
EMPTY = None
COVER = {"copy_first": 0, "take_back": 0, "exchange": 0, "stash": 0, "skip_marked": 0, "reader_confirm": 0, "reader_retry": 0}


class Mem:
  """Shared memory: slots hold (key, marked) or EMPTY; counters; a 2-entry stash."""

  def __init__(self, nb, homes):
    self.slots = [EMPTY] * (nb * SLOTS)
    self.stash = [EMPTY, EMPTY]
    self.moves = 0
    self.inflight = 0
    self.stash_count = 0
    self.homes = homes          # key -> (b1, b2)
    self.error = 0

  def bucket(self, b):
    return range(b * SLOTS, (b + 1) * SLOTS)


def cas(mem, where, idx, expect, new):
  arr = mem.slots if where == "s" else mem.stash
  old = arr[idx]
  if old == expect:
    arr[idx] = new
  return old


def inserter(mem, key, victim_order):
  """cuckoo_insert(e = key).  Every `yield` precedes exactly one atomic operation on shared memory.
  victim_order: the order in which the victims of a full bucket are tried (the device derives it from a hash)."""
  e = (key, False)
  b1, b2 = mem.homes[key]
  cur = b1
  carrying = False

  def placed():
    """The carried entry is in the table again: move counter first, then the in-flight count (two atomics)."""
    if carrying:
      mem.moves += 1
      yield
      mem.inflight -= 1

  for it in range(6):          # kMaxEvictions (48 on the device)
    # empty slots of cur (and of the alternate bucket on the first iteration)
    for which in range(2 if it == 0 else 1):
      b = cur if which == 0 else (b2 if cur == b1 else b1)
      for i in mem.bucket(b):
        yield
        o = mem.slots[i]
        if o is EMPTY:
          yield
          if cas(mem, "s", i, EMPTY, e) is EMPTY:
            yield
            yield from placed()
            return
    # copy-first displacement
    rescan = False
    order = [cur * SLOTS + ((victim_order + k) % SLOTS) for k in range(SLOTS)]
    for vp in order:
      yield
      v = mem.slots[vp]
      if v is EMPTY:
        yield
        if cas(mem, "s", vp, EMPTY, e) is EMPTY:
          yield
          yield from placed()
          return
        rescan = True
        break
      if v[1]:
        COVER["skip_marked"] += 1
        continue               # another mover owns this victim
      marked = (v[0], True)
      yield
      if cas(mem, "s", vp, v, marked) != v:
        rescan = True
        break
      a1, a2 = mem.homes[v[0]]
      va = a2 if cur == a1 else a1
      copied = False
      for i in mem.bucket(va):
        yield
        o = mem.slots[i]
        if o is EMPTY:
          yield
          if cas(mem, "s", i, EMPTY, v) is EMPTY:
            copied = True
            break
      if copied:
        yield
        mem.moves += 1
      want = e if copied else v
      yield
      old = cas(mem, "s", vp, marked, want)
      done = old == marked
      if copied and done:
        COVER["copy_first"] += 1
        yield
        yield from placed()
        return
      if copied:
        mem.error |= 4
      else:
        COVER["take_back"] += 1
    if rescan:
      continue
    # exchange fallback on the first victim
    vp = order[0]
    yield
    victim = mem.slots[vp]
    if victim is not EMPTY and victim[1]:
      continue
    first_carry = (not carrying) and victim is not EMPTY
    if first_carry:
      yield
      mem.inflight += 1
    yield
    mem.moves += 1
    yield
    if cas(mem, "s", vp, victim, e) != victim:
      if first_carry:
        yield
        mem.inflight -= 1
      continue
    if victim is EMPTY:
      yield
      yield from placed()
      return
    carrying = True
    COVER["exchange"] += 1
    e = victim
    b1, b2 = mem.homes[e[0]]
    cur = b2 if cur == b1 else b1
  # stash: the count goes up before the entry becomes visible there
  yield
  mem.stash_count += 1
  COVER["stash"] += 1
  for i in range(len(mem.stash)):
    yield
    if mem.stash[i] is EMPTY:
      yield
      if cas(mem, "t", i, EMPTY, e) is EMPTY:
        yield
        yield from placed()
        return
  mem.error |= 1
  yield
  yield from placed()


def probe(mem, key, order):
  """probe_lane: bucket `order[0]` then `order[1]`, then the stash when its count is non-zero.  One load per slot."""
  for b in order:
    for i in mem.bucket(b):
      yield
      o = mem.slots[i]
      if o is not EMPTY and o[0] == key:
        return True
  yield
  if mem.stash_count != 0:
    for i in range(len(mem.stash)):
      yield
      o = mem.stash[i]
      if o is not EMPTY and o[0] == key:
        return True
  return False


def reader(mem, key, result):
  """Fast probe, then probe_lane_confirm_miss: coherent probes bracketed by reads of the move / in-flight counters."""
  b1, b2 = mem.homes[key]
  found = yield from probe(mem, key, (b1, b2))
  if found:
    result.append(True)
    return
  yield
  ma = mem.moves
  COVER["reader_confirm"] += 1
  for _ in range(6):
    found = yield from probe(mem, key, (b1, b2))
    if found:
      result.append(True)
      return
    yield
    fl = mem.inflight          # in-flight count BEFORE the move counter (see probe_lane_confirm_miss)
    yield
    mb = mem.moves
    if mb == ma and fl == 0:
      result.append(False)
      return
    COVER["reader_retry"] += 1
    ma = mb
  result.append(None)          # gave up (bounded retries): not a miss verdict in the model


def run(schedule_fn, build):
  """Runs the actors of `build()` under a scheduler: schedule_fn(step, runnable ids) -> id."""
  mem, actors, check = build()
  gens = {i: a for i, a in enumerate(actors)}
  step = 0
  while gens:
    ids = sorted(gens)
    pick = schedule_fn(step, ids)
    try:
      next(gens[pick])
    except StopIteration:
      del gens[pick]
    step += 1
    assert step < 5000, "livelock in the model"
  check(mem)


def final_invariants(mem, all_keys):
  assert mem.error == 0, mem.error
  assert mem.inflight == 0
  seen = [s[0] for s in mem.slots if s is not EMPTY] + [s[0] for s in mem.stash if s is not EMPTY]
  assert sorted(seen) == sorted(all_keys), (seen, all_keys)        # every key exactly once: none lost, none twice
  assert not any(s[1] for s in mem.slots if s is not EMPTY)         # no mark left behind
  for i, s in enumerate(mem.slots):                                 # every key in one of ITS buckets
    if s is not EMPTY:
      assert i // SLOTS in mem.homes[s[0]]


def scenario(kind, reader_key, reader_home_swapped, vorders):
  """B0 and B1 are full, new keys x (and y) hash to (B0, B1).  kind 'free': B2 has room for displaced entries (copy-first
  path); 'tight': B2 has ONE free slot (second mover takes the mark back / uses the exchange fallback, chains, stash)."""
  def build():
    homes = {"a": (0, 2), "b": (0, 2), "c": (1, 2), "d": (1, 2), "x": (0, 1), "y": (1, 0), "e": (2, 3), "f": (3, 2), "g": (3, 2)}
    nb = 4
    mem = Mem(nb, dict(homes))
    mem.slots[0], mem.slots[1] = ("a", False), ("b", False)
    mem.slots[2], mem.slots[3] = ("c", False), ("d", False)
    resident = ["a", "b", "c", "d"]
    if kind == "tight":
      mem.slots[4] = ("e", False)
      mem.slots[6], mem.slots[7] = ("f", False), ("g", False)     # e's alternate bucket is full too: chains go on
      resident += ["e", "f", "g"]
    if reader_home_swapped:                                        # the reader's key lives in its SECOND bucket
      b1, b2 = mem.homes[reader_key]
      mem.homes[reader_key] = (b2, b1)
    result = []
    actors = [inserter(mem, "x", vorders[0]), inserter(mem, "y", vorders[1]), reader(mem, reader_key, result)]

    def check(m):
      final_invariants(m, resident + ["x", "y"])
      assert result and result[0] is not False, f"the reader missed resident key {reader_key}"
    return mem, actors, check
  return build


def scenario_three_movers(reader_key, vorders):
  """Three inserters and a reader on five buckets: x -> (B0, B1) and y -> (B1, B0) displace into B2, z -> (B2, B3) arrives in
  B2 while the copies land there — the constellation in which a plain copy-then-overwrite displacement leaves a key twice
  (two movers on one victim, a third moving the first copy)."""
  def build():
    homes = {"a": (0, 2), "b": (0, 2), "c": (1, 2), "d": (1, 2), "e": (2, 3), "f": (3, 4), "g": (3, 4),
             "x": (0, 1), "y": (1, 0), "z": (2, 3)}
    mem = Mem(5, dict(homes))
    for i, k in enumerate(["a", "b", "c", "d", "e", None, "f", "g"]):
      mem.slots[i] = (k, False) if k else EMPTY
    resident = ["a", "b", "c", "d", "e", "f", "g"]
    result = []
    actors = [inserter(mem, "x", vorders[0]), inserter(mem, "y", vorders[1]), inserter(mem, "z", vorders[2]),
              reader(mem, reader_key, result)]

    def check(m):
      final_invariants(m, resident + ["x", "y", "z"])
      assert result and result[0] is not False, f"the reader missed resident key {reader_key}"
    return mem, actors, check
  return build


def bounded_schedules(n_actors, max_preemptions, horizon):
  """All schedules that run one actor until it finishes (or until a chosen step), with <= max_preemptions switches at
  chosen steps: the classic context-bounded exploration.  Yields scheduler functions."""
  for order in itertools.permutations(range(n_actors)):
    for k in range(max_preemptions + 1):
      for points in itertools.combinations(range(1, horizon), k):
        for targets in itertools.product(range(n_actors), repeat=k):
          def fn(step, ids, order=order, points=points, targets=targets, state={}):
            if step == 0:
              state["cur"] = None
            if step in points:
              t = targets[points.index(step)]
              if t in ids:
                state["cur"] = t
            if state.get("cur") not in ids:
              state["cur"] = next(o for o in order if o in ids)
            return state["cur"]
          yield fn


@pytest.mark.parametrize("kind", ["free", "tight"])
@pytest.mark.parametrize("reader_key,swapped", [("a", False), ("a", True), ("c", True), ("e", False)])
def test_protocol_random_schedules(kind, reader_key, swapped):
  if kind == "free" and reader_key == "e":
    pytest.skip("key e is resident only in the tight scenario")
  rng = random.Random(zlib.crc32(repr((kind, reader_key, swapped)).encode()))   # stable across processes
  for trial in range(2500):
    vorders = (rng.randrange(SLOTS), rng.randrange(SLOTS))
    # bursty random scheduler: stays on an actor for a random number of steps (long races AND fine interleavings)
    state = {"cur": None, "left": 0}

    def fn(step, ids, state=state):
      if state["cur"] not in ids or state["left"] == 0:
        state["cur"] = rng.choice(ids)
        state["left"] = rng.choice((1, 1, 2, 3, 5, 8, 13))
      state["left"] -= 1
      return state["cur"]
    run(fn, scenario(kind, reader_key, swapped, vorders))


@pytest.mark.parametrize("reader_key", ["a", "e"])
def test_protocol_three_movers_random_schedules(reader_key):
  rng = random.Random(zlib.crc32(reader_key.encode()) + 3)
  for trial in range(3000):
    vorders = tuple(rng.randrange(SLOTS) for _ in range(3))
    state = {"cur": None, "left": 0}

    def fn(step, ids, state=state):
      if state["cur"] not in ids or state["left"] == 0:
        state["cur"] = rng.choice(ids)
        state["left"] = rng.choice((1, 1, 2, 3, 5, 8, 13))
      state["left"] -= 1
      return state["cur"]
    run(fn, scenario_three_movers(reader_key, vorders))


@pytest.mark.parametrize("kind", ["free", "tight"])
def test_protocol_all_schedules_with_two_preemptions(kind):
  n = 0
  for vorders in ((0, 0), (0, 1)):
    for fn in bounded_schedules(3, 2, 40):
      run(fn, scenario(kind, "a", True, vorders))
      n += 1
  assert n > 5000


def test_every_path_of_the_protocol_was_exercised():
  """Runs last in this file: the schedules above went through the copy-first path, the take-back of a mark, the exchange
  fallback with a carried entry, the stash, a mover skipping a marked victim, and readers that had to confirm / retry."""
  if COVER["copy_first"] == 0:      # run alone: generate the coverage first
    test_protocol_all_schedules_with_two_preemptions("tight")
    test_protocol_random_schedules("tight", "a", True)
  assert all(v > 0 for v in COVER.values()), COVER


def test_model_detects_a_broken_protocol():
  """The checker is not vacuous: a plain copy-then-overwrite displacement WITHOUT the move counter lets a reader miss a
  resident key (it reads the alternate bucket before the copy and the old slot after the overwrite)."""
  def build():
    homes = {"a": (2, 0), "b": (0, 2), "x": (0, 1), "c": (1, 2), "d": (1, 2)}
    mem = Mem(3, homes)
    mem.slots[0], mem.slots[1] = ("a", False), ("b", False)
    mem.slots[2], mem.slots[3] = ("c", False), ("d", False)
    result = []

    def naive_mover():
      yield
      mem.slots[4] = ("a", False)      # copy the victim to its alternate bucket ...
      yield
      mem.slots[0] = ("x", False)      # ... and overwrite the old slot; no counter is bumped

    def naive_reader():
      found = yield from probe(mem, "a", (2, 0))
      result.append(found)
    actors = [naive_mover(), naive_reader()]

    def check(m):
      assert result[0], "missed"
    return mem, actors, check

  failures = 0
  for fn in bounded_schedules(2, 2, 8):
    try:
      run(fn, build)
    except AssertionError:
      failures += 1
  assert failures > 0
