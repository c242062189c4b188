"""Checkpoints in the reference's on-disk format — host side of MonolithMultiHashTableSave / Restore
(ref: RT/ops/multi_hash_table_save_restore_ops.cc:105-388; NT/multi_hash_table_ops.py:409-420).

Files per shard i of n:  `<basename>-%05d-of-%05d`       Snappy-compressed TFRecords of EntryDump, table after
                                                          table (table order = sorted names, like the resource)
                         `<basename>.meta-%05d-of-%05d`  plain TFRecords of MultiHashTableMetadata
The record / container / protobuf encoding lives in csrc/ckpt.cu (mono_ckpt_*); this module moves rows
between the table (device export / restore kernels) and those writers, and mirrors the reference's file
validation (file_utils.cc:34-80), shard-count choice (:251-259) and TTL filter at save (:214-221).
The reference partitions a table over the shard files by its internal cuckoo buckets (unpinned); here an
entry goes to shard (uint64)fid mod n — any partition restores to the same table.
"""
import ctypes as C
import glob
import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

_SHARD_RE = re.compile(r"-(\d{5})-of-(\d{5})$")


def sharded_file_name(basename: str, shard: int, nshards: int) -> str:
  """ref: GetShardedFileName, RT/ops/file_utils.cc:26-32."""
  return "%s-%05d-of-%05d" % (basename, shard, nshards)


def sharded_meta_file_name(basename: str, shard: int, nshards: int) -> str:
  """ref: GetShardedMetadataFileName, multi_hash_table_save_restore_ops.cc:108-113."""
  return "%s.meta-%05d-of-%05d" % (basename, shard, nshards)


def validate_sharded_files(basename: str, filenames: Sequence[str]) -> int:
  """ref: ValidateShardedFiles, RT/ops/file_utils.cc:34-80.  Returns nshards."""
  show: List[bool] = []
  for f in filenames:
    if not f.startswith(basename):
      raise ValueError(f"Filename {f} doesn't belong to {basename}")
    m = _SHARD_RE.fullmatch(f[len(basename):])
    if not m:
      continue  # invalid files are ignored
    shard, nshards = int(m.group(1)), int(m.group(2))
    if not show:
      show = [False] * nshards
    if nshards != len(show):
      raise ValueError(f"Filename {f} doesn't match nshards. {len(show)}")
    if shard >= nshards:
      raise ValueError(f"Shard {shard} exceeds {nshards} for {f}")
    show[shard] = True
  if not show:
    raise ValueError(f"There is no valid sharded files for {basename}")
  for i, s in enumerate(show):
    if not s:
      raise ValueError(f"Shard {i} doesn't show up for {basename}")
  return len(show)


def segment_array(table_config):
  """ctypes mono_segment_cfg[] of one table (what the EntryDump codec needs: dims and optimizer kinds)."""
  segs = (_lib.SegmentCfg * len(table_config.segments))()
  for j, s in enumerate(table_config.segments):
    segs[j].dim = int(s.dim_size)
    segs[j].init_type = s.initializer.init_type
    segs[j].opt_type = s.optimizer.opt_type
  return segs


def expire_days_by_slot(table_config) -> np.ndarray:
  """int64[32768] expire time in days per slot (ref: slot_to_expire_time_, save op ctor :124-130)."""
  days = np.full(1 << 15, int(table_config.default_expire_time), np.int64)
  for slot, d in table_config.slot_expire_times.items():
    days[int(slot)] = int(d)
  return days


def _ck(rc):
  if rc != 0:
    raise RuntimeError("checkpoint: " + _lib.load().mono_ckpt_last_error().decode())


class ShardWriter:
  """One (data, meta) file pair."""

  def __init__(self, basename: str, shard: int, nshards: int, snappy: bool = True):
    self._h = None
    self._lib = _lib.load()
    h = C.c_void_p()
    _ck(self._lib.mono_ckpt_writer_open(sharded_file_name(basename, shard, nshards).encode(),
                                        sharded_meta_file_name(basename, shard, nshards).encode(),
                                        1 if snappy else 0, C.byref(h)))
    self._h = h

  def begin_table(self, name: str, segs):
    _ck(self._lib.mono_ckpt_writer_begin_table(self._h, name.encode(), segs, len(segs)))

  def add(self, ids: np.ndarray, rows: np.ndarray, max_update_ts: int = 0, expire_days: Optional[np.ndarray] = None) -> int:
    ids = np.ascontiguousarray(ids, np.int64)
    rows = np.ascontiguousarray(rows, np.float32)
    n = C.c_int64(0)
    ed = None if expire_days is None else np.ascontiguousarray(expire_days, np.int64).ctypes.data_as(C.c_void_p)
    _ck(self._lib.mono_ckpt_writer_add(self._h, ids.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.c_void_p),
                                       ids.size, int(max_update_ts), ed, C.byref(n)))
    return n.value

  def end_table(self):
    _ck(self._lib.mono_ckpt_writer_end_table(self._h))

  def close(self, commit: bool = True):
    if self._h:
      h, self._h = self._h, None
      _ck(self._lib.mono_ckpt_writer_close(h, 1 if commit else 0))

  def __del__(self):
    try:
      self.close(False)
    except Exception:
      pass


class ShardReader:

  def __init__(self, basename: str, shard: int, nshards: int, snappy: bool = True):
    self._h = None
    self._lib = _lib.load()
    h = C.c_void_p()
    _ck(self._lib.mono_ckpt_reader_open(sharded_file_name(basename, shard, nshards).encode(),
                                        sharded_meta_file_name(basename, shard, nshards).encode(),
                                        1 if snappy else 0, C.byref(h)))
    self._h = h

  def next_table(self) -> Optional[Tuple[str, int]]:
    name = C.create_string_buffer(1024)
    n, has = C.c_int64(0), C.c_int32(0)
    _ck(self._lib.mono_ckpt_reader_next_table(self._h, name, 1024, C.byref(n), C.byref(has)))
    return (name.value.decode(), n.value) if has.value else None

  def read(self, segs, width: int, max_n: int) -> Tuple[np.ndarray, np.ndarray]:
    ids = np.empty(max_n, np.int64)
    rows = np.empty((max_n, width), np.float32)
    n = C.c_int64(0)
    _ck(self._lib.mono_ckpt_reader_read(self._h, segs, len(segs), ids.ctypes.data_as(C.c_void_p),
                                        rows.ctypes.data_as(C.c_void_p), max_n, C.byref(n)))
    return ids[:n.value], rows[:n.value]

  def close(self):
    if self._h:
      h, self._h = self._h, None
      self._lib.mono_ckpt_reader_close(h)

  def __del__(self):
    self.close()


def pick_nshards(nshards: int, total_size: int) -> int:
  """ref: PickNshards, multi_hash_table_save_restore_ops.cc:251-259."""
  if nshards > 0:
    return nshards
  return int(min(4, max(1, total_size // 1000000)))


def save(table, basename: str, nshards: int = -1, snappy: bool = True, chunk: int = 1 << 18) -> Dict[str, int]:
  """Writes every live, non-expired entry of every table.  Returns entries written per table."""
  import torch
  dirname = os.path.dirname(basename)
  if dirname:
    os.makedirs(dirname, exist_ok=True)  # ref :153-154
  names = list(table.table_names)
  n = pick_nshards(nshards, sum(table.size(t) for t in names) if nshards <= 0 else 0)
  writers = [ShardWriter(basename, i, n, snappy) for i in range(n)]
  written = {}
  try:
    for name in names:
      cfg = table.configs[name].table_config
      segs = segment_array(cfg)
      days = expire_days_by_slot(cfg)
      max_ts = table.max_update_ts(name)
      for w in writers:
        w.begin_table(name, segs)
      cnt = 0
      for ids_d, rows_d in table.export(name, chunk=chunk):
        ids, rows = ids_d.cpu().numpy(), rows_d.cpu().numpy()
        if n == 1:
          cnt += writers[0].add(ids, rows, max_ts, days)
        else:
          shard = (ids.view(np.uint64) % np.uint64(n)).astype(np.int64)
          for i, w in enumerate(writers):
            m = shard == i
            if m.any():
              cnt += w.add(ids[m], rows[m], max_ts, days)
      for w in writers:
        w.end_table()
      written[name] = cnt
    for w in writers:
      w.close(True)
  except Exception:
    for w in writers:
      w.close(False)
    raise
  return written


def restore(table, basename: str, snappy: bool = True, chunk: int = 1 << 18) -> Dict[str, int]:
  """Upserts every entry of the checkpoint whose table exists here (others are skipped, ref :371-381).
  Returns entries read per table."""
  import torch
  files = glob.glob(glob.escape(basename) + "-*")
  n = validate_sharded_files(basename, files)
  names = set(table.table_names)
  read: Dict[str, int] = {}
  for i in range(n):
    r = ShardReader(basename, i, n, snappy)
    try:
      while True:
        nt = r.next_table()
        if nt is None:
          break
        name, _ = nt
        if name not in names:
          continue
        cfg = table.configs[name].table_config
        segs = segment_array(cfg)
        width = table.entry_width(name)
        max_ts = 0
        while True:
          ids, rows = r.read(segs, width, chunk)
          if ids.size == 0:
            break
          max_ts = max(max_ts, int(rows[:, -1].view(np.uint32).max()))
          table.restore_rows(name, torch.from_numpy(ids), torch.from_numpy(rows))
          read[name] = read.get(name, 0) + ids.size
          if ids.size < chunk:
            break
        table.note_update_ts(name, max_ts)
    finally:
      r.close()
  return read
