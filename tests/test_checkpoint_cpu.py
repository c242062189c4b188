"""Checkpoint files in the reference's format (csrc/ckpt.cu, monolith_b200/checkpoint.py) — CPU tests.

What pins what:
  * crc32c: RFC 3720 B.4 known answers; TFRecord framing: an independent bit-wise crc32c + struct packing here
  * Snappy codec: pyarrow's Snappy (both directions, literal / 1-, 2-, 4-byte-offset copies)
  * EntryDump / OptimizerDump / MultiHashTableMetadata wire bytes: the protobuf runtime, with message descriptors
    transcribed from embedding_hash_table.proto:45-50,139-142 and optimizer.proto:28-30,56-57,69-72,130-135,232-252
    (byte-for-byte equal to what the protobuf runtime serializes, and decoding packed and unpacked floats)
  * file layout: multi_hash_table_save_restore_ops.cc (names, metadata counts, TTL filter, unknown tables skipped)
The Snappy block container (u32 BE compressed length | raw block, per 256 KiB input buffer) is restated from TF's
snappy_outputbuffer.cc / snappy_inputbuffer.cc and has no TF-written fixture: unpinned.
"""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from monolith_b200 import _lib, checkpoint as ck

OPT_SGD, OPT_ADAGRAD, OPT_FTRL, OPT_ADAM = _lib.OPT_SGD, _lib.OPT_ADAGRAD, _lib.OPT_FTRL, _lib.OPT_ADAM


def lib():
  return _lib.load()


def segs_of(spec):
  arr = (_lib.SegmentCfg * len(spec))()
  for i, (dim, opt) in enumerate(spec):
    arr[i].dim, arr[i].opt_type = dim, opt
  return arr


def state_floats(spec):
  return sum({OPT_SGD: 0, OPT_ADAGRAD: d, OPT_FTRL: 2 * d, OPT_ADAM: 2 * d + 2}[o] for d, o in spec)


def crc32c_ref(data: bytes) -> int:
  c = 0xFFFFFFFF
  for b in data:
    c ^= b
    for _ in range(8):
      c = (c >> 1) ^ (0x82F63B78 & -(c & 1))
  return c ^ 0xFFFFFFFF


def masked_ref(data: bytes) -> int:
  c = crc32c_ref(data)
  return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def test_crc32c_known_answers():
  L = lib()
  crc = lambda b: L.mono_ckpt_crc32c(b, len(b))
  assert crc(b"123456789") == 0xE3069283
  assert crc(bytes(32)) == 0x8A9136AA              # RFC 3720 B.4
  assert crc(b"\xff" * 32) == 0x62A8AB43
  assert crc(bytes(range(32))) == 0x46DD794E
  assert crc(bytes(range(31, -1, -1))) == 0x113FDB5C
  rng = np.random.default_rng(0)
  for n in (0, 1, 7, 8, 9, 63, 1000):
    b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    assert crc(b) == crc32c_ref(b)
    assert L.mono_ckpt_masked_crc32c(b, len(b)) == masked_ref(b)


@pytest.mark.parametrize("kind", ["empty", "one", "text", "zeros", "random", "periodic_far", "mixed"])
def test_snappy_against_pyarrow(kind):
  pa = pytest.importorskip("pyarrow")
  codec = pa.Codec("snappy")
  rng = np.random.default_rng(7)
  data = {
      "empty": b"", "one": b"x", "text": b"the quick brown fox jumps over the lazy dog. " * 400,
      "zeros": bytes(300000), "random": rng.integers(0, 256, 200000, dtype=np.uint8).tobytes(),
      # period 70000 > 65535: only 4-byte-offset copies can express it (pyarrow side), ours falls back to literals
      "periodic_far": (rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()) * 3,
      "mixed": b"".join(rng.integers(0, 4, 50, dtype=np.uint8).tobytes() + rng.integers(0, 256, 13, dtype=np.uint8).tobytes()
                        for _ in range(3000)),
  }[kind]
  L = lib()
  out = C.create_string_buffer(len(data) + len(data) // 6 + 64)
  n = L.mono_ckpt_snappy_compress(data, len(data), out, len(out))
  assert n > 0
  ours = out.raw[:n]
  assert codec.decompress(ours, len(data), asbytes=True) == data if data else True       # ours -> pyarrow
  if kind in ("text", "zeros"):
    assert n < len(data) // 3                                                          # it does compress
  theirs = codec.compress(data, asbytes=True)                                           # pyarrow -> ours
  back = C.create_string_buffer(max(len(data), 1))
  m = L.mono_ckpt_snappy_uncompress(theirs, len(theirs), back, len(back))
  assert m == len(data) and back.raw[:m] == data
  m = L.mono_ckpt_snappy_uncompress(ours, n, back, len(back))
  assert m == len(data) and back.raw[:m] == data
  if len(theirs) > 8:                                                                   # corruption is reported
    assert L.mono_ckpt_snappy_uncompress(theirs[:-3], len(theirs) - 3, back, len(back)) < 0


# ---- protobuf messages transcribed from the reference .proto files ---------------------------------
def _proto_classes():
  pytest.importorskip("google.protobuf")
  from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
  F = descriptor_pb2.FieldDescriptorProto
  fd = descriptor_pb2.FileDescriptorProto(name="mono_ckpt_test.proto", package="t", syntax="proto2")

  def msg(name, fields):
    m = fd.message_type.add(name=name)
    for fname, num, typ, label, tname, packed in fields:
      f = m.field.add(name=fname, number=num, type=typ, label=label)
      if tname:
        f.type_name = ".t." + tname
      if packed:
        f.options.packed = True
    return m

  OPT, REP = F.LABEL_OPTIONAL, F.LABEL_REPEATED
  msg("AdagradOptimizerDump", [("norm", 1, F.TYPE_FLOAT, REP, None, False)])
  msg("SgdOptimizerDump", [])
  msg("FtrlOptimizerDump", [("zero", 1, F.TYPE_FLOAT, REP, None, False), ("norm", 2, F.TYPE_FLOAT, REP, None, False)])
  msg("AdamOptimizerDump", [("m", 1, F.TYPE_FLOAT, REP, None, False), ("v", 2, F.TYPE_FLOAT, REP, None, False),
                            ("beta1_power", 3, F.TYPE_FLOAT, OPT, None, False), ("beta2_power", 4, F.TYPE_FLOAT, OPT, None, False)])
  msg("SingleOptimizerDump", [("adagrad", 1, F.TYPE_MESSAGE, OPT, "AdagradOptimizerDump", False),
                              ("sgd", 2, F.TYPE_MESSAGE, OPT, "SgdOptimizerDump", False),
                              ("ftrl", 3, F.TYPE_MESSAGE, OPT, "FtrlOptimizerDump", False),
                              ("adam", 7, F.TYPE_MESSAGE, OPT, "AdamOptimizerDump", False)])
  msg("OptimizerDump", [("dump", 1, F.TYPE_MESSAGE, REP, "SingleOptimizerDump", False)])
  for name, packed in (("EntryDump", False), ("EntryDumpPacked", True)):
    msg(name, [("id", 1, F.TYPE_SFIXED64, OPT, None, False), ("num", 2, F.TYPE_FLOAT, REP, None, packed),
               ("opt", 3, F.TYPE_MESSAGE, OPT, "OptimizerDump", False),
               ("last_update_ts_sec", 4, F.TYPE_INT64, OPT, None, False)])
  msg("MultiHashTableMetadata", [("table_name", 1, F.TYPE_STRING, OPT, None, False),
                                 ("num_entries", 2, F.TYPE_UINT64, OPT, None, False)])
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fd)
  get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("t." + n))
  return {n: get(n) for n in ("EntryDump", "EntryDumpPacked", "MultiHashTableMetadata")}


SPEC = [(3, OPT_ADAGRAD), (2, OPT_SGD), (4, OPT_FTRL), (2, OPT_ADAM)]


def _row(rng, spec, ts):
  dim, st = sum(d for d, _ in spec), state_floats(spec)
  row = rng.standard_normal(dim + st + 2).astype(np.float32)
  row[dim + st:] = np.array([1, ts], np.uint32).view(np.float32)
  return row


def _fill_proto(m, fid, row, spec, ts):
  dim = sum(d for d, _ in spec)
  m.id = fid
  m.num.extend(row[:dim].tolist())
  st = row[dim:]
  for d, o in spec:
    s = m.opt.dump.add()
    if o == OPT_ADAGRAD:
      s.adagrad.norm.extend(st[:d].tolist()); st = st[d:]
    elif o == OPT_SGD:
      s.sgd.SetInParent()
    elif o == OPT_FTRL:
      s.ftrl.norm.extend(st[:d].tolist()); s.ftrl.zero.extend(st[d:2 * d].tolist()); st = st[2 * d:]
    else:
      s.adam.m.extend(st[:d].tolist()); s.adam.v.extend(st[d:2 * d].tolist())
      s.adam.beta1_power, s.adam.beta2_power = float(st[2 * d]), float(st[2 * d + 1]); st = st[2 * d + 2:]
  m.last_update_ts_sec = ts


def test_entry_dump_wire_bytes_match_protobuf():
  cls = _proto_classes()
  L, rng = lib(), np.random.default_rng(3)
  segs = segs_of(SPEC)
  for fid, ts in ((5 << 48 | 77, 1700000000), (-3, 0), (2**63 - 1, 2**32 - 1)):
    row = _row(rng, SPEC, ts)
    buf = C.create_string_buffer(4096)
    n = L.mono_ckpt_encode_entry(segs, len(SPEC), fid, row.ctypes.data_as(C.c_void_p), buf, len(buf))
    assert n > 0
    m = cls["EntryDump"]()
    _fill_proto(m, fid, row, SPEC, ts)
    assert buf.raw[:n] == m.SerializeToString()                      # byte for byte what protobuf writes
    for variant in ("EntryDump", "EntryDumpPacked"):                 # and we read both float encodings
      p = cls[variant]()
      _fill_proto(p, fid, row, SPEC, ts)
      wire = p.SerializeToString()
      out = np.zeros_like(row)
      fid_out = C.c_int64(0)
      assert L.mono_ckpt_decode_entry(segs, len(SPEC), wire, len(wire), C.byref(fid_out), out.ctypes.data_as(C.c_void_p)) == 0
      assert fid_out.value == fid
      np.testing.assert_array_equal(out.view(np.uint32), row.view(np.uint32))
  # missing optional fields: zeros / ts 0 (ref: restore sets last_update_ts_sec = 0 when absent, :377-379)
  m = cls["EntryDump"]()
  m.id = 9
  wire = m.SerializeToString()
  out = np.ones(sum(d for d, _ in SPEC) + state_floats(SPEC) + 2, np.float32)
  fid_out = C.c_int64(0)
  assert L.mono_ckpt_decode_entry(segs, len(SPEC), wire, len(wire), C.byref(fid_out), out.ctypes.data_as(C.c_void_p)) == 0
  assert fid_out.value == 9 and not out[:-2].any() and out[-2:].view(np.uint32).tolist() == [1, 0]
  assert L.mono_ckpt_decode_entry(segs, len(SPEC), b"\x0a\xff", 2, C.byref(fid_out), out.ctypes.data_as(C.c_void_p)) != 0


def _parse_tfrecords(stream: bytes):
  recs, p = [], 0
  while p < len(stream):
    (n,) = struct.unpack_from("<Q", stream, p)
    assert struct.unpack_from("<I", stream, p + 8)[0] == masked_ref(stream[p:p + 8])
    data = stream[p + 12:p + 12 + n]
    assert struct.unpack_from("<I", stream, p + 12 + n)[0] == masked_ref(data)
    recs.append(data)
    p += 16 + n
  return recs


@pytest.mark.parametrize("snappy", [True, False])
def test_file_layout_and_roundtrip(tmp_path, snappy):
  pa = pytest.importorskip("pyarrow")
  cls = _proto_classes()
  rng = np.random.default_rng(11)
  base = str(tmp_path / "ckpt" / "model.ckpt-100-emb")
  os.makedirs(os.path.dirname(base))
  tables = {"item": [(8, OPT_ADAGRAD)], "user": SPEC}
  n_rows = {"item": 9000, "user": 700}     # item: > 256 KiB of records -> several Snappy blocks
  data, nshards = {}, 2
  writers = [ck.ShardWriter(base, i, nshards, snappy) for i in range(nshards)]
  for name in sorted(tables):
    spec = tables[name]
    ids = rng.choice(1 << 40, n_rows[name], replace=False).astype(np.int64) | (np.int64(3) << 48)
    rows = np.stack([_row(rng, spec, 1000 + int(i % 50)) for i in range(ids.size)])
    data[name] = (ids, rows)
    for i, w in enumerate(writers):
      w.begin_table(name, segs_of(spec))
      m = (ids.view(np.uint64) % np.uint64(nshards)) == i
      assert w.add(ids[m], rows[m]) == int(m.sum())
      w.end_table()
  for w in writers:
    w.close(True)
  assert sorted(os.listdir(os.path.dirname(base))) == [
      "model.ckpt-100-emb-00000-of-00002", "model.ckpt-100-emb-00001-of-00002",
      "model.ckpt-100-emb.meta-00000-of-00002", "model.ckpt-100-emb.meta-00001-of-00002"]   # no temporaries left
  assert ck.validate_sharded_files(base, [os.path.join(os.path.dirname(base), f) for f in os.listdir(os.path.dirname(base))]) == 2
  # independent decode of shard 0: container -> TFRecords -> protobuf
  raw = open(ck.sharded_file_name(base, 0, nshards), "rb").read()
  if snappy:
    codec, stream, p, nblocks = pa.Codec("snappy"), b"", 0, 0
    while p < len(raw):
      # TF SnappyOutputBuffer layout: u32 BE compressed length | raw Snappy block (uncompressed size = Snappy varint preamble)
      clen, = struct.unpack_from(">I", raw, p)
      blk = raw[p + 4:p + 4 + clen]
      ulen, shift, q = 0, 0, 0
      while True:
        ulen |= (blk[q] & 0x7F) << shift
        shift += 7
        q += 1
        if not blk[q - 1] & 0x80:
          break
      assert ulen <= 256 * 1024                          # RecordWriter's Snappy input buffer
      stream += codec.decompress(blk, ulen, asbytes=True)
      p += 4 + clen
      nblocks += 1
    assert nblocks > 1 and len(raw) < len(stream)
  else:
    stream = raw
  recs = _parse_tfrecords(stream)
  metas = [cls["MultiHashTableMetadata"].FromString(r) for r in _parse_tfrecords(open(ck.sharded_meta_file_name(base, 0, nshards), "rb").read())]
  assert [m.table_name for m in metas] == ["item", "user"] and sum(m.num_entries for m in metas) == len(recs)
  off = 0
  for m in metas:
    ids, rows = data[m.table_name]
    sel = (ids.view(np.uint64) % np.uint64(nshards)) == 0
    assert m.num_entries == int(sel.sum())
    dim = sum(d for d, _ in tables[m.table_name])
    for fid, row, rec in zip(ids[sel], rows[sel], recs[off:off + m.num_entries]):
      e = cls["EntryDump"].FromString(rec)
      assert e.id == fid and e.last_update_ts_sec == int(row[-1:].view(np.uint32)[0])
      np.testing.assert_array_equal(np.array(e.num, np.float32), row[:dim])
      assert len(e.opt.dump) == len(tables[m.table_name])
    off += m.num_entries
  # our reader: every table back, bit for bit; a table the reader does not ask for is skipped
  got = {}
  for i in range(nshards):
    r = ck.ShardReader(base, i, nshards, snappy)
    while True:
      nt = r.next_table()
      if nt is None:
        break
      name, cnt = nt
      if name == "item" and i == 1:
        continue                                   # skip: next_table must step over its records
      spec = tables[name]
      width = sum(d for d, _ in spec) + state_floats(spec) + 2
      parts = []
      while True:
        ids, rows = r.read(segs_of(spec), width, 256)
        if ids.size == 0:
          break
        parts.append((ids.copy(), rows.copy()))
      assert sum(p[0].size for p in parts) == cnt
      got.setdefault(name, []).extend(parts)
    r.close()
  for name in tables:
    ids = np.concatenate([p[0] for p in got[name]])
    rows = np.concatenate([p[1] for p in got[name]])
    want_ids, want_rows = data[name]
    if name == "item":
      keep = (want_ids.view(np.uint64) % np.uint64(nshards)) == 0
      want_ids, want_rows = want_ids[keep], want_rows[keep]
    o, wo = np.argsort(ids), np.argsort(want_ids)
    np.testing.assert_array_equal(ids[o], want_ids[wo])
    np.testing.assert_array_equal(rows[o].view(np.uint32), want_rows[wo].view(np.uint32))


def test_ttl_filter_at_save_and_errors(tmp_path):
  base = str(tmp_path / "t")
  spec = [(2, OPT_SGD)]
  w = ck.ShardWriter(base, 0, 1)
  w.begin_table("t", segs_of(spec))
  ids = np.array([(1 << 48) | 1, (1 << 48) | 2, (2 << 48) | 3, (2 << 48) | 4], np.int64)
  rows = np.zeros((4, 4), np.float32)
  day = 24 * 3600
  rows[:, -1] = np.array([100, 100 + 5 * day, 100, 100 + 5 * day], np.uint32).view(np.float32)
  days = np.full(1 << 15, 36500, np.int64)
  days[1], days[2] = 5, 6
  # ref :214-221: dropped when max_update_ts - ts >= expire_days(slot) * 86400
  assert w.add(ids, rows, max_update_ts=100 + 5 * day, expire_days=days) == 3      # slot 1, ts 100: exactly 5 days -> dropped
  assert w.add(ids, rows, max_update_ts=100 + 6 * day, expire_days=days) == 2      # slot 2, ts 100: 6 days -> dropped too
  with pytest.raises(RuntimeError):
    w.close(True)                                                                   # table not ended
  assert not any(f.startswith("t") for f in os.listdir(tmp_path))                   # nothing committed, temporaries removed
  with pytest.raises(RuntimeError):
    ck.ShardReader(base, 0, 1)
  with pytest.raises(ValueError):
    ck.validate_sharded_files(base, [base + "-00000-of-00002"])                     # shard 1 missing
  with pytest.raises(ValueError):
    ck.validate_sharded_files(base, [base + "-00000-of-00002", base + "-00001-of-00003"])
  assert ck.validate_sharded_files(base, [base + "-00000-of-00001", base + "-junk", base + ".meta-00000-of-00001"]) == 1
  assert ck.pick_nshards(-1, 10) == 1 and ck.pick_nshards(-1, 2_500_000) == 2 and ck.pick_nshards(-1, 10**9) == 4
  assert ck.pick_nshards(3, 0) == 3
  # a corrupted data file is reported, not silently accepted
  w = ck.ShardWriter(base, 0, 1, snappy=False)
  w.begin_table("t", segs_of(spec)); w.add(ids, rows); w.end_table(); w.close(True)
  path = ck.sharded_file_name(base, 0, 1)
  blob = bytearray(open(path, "rb").read())
  blob[20] ^= 0x40
  open(path, "wb").write(bytes(blob))
  r = ck.ShardReader(base, 0, 1, snappy=False)
  assert r.next_table() == ("t", 4)
  with pytest.raises(RuntimeError):
    r.read(segs_of(spec), 4, 16)


@pytest.mark.parametrize("case", [
    # (default_expire_days, {slot: days}, ids, values, write_ts, expected lookup after save -> restore)
    dict(name="ttl_zero", default=0, slots={}, ids=[-1, 1], values=[1.0, 2.0], ts=0, expect=[0.0, 0.0]),          # hash_table_ops_test.py:381-395
    dict(name="ttl_not_zero", default=3600, slots={}, ids=[-1, 1], values=[1.0, 2.0], ts=0, expect=[1.0, 2.0]),   # :397-411
    dict(name="ttl_by_slots", default=3600, slots={1: 0, 2: 1}, ids=[1 << 48, 2 << 48], values=[1.0, 2.0], ts=100,
         expect=[0.0, 2.0]),                                                                                       # :413-466
], ids=lambda c: c["name"])
def test_reference_ttl_goldens_through_the_files(tmp_path, case):
  """The reference's save-time TTL tests, replayed at the file level: entries written at `ts` into a table whose
  max_update_ts is `ts`; what a restore would find = what the reader returns (absent -> lookup gives 0).
  FID -1 is a legal key and falls in slot 0x7fff (reader_util.h:36-38)."""
  spec = [(1, OPT_SGD)]
  days = np.full(1 << 15, case["default"], np.int64)
  for s_, d in case["slots"].items():
    days[s_] = d
  ids = np.array(case["ids"], np.int64)
  rows = np.zeros((ids.size, 3), np.float32)
  rows[:, 0] = case["values"]
  rows[:, 1:] = np.array([[1, case["ts"]]] * ids.size, np.uint32).view(np.float32)
  base = str(tmp_path / "table")
  w = ck.ShardWriter(base, 0, 1)
  w.begin_table("t", segs_of(spec))
  w.add(ids, rows, max_update_ts=case["ts"], expire_days=days)
  w.end_table()
  w.close(True)
  r = ck.ShardReader(base, 0, 1)
  name, cnt = r.next_table()
  got_ids, got_rows = r.read(segs_of(spec), 3, 16)
  assert name == "t" and cnt == got_ids.size
  table = {int(i): float(v) for i, v in zip(got_ids, got_rows[:, 0])}
  assert [table.get(int(i), 0.0) for i in ids] == case["expect"]
  assert r.next_table() is None
  # saving what was restored and restoring again changes nothing (the second half of :440-466)
  w = ck.ShardWriter(base + "_new", 0, 1)
  w.begin_table("t", segs_of(spec))
  w.add(got_ids, got_rows, max_update_ts=case["ts"], expire_days=days)
  w.end_table()
  w.close(True)
  r2 = ck.ShardReader(base + "_new", 0, 1)
  r2.next_table()
  ids2, rows2 = r2.read(segs_of(spec), 3, 16)
  np.testing.assert_array_equal(ids2, got_ids)
  np.testing.assert_array_equal(rows2.view(np.uint32), got_rows.view(np.uint32))


def test_restore_not_found(tmp_path):
  """ref: test_restore_not_found, hash_table_ops_test.py:468-476: restoring a basename without files is an error."""
  with pytest.raises(ValueError):
    ck.validate_sharded_files(str(tmp_path / "nothing_here"), [])


def test_snappy_and_entry_codec_fuzz():
  """Property checks (hypothesis): any byte string survives our Snappy codec in both directions against pyarrow's,
  and any float / id bit pattern survives the EntryDump codec (NaN payloads, infinities, denormals, -0.0 included)."""
  hyp = pytest.importorskip("hypothesis")
  st = pytest.importorskip("hypothesis.strategies")
  pa = pytest.importorskip("pyarrow")
  codec, L = pa.Codec("snappy"), lib()

  @hyp.settings(max_examples=150, deadline=None)
  @hyp.given(st.one_of(st.binary(max_size=3000),
                       st.builds(lambda a, n, b: a * n + b, st.binary(min_size=1, max_size=40), st.integers(1, 200), st.binary(max_size=20))))
  def snappy_roundtrip(data):
    out = C.create_string_buffer(len(data) + len(data) // 6 + 64)
    n = L.mono_ckpt_snappy_compress(data, len(data), out, len(out))
    assert n > 0
    if data:
      assert codec.decompress(out.raw[:n], len(data), asbytes=True) == data
    theirs = codec.compress(data, asbytes=True)
    back = C.create_string_buffer(max(len(data), 1))
    assert L.mono_ckpt_snappy_uncompress(theirs, len(theirs), back, len(back)) == len(data)
    assert back.raw[:len(data)] == data

  snappy_roundtrip()

  segs = segs_of(SPEC)
  width = sum(d for d, _ in SPEC) + state_floats(SPEC) + 2

  @hyp.settings(max_examples=150, deadline=None)
  @hyp.given(st.integers(-2**63, 2**63 - 1), st.lists(st.integers(0, 2**32 - 1), min_size=width - 2, max_size=width - 2),
             st.integers(0, 2**32 - 1))
  def entry_roundtrip(fid, bits, ts):
    row = np.array(bits + [1, ts], np.uint32).view(np.float32)
    buf = C.create_string_buffer(8192)
    n = L.mono_ckpt_encode_entry(segs, len(SPEC), fid, row.ctypes.data_as(C.c_void_p), buf, len(buf))
    assert n > 0
    out = np.zeros(width, np.float32)
    fid_out = C.c_int64(0)
    assert L.mono_ckpt_decode_entry(segs, len(SPEC), buf.raw[:n], n, C.byref(fid_out), out.ctypes.data_as(C.c_void_p)) == 0
    assert fid_out.value == fid and out.view(np.uint32).tolist() == row.view(np.uint32).tolist()

  entry_roundtrip()

  @hyp.settings(max_examples=100, deadline=None)
  @hyp.given(st.binary(max_size=200))
  def decoder_never_crashes_on_garbage(blob):
    out = np.zeros(width, np.float32)
    fid_out = C.c_int64(0)
    L.mono_ckpt_decode_entry(segs, len(SPEC), blob, len(blob), C.byref(fid_out), out.ctypes.data_as(C.c_void_p))   # any status
    back = C.create_string_buffer(4096)
    L.mono_ckpt_snappy_uncompress(blob, len(blob), back, len(back))                                                 # any status

  decoder_never_crashes_on_garbage()


def test_further_optimizer_dumps_wire_bytes_match_protobuf():
  """Momentum / RMSprop / RMSpropV2 / Adadelta / AMSGrad state in EntryDump: byte for byte what the protobuf runtime
  serializes for the reference's messages (optimizer.proto: MomentumOptimizerDump = SingleOptimizerDump field 9 {n},
  Rmsprop = 11 {n}, RmspropV2 = 12 {n}, Adadelta = 6 {accum, accum_update}, Amsgrad = 8 {m, v, vhat, beta powers})."""
  pytest.importorskip("google.protobuf")
  from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
  F = descriptor_pb2.FieldDescriptorProto
  fd = descriptor_pb2.FileDescriptorProto(name="mono_ckpt_more.proto", package="u", syntax="proto2")
  OPT, REP = F.LABEL_OPTIONAL, F.LABEL_REPEATED

  def msg(name, fields):
    m = fd.message_type.add(name=name)
    for fname, num, typ, label, tname in fields:
      f = m.field.add(name=fname, number=num, type=typ, label=label)
      if tname:
        f.type_name = ".u." + tname

  fl = lambda n, i, lab=REP: (n, i, F.TYPE_FLOAT, lab, None)
  msg("MomentumOptimizerDump", [fl("n", 1)])
  msg("RmspropOptimizerDump", [fl("n", 1)])
  msg("RmspropV2OptimizerDump", [fl("n", 1)])
  msg("AdadeltaOptimizerDump", [fl("accum", 1), fl("accum_update", 2)])
  msg("AmsgradOptimizerDump", [fl("m", 1), fl("v", 2), fl("vhat", 3), fl("beta1_power", 4, OPT), fl("beta2_power", 5, OPT)])
  msg("SingleOptimizerDump", [("adadelta", 6, F.TYPE_MESSAGE, OPT, "AdadeltaOptimizerDump"),
                              ("amsgrad", 8, F.TYPE_MESSAGE, OPT, "AmsgradOptimizerDump"),
                              ("momentum", 9, F.TYPE_MESSAGE, OPT, "MomentumOptimizerDump"),
                              ("rmsprop", 11, F.TYPE_MESSAGE, OPT, "RmspropOptimizerDump"),
                              ("rmspropv2", 12, F.TYPE_MESSAGE, OPT, "RmspropV2OptimizerDump")])
  msg("OptimizerDump", [("dump", 1, F.TYPE_MESSAGE, REP, "SingleOptimizerDump")])
  msg("EntryDump", [("id", 1, F.TYPE_SFIXED64, OPT, None), ("num", 2, F.TYPE_FLOAT, REP, None),
                    ("opt", 3, F.TYPE_MESSAGE, OPT, "OptimizerDump"), ("last_update_ts_sec", 4, F.TYPE_INT64, OPT, None)])
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fd)
  Entry = message_factory.GetMessageClass(pool.FindMessageTypeByName("u.EntryDump"))
  MOM, RMS, RMS2, ADD, AMS = _lib.OPT_MOMENTUM, _lib.OPT_RMSPROP, _lib.OPT_RMSPROPV2, _lib.OPT_ADADELTA, _lib.OPT_AMSGRAD
  spec = [(2, MOM), (3, RMS), (1, RMS2), (2, ADD), (3, AMS)]
  nstate = {MOM: 1, RMS: 1, RMS2: 1, ADD: 2}
  dim = sum(d for d, _ in spec)
  st_n = sum(3 * d + 2 if o == AMS else nstate[o] * d for d, o in spec)
  rng = np.random.default_rng(8)
  row = rng.standard_normal(dim + st_n + 2).astype(np.float32)
  fid, ts = (9 << 48) | 12345, 1700000123
  row[dim + st_n:] = np.array([1, ts], np.uint32).view(np.float32)
  L, segs = lib(), segs_of(spec)
  buf = C.create_string_buffer(4096)
  n = L.mono_ckpt_encode_entry(segs, len(spec), fid, row.ctypes.data_as(C.c_void_p), buf, len(buf))
  assert n > 0
  m = Entry()
  m.id = fid
  m.num.extend(row[:dim].tolist())
  st = row[dim:]
  for d, o in spec:
    s = m.opt.dump.add()
    if o == MOM:
      s.momentum.n.extend(st[:d].tolist()); st = st[d:]
    elif o == RMS:
      s.rmsprop.n.extend(st[:d].tolist()); st = st[d:]
    elif o == RMS2:
      s.rmspropv2.n.extend(st[:d].tolist()); st = st[d:]
    elif o == ADD:
      s.adadelta.accum.extend(st[:d].tolist()); s.adadelta.accum_update.extend(st[d:2 * d].tolist()); st = st[2 * d:]
    else:
      s.amsgrad.m.extend(st[:d].tolist()); s.amsgrad.v.extend(st[d:2 * d].tolist()); s.amsgrad.vhat.extend(st[2 * d:3 * d].tolist())
      s.amsgrad.beta1_power, s.amsgrad.beta2_power = float(st[3 * d]), float(st[3 * d + 1]); st = st[3 * d + 2:]
  m.last_update_ts_sec = ts
  wire = m.SerializeToString()
  assert buf.raw[:n] == wire
  out = np.zeros_like(row)
  fid_out = C.c_int64(0)
  assert L.mono_ckpt_decode_entry(segs, len(spec), wire, len(wire), C.byref(fid_out), out.ctypes.data_as(C.c_void_p)) == 0
  assert fid_out.value == fid
  np.testing.assert_array_equal(out.view(np.uint32), row.view(np.uint32))


def test_group_adagrad_and_moving_average_dump_wire_bytes():
  """GroupAdaGrad state = SingleOptimizerDump field 15 {grad_square_sum = 1 (float)} (optimizer.proto:100-102,245);
  MovingAverage::Save returns an EMPTY OptimizerDump (moving_average_optimizer.cc:54-57), so a moving-average segment
  contributes no SingleOptimizerDump at all and the following segments' dumps move up."""
  pytest.importorskip("google.protobuf")
  from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
  F = descriptor_pb2.FieldDescriptorProto
  fd = descriptor_pb2.FileDescriptorProto(name="mono_ckpt_group.proto", package="g", syntax="proto2")
  OPT, REP = F.LABEL_OPTIONAL, F.LABEL_REPEATED

  def msg(name, fields):
    m = fd.message_type.add(name=name)
    for fname, num, typ, label, tname in fields:
      f = m.field.add(name=fname, number=num, type=typ, label=label)
      if tname:
        f.type_name = ".g." + tname

  msg("AdagradOptimizerDump", [("norm", 1, F.TYPE_FLOAT, REP, None)])
  msg("GroupAdaGradOptimizerDump", [("grad_square_sum", 1, F.TYPE_FLOAT, OPT, None)])
  msg("SingleOptimizerDump", [("adagrad", 1, F.TYPE_MESSAGE, OPT, "AdagradOptimizerDump"),
                              ("group_adagrad", 15, F.TYPE_MESSAGE, OPT, "GroupAdaGradOptimizerDump")])
  msg("OptimizerDump", [("dump", 1, F.TYPE_MESSAGE, REP, "SingleOptimizerDump")])
  msg("EntryDump", [("id", 1, F.TYPE_SFIXED64, OPT, None), ("num", 2, F.TYPE_FLOAT, REP, None),
                    ("opt", 3, F.TYPE_MESSAGE, OPT, "OptimizerDump"), ("last_update_ts_sec", 4, F.TYPE_INT64, OPT, None)])
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fd)
  Entry = message_factory.GetMessageClass(pool.FindMessageTypeByName("g.EntryDump"))
  GA, MA, AG = _lib.OPT_GROUP_ADAGRAD, _lib.OPT_MOVING_AVERAGE, _lib.OPT_ADAGRAD
  spec = [(3, GA), (2, MA), (2, AG)]            # state: 1 | 0 | 2
  dim, st_n = 7, 3
  rng = np.random.default_rng(4)
  row = rng.standard_normal(dim + st_n + 2).astype(np.float32)
  fid, ts = (5 << 48) | 77, 1700000999
  row[dim + st_n:] = np.array([1, ts], np.uint32).view(np.float32)
  L, segs = lib(), segs_of(spec)
  buf = C.create_string_buffer(2048)
  n = L.mono_ckpt_encode_entry(segs, len(spec), fid, row.ctypes.data_as(C.c_void_p), buf, len(buf))
  assert n > 0
  m = Entry()
  m.id = fid
  m.num.extend(row[:dim].tolist())
  m.opt.dump.add().group_adagrad.grad_square_sum = float(row[dim])
  m.opt.dump.add().adagrad.norm.extend(row[dim + 1:dim + 3].tolist())
  m.last_update_ts_sec = ts
  wire = m.SerializeToString()
  assert buf.raw[:n] == wire
  out = np.zeros_like(row)
  fid_out = C.c_int64(0)
  assert L.mono_ckpt_decode_entry(segs, len(spec), wire, len(wire), C.byref(fid_out), out.ctypes.data_as(C.c_void_p)) == 0
  assert fid_out.value == fid
  np.testing.assert_array_equal(out.view(np.uint32), row.view(np.uint32))
