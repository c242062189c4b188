"""FID bit layout helpers — mirror of monolith/native_training/data/training_instance/cc/reader_util.h:31-69
(SURVEY §8 row a1).  FID v2 = slot:15 bits << 48 | signature:48 bits; v1 = slot:10 bits << 54 | signature:54 bits.
Work on Python ints, numpy int64 / uint64 arrays and torch int64 tensors alike (the arithmetic is done on the
two's-complement bit pattern, so FIDs with the top bit set — FID -1 is a legal key — behave as in the reference)."""
import numpy as np

FID_V1_MASK = (1 << 54) - 1
FID_V2_MASK = (1 << 48) - 1
MAX_SLOT_NUMBER = 1 << 15          # get_max_slot_number(), reader_util.h:61


def _wrap(x):
  """-> signed 64-bit (Python int) / unchanged array type."""
  if isinstance(x, int):
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >> 63 else x
  return x


def _u(x):
  """bit pattern as an unsigned quantity for shifts (arrays: logical shift via masks)."""
  return x & ((1 << 64) - 1) if isinstance(x, int) else x


def slot_id_v2(fid):
  """ref: slot_id_v2, reader_util.h:36-38: (fid >> 48) & 0x7fff."""
  if isinstance(fid, int):
    return (_u(fid) >> 48) & 0x7FFF
  return (fid >> 48) & 0x7FFF      # arithmetic shift then mask == logical shift then mask for the low 15 bits


def slot_id_v1(fid):
  """ref: slot_id_v1, reader_util.h:34: fid >> 54 on the unsigned pattern."""
  if isinstance(fid, int):
    return _u(fid) >> 54
  return (fid >> 54) & 0x3FF


def get_fid_v2(slot, signature):
  """ref: GetFidV2, reader_util.h:67-69."""
  if isinstance(slot, int) and isinstance(signature, int):
    return _wrap((slot << 48) | (signature & FID_V2_MASK))
  return (np.int64(1) * slot << 48) | (signature & FID_V2_MASK) if isinstance(signature, np.ndarray) else \
      (slot << 48) | (signature & FID_V2_MASK)


def get_fid_v1(slot, signature):
  """ref: GetFidV1, reader_util.h:63-65."""
  if isinstance(slot, int) and isinstance(signature, int):
    return _wrap((slot << 54) | (signature & FID_V1_MASK))
  return (slot << 54) | (signature & FID_V1_MASK)


def convert_fid_v1_to_v2(fid):
  """ref: convert_fid_v1_to_v2, reader_util.h:45-48: slot = fid >> 54, keep the low 48 signature bits."""
  return get_fid_v2(slot_id_v1(fid), fid)


def switch_slot_v2(fid, slot):
  """ref: switch_slot_v2, reader_util.h:55-57."""
  return get_fid_v2(slot, fid)
