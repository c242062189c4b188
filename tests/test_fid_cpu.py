"""FID bit layout helpers (monolith_b200/fid.py) against the formulas of reader_util.h:31-69 (row a1)."""
import numpy as np
import torch

from monolith_b200 import fid


def test_scalar_known_values():
  assert fid.get_fid_v2(7, 0x123456789) == (7 << 48) | 0x123456789
  assert fid.slot_id_v2((7 << 48) | 0x123456789) == 7
  assert fid.get_fid_v2(3, -1) == (3 << 48) | ((1 << 48) - 1)              # signature masked to 48 bits
  assert fid.slot_id_v2(-1) == 0x7FFF                                      # FID -1 is legal: slot 32767
  assert fid.get_fid_v2(0x7FFF, 5) == ((0x7FFF << 48) | 5)
  assert fid.get_fid_v2(0x8000, 5) == -(1 << 63) + 5                       # bit 63: wraps to a negative int64
  assert fid.slot_id_v2(fid.get_fid_v2(0x8000, 5)) == 0                    # ... and the 16th bit is not part of the slot
  assert fid.MAX_SLOT_NUMBER == 32768
  v1 = fid.get_fid_v1(991, 0xABCDEF0123)
  assert fid.slot_id_v1(v1) == 991 and fid.convert_fid_v1_to_v2(v1) == (991 << 48) | 0xABCDEF0123
  big = fid.get_fid_v1(5, (1 << 53) | 77)                                  # v1 signature bits above 48 are dropped
  assert fid.convert_fid_v1_to_v2(big) == (5 << 48) | 77
  assert fid.switch_slot_v2((7 << 48) | 99, 12) == (12 << 48) | 99


def test_arrays_and_tensors_match_scalars():
  rng = np.random.default_rng(0)
  f = rng.integers(-2**63, 2**63 - 1, 1000).astype(np.int64)
  want = [fid.slot_id_v2(int(x)) for x in f]
  assert fid.slot_id_v2(f).tolist() == want
  assert fid.slot_id_v2(torch.from_numpy(f)).tolist() == want
  slots = rng.integers(0, 1 << 15, 1000).astype(np.int64)
  sig = rng.integers(0, 1 << 48, 1000).astype(np.int64)
  want = [fid.get_fid_v2(int(s), int(g)) for s, g in zip(slots, sig)]
  assert fid.get_fid_v2(slots, sig).tolist() == want
  assert fid.get_fid_v2(torch.from_numpy(slots), torch.from_numpy(sig)).tolist() == want
  assert fid.slot_id_v2(fid.get_fid_v2(slots, sig)).tolist() == slots.tolist()
