"""Table / optimizer / initializer configuration.

Host-side mirror of monolith/native_training/entry.py (same class names and argument meaning,
ref: entry.py:54-75 SgdOptimizer, :77-113 AdagradOptimizer, :136-180 AdamOptimizer, :365-393
FtrlOptimizer, :404-442 initializers, :514-538 CombineAsSegment, :549-564 CuckooHashTableConfig,
:566-584 HashTableConfigInstance).  The reference serialises these into an
EmbeddingHashTableConfig proto (runtime/hash_table/embedding_hash_table.proto:23-96); here they
flatten into the plain structs of include/mono_emb.h.  Defaults are the proto defaults
(runtime/hash_table/optimizer/optimizer.proto:19-128, initializer/initializer_config.proto:19-45).
"""
import dataclasses
from typing import Callable, Dict, List, Optional, Sequence, Union

from . import _lib


class Optimizer:
  opt_type = -1

  def params(self) -> List[float]:
    raise NotImplementedError


def _d(v, default):
  return default if v is None else v


@dataclasses.dataclass
class SgdOptimizer(Optimizer):
  learning_rate: Optional[float] = None  # proto default 0.01
  opt_type = _lib.OPT_SGD

  def params(self):
    return [0.0] * 6


@dataclasses.dataclass
class AdagradOptimizer(Optimizer):
  learning_rate: Optional[float] = None  # proto default 0.001
  initial_accumulator_value: Optional[float] = None  # proto default 0.1
  hessian_compression_times: int = 1
  warmup_steps: int = 0
  weight_decay_factor: float = 0.0
  opt_type = _lib.OPT_ADAGRAD

  def params(self):
    if self.hessian_compression_times != 1:
      raise NotImplementedError("hessian_compression_times != 1 is outside the hot-path scope")
    return [_d(self.initial_accumulator_value, 0.1), self.weight_decay_factor, 0.0, 0.0, 0.0, 0.0]


@dataclasses.dataclass
class FtrlOptimizer(Optimizer):
  learning_rate: Optional[float] = None  # proto default 0.01
  initial_accumulator_value: Optional[float] = None  # 0.1
  beta: Optional[float] = None  # 0.0
  warmup_steps: int = 0
  l1_regularization: Optional[float] = None  # 0.0
  l2_regularization: Optional[float] = None  # 0.0
  opt_type = _lib.OPT_FTRL

  def params(self):
    return [_d(self.initial_accumulator_value, 0.1), _d(self.beta, 0.0),
            _d(self.l1_regularization, 0.0), _d(self.l2_regularization, 0.0), 0.0, 0.0]


@dataclasses.dataclass
class AdamOptimizer(Optimizer):
  learning_rate: Optional[float] = None  # 0.01
  beta1: float = 0.9
  beta2: float = 0.99
  use_beta1_warmup: bool = False
  weight_decay_factor: float = 0.0
  use_nesterov: bool = False
  epsilon: float = 0.01
  warmup_steps: int = 0
  opt_type = _lib.OPT_ADAM

  def params(self):
    if self.use_beta1_warmup:
      raise NotImplementedError("use_beta1_warmup is not read by the reference kernel either")
    return [self.beta1, self.beta2, self.epsilon, self.weight_decay_factor,
            1.0 if self.use_nesterov else 0.0, 0.0]


@dataclasses.dataclass
class MomentumOptimizer(Optimizer):
  """ref: NT/entry.py:227-245 / MomentumOptimizerConfig (optimizer.proto)."""
  learning_rate: Optional[float] = None  # 0.01
  weight_decay_factor: float = 0.0
  use_nesterov: bool = False
  momentum: float = 0.9
  warmup_steps: int = 0
  opt_type = _lib.OPT_MOMENTUM

  def params(self):
    return [self.momentum, self.weight_decay_factor, 1.0 if self.use_nesterov else 0.0, 0.0, 0.0, 0.0]


@dataclasses.dataclass
class RmspropOptimizer(Optimizer):
  """ref: NT/entry.py:259-271.  NOTE the reference kernel reads the CONFIG's learning rate, not the per-call one
  (rmsprop_optimizer.cc:62)."""
  learning_rate: Optional[float] = None  # 0.01
  weight_decay_factor: float = 0.0
  momentum: float = 0.9
  opt_type = _lib.OPT_RMSPROP

  def params(self):
    return [self.momentum, self.weight_decay_factor, 0.01 if self.learning_rate is None else self.learning_rate, 0.0, 0.0, 0.0]


@dataclasses.dataclass
class RmspropV2Optimizer(Optimizer):
  """ref: NT/entry.py:273-285."""
  learning_rate: Optional[float] = None  # 0.01
  weight_decay_factor: float = 0.0
  momentum: float = 0.9
  opt_type = _lib.OPT_RMSPROPV2

  def params(self):
    return [self.momentum, self.weight_decay_factor, 0.0, 0.0, 0.0, 0.0]


@dataclasses.dataclass
class AdadeltaOptimizer(Optimizer):
  """ref: NT/entry.py:115-134."""
  learning_rate: Optional[float] = None  # 0.01
  weight_decay_factor: float = 0.0
  averaging_ratio: float = 0.9
  epsilon: float = 0.01
  warmup_steps: int = 0
  opt_type = _lib.OPT_ADADELTA

  def params(self):
    return [self.averaging_ratio, self.epsilon, self.weight_decay_factor, 0.0, 0.0, 0.0]


@dataclasses.dataclass
class AmsgradOptimizer(Optimizer):
  """ref: NT/entry.py:182-205."""
  learning_rate: Optional[float] = None  # 0.01
  beta1: float = 0.9
  beta2: float = 0.99
  weight_decay_factor: float = 0.0
  use_nesterov: bool = False
  epsilon: float = 0.01
  warmup_steps: int = 0
  opt_type = _lib.OPT_AMSGRAD

  def params(self):
    return [self.beta1, self.beta2, self.epsilon, self.weight_decay_factor, 1.0 if self.use_nesterov else 0.0, 0.0]


@dataclasses.dataclass
class MovingAverageOptimizer(Optimizer):
  """ref: NT/entry.py:247-255 / moving_average_optimizer.cc:44-52: w = momentum * w + (1 - momentum) * grad; no state, the
  learning rate is ignored."""
  momentum: float = 0.9
  learning_rate: Optional[float] = None
  opt_type = _lib.OPT_MOVING_AVERAGE

  def params(self):
    return [self.momentum, 0.0, 0.0, 0.0, 0.0, 0.0]


@dataclasses.dataclass
class GroupAdaGradOptimizer(Optimizer):
  """ref: GroupAdaGradOptimizerConfig (optimizer.proto:90-98) / group_adagrad_optimizer.cc:50-89 (no Python wrapper in
  the reference's entry.py; reachable there through the proto).  The segment is one group."""
  learning_rate: Optional[float] = None  # 0.01
  beta: float = 0.0
  initial_accumulator_value: float = 0.1
  l2_regularization_strength: float = 0.0
  weight_decay_factor: float = 0.0
  warmup_steps: int = 0
  opt_type = _lib.OPT_GROUP_ADAGRAD

  def params(self):
    return [self.initial_accumulator_value, self.weight_decay_factor, self.beta, self.l2_regularization_strength, 0.0, 0.0]


_DEFAULT_LR = {_lib.OPT_MOVING_AVERAGE: 0.01, _lib.OPT_GROUP_ADAGRAD: 0.01, _lib.OPT_SGD: 0.01, _lib.OPT_ADAGRAD: 0.001, _lib.OPT_FTRL: 0.01, _lib.OPT_ADAM: 0.01, _lib.OPT_MOMENTUM: 0.01,
               _lib.OPT_RMSPROP: 0.01, _lib.OPT_RMSPROPV2: 0.01, _lib.OPT_ADADELTA: 0.01, _lib.OPT_AMSGRAD: 0.01}


class Initializer:
  init_type = _lib.INIT_ZEROS
  a = 0.0
  b = 0.0


class ZerosInitializer(Initializer):
  pass


class OnesInitializer(Initializer):
  init_type = _lib.INIT_ONES


class ConstantsInitializer(Initializer):
  init_type = _lib.INIT_CONSTANT

  def __init__(self, constant: float):
    self.a = float(constant)


class RandomUniformInitializer(Initializer):
  init_type = _lib.INIT_UNIFORM

  def __init__(self, minval=None, maxval=None):
    self.a = -0.05 if minval is None else float(minval)
    self.b = 0.05 if maxval is None else float(maxval)


@dataclasses.dataclass
class Segment:
  """ref: EntryConfig.Segment (embedding_hash_table.proto:24-33)."""
  dim_size: int
  initializer: Initializer
  optimizer: Optimizer


def CombineAsSegment(dim_size: int, initializer: Initializer, optimizer: Optimizer,
                     compressor=None) -> Segment:
  """ref: entry.py:514-538 (compressors are serving-side and out of scope: ignored)."""
  return Segment(dim_size, initializer, optimizer)


@dataclasses.dataclass
class CuckooHashTableConfig:
  """ref: entry.py:549-564."""
  initial_capacity: int = 1
  feature_evict_every_n_hours: int = 0


@dataclasses.dataclass
class TableConfig:
  """Plain equivalent of EmbeddingHashTableConfig (embedding_hash_table.proto:68-91)."""
  segments: List[Segment]
  initial_capacity: int = 1
  default_expire_time: int = 36500  # days (SlotExpireTimeConfig, proto:54-64)
  slot_expire_times: Dict[int, int] = dataclasses.field(default_factory=dict)
  init_seed: int = 0

  @property
  def dim_size(self) -> int:
    return sum(s.dim_size for s in self.segments)


class HashTableConfigInstance:
  """ref: entry.py:566-630: a table config plus one learning rate (value or callable) per segment."""

  def __init__(self, table_config: TableConfig,
               learning_rate_fns: Optional[Sequence[Union[float, Callable[[], float]]]] = None):
    self._table_config = table_config
    if learning_rate_fns is None:
      learning_rate_fns = [
          _d(s.optimizer.learning_rate, _DEFAULT_LR[s.optimizer.opt_type])
          for s in table_config.segments
      ]
    if len(learning_rate_fns) != len(table_config.segments):
      raise ValueError("one learning rate per segment is required")
    self._learning_rate_fns = list(learning_rate_fns)

  @property
  def table_config(self) -> TableConfig:
    return self._table_config

  @property
  def learning_rate_fns(self):
    return self._learning_rate_fns

  def call_learning_rate_fns(self) -> List[float]:
    """ref: entry.py:600-617."""
    if not self._learning_rate_fns:
      raise Exception("Learning_rate_fns must be not empty.")
    return [float(fn() if callable(fn) else fn) for fn in self._learning_rate_fns]


def to_c_table_cfgs(configs: Dict[str, HashTableConfigInstance]):
  """Flattens named configs into a ctypes array of mono_table_cfg (keeps the buffers alive)."""
  import ctypes as C
  names = list(configs.keys())
  arr = (_lib.TableCfg * len(names))()
  keep = []
  for i, name in enumerate(names):
    tc = configs[name].table_config
    segs = (_lib.SegmentCfg * len(tc.segments))()
    for j, s in enumerate(tc.segments):
      segs[j].dim = int(s.dim_size)
      segs[j].init_type = s.initializer.init_type
      segs[j].init_a = float(s.initializer.a)
      segs[j].init_b = float(s.initializer.b)
      segs[j].opt_type = s.optimizer.opt_type
      for q, v in enumerate(s.optimizer.params()):
        segs[j].opt_p[q] = float(v)
    slots = sorted(tc.slot_expire_times.items())
    slot_ids = (C.c_uint32 * max(len(slots), 1))(*[k for k, _ in slots])
    slot_days = (C.c_uint32 * max(len(slots), 1))(*[v for _, v in slots])
    bname = name.encode()
    arr[i].name = bname
    arr[i].n_segments = len(tc.segments)
    arr[i].segments = segs
    arr[i].initial_capacity = int(tc.initial_capacity)
    arr[i].default_expire_days = int(tc.default_expire_time)
    arr[i].n_slot_expire = len(slots)
    arr[i].slot_ids = slot_ids
    arr[i].slot_expire_days = slot_days
    arr[i].init_seed = int(tc.init_seed)
    keep += [segs, slot_ids, slot_days, bname]
  return arr, keep
