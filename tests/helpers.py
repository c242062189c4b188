"""Shared test helpers: table configs."""
from monolith_b200 import entry


def opt_from(name, params=None, lr=None):
  params = params or {}
  if name == "sgd":
    return entry.SgdOptimizer(learning_rate=lr)
  if name == "adagrad":
    return entry.AdagradOptimizer(learning_rate=lr, initial_accumulator_value=params.get("initial_accumulator_value"),
                                  weight_decay_factor=params.get("weight_decay_factor", 0.0))
  if name == "ftrl":
    return entry.FtrlOptimizer(learning_rate=lr, initial_accumulator_value=params.get("initial_accumulator_value"),
                               beta=params.get("beta"), l1_regularization=params.get("l1"),
                               l2_regularization=params.get("l2"))
  if name == "adam":
    return entry.AdamOptimizer(learning_rate=lr, beta1=params.get("beta1", 0.9), beta2=params.get("beta2", 0.99),
                               epsilon=params.get("epsilon", 0.01),
                               weight_decay_factor=params.get("weight_decay_factor", 0.0),
                               use_nesterov=params.get("use_nesterov", False))
  if name == "momentum":
    return entry.MomentumOptimizer(learning_rate=lr, weight_decay_factor=params.get("weight_decay_factor", 0.0),
                                   use_nesterov=params.get("use_nesterov", False), momentum=params.get("momentum", 0.9))
  if name == "rmsprop":
    return entry.RmspropOptimizer(learning_rate=params.get("learning_rate", lr), weight_decay_factor=params.get("weight_decay_factor", 0.0),
                                  momentum=params.get("momentum", 0.9))
  if name == "rmspropv2":
    return entry.RmspropV2Optimizer(learning_rate=lr, weight_decay_factor=params.get("weight_decay_factor", 0.0),
                                    momentum=params.get("momentum", 0.9))
  if name == "adadelta":
    return entry.AdadeltaOptimizer(learning_rate=lr, weight_decay_factor=params.get("weight_decay_factor", 0.0),
                                   averaging_ratio=params.get("averaging_ratio", 0.9), epsilon=params.get("epsilon", 0.01))
  if name == "amsgrad":
    return entry.AmsgradOptimizer(learning_rate=lr, beta1=params.get("beta1", 0.9), beta2=params.get("beta2", 0.99),
                                  epsilon=params.get("epsilon", 0.01), weight_decay_factor=params.get("weight_decay_factor", 0.0),
                                  use_nesterov=params.get("use_nesterov", False))
  if name == "moving_average":
    return entry.MovingAverageOptimizer(momentum=params.get("momentum", 0.9), learning_rate=lr)
  if name == "group_adagrad":
    return entry.GroupAdaGradOptimizer(learning_rate=lr, beta=params.get("beta", 0.0),
                                       initial_accumulator_value=params.get("initial_accumulator_value", 0.1),
                                       l2_regularization_strength=params.get("l2", 0.0),
                                       weight_decay_factor=params.get("weight_decay_factor", 0.0))
  raise ValueError(name)


def table(segments, lrs=None, capacity=1, init=None, **kw):
  """segments: list of (dim, opt_name, params)."""
  segs = []
  for i, (dim, name, params) in enumerate(segments):
    ini = init[i] if isinstance(init, (list, tuple)) else (init or entry.ZerosInitializer())
    segs.append(entry.CombineAsSegment(dim, ini, opt_from(name, params)))
  tc = entry.TableConfig(segments=segs, initial_capacity=capacity, **kw)
  return entry.HashTableConfigInstance(tc, lrs)


def sgd_table(dim, lr=1.0, **kw):
  """ref: test_utils.generate_test_hash_table_config / hash_table_ops.vocab_hash_table: SGD, zeros init."""
  return table([(dim, "sgd", {})], [lr], **kw)
