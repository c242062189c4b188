"""monolith_b200 — B200-native collisionless-embedding engine behind Monolith's MultiHashTable op surface.

Only the embedding hot path is here: hand-written sm_100a CUDA in csrc/ behind the
C ABI of include/mono_emb.h, and a host-side mirror of the reference's Python interface
(monolith/native_training/{multi_hash_table_ops,distribution_ops,distributed_ps,entry}.py).
There is no CPU fallback: compute entry points raise when the CUDA library or a GPU is missing.
"""
from . import entry  # noqa: F401
from .multi_hash_table_ops import MultiHashTable  # noqa: F401

__all__ = ["entry", "MultiHashTable"]
