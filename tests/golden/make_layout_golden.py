"""Generates tests/golden/ref_layout_pooling.npz with the REFERENCE's own numpy pooling oracle.

The reference's test file monolith/native_training/fused_embedding_to_layout_test.py defines `pooling(pooling_type,
in_data, max_length)` (:91-116), the numpy oracle its FusedEmbeddingToLayout tests compare the op against.  The
module imports TensorFlow, which is not installed here, so the function is lifted out of the file with `ast`
(source untouched, nothing copied into this repo) and executed on seeded inputs; inputs and its outputs are
stored as a small fixture.  Run in the build container (needs /root/reference):

    python tests/golden/make_layout_golden.py
"""
import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/monolith/native_training/fused_embedding_to_layout_test.py"


class PoolingType:  # stand-in for the proto enum the reference function compares against
  SUM, MEAN, FIRSTN = 0, 1, 2


def reference_pooling():
  tree = ast.parse(open(REF).read())
  fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "pooling")
  ns = {"np": np, "PoolingType": PoolingType}
  exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
  return ns["pooling"]


def main():
  pooling = reference_pooling()
  rng = np.random.default_rng(21)
  out = {}
  cases = [(PoolingType.SUM, 4, 0), (PoolingType.SUM, 16, 0), (PoolingType.MEAN, 8, 0), (PoolingType.MEAN, 1, 0),
           (PoolingType.FIRSTN, 3, 5), (PoolingType.FIRSTN, 16, 2)]
  for ci, (pt, dim, max_len) in enumerate(cases):
    R, B = 50, 40
    table = rng.standard_normal((R, dim)).astype(np.float32)
    lens = rng.integers(1, 8, B)                      # the reference oracle needs >= 1 row per sample
    idx = rng.integers(0, R, int(lens.sum()))
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    res = []
    for b in range(B):
      rows = [table[i].copy() for i in idx[offs[b]:offs[b + 1]]]
      res.append(np.asarray(pooling(pt, rows, max_len), np.float32))
    out[f"c{ci}_pooling"] = np.int64(pt)
    out[f"c{ci}_max_len"] = np.int64(max_len)
    out[f"c{ci}_table"] = table
    out[f"c{ci}_idx"] = idx.astype(np.int64)
    out[f"c{ci}_offs"] = offs
    out[f"c{ci}_expect"] = np.stack(res)
  out["n_cases"] = np.int64(len(cases))
  np.savez(os.path.join(HERE, "ref_layout_pooling.npz"), **out)
  print("wrote ref_layout_pooling.npz")


if __name__ == "__main__":
  main()
