"""Generates tests/golden/ref_sharding_sparse_fids.npz with the REFERENCE's own Python oracle of
ShardingSparseFids (the op behind SURVEY §8 row a2).

monolith/native_training/data/parse_sparse_feature_test.py holds, as methods of its test class, a pure-Python
model of the op: `get_feature_cfg` (:87-140, output index = table_index * ps_num + shard for version 3),
`handle_feature` (:142-160, shard = fid % ps_num, per-(feature, shard) first-occurrence ordinals) and
`get_offset_result` (:162-240, per-(table, shard) lists = the features' lists in feature order, fid_offset =
index << 32 | feature_pre_offset + ordinal * dims_sum, feature_offset, nfl_offset with the shared flag).  The
module imports TensorFlow and generated protos, so the methods are lifted out with `ast` (nothing is copied into
this repo) and run on seeded inputs; inputs and outputs are stored as a fixture.

    python tests/golden/make_sharding_golden.py        # needs /root/reference
"""
import ast
import logging
import os
import types
from collections import defaultdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/monolith/native_training/data/parse_sparse_feature_test.py"
WANTED = ("fid_v1_to_v2", "get_pre_output_offset", "get_feature_cfg", "handle_feature", "get_offset_result")


def reference_model():
  tree = ast.parse(open(REF).read())
  cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DataOpsV2Test")
  fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
  assert len(fns) == len(WANTED)
  ns = {"defaultdict": defaultdict, "logging": logging, "print": lambda *a, **k: None}
  exec(compile(ast.Module(body=fns, type_ignores=[]), REF, "exec"), ns)
  model = types.SimpleNamespace(mask=(1 << 48) - 1, version=3)
  for name in WANTED:
    setattr(model, name, types.MethodType(ns[name], model))
  return model


def main():
  model = reference_model()
  rng = np.random.default_rng(33)
  out = {}
  cases = [dict(n_feat=5, n_tab=2, N=3, B=6, max_len=4, shared=()),
           dict(n_feat=9, n_tab=3, N=3, B=17, max_len=6, shared=(2, 7)),
           dict(n_feat=4, n_tab=4, N=1, B=5, max_len=3, shared=()),
           dict(n_feat=12, n_tab=3, N=8, B=40, max_len=9, shared=(0,))]
  for ci, c in enumerate(cases):
    names = [f"f_{chr(97 + (7 * i) % 26)}{i}" for i in range(c["n_feat"])]          # not in sorted order on purpose
    cfgs = types.SimpleNamespace(feature_configs={
        n: types.SimpleNamespace(table=f"table_{i % c['n_tab']}", slice_dims=list(rng.integers(1, 6, 3)))
        for i, n in enumerate(names)})
    feature_cfg, table_cfg, feature_name_sort, table_name_sort = model.get_feature_cfg(cfgs, c["N"])
    fid_map_t, fid_map_unique_t, fid_map_unique_map = defaultdict(list), defaultdict(list), defaultdict(dict)
    fid_offset_map, fid_offset_map_unique = defaultdict(list), defaultdict(list)
    shared = {names[i] for i in c["shared"]}
    flat, splits = {}, {}
    for fi, name in enumerate(names):
      slot = 100 + fi
      rows = 1 if name in shared else c["B"]
      lens = rng.integers(0, c["max_len"] + 1, rows)
      vals = ((np.int64(slot) << 48) | rng.integers(0, 12, int(lens.sum()))).astype(np.int64)   # small vocab: many repeats
      flat[name], splits[name] = vals, np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
      f_cfg, t_cfg = feature_cfg[name], table_cfg[feature_cfg[name]["table_name"]]
      for b in range(rows):
        l1, l2 = [], []
        model.handle_feature([], [int(v) for v in vals[splits[name][b]:splits[name][b + 1]]], f_cfg, t_cfg, c["N"],
                             l1, l2, fid_map_t, fid_map_unique_map, fid_map_unique_t)
        fid_offset_map[name].append(l1)
        fid_offset_map_unique[name].append(l2)
    nfl, feat_off, fid_off, fid_off_unique, tab_lists, tab_lists_unique = model.get_offset_result(
        feature_name_sort, table_name_sort, c["N"], feature_cfg, table_cfg, fid_offset_map, fid_offset_map_unique,
        fid_map_t, fid_map_unique_t, shared)
    out[f"c{ci}_names"] = np.array(names)
    out[f"c{ci}_tables"] = np.array([cfgs.feature_configs[n].table for n in names])
    out[f"c{ci}_dims_sum"] = np.array([int(sum(cfgs.feature_configs[n].slice_dims)) for n in names], np.int64)
    out[f"c{ci}_shared"] = np.array([n in shared for n in names])
    out[f"c{ci}_N"] = np.int64(c["N"])
    out[f"c{ci}_B"] = np.int64(c["B"])
    for n in names:
      out[f"c{ci}_fids_{n}"] = flat[n]
      out[f"c{ci}_splits_{n}"] = splits[n]
    out[f"c{ci}_nfl_offset"] = np.array(nfl, np.uint32)
    out[f"c{ci}_feature_offset"] = np.array(feat_off, np.int32)
    out[f"c{ci}_fid_offset_unique"] = np.array(fid_off_unique, np.uint64)
    out[f"c{ci}_fid_offset_all"] = np.array(fid_off, np.uint64)
    for ti, t in enumerate(table_name_sort):
      for n in range(c["N"]):
        out[f"c{ci}_list_{ti}_{n}"] = np.array(tab_lists_unique[f"{t}:{n}"], np.int64)
    out[f"c{ci}_n_tables"] = np.int64(len(table_name_sort))
  out["n_cases"] = np.int64(len(cases))
  np.savez(os.path.join(HERE, "ref_sharding_sparse_fids.npz"), **out)
  print("wrote ref_sharding_sparse_fids.npz")


if __name__ == "__main__":
  main()
