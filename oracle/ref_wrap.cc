// TEST INFRASTRUCTURE ONLY.  Thin C wrappers that compile two header-only pieces of the REAL
// reference, from where they lie under /root/reference (never copied into this repo):
//   * runtime/hash_table/optimizer/avx_utils.h   (Adagrad math, AVX2+FMA path)
//   * data/kernels/internal/uniq_hashtable.h     (first-occurrence dedup used by ShardingSparseFids)
// Output goes to oracle/_ref/libmonoref.so (git-ignored).  Used to validate oracle/oracle.cc.
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>
#include "monolith/native_training/runtime/hash_table/optimizer/avx_utils.h"
#include "monolith/native_training/data/kernels/internal/uniq_hashtable.h"

extern "C" {
// ref: avx_utils.h:238-245 (dispatches to Avx256AdagradOptimize when _ENABLE_AVX && __AVX__)
void ref_adagrad(float* num, float* norm, const float* grad, int64_t len, float lr, float wd) {
  monolith::hash_table::AdagradOptimize(num, norm, grad, (size_t)len, lr, wd);
}
void ref_adagrad_baseline(float* num, float* norm, const float* grad, int64_t len, float lr, float wd) {
  monolith::hash_table::BaselineAdagradOptimize(num, norm, grad, (size_t)len, lr, wd);
}
// ref: MultiShardUniqHashTable::uniq_fid (uniq_hashtable.h:240-247): returns, per occurrence, the
// first-occurrence ordinal of the fid inside its shard list.
void ref_uniq_fid(const uint64_t* fids, int64_t n, int num_shards, int64_t* uniq_idx_out,
                  int64_t* shard_sizes_out) {
  tensorflow::monolith_tf::UniqHashTable ht;
  tensorflow::monolith_tf::MultiShardUniqHashTable mt;
  mt.init(&ht);
  mt.resize(num_shards);
  mt.reset();
  for (int64_t i = 0; i < n; ++i) {
    int shard = (int)(fids[i] % (uint64_t)num_shards);
    uniq_idx_out[i] = (int64_t)mt.uniq_fid(fids[i], shard);
  }
  for (int s = 0; s < num_shards; ++s) shard_sizes_out[s] = mt.fid_num(s);
}
}
