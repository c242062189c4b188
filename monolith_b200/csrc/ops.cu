// ops.cu — the hot-path kernels: lookup, fused lookup+pool, and the upsert family
// (optimize / assign / assign_add / reinitialize / restore) with the fused sparse optimizers.
//
// Work decomposition: G lanes (G = 4..32, picked from the row width: one 16-byte vector per lane)
// cooperate on one id.  4 lanes fetch the 64-byte bucket with one 128-bit load each; the match is
// found with a warp ballot; all G lanes then move the row with 128-bit loads/stores.
//
// Why no TMA here: a row is 32..512 bytes at a data-dependent address.  cp.async.bulk needs one
// elected thread + an mbarrier round trip per row and pays off from ~1 KB per copy; at dim 32 a row is
// a single 128-byte line that 8 lanes fetch with one LDG.128 each.  The kernels are bound by random
// 64 B + 128 B HBM accesses in flight, so occupancy (memory-level parallelism), not staging, is the
// lever.
#include <algorithm>
#include <cstring>

#include "engine.h"

namespace mono {

// ------------------------------------------------------------------------------------------
// optimizer math (bit-exact with oracle/oracle.cc; every step is an explicit IEEE op so that
// nvcc never contracts a*b+c on its own)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int state_floats_dev(const SegDev& s) {
  switch (s.opt_type) {
    case MONO_OPT_ADAGRAD: return s.dim;
    case MONO_OPT_FTRL: return 2 * s.dim;
    case MONO_OPT_ADAM: return 2 * s.dim + 2;
    default: return 0;
  }
}

__device__ __forceinline__ float init_emb_value(const TableDev* t, const SegDev& s, int64_t key,
                                                int col) {
  switch (s.init_type) {
    case MONO_INIT_ONES: return 1.0f;
    case MONO_INIT_CONSTANT: return s.init_a;
    case MONO_INIT_UNIFORM: return uniform_init(t->seed, key, col, s.init_a, s.init_b);
    default: return 0.0f;
  }
}

// initial value of state float `lj` (local index inside the segment's state block)
// ref: adagrad_optimizer.cc:47-52, ftrl_optimizer.cc:45-52, adam_optimizer.cc:44-55
__device__ __forceinline__ float init_state_value(const SegDev& s, int lj) {
  switch (s.opt_type) {
    case MONO_OPT_ADAGRAD: return s.p[0];
    case MONO_OPT_FTRL: return lj < s.dim ? s.p[0] : 0.0f;
    case MONO_OPT_ADAM: return lj < 2 * s.dim ? 0.0f : (lj == 2 * s.dim ? s.p[0] : s.p[1]);
    default: return 0.0f;
  }
}

// One element of one optimizer step.  a/b are the element's state values
// (Adagrad: a = norm; FTRL: a = norm, b = zero; Adam: a = m, b = v).  `lr` is the slice learning rate
// (for Adam: the bias-corrected lr_t of the row).  `avx_form` selects the reference's AVX-path
// arithmetic for Adagrad (first floor(dim/8)*8 lanes, ref: avx_utils.h:96-119) vs the baseline form.
__device__ __forceinline__ void opt_elem(const SegDev& s, bool avx_form, float lr, float g, float& w,
                                         float& a, float& b) {
  switch (s.opt_type) {
    case MONO_OPT_SGD:  // sgd_optimizer.cc:46-48
      w = __fsub_rn(w, __fmul_rn(lr, g));
      break;
    case MONO_OPT_ADAGRAD: {
      const float wd = s.p[1];
      if (avx_form) {  // avx_utils.h:106-113
        float ug = __fmaf_rn(wd, w, g);
        float nn = __fmaf_rn(ug, ug, a);
        a = nn;
        float eff = __fdiv_rn(lr, __fsqrt_rn(nn));
        w = __fmaf_rn(-eff, g, w);
      } else {  // avx_utils.h:31-37
        float gg = __fadd_rn(g, __fmul_rn(wd, w));
        a = __fadd_rn(a, __fmul_rn(gg, gg));
        float eff = __fdiv_rn(lr, __fsqrt_rn(a));
        w = __fsub_rn(w, __fmul_rn(eff, gg));
      }
      break;
    }
    case MONO_OPT_FTRL: {  // ftrl_optimizer.cc:62-75
      const float beta = s.p[1], l1 = s.p[2], l2 = s.p[3];
      float norm_new = __fadd_rn(a, __fmul_rn(g, g));
      float sigma = __fdiv_rn(__fsub_rn(__fsqrt_rn(norm_new), __fsqrt_rn(a)), lr);
      b = __fadd_rn(b, __fsub_rn(g, __fmul_rn(sigma, w)));
      a = norm_new;
      if (fabsf(b) > l1) {
        float sb = signbit(b) ? 1.0f : 0.0f;
        float num = __fmul_rn(lr, __fsub_rn(__fmul_rn(sb, l1), b));
        float den = __fadd_rn(__fadd_rn(__fsqrt_rn(a), beta), __fmul_rn(l2, lr));
        w = __fdiv_rn(num, den);
      } else {
        w = 0.0f;
      }
      break;
    }
    case MONO_OPT_ADAM: {  // adam_optimizer.cc:65-80
      const float beta1 = s.p[0], beta2 = s.p[1], eps = s.p[2], wd = s.p[3];
      const bool nesterov = s.p[4] != 0.0f;
      float cur = __fadd_rn(g, __fmul_rn(wd, w));
      float new_m = __fadd_rn(a, __fmul_rn(__fsub_rn(cur, a), __fsub_rn(1.0f, beta1)));
      float new_v = __fadd_rn(b, __fmul_rn(__fsub_rn(__fmul_rn(cur, cur), b), __fsub_rn(1.0f, beta2)));
      float den = __fadd_rn(__fsqrt_rn(new_v), eps);
      if (nesterov) {
        float t1 = __fadd_rn(__fmul_rn(cur, __fsub_rn(1.0f, beta1)), __fmul_rn(beta1, new_m));
        w = __fsub_rn(w, __fdiv_rn(__fmul_rn(t1, lr), den));
      } else {
        w = __fsub_rn(w, __fdiv_rn(__fmul_rn(new_m, lr), den));
      }
      a = new_m;
      b = new_v;
      break;
    }
  }
}

__device__ __forceinline__ float adam_lr(float lr0, float b1p, float b2p) {  // adam_optimizer.cc:63
  return __fdiv_rn(__fmul_rn(lr0, __fsqrt_rn(__fsub_rn(1.0f, b2p))), __fsub_rn(1.0f, b1p));
}

__device__ __forceinline__ int seg_of_col(const TableDev* t, int c) {
  int s = 0;
  for (int i = 1; i < t->num_segs; ++i)
    if (c >= t->segs[i].col_begin) s = i;
  return s;
}
__device__ __forceinline__ int seg_of_state(const TableDev* t, int j) {
  int s = 0;
  for (int i = 1; i < t->num_segs; ++i)
    if (j >= t->segs[i].state_off) s = i;
  return s;
}

// Apply operation OP to row `row` of table t with G lanes.  `fresh` == row was just allocated for a
// key that was absent (ref: UpsertEntry init_fn: Init then fn, cuckoo_embedding_hash_table.cc:346-353).
// vals points at this id's dim floats (grad / value); for kOpRestore at dim+state+2 floats.
template <int G, int OP>
__device__ __forceinline__ void apply_row(const TableDev* __restrict__ t, uint32_t row, int64_t key,
                                          const float* __restrict__ vals,
                                          const float* __restrict__ lr, bool fresh) {
  const int gl = Group<G>::gl();
  const int D = t->dim;
  float* __restrict__ w_row = t->emb + (size_t)row * t->emb_stride;
  float* __restrict__ s_row = t->state + (size_t)row * t->state_stride;
  const bool init_all = fresh || OP == kOpReinit;

  if (OP == kOpOptimize) {
    // ---- fast path: one segment, dim % 4 == 0: 128-bit accesses on w, state and grad ----
    const SegDev& s0 = t->segs[0];
    if (t->num_segs == 1 && (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(vals) & 15) == 0)) {
      const float lr0 = lr[0];
      float lrt = lr0;
      float b1p = 0.f, b2p = 0.f;
      if (s0.opt_type == MONO_OPT_ADAM) {
        b1p = init_all ? s0.p[0] : s_row[2 * D];
        b2p = init_all ? s0.p[1] : s_row[2 * D + 1];
        lrt = adam_lr(lr0, b1p, b2p);
      }
      const int d8 = D & ~7;
      for (int c = gl * 4; c < D; c += G * 4) {
        float4 g4 = __ldg(reinterpret_cast<const float4*>(vals + c));
        float4 w4, a4 = make_float4(0, 0, 0, 0), b4 = make_float4(0, 0, 0, 0);
        if (init_all) {
          w4.x = init_emb_value(t, s0, key, c);
          w4.y = init_emb_value(t, s0, key, c + 1);
          w4.z = init_emb_value(t, s0, key, c + 2);
          w4.w = init_emb_value(t, s0, key, c + 3);
          a4.x = a4.y = a4.z = a4.w = init_state_value(s0, 0);
          b4.x = b4.y = b4.z = b4.w = init_state_value(s0, D);
        } else {
          w4 = *reinterpret_cast<const float4*>(w_row + c);
          if (s0.opt_type != MONO_OPT_SGD) a4 = *reinterpret_cast<const float4*>(s_row + c);
          if (s0.opt_type == MONO_OPT_FTRL || s0.opt_type == MONO_OPT_ADAM)
            b4 = *reinterpret_cast<const float4*>(s_row + D + c);
        }
        const bool avx = c < d8;  // c is a multiple of 4 and d8 of 8: the 4 lanes agree
        opt_elem(s0, avx, lrt, g4.x, w4.x, a4.x, b4.x);
        opt_elem(s0, avx, lrt, g4.y, w4.y, a4.y, b4.y);
        opt_elem(s0, avx, lrt, g4.z, w4.z, a4.z, b4.z);
        opt_elem(s0, avx, lrt, g4.w, w4.w, a4.w, b4.w);
        *reinterpret_cast<float4*>(w_row + c) = w4;
        if (s0.opt_type != MONO_OPT_SGD) *reinterpret_cast<float4*>(s_row + c) = a4;
        if (s0.opt_type == MONO_OPT_FTRL || s0.opt_type == MONO_OPT_ADAM)
          *reinterpret_cast<float4*>(s_row + D + c) = b4;
      }
      if (s0.opt_type == MONO_OPT_ADAM) {  // adam_optimizer.cc:82-83
        __syncwarp(Group<G>::mask());  // every lane has read the old powers
        if (gl == 0) {
          s_row[2 * D] = __fmul_rn(b1p, s0.p[0]);
          s_row[2 * D + 1] = __fmul_rn(b2p, s0.p[1]);
        }
      }
      return;
    }
  }

  // ---- generic path: any segment mix, any dim; one float per lane per step ----
  if (OP == kOpRestore) {
    for (int c = gl; c < D; c += G) w_row[c] = vals[c];
    for (int j = gl; j < t->state_dim; j += G) s_row[j] = vals[D + j];
    return;
  }
  for (int c = gl; c < D; c += G) {
    const int si = seg_of_col(t, c);
    const SegDev& s = t->segs[si];
    const int lc = c - s.col_begin;
    float w = init_all ? init_emb_value(t, s, key, c) : w_row[c];
    if (OP == kOpAssign) {
      w = vals[c];
    } else if (OP == kOpAssignAdd) {
      w = __fadd_rn(w, vals[c]);
    } else if (OP == kOpOptimize) {
      float a = 0.f, b = 0.f, lrt = lr[si];
      float* sp = s_row + s.state_off;
      if (s.opt_type == MONO_OPT_ADAGRAD) {
        a = init_all ? s.p[0] : sp[lc];
      } else if (s.opt_type == MONO_OPT_FTRL) {
        a = init_all ? s.p[0] : sp[lc];
        b = init_all ? 0.0f : sp[s.dim + lc];
      } else if (s.opt_type == MONO_OPT_ADAM) {
        a = init_all ? 0.0f : sp[lc];
        b = init_all ? 0.0f : sp[s.dim + lc];
        float b1p = init_all ? s.p[0] : sp[2 * s.dim];
        float b2p = init_all ? s.p[1] : sp[2 * s.dim + 1];
        lrt = adam_lr(lrt, b1p, b2p);
      }
      opt_elem(s, lc < (s.dim & ~7), lrt, vals[c], w, a, b);
      if (s.opt_type == MONO_OPT_ADAGRAD) {
        sp[lc] = a;
      } else if (s.opt_type == MONO_OPT_FTRL || s.opt_type == MONO_OPT_ADAM) {
        sp[lc] = a;
        sp[s.dim + lc] = b;
      }
    }
    w_row[c] = w;
  }
  if (OP == kOpOptimize) {
    // per-row beta powers advance once per step (after every lane has read the old values)
    __syncwarp(Group<G>::mask());
    if (gl == 0) {
      for (int si = 0; si < t->num_segs; ++si) {
        const SegDev& s = t->segs[si];
        if (s.opt_type != MONO_OPT_ADAM) continue;
        float* sp = s_row + s.state_off;
        float b1p = init_all ? s.p[0] : sp[2 * s.dim];
        float b2p = init_all ? s.p[1] : sp[2 * s.dim + 1];
        sp[2 * s.dim] = __fmul_rn(b1p, s.p[0]);
        sp[2 * s.dim + 1] = __fmul_rn(b2p, s.p[1]);
      }
    }
  } else if (init_all) {
    // assign / assign_add / reinitialize on a fresh (or re-initialised) row: optimizer Init
    for (int j = gl; j < t->state_dim; j += G) {
      const SegDev& s = t->segs[seg_of_state(t, j)];
      s_row[j] = init_state_value(s, j - s.state_off);
    }
  }
}

// ------------------------------------------------------------------------------------------
// lookup
// ------------------------------------------------------------------------------------------
// ref: MultiHashTableLookupOp::Compute / FusedLookupOp (multi_hash_table_lookup_op.cc:37-88,143-197)
// One launch serves every (shard, table) segment of the call.
template <int G>
__global__ void __launch_bounds__(kThreads)
lookup_kernel(const TableDev* __restrict__ tables, const CallSeg* __restrict__ segs, int nsegs,
              const int64_t* __restrict__ ids, int64_t n_total, float* __restrict__ out) {
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = Group<G>::gl();
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * GPW;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * GPW;
       wbase < n_total; wbase += wstride) {
    const int64_t i = wbase + lane / G;
    const bool active = i < n_total;
    int si = 0;
    const TableDev* t = tables;
    int64_t key = 0;
    uint32_t stash = 0;
    if (active) {
      si = nsegs > 1 ? find_seg(segs, nsegs, i) : 0;
      t = tables + segs[si].table;
      key = __ldg(ids + i);
      stash = t->ctrs[kCtrStash];
    }
    Probe pr = probe_key<G, 0>(t, key, active, stash);
    if (!active) continue;
    const int D = t->dim;
    float* dst = out + segs[si].val_off + (i - segs[si].id_begin) * D;
    const float* src = t->emb + (size_t)pr.row * t->emb_stride;
    const bool hit = pr.row != kEmptyRow;
    if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      for (int c = gl * 4; c < D; c += G * 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hit) v = __ldg(reinterpret_cast<const float4*>(src + c));
        *reinterpret_cast<float4*>(dst + c) = v;
      }
    } else {
      for (int c = gl; c < D; c += G) dst[c] = hit ? __ldg(src + c) : 0.0f;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
contains_kernel(const TableDev* __restrict__ t, const int64_t* __restrict__ ids, int64_t n,
                uint8_t* __restrict__ out) {
  constexpr int G = 4, GPW = 8;
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * GPW;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * GPW; wbase < n;
       wbase += wstride) {
    const int64_t i = wbase + lane / G;
    const bool active = i < n;
    int64_t key = active ? ids[i] : 0;
    Probe pr = probe_key<G, 0>(t, key, active, t->ctrs[kCtrStash]);
    if (active && Group<G>::gl() == 0) out[i] = pr.row != kEmptyRow;
  }
}

// flat entry dump [emb | state | found | ts]
__global__ void __launch_bounds__(kThreads)
lookup_entry_kernel(const TableDev* __restrict__ t, const int64_t* __restrict__ ids, int64_t n,
                    float* __restrict__ out) {
  constexpr int G = 8, GPW = 4;
  const int lane = threadIdx.x & 31, gl = Group<G>::gl();
  const int W = t->dim + t->state_dim + 2;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * GPW;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * GPW; wbase < n;
       wbase += wstride) {
    const int64_t i = wbase + lane / G;
    const bool active = i < n;
    int64_t key = active ? ids[i] : 0;
    Probe pr = probe_key<G, 1>(t, key, active, t->ctrs[kCtrStash]);
    if (!active) continue;
    float* dst = out + i * W;
    if (pr.row == kEmptyRow) {
      for (int c = gl; c < W; c += G) dst[c] = 0.0f;
    } else {
      const float* w = t->emb + (size_t)pr.row * t->emb_stride;
      const float* s = t->state + (size_t)pr.row * t->state_stride;
      for (int c = gl; c < t->dim; c += G) dst[c] = w[c];
      for (int c = gl; c < t->state_dim; c += G) dst[t->dim + c] = s[c];
      if (gl == 0) {
        dst[t->dim + t->state_dim] = __uint_as_float(1u);
        dst[t->dim + t->state_dim + 1] = __uint_as_float(ld_entry(pr.slot).ts);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// fused lookup + pool (forward headline kernel)
// ------------------------------------------------------------------------------------------
// One lane group per pooled output row: walk the row's FIDs in order, probe, gather the 16-byte
// vectors of the embedding row and accumulate in registers; one coalesced store per output row.
// Terms are added in FID order => bit-exact with the CPU reference's pooling
// (ref: OptimizedSumpooling, fused_embedding_to_layout.cc:26-59; ReduceSumOp, reduce_op.cc:29-51).
template <int G, int NV>  // NV = 16-byte vectors per lane (dim <= 4*G*NV)
__global__ void __launch_bounds__(kThreads)
lookup_pool_kernel(const TableDev* __restrict__ t, const int64_t* __restrict__ fids,
                   const int32_t* __restrict__ row_offsets, int64_t n_rows, int pooling,
                   float* __restrict__ out, int64_t out_stride, int out_col) {
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = Group<G>::gl();
  const int D = t->dim;
  const uint32_t stash = t->ctrs[kCtrStash];
  const float* __restrict__ emb = t->emb;
  const uint32_t stride = t->emb_stride;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * GPW;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * GPW;
       wbase < n_rows; wbase += wstride) {
    const int64_t r = wbase + lane / G;
    const bool ractive = r < n_rows;
    int64_t b = 0, e = 0;
    if (ractive) {
      b = row_offsets ? row_offsets[r] : r;
      e = row_offsets ? row_offsets[r + 1] : r + 1;
    }
    const int n = (int)(e - b);
    const int nmax = __reduce_max_sync(0xffffffffu, n);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float fn = (float)n;
    for (int j = 0; j < nmax; ++j) {
      const bool active = j < n;
      int64_t key = active ? __ldg(fids + b + j) : 0;
      Probe pr = probe_key<G, 0>(t, key, active, stash);
      if (!active) continue;
      const float* src = emb + (size_t)pr.row * stride;
      const bool hit = pr.row != kEmptyRow;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = (v * G + gl) * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hit && c < D) x = __ldg(reinterpret_cast<const float4*>(src + c));
        if (pooling == MONO_POOL_MEAN) {
          x.x = __fdiv_rn(x.x, fn); x.y = __fdiv_rn(x.y, fn);
          x.z = __fdiv_rn(x.z, fn); x.w = __fdiv_rn(x.w, fn);
        }
        if (j == 0) {
          acc[v] = x;
        } else {
          acc[v].x = __fadd_rn(acc[v].x, x.x); acc[v].y = __fadd_rn(acc[v].y, x.y);
          acc[v].z = __fadd_rn(acc[v].z, x.z); acc[v].w = __fadd_rn(acc[v].w, x.w);
        }
      }
    }
    if (!ractive) continue;
    float* dst = out + r * out_stride + out_col;
    const bool vec_ok = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = (v * G + gl) * 4;
      if (c >= D) continue;
      if (vec_ok) {
        *reinterpret_cast<float4*>(dst + c) = acc[v];
      } else {
        const float a[4] = {acc[v].x, acc[v].y, acc[v].z, acc[v].w};
        for (int q = 0; q < 4 && c + q < D; ++q) dst[c + q] = a[q];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// upsert family
// ------------------------------------------------------------------------------------------
struct UpsertArgs {
  const TableDev* tables;
  const CallSeg* segs;
  int nsegs;
  const int64_t* ids;
  const uint32_t* idx_list;  // optional indirection: process ids[idx_list[j]]
  int64_t n;                 // number of items to process
  const uint32_t* n_dev;     // optional: item count on device (overrides n when non-null)
  const float* vals;
  const float* lr;           // device copy of the call's learning rates
  uint32_t update_ts;
  uint32_t* miss_ctr;        // global miss counter of the call
  uint32_t* miss_list;       // positions (into ids) that missed
  int32_t* status;           // reinitialize only
  int val_width_extra;       // kOpRestore: row width = dim + state + 2
};

// Pass 1: ids that are present.  Probe, bump the expiry timestamp in the bucket entry, apply the
// op in place; misses are compacted (warp ballot + one atomic per warp) into miss_list.
template <int G, int OP>
__global__ void __launch_bounds__(kThreads) upsert_hit_kernel(UpsertArgs a) {
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = Group<G>::gl();
  const int64_t n = a.n_dev ? (int64_t)*a.n_dev : a.n;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * GPW;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * GPW; wbase < n;
       wbase += wstride) {
    const int64_t j = wbase + lane / G;
    const bool active = j < n;
    int64_t i = 0, key = 0;
    int si = 0;
    const TableDev* t = a.tables;
    uint32_t stash = 0;
    if (active) {
      i = a.idx_list ? (int64_t)a.idx_list[j] : j;
      si = a.nsegs > 1 ? find_seg(a.segs, a.nsegs, i) : 0;
      t = a.tables + a.segs[si].table;
      key = a.ids[i];
      stash = t->ctrs[kCtrStash];
    }
    Probe pr = probe_key<G, 1>(t, key, active, stash);
    const bool miss = active && pr.row == kEmptyRow;
    const uint32_t mbal = __ballot_sync(0xffffffffu, miss && gl == 0);
    if (mbal) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(a.miss_ctr, (uint32_t)__popc(mbal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (miss && gl == 0) a.miss_list[base + __popc(mbal & ((1u << lane) - 1u))] = (uint32_t)i;
    }
    if (!active || miss) continue;
    const CallSeg sg = a.segs[si];
    const int width = OP == kOpRestore ? t->dim + t->state_dim + 2 : t->dim;
    const float* v = a.vals ? a.vals + sg.val_off + (i - sg.id_begin) * width : nullptr;
    if (gl == 0) {
      uint32_t ts = a.update_ts;
      if (OP == kOpRestore) ts = __float_as_uint(v[t->dim + t->state_dim + 1]);
      pr.slot->ts = ts;  // ref: entry.SetTimestamp(update_time), cuckoo_embedding_hash_table.cc:243
      if (OP == kOpReinit) a.status[i] = 1;
    }
    apply_row<G, OP>(t, pr.row, key, v, a.lr + sg.lr_off, false);
  }
}

// Pass 2: absent ids (unique within the call): allocate a row (free list first, then bump),
// initialise + apply the op, then publish the entry with the lock-free cuckoo insert.
template <int G, int OP>
__global__ void __launch_bounds__(kThreads) upsert_miss_kernel(UpsertArgs a) {
  constexpr int GPW = 32 / G;
  const int lane = threadIdx.x & 31, gl = Group<G>::gl();
  const int64_t n = (int64_t)*a.miss_ctr;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * GPW;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * GPW; wbase < n;
       wbase += wstride) {
    const int64_t j = wbase + lane / G;
    const bool active = j < n;
    if (!active) continue;  // no warp-wide votes below
    const int64_t i = a.miss_list[j];
    const int si = a.nsegs > 1 ? find_seg(a.segs, a.nsegs, i) : 0;
    const CallSeg sg = a.segs[si];
    const TableDev* t = a.tables + sg.table;
    const int64_t key = a.ids[i];
    uint32_t row = 0;
    if (gl == 0) {
      const uint32_t ticket = atomicAdd(t->ctrs + kCtrMiss, 1u);
      const uint32_t fc = t->ctrs[kCtrFree];  // stable during this kernel (finalize updates it)
      row = ticket < fc ? t->free_list[fc - 1 - ticket] : t->ctrs[kCtrBump] + (ticket - fc);
      if (row >= t->row_cap) {
        atomicOr(t->ctrs + kCtrError, 2u);
        row = kEmptyRow;
      }
    }
    row = __shfl_sync(Group<G>::mask(), row, Group<G>::base());
    if (row == kEmptyRow) continue;
    const int width = OP == kOpRestore ? t->dim + t->state_dim + 2 : t->dim;
    const float* v = a.vals ? a.vals + sg.val_off + (i - sg.id_begin) * width : nullptr;
    apply_row<G, OP>(t, row, key, v, a.lr + sg.lr_off, true);
    if (gl == 0) {
      Entry e;
      e.key = key;
      e.row = row;
      e.ts = OP == kOpRestore ? __float_as_uint(v[t->dim + t->state_dim + 1]) : a.update_ts;
      cuckoo_insert(t, e);
      if (OP == kOpReinit) a.status[i] = 0;
    }
  }
}

// Fold the call's per-table miss tickets into the allocator counters.
__global__ void upsert_finalize_kernel(const TableDev* tables, const int32_t* table_ids, int ntab,
                                       uint32_t* miss_ctr, uint32_t update_ts) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q == 0) *miss_ctr = 0;
  if (q >= ntab) return;
  const TableDev* t = tables + table_ids[q];
  uint32_t m = t->ctrs[kCtrMiss];
  uint32_t fc = t->ctrs[kCtrFree];
  if (m <= fc) {
    t->ctrs[kCtrFree] = fc - m;
  } else {
    t->ctrs[kCtrFree] = 0;
    t->ctrs[kCtrBump] += m - fc;
  }
  t->ctrs[kCtrSize] += m;
  t->ctrs[kCtrMiss] = 0;
  if (update_ts > t->ctrs[kCtrMaxTs]) t->ctrs[kCtrMaxTs] = update_ts;
}

// ---- duplicate handling (ids not guaranteed unique) ------------------------------------------
struct SetEntry {  // scratch open-addressing set keyed by (table, fid)
  int64_t key;
  int32_t table;
  int32_t first_pos;
};

__device__ __forceinline__ bool cas_set(SetEntry* addr, const SetEntry& val) {
  Entry cmp = empty_entry();
  Entry v;
  v.key = val.key;
  v.row = (uint32_t)val.table;
  v.ts = (uint32_t)val.first_pos;
  return cas_entry(reinterpret_cast<Entry*>(addr), cmp, v);
}

// every pending position claims / joins the set slot of its (table, fid) and lowers first_pos
__global__ void __launch_bounds__(kThreads)
dup_claim_kernel(const CallSeg* __restrict__ segs, int nsegs, const int64_t* __restrict__ ids,
                 const uint32_t* __restrict__ pending, int64_t n, SetEntry* set, uint32_t mask,
                 uint32_t* __restrict__ slot_of) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n;
       j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = pending ? (int64_t)pending[j] : j;
    const int si = nsegs > 1 ? find_seg(segs, nsegs, i) : 0;
    const int table = segs[si].table;
    const int64_t key = ids[i];
    uint32_t s = (uint32_t)(mix64((uint64_t)key * 31 + table) >> 20) & mask;
    while (true) {
      Entry e = ld_entry_cg(reinterpret_cast<Entry*>(set + s));
      if (e.row == kEmptyRow && e.ts == 0xFFFFFFFFu && e.key == -1) {  // empty (table == -1)
        SetEntry ne;
        ne.key = key;
        ne.table = table;
        ne.first_pos = (int32_t)i;
        if (cas_set(set + s, ne)) break;
        continue;  // lost the race: re-read the same slot
      }
      if (e.key == key && (int32_t)e.row == table) {
        atomicMin(&set[s].first_pos, (int32_t)i);
        break;
      }
      s = (s + 1) & mask;
    }
    slot_of[j] = s;
  }
}

// split pending into leaders (lowest pending position of their key) and the rest
__global__ void __launch_bounds__(kThreads)
dup_split_kernel(const uint32_t* __restrict__ pending, int64_t n, const SetEntry* __restrict__ set,
                 const uint32_t* __restrict__ slot_of, uint32_t* leaders, uint32_t* rest,
                 uint32_t* ctr /*[0]=leaders,[1]=rest*/, uint32_t* leader_of /*optional, by position*/) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n;
       j += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t i = pending ? pending[j] : (uint32_t)j;
    const int32_t fp = set[slot_of[j]].first_pos;
    if (leader_of) leader_of[i] = (uint32_t)fp;
    if ((uint32_t)fp == i) leaders[atomicAdd(ctr, 1u)] = i;
    else rest[atomicAdd(ctr + 1, 1u)] = i;
  }
}

// acc_row(leader_of[i]) += vals_row(i) for this round's items (one item per key per round)
template <int G>
__global__ void __launch_bounds__(kThreads)
dup_accumulate_kernel(const TableDev* __restrict__ tables, const CallSeg* __restrict__ segs, int nsegs,
                      const uint32_t* __restrict__ items, const uint32_t* n_dev,
                      const uint32_t* __restrict__ leader_of, const float* __restrict__ vals,
                      float* __restrict__ acc) {
  const int64_t n = *n_dev;
  const int gl = Group<G>::gl();
  for (int64_t j = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / G; j < n;
       j += (int64_t)gridDim.x * blockDim.x / G) {
    const int64_t i = items[j];
    const int64_t l = leader_of[i];
    const int si = nsegs > 1 ? find_seg(segs, nsegs, i) : 0;
    const CallSeg sg = segs[si];
    const int D = tables[sg.table].dim;
    const float* src = vals + sg.val_off + (i - sg.id_begin) * D;
    float* dst = acc + sg.val_off + (l - sg.id_begin) * D;  // leader is in the same segment? see host
    for (int c = gl; c < D; c += G) dst[c] = __fadd_rn(dst[c], src[c]);
  }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
static int pick_group(int max_dim) {
  int v = (max_dim + 3) / 4;
  int g = 4;
  while (g < v && g < 32) g <<= 1;
  return g;
}

struct CallBlob {  // device pointers into the staged per-call descriptor block
  const CallSeg* segs;
  const float* lr;
  const int32_t* table_ids;
  int ntab;
};

static CallBlob stage_call(mono_mtable* mt, const CallSeg* h_segs, int nsegs, const float* lr_host,
                           int n_lr, cudaStream_t s) {
  size_t seg_bytes = sizeof(CallSeg) * (size_t)nsegs;
  size_t lr_bytes = sizeof(float) * (size_t)std::max(n_lr, 1);
  std::vector<int32_t> tabs;
  for (int i = 0; i < nsegs; ++i)
    if (std::find(tabs.begin(), tabs.end(), h_segs[i].table) == tabs.end())
      tabs.push_back(h_segs[i].table);
  size_t lr_off = (seg_bytes + 15) & ~(size_t)15;
  size_t tab_off = (lr_off + lr_bytes + 15) & ~(size_t)15;
  size_t total = tab_off + sizeof(int32_t) * tabs.size();
  if (total > StageRing::kBlockBytes) throw ArgError("too many segments in one call");
  int b = mt->ring.acquire();
  char* h = mt->ring.h(b);
  std::memcpy(h, h_segs, seg_bytes);
  if (n_lr > 0) std::memcpy(h + lr_off, lr_host, sizeof(float) * n_lr);
  std::memcpy(h + tab_off, tabs.data(), sizeof(int32_t) * tabs.size());
  mt->ring.commit(b, total, s);
  CallBlob cb;
  cb.segs = reinterpret_cast<const CallSeg*>(mt->ring.d(b));
  cb.lr = reinterpret_cast<const float*>(mt->ring.d(b) + lr_off);
  cb.table_ids = reinterpret_cast<const int32_t*>(mt->ring.d(b) + tab_off);
  cb.ntab = (int)tabs.size();
  return cb;
}

static int max_dim_of(mono_mtable* mt, const CallSeg* h_segs, int nsegs) {
  int m = 1;
  for (int i = 0; i < nsegs; ++i) m = std::max(m, mt->tables[h_segs[i].table].dim);
  return m;
}

void launch_lookup(mono_mtable* mt, const CallSeg* h_segs, int nsegs, const int64_t* ids_dev,
                   int64_t n_total, float* out_dev, cudaStream_t s) {
  if (n_total <= 0) return;
  upload_tables(mt, s);
  CallBlob cb = stage_call(mt, h_segs, nsegs, nullptr, 0, s);
  const int G = pick_group(max_dim_of(mt, h_segs, nsegs));
  const int grid = grid_for(n_total, kThreads / G);
#define L(GG) lookup_kernel<GG><<<grid, kThreads, 0, s>>>(mt->d_tables, cb.segs, nsegs, ids_dev, n_total, out_dev)
  switch (G) {
    case 4: L(4); break;
    case 8: L(8); break;
    case 16: L(16); break;
    default: L(32); break;
  }
#undef L
  MONO_CHECK_LAUNCH();
}

void launch_lookup_pool(mono_mtable* mt, int k, const int64_t* fids_dev, const int32_t* row_offsets,
                        int64_t n_rows, int pooling, float* out, int64_t out_stride, int out_col,
                        cudaStream_t s) {
  if (n_rows <= 0) return;
  if (pooling != MONO_POOL_SUM && pooling != MONO_POOL_MEAN)
    throw ArgError("lookup_pool supports SUM and MEAN pooling");
  upload_tables(mt, s);
  const int D = mt->tables[k].dim;
  if (D > 512) throw ArgError("lookup_pool supports dim <= 512");
  const int G = pick_group(D);
  const int nv = (D + 4 * G - 1) / (4 * G);
  const int grid = grid_for(n_rows, kThreads / G);
  const TableDev* t = mt->d_tables + k;
#define LP(GG, NV) lookup_pool_kernel<GG, NV><<<grid, kThreads, 0, s>>>(t, fids_dev, row_offsets, n_rows, pooling, out, out_stride, out_col)
  if (G == 4) LP(4, 1);
  else if (G == 8) LP(8, 1);
  else if (G == 16) LP(16, 1);
  else if (nv == 1) LP(32, 1);
  else if (nv == 2) LP(32, 2);
  else LP(32, 4);
#undef LP
  MONO_CHECK_LAUNCH();
}

void launch_contains(mono_mtable* mt, int k, const int64_t* ids, int64_t n, uint8_t* out,
                     cudaStream_t s) {
  if (n <= 0) return;
  upload_tables(mt, s);
  contains_kernel<<<grid_for(n, kThreads / 4), kThreads, 0, s>>>(mt->d_tables + k, ids, n, out);
  MONO_CHECK_LAUNCH();
}

void launch_lookup_entry(mono_mtable* mt, int k, const int64_t* ids, int64_t n, float* out,
                         cudaStream_t s) {
  if (n <= 0) return;
  upload_tables(mt, s);
  lookup_entry_kernel<<<grid_for(n, kThreads / 8), kThreads, 0, s>>>(mt->d_tables + k, ids, n, out);
  MONO_CHECK_LAUNCH();
}

template <int G, int OP>
static void launch_upsert_pair(const UpsertArgs& a, int64_t n_upper, cudaStream_t s) {
  const int grid = grid_for(n_upper, kThreads / G);
  upsert_hit_kernel<G, OP><<<grid, kThreads, 0, s>>>(a);
  MONO_CHECK_LAUNCH();
  upsert_miss_kernel<G, OP><<<grid, kThreads, 0, s>>>(a);
  MONO_CHECK_LAUNCH();
}

template <int OP>
static void launch_upsert_g(int G, const UpsertArgs& a, int64_t n_upper, cudaStream_t s) {
  switch (G) {
    case 4: launch_upsert_pair<4, OP>(a, n_upper, s); break;
    case 8: launch_upsert_pair<8, OP>(a, n_upper, s); break;
    case 16: launch_upsert_pair<16, OP>(a, n_upper, s); break;
    default: launch_upsert_pair<32, OP>(a, n_upper, s); break;
  }
}

static void launch_upsert(UpsertOp op, int G, const UpsertArgs& a, int64_t n_upper, cudaStream_t s) {
  switch (op) {
    case kOpOptimize: launch_upsert_g<kOpOptimize>(G, a, n_upper, s); break;
    case kOpAssign: launch_upsert_g<kOpAssign>(G, a, n_upper, s); break;
    case kOpAssignAdd: launch_upsert_g<kOpAssignAdd>(G, a, n_upper, s); break;
    case kOpReinit: launch_upsert_g<kOpReinit>(G, a, n_upper, s); break;
    case kOpRestore: launch_upsert_g<kOpRestore>(G, a, n_upper, s); break;
  }
}

void run_upsert(mono_mtable* mt, UpsertOp op, const CallSeg* h_segs, int nsegs,
                const int64_t* ids_dev, int64_t n_total, const float* vals_dev,
                const float* lr_host, int n_lr, int64_t update_time, bool unique, bool dedup_sum,
                int32_t* status_dev, cudaStream_t s) {
  if (n_total <= 0 || nsegs <= 0) return;
  // The kernels walk positions [0, n): rebase the call on the id range its segments cover
  // (fused_optimize passes one shard's segments of a larger id array).
  std::vector<CallSeg> rebased(h_segs, h_segs + nsegs);
  {
    const int64_t i0 = rebased.front().id_begin;
    for (int i = 0; i < nsegs; ++i) {
      if (i > 0 && rebased[i].id_begin != h_segs[i - 1].id_end)
        throw ArgError("call segments must be contiguous");
      rebased[i].id_begin -= i0;
      rebased[i].id_end -= i0;
    }
    ids_dev += i0;
    if (status_dev) status_dev += i0;
    n_total = rebased.back().id_end;
    h_segs = rebased.data();
  }
  if (n_total >= (int64_t)1 << 31) throw ArgError("more than 2^31 ids in one call");
  // capacity first (may rehash / grow and dirty the table descriptors)
  std::vector<uint64_t> per_table(mt->tables.size(), 0);
  for (int i = 0; i < nsegs; ++i)
    per_table[h_segs[i].table] += (uint64_t)(h_segs[i].id_end - h_segs[i].id_begin);
  for (size_t k = 0; k < per_table.size(); ++k)
    if (per_table[k]) ensure_capacity(mt, (int)k, per_table[k], s);
  upload_tables(mt, s);
  CallBlob cb = stage_call(mt, h_segs, nsegs, lr_host, n_lr, s);
  const int G = pick_group(max_dim_of(mt, h_segs, nsegs));

  // scratch: [miss_ctr (16 B) | miss_list u32[n_total]]
  char* ws = (char*)mt->ws_miss.get(64 + sizeof(uint32_t) * (size_t)n_total, s);
  uint32_t* miss_ctr = reinterpret_cast<uint32_t*>(ws);
  uint32_t* miss_list = reinterpret_cast<uint32_t*>(ws + 64);
  MONO_CUDA(cudaMemsetAsync(miss_ctr, 0, 64, s));

  UpsertArgs a;
  a.tables = mt->d_tables;
  a.segs = cb.segs;
  a.nsegs = nsegs;
  a.ids = ids_dev;
  a.idx_list = nullptr;
  a.n = n_total;
  a.n_dev = nullptr;
  a.vals = vals_dev;
  a.lr = cb.lr;
  a.update_ts = (uint32_t)update_time;
  a.miss_ctr = miss_ctr;
  a.miss_list = miss_list;
  a.status = status_dev;
  a.val_width_extra = 0;

  auto finalize = [&]() {
    upsert_finalize_kernel<<<(cb.ntab + 63) / 64, 64, 0, s>>>(mt->d_tables, cb.table_ids, cb.ntab,
                                                             miss_ctr, (uint32_t)update_time);
    MONO_CHECK_LAUNCH();
  };

  if (unique) {
    launch_upsert(op, G, a, n_total, s);
    finalize();
  } else {
    // Sequential semantics for duplicates (ref: per-id loops in BatchOptimize / AssignAdd):
    // round r applies the r-th occurrence of every key; within a round keys are unique.
    uint32_t cap = 1024;
    while (cap < 2 * (uint64_t)n_total) cap <<= 1;
    SetEntry* set = (SetEntry*)mt->ws_a.get(sizeof(SetEntry) * (size_t)cap, s);
    uint32_t* slot_of = (uint32_t*)mt->ws_b.get(sizeof(uint32_t) * (size_t)n_total, s);
    uint32_t* lists = (uint32_t*)mt->ws_c.get(sizeof(uint32_t) * (size_t)n_total * 3 + 64, s);
    uint32_t* ctr = lists;  // [0] leaders [1] rest
    uint32_t* leaders = lists + 16;
    uint32_t* bufA = leaders + n_total;
    uint32_t* bufB = bufA + n_total;
    uint32_t* leader_of = nullptr;
    float* acc = nullptr;
    uint32_t* round0_leaders = nullptr;
    uint32_t n_round0 = 0;
    size_t val_floats = 0;
    if (dedup_sum) {
      if (op != kOpOptimize) throw ArgError("dedup_sum applies to optimize only");
      {  // the reference dedups per BatchOptimize call = per segment: one segment per table here
        std::vector<int> seen(mt->tables.size(), 0);
        for (int i = 0; i < nsegs; ++i)
          if (h_segs[i].id_end > h_segs[i].id_begin && seen[h_segs[i].table]++)
            throw ArgError("dedup_sum needs at most one segment per table per call");
      }
      for (int i = 0; i < nsegs; ++i)
        val_floats = std::max<size_t>(val_floats, (size_t)h_segs[i].val_off +
                                                     (size_t)(h_segs[i].id_end - h_segs[i].id_begin) *
                                                         mt->tables[h_segs[i].table].dim);
      leader_of = (uint32_t*)mt->ws_d.get(sizeof(uint32_t) * (size_t)n_total * 2, s);
      round0_leaders = leader_of + n_total;
      acc = (float*)mt->ws_e.get(sizeof(float) * val_floats, s);
      MONO_CUDA(cudaMemcpyAsync(acc, vals_dev, sizeof(float) * val_floats, cudaMemcpyDeviceToDevice, s));
    }
    const uint32_t* pending = nullptr;
    int64_t n_pending = n_total;
    uint32_t* rest = bufA;
    int round = 0;
    while (n_pending > 0) {
      MONO_CUDA(cudaMemsetAsync(set, 0xFF, sizeof(SetEntry) * (size_t)cap, s));
      MONO_CUDA(cudaMemsetAsync(ctr, 0, 64, s));
      dup_claim_kernel<<<grid_for(n_pending, kThreads), kThreads, 0, s>>>(
          cb.segs, nsegs, ids_dev, pending, n_pending, set, cap - 1, slot_of);
      MONO_CHECK_LAUNCH();
      dup_split_kernel<<<grid_for(n_pending, kThreads), kThreads, 0, s>>>(
          pending, n_pending, set, slot_of, leaders, rest, ctr,
          (dedup_sum && round == 0) ? leader_of : nullptr);
      MONO_CHECK_LAUNCH();
      MONO_CUDA(cudaMemcpyAsync(mt->h_flag, ctr, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
      MONO_CUDA(cudaStreamSynchronize(s));
      const uint32_t n_lead = mt->h_flag[0], n_rest = mt->h_flag[1];
      if (dedup_sum) {
        if (round == 0) {
          MONO_CUDA(cudaMemcpyAsync(round0_leaders, leaders, sizeof(uint32_t) * n_lead,
                                    cudaMemcpyDeviceToDevice, s));
          n_round0 = n_lead;
        } else {
          switch (G) {
            case 4: dup_accumulate_kernel<4><<<grid_for(n_lead, kThreads / 4), kThreads, 0, s>>>(mt->d_tables, cb.segs, nsegs, leaders, ctr, leader_of, vals_dev, acc); break;
            case 8: dup_accumulate_kernel<8><<<grid_for(n_lead, kThreads / 8), kThreads, 0, s>>>(mt->d_tables, cb.segs, nsegs, leaders, ctr, leader_of, vals_dev, acc); break;
            case 16: dup_accumulate_kernel<16><<<grid_for(n_lead, kThreads / 16), kThreads, 0, s>>>(mt->d_tables, cb.segs, nsegs, leaders, ctr, leader_of, vals_dev, acc); break;
            default: dup_accumulate_kernel<32><<<grid_for(n_lead, kThreads / 32), kThreads, 0, s>>>(mt->d_tables, cb.segs, nsegs, leaders, ctr, leader_of, vals_dev, acc); break;
          }
          MONO_CHECK_LAUNCH();
          // ctr is reused next round: the kernel above must finish reading it first (same stream)
        }
      } else {
        a.idx_list = leaders;
        a.n = n_lead;
        launch_upsert(op, G, a, n_lead, s);
        finalize();
      }
      pending = rest;
      rest = (rest == bufA) ? bufB : bufA;
      n_pending = n_rest;
      ++round;
    }
    if (dedup_sum) {
      a.idx_list = round0_leaders;
      a.n = n_round0;
      a.vals = acc;
      launch_upsert(op, G, a, n_round0, s);
      finalize();
    }
  }
  for (size_t k = 0; k < per_table.size(); ++k) {
    if (!per_table[k]) continue;
    HostTable& t = mt->tables[k];
    t.issued_total += per_table[k];
    t.max_update_ts = std::max<int64_t>(t.max_update_ts, update_time);
    request_snapshot(mt, (int)k, s);
  }
}

}  // namespace mono
