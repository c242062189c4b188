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
template <int OPT>
__device__ __forceinline__ void opt_elem_t(const float* __restrict__ p, bool avx_form, float lr, float g,
                                           float& w, float& a, float& b) {
  if (OPT == MONO_OPT_SGD) {  // sgd_optimizer.cc:46-48
    w = __fsub_rn(w, __fmul_rn(lr, g));
  } else if (OPT == MONO_OPT_ADAGRAD) {
    const float wd = p[1];
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
  } else if (OPT == MONO_OPT_FTRL) {  // ftrl_optimizer.cc:62-75
    const float beta = p[1], l1 = p[2], l2 = p[3];
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
  } else if (OPT == MONO_OPT_ADAM) {  // adam_optimizer.cc:65-80
    const float beta1 = p[0], beta2 = p[1], eps = p[2], wd = p[3];
    const bool nesterov = p[4] != 0.0f;
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
  }
}

__device__ __forceinline__ void opt_elem(const SegDev& s, bool avx_form, float lr, float g, float& w,
                                         float& a, float& b) {
  switch (s.opt_type) {
    case MONO_OPT_SGD: opt_elem_t<MONO_OPT_SGD>(s.p, avx_form, lr, g, w, a, b); break;
    case MONO_OPT_ADAGRAD: opt_elem_t<MONO_OPT_ADAGRAD>(s.p, avx_form, lr, g, w, a, b); break;
    case MONO_OPT_FTRL: opt_elem_t<MONO_OPT_FTRL>(s.p, avx_form, lr, g, w, a, b); break;
    case MONO_OPT_ADAM: opt_elem_t<MONO_OPT_ADAM>(s.p, avx_form, lr, g, w, a, b); break;
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
//
// Two phases per warp-tile of 32 ids, designed for memory-level parallelism (the op is a chain of
// dependent random HBM reads: FID -> 64 B bucket -> 128 B row):
//   A. lane-per-key probe: every lane resolves its own FID (4 x LDG.128 = the whole 64-byte bucket),
//      so a warp keeps 32 independent bucket reads in flight with ~10 live registers per lane;
//      the few lanes that miss probe the alternate bucket (and the stash when non-empty).
//   B. group-per-row copy: row indices are handed around with shuffles; G lanes move one row with
//      16-byte vectors, kRowsInFlight rows per group issued back to back before the first store.
// Used by lookup / fused_lookup (multi-segment) and by lookup_pool when every pooled row has exactly
// one FID (the per-slot case of Criteo/MovieLens/DCN-shaped inputs, where pooling is the identity).
__device__ __forceinline__ uint32_t probe_lane(const TableDev* __restrict__ t, int64_t key) {
  const Entry* __restrict__ buckets = t->buckets;
  uint32_t b1, b2;
  bucket_pair(key, t->num_buckets, b1, b2);
  uint32_t row = kEmptyRow;
  {
    const Entry* p = buckets + (size_t)b1 * kBucketSlots;
    Entry e0 = ld_entry_nc(p), e1 = ld_entry_nc(p + 1), e2 = ld_entry_nc(p + 2), e3 = ld_entry_nc(p + 3);
    if (e0.key == key && e0.row < kTombRow) row = e0.row;
    if (e1.key == key && e1.row < kTombRow) row = e1.row;
    if (e2.key == key && e2.row < kTombRow) row = e2.row;
    if (e3.key == key && e3.row < kTombRow) row = e3.row;
  }
  if (row == kEmptyRow) {
    const Entry* p = buckets + (size_t)b2 * kBucketSlots;
    Entry e0 = ld_entry_nc(p), e1 = ld_entry_nc(p + 1), e2 = ld_entry_nc(p + 2), e3 = ld_entry_nc(p + 3);
    if (e0.key == key && e0.row < kTombRow) row = e0.row;
    if (e1.key == key && e1.row < kTombRow) row = e1.row;
    if (e2.key == key && e2.row < kTombRow) row = e2.row;
    if (e3.key == key && e3.row < kTombRow) row = e3.row;
    if (row == kEmptyRow && t->ctrs[kCtrStash] != 0) {
      const uint32_t mask = t->stash_cap - 1;
      const uint32_t s = (uint32_t)(mix64((uint64_t)key) >> 17) & mask;
      for (uint32_t i = 0; i <= mask; ++i) {
        Entry e = ld_entry_cg(t->stash + ((s + i) & mask));
        if (e.row == kEmptyRow) break;
        if (e.key == key && e.row < kTombRow) { row = e.row; break; }
      }
    }
  }
  return row;
}

// store one 16-byte vector of a row (vector store when the row base is 16-byte aligned)
__device__ __forceinline__ void store_vec(float* dst, int c, int D, const float4& x, bool vec_ok) {
  if (c >= D) return;
  if (vec_ok) {
    __stcs(reinterpret_cast<float4*>(dst + c), x);  // streaming: written once, consumed later
  } else {
    const float a[4] = {x.x, x.y, x.z, x.w};
    for (int w = 0; w < 4 && c + w < D; ++w) dst[c + w] = a[w];
  }
}

// SINGLE: the call has one segment (one table): table fields and the output base are warp-uniform
// and hoisted, which keeps the kernel at ~40 registers => 6 resident blocks (48 warps) per SM.
template <int G, bool SINGLE>
__global__ void __launch_bounds__(kThreads, SINGLE ? 6 : 4)
lookup_kernel(const TableDev* __restrict__ tables, const CallSeg* __restrict__ segs, int nsegs,
              const int64_t* __restrict__ ids, int64_t n_total, float* __restrict__ out,
              int64_t out_stride /* <= 0: rows packed at dim floats */, int out_col) {
  constexpr int RPI = 32 / G;   // rows copied per iteration of a warp
  constexpr int ITERS = G;      // iterations to drain the 32 resolved rows
  constexpr int UNR = 4;        // row loads in flight per lane
  const int lane = threadIdx.x & 31, gl = Group<G>::gl(), grp = lane / G;
  const int c = gl * 4;
  // SINGLE: hoisted table / output description
  const TableDev* t0 = tables + segs[0].table;
  const int D0 = t0->dim;
  const float* __restrict__ emb0 = t0->emb;
  const uint32_t stride0 = t0->emb_stride;
  const int64_t rs0 = out_stride > 0 ? out_stride : D0;
  float* const base0 = out + segs[0].val_off - segs[0].id_begin * rs0 + out_col;
  const bool vec0 = (D0 & 3) == 0 && (rs0 & 3) == 0 && (reinterpret_cast<uintptr_t>(base0) & 15) == 0;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * 32;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 32;
       wbase < n_total; wbase += wstride) {
    // ---- phase A: lane-per-key probe ----
    const int64_t i = wbase + lane;
    int si = 0;
    uint32_t row = kEmptyRow;
    if (i < n_total) {
      if (!SINGLE) si = find_seg(segs, nsegs, i);
      row = probe_lane(SINGLE ? t0 : tables + segs[si].table, __ldg(ids + i));
    }
    // ---- phase B: group-per-row copy, UNR rows in flight ----
#pragma unroll
    for (int it0 = 0; it0 < ITERS; it0 += UNR) {
      float4 x[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int src_lane = (it0 + u) * RPI + grp;
        const uint32_t r = __shfl_sync(0xffffffffu, row, src_lane);
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (SINGLE) {
          if (r != kEmptyRow && c < D0)
            x[u] = __ldg(reinterpret_cast<const float4*>(emb0 + (size_t)r * stride0 + c));
        } else {
          const int s_r = __shfl_sync(0xffffffffu, si, src_lane);
          if (r != kEmptyRow) {
            const TableDev* t = tables + segs[s_r].table;
            if (c < t->dim) x[u] = __ldg(reinterpret_cast<const float4*>(t->emb + (size_t)r * t->emb_stride + c));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int src_lane = (it0 + u) * RPI + grp;
        const int64_t ir = wbase + src_lane;
        // shuffles read lanes outside the group: all 32 lanes take part, before any divergence
        const uint32_t r_st = __shfl_sync(0xffffffffu, row, src_lane);
        const int s_st = SINGLE ? 0 : __shfl_sync(0xffffffffu, si, src_lane);
        if (ir >= n_total) continue;
        if (SINGLE) {
          float* dst = base0 + ir * rs0;
          store_vec(dst, c, D0, x[u], vec0);
          if (D0 > 4 * G) {  // wide rows (dim > 128, G == 32): remaining vectors of the row
            const uint32_t r = r_st;
            for (int cc = c + 4 * G; cc < D0; cc += 4 * G) {
              float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
              if (r != kEmptyRow) y = __ldg(reinterpret_cast<const float4*>(emb0 + (size_t)r * stride0 + cc));
              store_vec(dst, cc, D0, y, vec0);
            }
          }
        } else {
          const CallSeg sg = segs[s_st];
          const TableDev* t = tables + sg.table;
          const int D = t->dim;
          const int64_t rs = out_stride > 0 ? out_stride : D;
          float* dst = out + sg.val_off + (ir - sg.id_begin) * rs + out_col;
          const bool vec_ok = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
          store_vec(dst, c, D, x[u], vec_ok);
          if (D > 4 * G) {
            const uint32_t r = r_st;
            for (int cc = c + 4 * G; cc < D; cc += 4 * G) {
              float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
              if (r != kEmptyRow) y = __ldg(reinterpret_cast<const float4*>(t->emb + (size_t)r * t->emb_stride + cc));
              store_vec(dst, cc, D, y, vec_ok);
            }
          }
        }
      }
    }
  }
}

// Owner lookup fused with the row exchange of the sharded forward (replaces FusedLookup + the
// embedding all-to-all, ref: distributed_ps_sync.py:100-118): ids[] is the concatenation of the FID
// buckets the requesters stored into this rank's window; the row of id i is written straight into the
// window of the requester that asked for it (one 128-bit NVLink store per lane, a whole row per group).
// Same two phases as lookup_kernel<G, true>.
template <int G>
__global__ void __launch_bounds__(kThreads, 6)
lookup_push_kernel(const TableDev* __restrict__ t0, const int64_t* __restrict__ ids, int64_t n_total,
                   const PeerOut po) {
  constexpr int RPI = 32 / G;
  constexpr int ITERS = G;
  constexpr int UNR = 4;
  const int lane = threadIdx.x & 31, gl = Group<G>::gl(), grp = lane / G;
  const int c = gl * 4;
  const int D0 = t0->dim;
  const float* __restrict__ emb0 = t0->emb;
  const uint32_t stride0 = t0->emb_stride;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * 32;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 32;
       wbase < n_total; wbase += wstride) {
    const int64_t i = wbase + lane;
    uint32_t row = kEmptyRow;
    if (i < n_total) row = probe_lane(t0, __ldg(ids + i));
#pragma unroll
    for (int it0 = 0; it0 < ITERS; it0 += UNR) {
      float4 x[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const uint32_t r = __shfl_sync(0xffffffffu, row, (it0 + u) * RPI + grp);
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r != kEmptyRow && c < D0)
          x[u] = __ldg(reinterpret_cast<const float4*>(emb0 + (size_t)r * stride0 + c));
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int src_lane = (it0 + u) * RPI + grp;
        const int64_t ir = wbase + src_lane;
        const uint32_t r_st = __shfl_sync(0xffffffffu, row, src_lane);
        if (ir >= n_total) continue;
        const int pr = peer_part(po, ir);
        float* dst = reinterpret_cast<float*>(po.base[pr]) + (ir - po.start[pr]) * D0;
        if (c < D0) *reinterpret_cast<float4*>(dst + c) = x[u];
        for (int cc = c + 4 * G; cc < D0; cc += 4 * G) {  // wide rows (dim > 128)
          float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r_st != kEmptyRow) y = __ldg(reinterpret_cast<const float4*>(emb0 + (size_t)r_st * stride0 + cc));
          *reinterpret_cast<float4*>(dst + cc) = y;
        }
      }
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
template <int G, int NV, int U>  // NV = 16-byte vectors per lane (dim <= 4*G*NV); U = rows in flight
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
  // a warp owns GPW*U consecutive pooled rows per iteration; group g takes rows g, g+GPW, ...
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * GPW * U;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * GPW * U;
       wbase < n_rows; wbase += wstride) {
    int64_t r[U], b[U];
    int n[U];
    int nloc = 0;
#pragma unroll
    for (int q = 0; q < U; ++q) {
      r[q] = wbase + q * GPW + lane / G;
      b[q] = 0;
      n[q] = 0;
      if (r[q] < n_rows) {
        b[q] = row_offsets ? row_offsets[r[q]] : r[q];
        n[q] = row_offsets ? (int)(row_offsets[r[q] + 1] - b[q]) : 1;
      }
      nloc = max(nloc, n[q]);
    }
    const int nmax = row_offsets ? __reduce_max_sync(0xffffffffu, nloc) : 1;
    float4 acc[U][NV];
#pragma unroll
    for (int q = 0; q < U; ++q)
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[q][v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < nmax; ++j) {
      int64_t key[U];
      bool active[U];
      uint32_t row[U];
      Entry* slot[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        active[q] = j < n[q];
        key[q] = active[q] ? __ldg(fids + b[q] + j) : 0;
      }
      probe_keys<G, 0, U>(t, key, active, stash, row, slot);
      float4 x[U][NV];
#pragma unroll
      for (int q = 0; q < U; ++q)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = (v * G + gl) * 4;
          x[q][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (active[q] && row[q] != kEmptyRow && c < D)
            x[q][v] = __ldg(reinterpret_cast<const float4*>(emb + (size_t)row[q] * stride + c));
        }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        if (!active[q]) continue;
        const float fn = (float)n[q];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float4 xv = x[q][v];
          if (pooling == MONO_POOL_MEAN) {
            xv.x = __fdiv_rn(xv.x, fn); xv.y = __fdiv_rn(xv.y, fn);
            xv.z = __fdiv_rn(xv.z, fn); xv.w = __fdiv_rn(xv.w, fn);
          }
          if (j == 0) {
            acc[q][v] = xv;
          } else {
            acc[q][v].x = __fadd_rn(acc[q][v].x, xv.x); acc[q][v].y = __fadd_rn(acc[q][v].y, xv.y);
            acc[q][v].z = __fadd_rn(acc[q][v].z, xv.z); acc[q][v].w = __fadd_rn(acc[q][v].w, xv.w);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < U; ++q) {
      if (r[q] >= n_rows) continue;
      float* dst = out + r[q] * out_stride + out_col;
      const bool vec_ok = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = (v * G + gl) * 4;
        if (c >= D) continue;
        if (vec_ok) {
          __stcs(reinterpret_cast<float4*>(dst + c), acc[q][v]);  // streaming: written once, read later
        } else {
          const float a[4] = {acc[q][v].x, acc[q][v].y, acc[q][v].z, acc[q][v].w};
          for (int w = 0; w < 4 && c + w < D; ++w) dst[c + w] = a[w];
        }
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
  const uint32_t* idx_list;  // optional indirection: item j is ids[idx_list[j]]
  int64_t n;                 // number of items to process
  const uint32_t* n_dev;     // optional: item count on device (overrides n when non-null)
  const float* vals;
  const float* lr;           // device copy of the call's learning rates
  uint32_t update_ts;
  uint32_t* miss_ctr;        // global miss counter of the call
  uint32_t* miss_list;       // item indices j that missed
  uint32_t* rowidx;          // per item j: resolved row (bit 31 = freshly inserted)
  int32_t* status;           // reinitialize only
  int64_t pos0 = 0;          // apply pass without idx_list: item j is position pos0 + j
};
constexpr uint32_t kFreshBit = 0x80000000u;

// probe that also returns the matching entry's address (for the timestamp bump)
__device__ __forceinline__ uint32_t probe_lane_slot(const TableDev* __restrict__ t, int64_t key,
                                                    Entry** slot) {
  Entry* buckets = t->buckets;
  uint32_t b1, b2;
  bucket_pair(key, t->num_buckets, b1, b2);
  uint32_t row = kEmptyRow;
#pragma unroll
  for (int round = 0; round < 2; ++round) {
    Entry* p = buckets + (size_t)(round == 0 ? b1 : b2) * kBucketSlots;
    Entry e0 = ld_entry(p), e1 = ld_entry(p + 1), e2 = ld_entry(p + 2), e3 = ld_entry(p + 3);
    if (e0.key == key && e0.row < kTombRow) { row = e0.row; *slot = p; }
    if (e1.key == key && e1.row < kTombRow) { row = e1.row; *slot = p + 1; }
    if (e2.key == key && e2.row < kTombRow) { row = e2.row; *slot = p + 2; }
    if (e3.key == key && e3.row < kTombRow) { row = e3.row; *slot = p + 3; }
    if (row != kEmptyRow) return row;
  }
  if (t->ctrs[kCtrStash] != 0) {
    const uint32_t mask = t->stash_cap - 1;
    const uint32_t s = (uint32_t)(mix64((uint64_t)key) >> 17) & mask;
    for (uint32_t i = 0; i <= mask; ++i) {
      Entry* p = t->stash + ((s + i) & mask);
      Entry e = ld_entry_cg(p);
      if (e.row == kEmptyRow) break;
      if (e.key == key && e.row < kTombRow) { *slot = p; return e.row; }
    }
  }
  return kEmptyRow;
}

__device__ __forceinline__ uint32_t restore_ts(const UpsertArgs& a, const CallSeg& sg, const TableDev* t,
                                               int64_t i) {
  const int width = t->dim + t->state_dim + 2;
  return __float_as_uint(a.vals[sg.val_off + (i - sg.id_begin) * width + t->dim + t->state_dim + 1]);
}

// Pass 1 — resolve present keys, lane per key (32 independent probe chains per warp): row index to
// rowidx[j], expiry timestamp bumped in the bucket entry (ref: entry.SetTimestamp(update_time),
// cuckoo_embedding_hash_table.cc:243); misses are compacted with a warp ballot + one atomic per warp.
template <bool RESTORE>
__global__ void __launch_bounds__(kThreads) resolve_hit_kernel(UpsertArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t n = a.n_dev ? (int64_t)*a.n_dev : a.n;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * 32;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 32; wbase < n;
       wbase += wstride) {
    const int64_t j = wbase + lane;
    bool miss = false;
    if (j < n) {
      const int64_t i = a.idx_list ? (int64_t)a.idx_list[j] : j;
      const CallSeg sg = a.segs[a.nsegs > 1 ? find_seg(a.segs, a.nsegs, i) : 0];
      const TableDev* t = a.tables + sg.table;
      Entry* slot = nullptr;
      const uint32_t row = probe_lane_slot(t, a.ids[i], &slot);
      if (row == kEmptyRow) {
        miss = true;
      } else {
        slot->ts = RESTORE ? restore_ts(a, sg, t, i) : a.update_ts;
        a.rowidx[j] = row;
      }
    }
    const uint32_t mbal = __ballot_sync(0xffffffffu, miss);
    if (mbal) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(a.miss_ctr, (uint32_t)__popc(mbal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (miss) a.miss_list[base + __popc(mbal & ((1u << lane) - 1u))] = (uint32_t)j;
    }
  }
}

// Pass 2 — absent keys (unique within the call), thread per key: take a row (free list first, then
// the bump allocator) and publish {fid, row, ts} with the lock-free cuckoo insert.  The row contents
// are written by the apply pass (fresh bit set).
template <bool RESTORE>
__global__ void __launch_bounds__(kThreads) resolve_miss_kernel(UpsertArgs a) {
  const int64_t n = (int64_t)*a.miss_ctr;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = a.miss_list[q];
    const int64_t i = a.idx_list ? (int64_t)a.idx_list[j] : j;
    const CallSeg sg = a.segs[a.nsegs > 1 ? find_seg(a.segs, a.nsegs, i) : 0];
    const TableDev* t = a.tables + sg.table;
    const uint32_t ticket = atomicAdd(t->ctrs + kCtrMiss, 1u);
    const uint32_t fc = t->ctrs[kCtrFree];  // stable during this kernel (finalize updates it)
    const uint32_t row = ticket < fc ? t->free_list[fc - 1 - ticket] : t->ctrs[kCtrBump] + (ticket - fc);
    if (row >= t->row_cap) {
      atomicOr(t->ctrs + kCtrError, 2u);
      a.rowidx[j] = kEmptyRow;
      continue;
    }
    Entry e;
    e.key = a.ids[i];
    e.row = row;
    e.ts = RESTORE ? restore_ts(a, sg, t, i) : a.update_ts;
    cuckoo_insert(t, e);
    a.rowidx[j] = row | kFreshBit;
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

// Pass 3 — apply the op to the resolved rows, lane group per row (pure streaming over w / state /
// value rows: no hash logic left here).
template <int G, int OP>
__global__ void __launch_bounds__(kThreads) upsert_apply_kernel(UpsertArgs a) {
  const int gl = Group<G>::gl();
  const int64_t n = a.n_dev ? (int64_t)*a.n_dev : a.n;
  const int64_t gstride = (int64_t)gridDim.x * (kThreads / G);
  for (int64_t j = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G; j < n; j += gstride) {
    const uint32_t ri = a.rowidx[j];
    if (ri == kEmptyRow) continue;  // row slab overflow was flagged
    const int64_t i = a.idx_list ? (int64_t)a.idx_list[j] : j + a.pos0;
    const CallSeg sg = a.segs[a.nsegs > 1 ? find_seg(a.segs, a.nsegs, i) : 0];
    const TableDev* t = a.tables + sg.table;
    const bool fresh = (ri & kFreshBit) != 0;
    const int width = OP == kOpRestore ? t->dim + t->state_dim + 2 : t->dim;
    const float* v = a.vals ? a.vals + sg.val_off + (i - sg.id_begin) * width : nullptr;
    if (OP == kOpReinit && gl == 0) a.status[i] = fresh ? 0 : 1;
    apply_row<G, OP>(t, ri & ~kFreshBit, a.ids[i], v, a.lr + sg.lr_off, fresh);
  }
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
                 uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ n_dev = nullptr) {
  if (n_dev) n = (int64_t)*n_dev;
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
        if ((int32_t)e.ts > (int32_t)i) atomicMin(&set[s].first_pos, (int32_t)i);
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

// Misses of a call whose segments are each duplicate-free but may share FIDs with one another (the
// owner side of the sharded backward: one segment per requesting rank).  The lowest position of a
// missing FID inserts it (leaders -> resolve_miss_kernel); the others take the leader's row, not fresh.
__global__ void __launch_bounds__(kThreads)
miss_split_kernel(const uint32_t* __restrict__ miss_list, const uint32_t* __restrict__ n_miss,
                  const SetEntry* __restrict__ set, const uint32_t* __restrict__ slot_of,
                  uint32_t* __restrict__ leaders, uint32_t* __restrict__ followers,
                  uint32_t* __restrict__ fol_leader, uint32_t* ctr /*[1] = leaders, [2] = followers*/) {
  const int64_t n = *n_miss;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t i = miss_list[q];
    const uint32_t fp = (uint32_t)set[slot_of[q]].first_pos;
    if (fp == i) {
      leaders[atomicAdd(ctr + 1, 1u)] = i;
    } else {
      const uint32_t k = atomicAdd(ctr + 2, 1u);
      followers[k] = i;
      fol_leader[k] = fp;
    }
  }
}

__global__ void __launch_bounds__(kThreads)
miss_follow_kernel(const uint32_t* __restrict__ followers, const uint32_t* __restrict__ fol_leader,
                   const uint32_t* __restrict__ n_fol, uint32_t* __restrict__ rowidx) {
  const int64_t n = *n_fol;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t r = rowidx[fol_leader[k]];
    rowidx[followers[k]] = r == kEmptyRow ? r : (r & ~kFreshBit);
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

static void launch_lookup_staged(mono_mtable* mt, const CallSeg* d_segs, int nsegs, int G,
                                 const int64_t* ids_dev, int64_t n_total, float* out_dev,
                                 int64_t out_stride, int out_col, cudaStream_t s) {
#define L(GG)                                                                                      \
  if (nsegs == 1)                                                                                  \
    lookup_kernel<GG, true><<<resident_grid(lookup_kernel<GG, true>, n_total, kThreads), kThreads, 0, s>>>( \
        mt->d_tables, d_segs, nsegs, ids_dev, n_total, out_dev, out_stride, out_col);              \
  else                                                                                             \
    lookup_kernel<GG, false><<<resident_grid(lookup_kernel<GG, false>, n_total, kThreads), kThreads, 0, s>>>( \
        mt->d_tables, d_segs, nsegs, ids_dev, n_total, out_dev, out_stride, out_col)
  switch (G) {
    case 4: L(4); break;
    case 8: L(8); break;
    case 16: L(16); break;
    default: L(32); break;
  }
#undef L
  MONO_CHECK_LAUNCH();
}

void launch_lookup(mono_mtable* mt, const CallSeg* h_segs, int nsegs, const int64_t* ids_dev,
                   int64_t n_total, float* out_dev, cudaStream_t s) {
  if (n_total <= 0) return;
  upload_tables(mt, s);
  // merge adjacent segments of the same table whose ids and output rows are contiguous (fused_lookup
  // of a single table: N shard segments collapse into one -> hoisted single-table kernel)
  std::vector<CallSeg> merged;
  for (int i = 0; i < nsegs; ++i) {
    if (!merged.empty()) {
      CallSeg& b = merged.back();
      const int D = mt->tables[b.table].dim;
      if (b.table == h_segs[i].table && b.id_end == h_segs[i].id_begin &&
          b.val_off + (b.id_end - b.id_begin) * D == h_segs[i].val_off) {
        b.id_end = h_segs[i].id_end;
        continue;
      }
    }
    merged.push_back(h_segs[i]);
  }
  h_segs = merged.data();
  nsegs = (int)merged.size();
  CallBlob cb = stage_call(mt, h_segs, nsegs, nullptr, 0, s);
  launch_lookup_staged(mt, cb.segs, nsegs, pick_group(max_dim_of(mt, h_segs, nsegs)), ids_dev, n_total,
                       out_dev, 0, 0, s);
}

void launch_lookup_push(mono_mtable* mt, int k, const int64_t* ids_dev, int64_t n_total, const PeerOut& po,
                        cudaStream_t s) {
  if (n_total <= 0) return;
  const int D = mt->tables[k].dim;
  if (D & 3) throw ArgError("lookup_push needs dim % 4 == 0 (16-byte peer stores)");
  if (po.start[po.n] != n_total) throw ArgError("lookup_push: counts do not add up to the id count");
  upload_tables(mt, s);
  const TableDev* t = mt->d_tables + k;
#define LPUSH(GG)                                                                                           \
  lookup_push_kernel<GG><<<resident_grid(lookup_push_kernel<GG>, n_total, kThreads), kThreads, 0, s>>>(t, ids_dev, \
                                                                                                       n_total, po)
  switch (pick_group(D)) {
    case 4: LPUSH(4); break;
    case 8: LPUSH(8); break;
    case 16: LPUSH(16); break;
    default: LPUSH(32); break;
  }
#undef LPUSH
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
  if (row_offsets == nullptr) {
    // one FID per pooled row: SUM and MEAN are the identity (x / 1 == x): probe + gather only
    CallSeg sg;
    sg.id_begin = 0;
    sg.id_end = n_rows;
    sg.val_off = 0;
    sg.table = k;
    sg.lr_off = 0;
    CallBlob cb = stage_call(mt, &sg, 1, nullptr, 0, s);
    launch_lookup_staged(mt, cb.segs, 1, G, fids_dev, n_rows, out, out_stride, out_col, s);
    return;
  }
  const TableDev* t = mt->d_tables + k;
  // U rows in flight per lane group: enough independent HBM round trips per warp to cover latency
#define LP(GG, NV, UU)                                                                            \
  lookup_pool_kernel<GG, NV, UU>                                                                   \
      <<<resident_grid(lookup_pool_kernel<GG, NV, UU>, n_rows, (kThreads / GG) * UU), kThreads, 0, s>>>( \
          t, fids_dev, row_offsets, n_rows, pooling, out, out_stride, out_col)
  if (G == 4) LP(4, 1, 4);
  else if (G == 8) LP(8, 1, 4);
  else if (G == 16) LP(16, 1, 4);
  else if (nv == 1) LP(32, 1, 4);
  else if (nv == 2) LP(32, 2, 2);
  else LP(32, 4, 1);
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
  constexpr bool R = OP == kOpRestore;
  resolve_hit_kernel<R><<<resident_grid(resolve_hit_kernel<R>, n_upper, kThreads), kThreads, 0, s>>>(a);
  MONO_CHECK_LAUNCH();
  resolve_miss_kernel<R><<<resident_grid(resolve_miss_kernel<R>, n_upper, kThreads), kThreads, 0, s>>>(a);
  MONO_CHECK_LAUNCH();
  upsert_apply_kernel<G, OP>
      <<<resident_grid(upsert_apply_kernel<G, OP>, n_upper, kThreads / G), kThreads, 0, s>>>(a);
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

  // scratch: [miss_ctr (64 B) | miss_list u32[n_total] | rowidx u32[n_total]]
  char* ws = (char*)mt->ws_miss.get(64 + 2 * sizeof(uint32_t) * (size_t)n_total, s);
  uint32_t* miss_ctr = reinterpret_cast<uint32_t*>(ws);
  uint32_t* miss_list = reinterpret_cast<uint32_t*>(ws + 64);
  uint32_t* rowidx = miss_list + n_total;
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
  a.rowidx = rowidx;
  a.status = status_dev;

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
    if (n_total > ((int64_t)1 << 30)) throw ArgError("more than 2^30 non-unique ids in one call");
    uint32_t cap = 1024;
    while ((uint64_t)cap < 2 * (uint64_t)n_total) cap <<= 1;  // <= 2^31: n_total <= 2^30 checked above
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


// Optimize over several groups of segments (the shards of a fused call) where the ids are unique inside
// a group but a FID may appear once per group: group g is applied after group g-1 (the reference's
// shard-by-shard order, multi_hash_table_update_op.cc:286-300).  One resolve pass serves all groups and
// nothing returns to the host; only the apply pass runs once per group.
void run_upsert_groups(mono_mtable* mt, const CallSeg* h_segs, int nsegs, const int64_t* group_begin,
                       int ngroups, const int64_t* ids_dev, const float* vals_dev, const float* lr_host,
                       int n_lr, int64_t update_time, cudaStream_t s) {
  if (nsegs <= 0 || ngroups <= 0) return;
  const int64_t n_total = h_segs[nsegs - 1].id_end;
  if (n_total <= 0) return;
  if (h_segs[0].id_begin != 0) throw ArgError("grouped optimize: segments must start at id 0");
  for (int i = 1; i < nsegs; ++i)
    if (h_segs[i].id_begin != h_segs[i - 1].id_end) throw ArgError("call segments must be contiguous");
  if (n_total >= (int64_t)1 << 31) throw ArgError("more than 2^31 ids in one call");
  std::vector<uint64_t> per_table(mt->tables.size(), 0);
  for (int i = 0; i < nsegs; ++i)
    per_table[h_segs[i].table] += (uint64_t)(h_segs[i].id_end - h_segs[i].id_begin);
  for (size_t k = 0; k < per_table.size(); ++k)
    if (per_table[k]) ensure_capacity(mt, (int)k, per_table[k], s);
  upload_tables(mt, s);
  CallBlob cb = stage_call(mt, h_segs, nsegs, lr_host, n_lr, s);
  const int G = pick_group(max_dim_of(mt, h_segs, nsegs));

  if (n_total > ((int64_t)1 << 30)) throw ArgError("more than 2^30 ids in one grouped optimize call");
  uint32_t cap = 1024;
  while ((uint64_t)cap < 2 * (uint64_t)n_total) cap <<= 1;
  // scratch: [ctr 64 B | miss_list | rowidx | leaders | followers | fol_leader | slot_of] u32[n_total] each
  char* ws = (char*)mt->ws_miss.get(64 + 6 * sizeof(uint32_t) * (size_t)n_total, s);
  uint32_t* ctr = reinterpret_cast<uint32_t*>(ws);  // [0] misses [1] leaders [2] followers
  uint32_t* miss_list = reinterpret_cast<uint32_t*>(ws + 64);
  uint32_t* rowidx = miss_list + n_total;
  uint32_t* leaders = rowidx + n_total;
  uint32_t* followers = leaders + n_total;
  uint32_t* fol_leader = followers + n_total;
  uint32_t* slot_of = fol_leader + n_total;
  SetEntry* set = (SetEntry*)mt->ws_a.get(sizeof(SetEntry) * (size_t)cap, s);
  MONO_CUDA(cudaMemsetAsync(ctr, 0, 64, s));
  MONO_CUDA(cudaMemsetAsync(set, 0xFF, sizeof(SetEntry) * (size_t)cap, s));

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
  a.miss_ctr = ctr;
  a.miss_list = miss_list;
  a.rowidx = rowidx;
  a.status = nullptr;
  resolve_hit_kernel<false><<<resident_grid(resolve_hit_kernel<false>, n_total, kThreads), kThreads, 0, s>>>(a);
  MONO_CHECK_LAUNCH();
  // misses are few in steady state: small grids, counts stay on the device
  const int gm = (int)std::min<int64_t>(148 * 2, (n_total + kThreads - 1) / kThreads);
  dup_claim_kernel<<<gm, kThreads, 0, s>>>(cb.segs, nsegs, ids_dev, miss_list, 0, set, cap - 1, slot_of, ctr);
  MONO_CHECK_LAUNCH();
  miss_split_kernel<<<gm, kThreads, 0, s>>>(miss_list, ctr, set, slot_of, leaders, followers, fol_leader, ctr);
  MONO_CHECK_LAUNCH();
  UpsertArgs am = a;
  am.miss_list = leaders;
  am.miss_ctr = ctr + 1;
  resolve_miss_kernel<false><<<gm, kThreads, 0, s>>>(am);
  MONO_CHECK_LAUNCH();
  miss_follow_kernel<<<gm, kThreads, 0, s>>>(followers, fol_leader, ctr + 2, rowidx);
  MONO_CHECK_LAUNCH();
  upsert_finalize_kernel<<<(cb.ntab + 63) / 64, 64, 0, s>>>(mt->d_tables, cb.table_ids, cb.ntab, ctr,
                                                           (uint32_t)update_time);
  MONO_CHECK_LAUNCH();
  for (int g = 0; g < ngroups; ++g) {
    const int64_t b = group_begin[g], e = group_begin[g + 1];
    if (e <= b) continue;
    UpsertArgs ag = a;
    ag.pos0 = b;
    ag.n = e - b;
    ag.rowidx = rowidx + b;
    switch (G) {
      case 4: upsert_apply_kernel<4, kOpOptimize><<<resident_grid(upsert_apply_kernel<4, kOpOptimize>, ag.n, kThreads / 4), kThreads, 0, s>>>(ag); break;
      case 8: upsert_apply_kernel<8, kOpOptimize><<<resident_grid(upsert_apply_kernel<8, kOpOptimize>, ag.n, kThreads / 8), kThreads, 0, s>>>(ag); break;
      case 16: upsert_apply_kernel<16, kOpOptimize><<<resident_grid(upsert_apply_kernel<16, kOpOptimize>, ag.n, kThreads / 16), kThreads, 0, s>>>(ag); break;
      default: upsert_apply_kernel<32, kOpOptimize><<<resident_grid(upsert_apply_kernel<32, kOpOptimize>, ag.n, kThreads / 32), kThreads, 0, s>>>(ag); break;
    }
    MONO_CHECK_LAUNCH();
  }
  for (size_t k = 0; k < per_table.size(); ++k) {
    if (!per_table[k]) continue;
    HostTable& t = mt->tables[k];
    t.issued_total += per_table[k];
    t.max_update_ts = std::max<int64_t>(t.max_update_ts, update_time);
    request_snapshot(mt, (int)k, s);
  }
}

// ==========================================================================================
// Fused backward: pooled-grad scatter + sparse optimizer + expiry bump, deterministic, no float
// atomics.  Replaces (for one table) the reference's
//   ScatterGrad / BackwardBatchKernel (fused_embedding_to_layout.h:286-347, .cu.cc:337-381: float
//   atomicAdd per occurrence into a per-unique-FID grad buffer)  +  MultiHashTableOptimize.
// Pipeline (all on the caller's stream, no host sync):
//   1 claim   : occurrences -> scratch-set slot (same slot <=> same FID)
//   2 sort    : stable LSD radix sort of (slot, position) pairs, 8 bits per pass: occurrences of a
//               FID become one contiguous run, in position order
//   3 runs    : run starts -> compact run list (one run per unique FID)
//   4 resolve : lane-per-key probe of the run's FID (+ insert when absent), timestamp bump
//   5 reduce+update: a lane group walks a run, sums the pooled-grad rows IN POSITION ORDER (the CPU
//               reference's order) in registers and applies the optimizer in place: the per-unique
//               grad buffer is never materialised.  Runs longer than kShortRun (hot FIDs of Zipf
//               batches) are split into kSubRun-sized pieces reduced by whole blocks, then combined
//               in piece order (fixed association => run-to-run bit-stable).
// ==========================================================================================
constexpr int kSortTile = kThreads * 8;
constexpr int kFastRun = 4;
constexpr int kShortRun = 64;
constexpr int kSubRun = 1024;

// MODE 0: per-block digit histogram (+ global digit totals).  MODE 1: stable scatter using the
// row-scanned histogram and the digit totals.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
radix_pass_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, int64_t n,
                  int shift, int pre_shift /* first pass only: key = raw >> pre_shift */,
                  int32_t* __restrict__ blk_cnt, int32_t* __restrict__ dtot, int nblk,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                  const uint32_t* __restrict__ n_dev /* optional: element count on the device */) {
  if (n_dev) n = (int64_t)*n_dev;
  __shared__ int32_t wcnt[kThreads / 32][256];
  __shared__ int32_t bbase[256];
  __shared__ int32_t dbase[256];
  __shared__ int32_t wtot[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int NW = kThreads / 32;
  constexpr int kPerWarp = kSortTile / NW;
  if (MODE == 1) {  // exclusive scan of the 256 digit totals (kThreads == 256: one digit per thread)
    const int v = dtot[threadIdx.x];
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wtot[w] = x;
    __syncthreads();
    int off = 0;
    for (int ww = 0; ww < w; ++ww) off += wtot[ww];
    dbase[threadIdx.x] = off + x - v;
    __syncthreads();
  }
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t wbeg = (int64_t)blk * kSortTile + (int64_t)w * kPerWarp;
    for (int d = lane; d < 256; d += 32) wcnt[w][d] = 0;
    __syncwarp();
    for (int c = 0; c < kPerWarp; c += 32) {
      const int64_t i = wbeg + c + lane;
      const int dg = i < n ? (int)(((keys_in[i] >> pre_shift) >> shift) & 255u) : -1;
      const uint32_t same = __match_any_sync(0xffffffffu, dg);
      if (dg >= 0 && lane == (__ffs(same) - 1)) wcnt[w][dg] += __popc(same);
      __syncwarp();
    }
    __syncthreads();
    for (int d = threadIdx.x; d < 256; d += blockDim.x) {
      int run = 0;
      for (int ww = 0; ww < NW; ++ww) {
        const int v = wcnt[ww][d];
        wcnt[ww][d] = run;
        run += v;
      }
      if (MODE == 0) {
        blk_cnt[(size_t)d * nblk + blk] = run;
        if (run) atomicAdd(dtot + d, run);
      } else {
        bbase[d] = blk_cnt[(size_t)d * nblk + blk] + dbase[d];
      }
    }
    __syncthreads();
    if (MODE == 1) {
      for (int c = 0; c < kPerWarp; c += 32) {
        const int64_t i = wbeg + c + lane;
        uint32_t key = 0, val = 0;
        int dg = -1;
        if (i < n) {
          key = keys_in[i] >> pre_shift;
          val = vals_in ? vals_in[i] : (uint32_t)i;
          dg = (int)((key >> shift) & 255u);
        }
        const uint32_t same = __match_any_sync(0xffffffffu, dg);
        if (dg >= 0) {
          const int r = bbase[dg] + wcnt[w][dg] + __popc(same & ((1u << lane) - 1u));
          keys_out[r] = key;
          vals_out[r] = val;
        }
        __syncwarp();
        if (dg >= 0 && lane == (__ffs(same) - 1)) wcnt[w][dg] += __popc(same);
        __syncwarp();
      }
    }
    __syncthreads();
  }
}

// per-tile digit histogram (the counting half of a pass): shared-memory atomics, one tile per block
// iteration, same tiling as the scatter kernel (radix_pass_kernel<1>)
__global__ void __launch_bounds__(kThreads)
radix_hist_kernel(const uint32_t* __restrict__ keys_in, int64_t n, int shift, int pre_shift,
                  int32_t* __restrict__ blk_cnt, int32_t* __restrict__ dtot, int nblk) {
  static_assert(kThreads == 256, "one digit per thread");
  __shared__ int32_t cnt[256];
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blk * kSortTile;
#pragma unroll
    for (int c = 0; c < kSortTile; c += kThreads) {
      const int64_t i = base + c + threadIdx.x;
      if (i < n) atomicAdd(&cnt[((keys_in[i] >> pre_shift) >> shift) & 255u], 1);
    }
    __syncthreads();
    const int v = cnt[threadIdx.x];
    blk_cnt[(size_t)threadIdx.x * nblk + blk] = v;
    if (v) atomicAdd(dtot + threadIdx.x, v);
    __syncthreads();
  }
}

// exclusive scan of every digit row blk_cnt[d][0..nblk): one 1024-thread block per digit
__global__ void __launch_bounds__(1024) radix_rowscan_kernel(int32_t* __restrict__ blk_cnt, int nblk) {
  __shared__ int32_t wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int32_t* row = blk_cnt + (size_t)blockIdx.x * nblk;
  int carry = 0;
  for (int base = 0; base < nblk; base += 1024) {
    const int idx = base + threadIdx.x;
    const int v = idx < nblk ? row[idx] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      int t = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      wsum[lane] = t;
    }
    __syncthreads();
    if (idx < nblk) row[idx] = carry + (w ? wsum[w - 1] : 0) + x - v;
    carry += wsum[31];
    __syncthreads();
  }
}

// ---- ordered run compaction: run j = j-th distinct slot of the sorted array ----
__device__ __forceinline__ bool is_run_start(const uint32_t* __restrict__ skeys, int64_t i, int64_t n) {
  return i < n && (i == 0 || skeys[i] != skeys[i - 1]);
}

__global__ void __launch_bounds__(kThreads)
runs_count_kernel(const uint32_t* __restrict__ skeys, int64_t n, int nblk, uint32_t* __restrict__ blk_runs) {
  __shared__ uint32_t wc[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    uint32_t cnt = 0;
    for (int c = 0; c < kSortTile; c += kThreads)
      cnt += __popc(__ballot_sync(0xffffffffu, is_run_start(skeys, (int64_t)blk * kSortTile + c + threadIdx.x, n)));
    if (lane == 0) wc[w] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int ww = 0; ww < kThreads / 32; ++ww) t += wc[ww];
      blk_runs[blk] = t;
    }
    __syncthreads();
  }
}

// one block: exclusive scan of blk_runs, total -> n_runs, sentinel run_start[n_runs] = n
__global__ void __launch_bounds__(1024)
runs_scan_kernel(uint32_t* __restrict__ blk_runs, int nblk, uint32_t* __restrict__ n_runs,
                 uint32_t* __restrict__ run_start, int64_t n) {
  __shared__ uint32_t wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per = (nblk + 1023) / 1024;
  const int b = min(nblk, (int)threadIdx.x * per), e = min(nblk, b + per);
  uint32_t sum = 0;
  for (int i = b; i < e; ++i) sum += blk_runs[i];
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  if (w == 0) {
    uint32_t t = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    wsum[lane] = t;
  }
  __syncthreads();
  uint32_t run = (w ? wsum[w - 1] : 0) + x - sum;
  for (int i = b; i < e; ++i) {
    const uint32_t v = blk_runs[i];
    blk_runs[i] = run;
    run += v;
  }
  if (threadIdx.x == 1023) {
    *n_runs = wsum[31];
    run_start[wsum[31]] = (uint32_t)n;
  }
}

__global__ void __launch_bounds__(kThreads)
runs_write_kernel(const uint32_t* __restrict__ skeys, const uint32_t* __restrict__ perm, int64_t n, int nblk,
                  const uint32_t* __restrict__ blk_runs, uint32_t* __restrict__ run_start,
                  uint32_t* __restrict__ run_first_pos, uint32_t* __restrict__ run_of_sorted /* optional */,
                  const Entry* __restrict__ claim_set /* optional: rowidx[j] = row parked in run j's set entry */,
                  uint32_t* __restrict__ rowidx) {
  __shared__ uint32_t wc[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int NW = kThreads / 32;
  constexpr int kPerWarp = kSortTile / NW;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t wbeg = (int64_t)blk * kSortTile + (int64_t)w * kPerWarp;
    uint32_t cnt = 0;
    for (int c = 0; c < kPerWarp; c += 32)
      cnt += __popc(__ballot_sync(0xffffffffu, is_run_start(skeys, wbeg + c + lane, n)));
    if (lane == 0) wc[w] = cnt;
    __syncthreads();
    uint32_t base = blk_runs[blk];
    for (int ww = 0; ww < w; ++ww) base += wc[ww];
    for (int c = 0; c < kPerWarp; c += 32) {
      const int64_t i = wbeg + c + lane;
      const bool st = is_run_start(skeys, i, n);
      const uint32_t bal = __ballot_sync(0xffffffffu, st);
      if (st) {
        const uint32_t j = base + __popc(bal & ((1u << lane) - 1u));
        run_start[j] = (uint32_t)i;
        run_first_pos[j] = perm[i];
        if (claim_set) rowidx[j] = claim_set[skeys[i]].row;
      }
      if (run_of_sorted && i < n) run_of_sorted[i] = base + __popc(bal & (0xffffffffu >> (31 - lane))) - 1;
      base += __popc(bal);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads)
occ_row_kernel(const int32_t* __restrict__ row_offsets, int64_t n_rows, uint32_t* __restrict__ occ_row) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < n_rows;
       r += (int64_t)gridDim.x * blockDim.x)
    for (int m = row_offsets[r]; m < row_offsets[r + 1]; ++m) occ_row[m] = (uint32_t)r;
}

struct BwdArgs {
  TableDev td;          // by value: lives in the kernel-parameter constant bank (uniform, free reads)
  const TableDev* t;    // device copy, for the generic apply_row path
  const int64_t* fids;
  const uint32_t* skeys;
  const uint32_t* perm;
  int64_t n;  // occurrences
  const uint32_t* n_runs;
  const uint32_t* run_start;
  const uint32_t* run_first_pos;
  const uint32_t* rowidx;       // per run (bit 31 = fresh)
  const uint32_t* occ_row;      // occurrence -> pooled row (null: identity)
  const int32_t* row_offsets;   // for MEAN (null: n == 1)
  int pooling;
  const float* pooled_grad;
  int64_t grad_stride;
  int grad_col;
  const float* lr;              // device, slice learning rates of the table
  // long runs
  uint32_t* n_long;             // counters: n_long, n_med = n_long + 1, work_ctr = n_long + 2 (zeroed per call)
  uint32_t* med_list;           // runs of kFastRun < len <= kShortRun (run index j)
  uint32_t* long_list;          // run index j
  uint32_t* long_len;
  uint32_t* long_sub_base;      // exclusive prefix of sub-piece counts (+ total at [n_long])
  uint2* piece_desc;            // per piece: {first sorted index, length}
  float* partial;               // [sub pieces][D]
  float* ugrad;                 // [runs][D] summed gradient of every run (= unique FID), run order
  float* scratch;               // alias of ugrad for the generic (multi-segment) apply path
};

// Where the gradient rows come from: copied out of the kernel parameters once per thread so that the
// inner loops do not re-read the constant bank (ncu showed LDCU stalls inside the unrolled loads).
struct GradSrc {
  const float* base;            // pooled_grad + grad_col + lane column
  int64_t stride;
  const uint32_t* perm;
  const uint32_t* occ_row;
  const int32_t* row_offsets;
  bool mean;
};
__device__ __forceinline__ GradSrc make_grad_src(const BwdArgs& a, int c) {
  GradSrc g;
  g.base = a.pooled_grad + a.grad_col + c;
  g.stride = a.grad_stride;
  g.perm = a.perm;
  g.occ_row = a.occ_row;
  g.row_offsets = a.row_offsets;
  g.mean = a.pooling == MONO_POOL_MEAN && a.row_offsets != nullptr;
  return g;
}
// gradient row of occurrence m, columns c..c+3
__device__ __forceinline__ float4 occ_grad4(const GradSrc& gs, uint32_t m) {
  const uint32_t r = gs.occ_row ? gs.occ_row[m] : m;
  float4 g = __ldg(reinterpret_cast<const float4*>(gs.base + (size_t)r * gs.stride));
  if (gs.mean) {
    const float fn = (float)(gs.row_offsets[r + 1] - gs.row_offsets[r]);
    g.x = __fdiv_rn(g.x, fn); g.y = __fdiv_rn(g.y, fn); g.z = __fdiv_rn(g.z, fn); g.w = __fdiv_rn(g.w, fn);
  }
  return g;
}
__device__ __forceinline__ void add4(float4& a, const float4& b) {
  a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
}

// Row state prefetched BEFORE the gradient gather so that the w / optimizer-state reads overlap the
// random gradient-row reads (single-segment fast path).
struct RowPre {
  float4 w4, a4, b4;
  float b1p, b2p;
};

template <int G, int OPT>
__device__ __forceinline__ RowPre bwd_prefetch(const BwdArgs& a, uint32_t ri, int c) {
  RowPre p;
  p.w4 = p.a4 = p.b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  p.b1p = p.b2p = 0.f;
  if (OPT < 0 || (ri & kFreshBit)) return p;
  const int D = a.td.dim;
  const uint32_t row = ri & ~kFreshBit;
  const float* __restrict__ w_row = a.td.emb + (size_t)row * a.td.emb_stride;
  const float* __restrict__ s_row = a.td.state + (size_t)row * a.td.state_stride;
  if (c < D) {
    p.w4 = *reinterpret_cast<const float4*>(w_row + c);
    if (OPT != MONO_OPT_SGD) p.a4 = *reinterpret_cast<const float4*>(s_row + c);
    if (OPT == MONO_OPT_FTRL || OPT == MONO_OPT_ADAM) p.b4 = *reinterpret_cast<const float4*>(s_row + D + c);
  }
  if (OPT == MONO_OPT_ADAM) {
    p.b1p = s_row[2 * D];
    p.b2p = s_row[2 * D + 1];
  }
  return p;
}

// Apply the optimizer to row `ri` with the summed gradient held in registers (one float4 per lane,
// dim <= 4*G).  OPT >= 0: single-segment table with that optimizer (compile-time specialised);
// OPT < 0: any segment mix, staged through scratch + apply_row.
template <int G, int OPT>
__device__ __forceinline__ void bwd_apply(const BwdArgs& a, uint32_t j, uint32_t ri, float4 g4, int c,
                                          RowPre pre) {
  const int D = a.td.dim;
  const uint32_t row = ri & ~kFreshBit;
  const bool fresh = (ri & kFreshBit) != 0;
  if (OPT >= 0) {
    const SegDev& s0 = a.td.segs[0];
    float* __restrict__ w_row = a.td.emb + (size_t)row * a.td.emb_stride;
    float* __restrict__ s_row = a.td.state + (size_t)row * a.td.state_stride;
    float lrt = a.lr[0];
    float4 w4 = pre.w4, a4 = pre.a4, b4 = pre.b4;
    float b1p = pre.b1p, b2p = pre.b2p;
    if (fresh) {  // the key is only needed to initialise a new row
      const int64_t key = a.fids[a.run_first_pos[j]];
      w4.x = init_emb_value(&a.td, s0, key, c); w4.y = init_emb_value(&a.td, s0, key, c + 1);
      w4.z = init_emb_value(&a.td, s0, key, c + 2); w4.w = init_emb_value(&a.td, s0, key, c + 3);
      a4.x = a4.y = a4.z = a4.w = init_state_value(s0, 0);
      b4.x = b4.y = b4.z = b4.w = init_state_value(s0, D);
      b1p = s0.p[0];
      b2p = s0.p[1];
    }
    if (OPT == MONO_OPT_ADAM) lrt = adam_lr(lrt, b1p, b2p);
    if (c < D) {
      const bool avx = c < (D & ~7);
      opt_elem_t<OPT>(s0.p, avx, lrt, g4.x, w4.x, a4.x, b4.x);
      opt_elem_t<OPT>(s0.p, avx, lrt, g4.y, w4.y, a4.y, b4.y);
      opt_elem_t<OPT>(s0.p, avx, lrt, g4.z, w4.z, a4.z, b4.z);
      opt_elem_t<OPT>(s0.p, avx, lrt, g4.w, w4.w, a4.w, b4.w);
      *reinterpret_cast<float4*>(w_row + c) = w4;
      if (OPT != MONO_OPT_SGD) *reinterpret_cast<float4*>(s_row + c) = a4;
      if (OPT == MONO_OPT_FTRL || OPT == MONO_OPT_ADAM) *reinterpret_cast<float4*>(s_row + D + c) = b4;
    }
    if (OPT == MONO_OPT_ADAM) {
      __syncwarp(Group<G>::mask());  // every lane has read the old powers (in the prefetch)
      if (Group<G>::gl() == 0) {
        s_row[2 * D] = __fmul_rn(b1p, s0.p[0]);
        s_row[2 * D + 1] = __fmul_rn(b2p, s0.p[1]);
      }
    }
  } else {
    float* sc = a.scratch + (size_t)j * D;
    if (c < D) *reinterpret_cast<float4*>(sc + c) = g4;
    __syncwarp(Group<G>::mask());
    apply_row<G, kOpOptimize>(a.t, row, a.fids[a.run_first_pos[j]], sc, a.lr, fresh);
  }
}

// sum of the gradient rows of occurrences [s, s+len) of the sorted order, in that order.
// The group's lanes fetch `perm` cooperatively (one coalesced load per G occurrences) and keep UNR
// independent gradient-row loads in flight; the adds stay in position order.
template <int G, int UNR>
__device__ __forceinline__ float4 sum_grad_rows(const GradSrc& gs, uint32_t s, uint32_t len, bool in) {
  const int gl = Group<G>::gl(), gb = Group<G>::base();
  const uint32_t gmask = Group<G>::mask();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // the next G permutation entries are requested before the current G rows, so the row reads of
  // chunk q never wait for an index load (only the very first chunk pays the two-level chain)
  uint32_t m_cur = (uint32_t)gl < len ? gs.perm[s + gl] : 0u;
  for (uint32_t q0 = 0; q0 < len; q0 += G) {
    const uint32_t m_next = (q0 + G + gl < len) ? gs.perm[s + q0 + G + gl] : 0u;
    const int cnt = (int)min((uint32_t)G, len - q0);
    for (int u0 = 0; u0 < cnt; u0 += UNR) {
      float4 g[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const uint32_t m = __shfl_sync(gmask, m_cur, gb + min(u0 + u, G - 1));
        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u0 + u < cnt && in) g[u] = occ_grad4(gs, m);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (u0 + u >= cnt) continue;
        if (q0 + u0 + u == 0) acc = g[u]; else add4(acc, g[u]);
      }
    }
    m_cur = m_next;
  }
  return acc;
}

// destination of run j's summed row: ugrad[j], or (po.n != 0: the sharded backward's fused gradient
// exchange) row j of an owner-bucketed list whose part r lives in rank r's peer window.
__device__ __forceinline__ float* run_dst(float* ugrad, const PeerOut& po, int64_t j, int D) {
  if (po.n == 0) return ugrad + (size_t)j * D;
  const int r = peer_part(po, j);
  return reinterpret_cast<float*>(po.base[r]) + (j - po.start[r]) * D;
}

// Fast runs (<= kFastRun occurrences: the bulk of a Zipf batch).  Warp tile of 32 runs, everything that
// can be fetched lane-parallel is: (A) lane l reads run l's bounds and its <= 4 permutation entries
// (coalesced, independent), medium / long runs are appended to their lists with one atomic per warp;
// (B) CH runs per group are reduced together: for occurrence k = 0..3 all CH gradient rows are requested
// before the first add, so a warp keeps up to 16 independent 128-byte row reads in flight (the op is
// bound by random row reads).  Adds stay in occurrence order (bit-exact with the reference).
template <int G>
__global__ void __launch_bounds__(kThreads, 4) run_sum_kernel(BwdArgs a, const PeerOut po) {
  constexpr int RPI = 32 / G;            // groups per warp
  constexpr int FL = kFastRun;
  constexpr int CH = 4;                  // runs reduced together by one group (G >= 4 runs per group and tile)
  const int lane = threadIdx.x & 31, gl = Group<G>::gl(), grp = lane / G;
  const int c = gl * 4;
  const int64_t nr = *a.n_runs;
  const int D = a.td.dim;
  const bool in = c < D;
  const GradSrc gs = make_grad_src(a, c);
  const uint32_t* __restrict__ run_start = a.run_start;
  float* __restrict__ ugrad = a.ugrad;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * 32;
  for (int64_t wbase = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 32; wbase < nr;
       wbase += wstride) {
    // ---- phase A: lane per run ----
    const int64_t j_l = wbase + lane;
    uint32_t s_l = 0, len_l = 0;
    if (j_l < nr) {
      s_l = run_start[j_l];
      len_l = run_start[j_l + 1] - s_l;
    }
    const uint32_t fast_len = len_l <= (uint32_t)FL ? len_l : 0u;  // 0: not handled here
    uint32_t m_l[FL];
#pragma unroll
    for (int k = 0; k < FL; ++k) m_l[k] = (uint32_t)k < fast_len ? gs.perm[s_l + k] : 0u;
    {
      const bool med = len_l > (uint32_t)FL && len_l <= (uint32_t)kShortRun;
      const bool lng = len_l > (uint32_t)kShortRun;
      const uint32_t bm = __ballot_sync(0xffffffffu, med), bl = __ballot_sync(0xffffffffu, lng);
      if (bm) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.n_long + 1, (uint32_t)__popc(bm));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (med) a.med_list[base + __popc(bm & ((1u << lane) - 1u))] = (uint32_t)j_l;
      }
      if (bl) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.n_long, (uint32_t)__popc(bl));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lng) a.long_list[base + __popc(bl & ((1u << lane) - 1u))] = (uint32_t)j_l;
      }
    }
    // ---- phase B: group per run, CH runs at a time ----
#pragma unroll 1
    for (int it0 = 0; it0 < G; it0 += CH) {
      float4 acc[CH];
      uint32_t len[CH];
#pragma unroll
      for (int t = 0; t < CH; ++t) len[t] = __shfl_sync(0xffffffffu, fast_len, (it0 + t) * RPI + grp);
#pragma unroll
      for (int k = 0; k < FL; ++k) {
        float4 x[CH];
#pragma unroll
        for (int t = 0; t < CH; ++t) {
          const uint32_t m = __shfl_sync(0xffffffffu, m_l[k], (it0 + t) * RPI + grp);
          x[t] = make_float4(0.f, 0.f, 0.f, 0.f);
          if ((uint32_t)k < len[t] && in) x[t] = occ_grad4(gs, m);
        }
#pragma unroll
        for (int t = 0; t < CH; ++t) {
          if (k == 0) acc[t] = x[t];
          else if ((uint32_t)k < len[t]) add4(acc[t], x[t]);
        }
      }
#pragma unroll
      for (int t = 0; t < CH; ++t) {
        const int64_t j = wbase + (it0 + t) * RPI + grp;
        if (len[t] != 0 && in) *reinterpret_cast<float4*>(run_dst(ugrad, po, j, D) + c) = acc[t];
      }
    }
  }
}

// Medium runs (kFastRun < len <= kShortRun), one group per run, summed in position order.  Launched
// with one group per POSSIBLE medium run (the count lives on the device): the hardware block scheduler
// balances the uneven run lengths, blocks past the count exit at once.
template <int G>
__global__ void __launch_bounds__(kThreads) run_sum_med_kernel(BwdArgs a, const PeerOut po) {
  const int gl = Group<G>::gl(), c = gl * 4;
  const int64_t q = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G;
  if (q >= (int64_t)a.n_long[1]) return;
  const int D = a.td.dim;
  const bool in = c < D;
  const GradSrc gs = make_grad_src(a, c);
  const uint32_t j = a.med_list[q];
  const uint32_t s = a.run_start[j];
  const float4 acc = sum_grad_rows<G, (G >= 8 ? 8 : 4)>(gs, s, a.run_start[j + 1] - s, in);
  if (in) *reinterpret_cast<float4*>(run_dst(a.ugrad, po, j, D) + c) = acc;
}

// Apply: group per run; rowidx[j], ugrad[j] and the row's w / state are all independent loads.
template <int G, int OPT>
__global__ void __launch_bounds__(kThreads) runs_apply_kernel(BwdArgs a) {
  const int gl = Group<G>::gl();
  const int c = gl * 4;
  const int64_t nr = *a.n_runs;
  const int D = a.td.dim;
  const int64_t gstride = (int64_t)gridDim.x * (kThreads / G);
  for (int64_t j = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G; j < nr; j += gstride) {
    const uint32_t ri = a.rowidx[j];
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) g4 = __ldcs(reinterpret_cast<const float4*>(a.ugrad + (size_t)j * D + c));
    if (ri == kEmptyRow) continue;
    const RowPre pre = bwd_prefetch<G, OPT>(a, ri, c);
    bwd_apply<G, OPT>(a, (uint32_t)j, ri, g4, c, pre);
  }
}

// one block: piece counts of the long runs and their exclusive prefix
__global__ void __launch_bounds__(1024) long_prep_kernel(BwdArgs a) {
  __shared__ uint32_t carry;
  __shared__ uint32_t wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t nl = *a.n_long;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nl; base += blockDim.x) {
    const uint32_t q = base + threadIdx.x;
    uint32_t pieces = 0;
    if (q < nl) {
      const uint32_t j = a.long_list[q];
      const uint32_t len = a.run_start[j + 1] - a.run_start[j];
      a.long_len[q] = len;
      pieces = (len + kSubRun - 1) / kSubRun;
    }
    uint32_t x = pieces;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      uint32_t t = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      wsum[lane] = t;
    }
    __syncthreads();
    const uint32_t excl = carry + (w ? wsum[w - 1] : 0) + x - pieces;
    if (q < nl) {
      a.long_sub_base[q] = excl;
      const uint32_t s0 = a.run_start[a.long_list[q]], len = a.long_len[q];
      for (uint32_t p = 0; p < pieces; ++p)
        a.piece_desc[excl + p] = make_uint2(s0 + p * kSubRun, min((uint32_t)kSubRun, len - p * kSubRun));
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = carry + wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) a.long_sub_base[nl] = carry;
}

// block per piece of a long run: kThreads/G groups reduce contiguous slices in order, then the
// slices are combined in order (fixed association: deterministic)
template <int G>
__global__ void __launch_bounds__(kThreads, 4) pool_bwd_long_partial_kernel(BwdArgs a) {
  constexpr int NG = kThreads / G;
  __shared__ float4 sm[NG][G];
  const int gl = Group<G>::gl(), g = threadIdx.x / G, c = gl * 4;
  const uint32_t total = a.long_sub_base[*a.n_long];
  const int D = a.td.dim;
  const bool in = c < D;
  const GradSrc gs = make_grad_src(a, c);
  const uint2* __restrict__ desc = a.piece_desc;
  float* __restrict__ partial = a.partial + c;
  __shared__ uint32_t next_piece;
  while (true) {  // pieces differ 16x in size: blocks pull the next one from a device counter
    if (threadIdx.x == 0) next_piece = atomicAdd(a.n_long + 2, 1u);
    __syncthreads();
    const uint32_t wi = next_piece;
    if (wi >= total) break;
    const uint2 d = desc[wi];
    const uint32_t len = d.y;
    const uint32_t per = (len + NG - 1) / NG;
    const uint32_t b = min(len, g * per), e = min(len, b + per);
    sm[g][gl] = sum_grad_rows<G, (G >= 8 ? 8 : 4)>(gs, d.x + b, e - b, in);
    __syncthreads();
    if (g == 0) {
      float4 t = sm[0][gl];
      for (int k = 1; k < NG; ++k)
        if ((uint32_t)k * per < len) add4(t, sm[k][gl]);
      if (in) *reinterpret_cast<float4*>(partial + (size_t)wi * D) = t;
    }
    __syncthreads();
  }
}

// group per long run: combine its pieces in order into ugrad[j]
template <int G>
__global__ void __launch_bounds__(kThreads) pool_bwd_long_final_kernel(BwdArgs a, const PeerOut po) {
  const int gl = Group<G>::gl(), c = gl * 4;
  const uint32_t nl = *a.n_long;
  const int D = a.td.dim;
  const int64_t gstride = (int64_t)gridDim.x * (kThreads / G);
  for (int64_t q = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G; q < nl; q += gstride) {
    const uint32_t j = a.long_list[q];
    const uint32_t b = a.long_sub_base[q], e = a.long_sub_base[q + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t w0 = b; w0 < e; w0 += 8) {
      float4 p[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w0 + u < e && c < D) p[u] = *reinterpret_cast<const float4*>(a.partial + (size_t)(w0 + u) * D + c);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (w0 + u >= e) continue;
        if (w0 + u == b) acc = p[u]; else add4(acc, p[u]);
      }
    }
    if (c < D) *reinterpret_cast<float4*>(run_dst(a.ugrad, po, j, D) + c) = acc;
  }
}

struct SortWs {
  uint32_t *k0, *v0, *k1, *v1;  // k0 holds the input keys; values are implicit positions
  int32_t* blk_cnt;             // [256][nblk]
  int32_t* dtot;                // [passes][256], zeroed
  uint32_t* blk_runs;           // [nblk]
  uint32_t* run_start;          // [M + 1]
  uint32_t* run_first_pos;      // [M]
  uint32_t* n_runs;             // device counter, zeroed
  uint32_t* run_of_sorted = nullptr;  // optional [M]: run index of every sorted element
  const Entry* claim_set = nullptr;   // optional: the claim set (keys are its slots) holding resolved rows ...
  uint32_t* rowidx = nullptr;         // ... gathered per run into rowidx[j]
};

// stable LSD radix sort of (key >> pre_shift, position) over `bits` key bits, then the ordered run
// list (one run per distinct key).  Everything stays on the stream; counts stay on the device.
static void sort_and_runs(const SortWs& w, int64_t M, int bits, int pre_shift, const uint32_t** skeys_out,
                          const uint32_t** perm_out, cudaStream_t s) {
  const int passes = std::max(1, (bits + 7) / 8);
  const int nblk = (int)((M + kSortTile - 1) / kSortTile);
  const uint32_t* vin = nullptr;  // first pass: value = position
  uint32_t *kin = w.k0, *kout = w.k1, *vout = w.v1;
  const int gh = resident_grid(radix_hist_kernel, nblk, 1);
  const int gs = resident_grid(radix_pass_kernel<1>, nblk, 1);
  for (int p = 0; p < passes; ++p) {
    int32_t* dt = w.dtot + 256 * p;
    const int ps = p == 0 ? pre_shift : 0;
    radix_hist_kernel<<<gh, kThreads, 0, s>>>(kin, M, 8 * p, ps, w.blk_cnt, dt, nblk);
    MONO_CHECK_LAUNCH();
    radix_rowscan_kernel<<<256, 1024, 0, s>>>(w.blk_cnt, nblk);
    MONO_CHECK_LAUNCH();
    radix_pass_kernel<1><<<gs, kThreads, 0, s>>>(kin, vin, M, 8 * p, ps, w.blk_cnt, dt, nblk, kout, vout, nullptr);
    MONO_CHECK_LAUNCH();
    vin = vout;
    std::swap(kin, kout);
    vout = (vout == w.v1) ? w.v0 : w.v1;
  }
  runs_count_kernel<<<resident_grid(runs_count_kernel, nblk, 1), kThreads, 0, s>>>(kin, M, nblk, w.blk_runs);
  MONO_CHECK_LAUNCH();
  runs_scan_kernel<<<1, 1024, 0, s>>>(w.blk_runs, nblk, w.n_runs, w.run_start, M);
  MONO_CHECK_LAUNCH();
  runs_write_kernel<<<resident_grid(runs_write_kernel, nblk, 1), kThreads, 0, s>>>(
      kin, vin, M, nblk, w.blk_runs, w.run_start, w.run_first_pos, w.run_of_sorted, w.claim_set, w.rowidx);
  MONO_CHECK_LAUNCH();
  *skeys_out = kin;
  *perm_out = vin;
}

// run reduction: fast runs inline, medium runs, then the long runs' pieces; ugrad[j] (or the peer
// window row of run j) holds the sum of run j afterwards.  The three counters at a.n_long must be zero.
static void launch_reduce(const BwdArgs& a, const PeerOut& po, int G, int64_t M, size_t max_pieces, cudaStream_t s) {
  const int64_t med_max = M / (kFastRun + 1) + 1;
#define RED(GG)                                                                                               \
  run_sum_kernel<GG><<<resident_grid(run_sum_kernel<GG>, M, kThreads), kThreads, 0, s>>>(a, po);              \
  MONO_CHECK_LAUNCH();                                                                                        \
  run_sum_med_kernel<GG><<<(unsigned)((med_max + kThreads / GG - 1) / (kThreads / GG)), kThreads, 0, s>>>(a, po); \
  MONO_CHECK_LAUNCH();                                                                                        \
  long_prep_kernel<<<1, 1024, 0, s>>>(a);                                                                     \
  MONO_CHECK_LAUNCH();                                                                                        \
  pool_bwd_long_partial_kernel<GG>                                                                            \
      <<<resident_grid(pool_bwd_long_partial_kernel<GG>, (int64_t)max_pieces, 1), kThreads, 0, s>>>(a);       \
  MONO_CHECK_LAUNCH();                                                                                        \
  pool_bwd_long_final_kernel<GG><<<148 * 2, kThreads, 0, s>>>(a, po);                                         \
  MONO_CHECK_LAUNCH()
  switch (G) {
    case 4: RED(4); break;
    case 8: RED(8); break;
    case 16: RED(16); break;
    default: RED(32); break;
  }
#undef RED
}

// Claim set entry (16 B, viewed as Entry): key = FID, ts = epoch of the call that claimed it (any other value =
// empty: the set is never cleared, engine.h ClaimSet), row = the FID's resolved table row (RESOLVE) or 0.
// Claimed entries never change during the kernel, so every read may be served by L1 (plain ld.global): the hot
// FIDs of a Zipf batch (8 % of the occurrences hit ONE slot) are answered per SM instead of serialising on one
// L2 slice.  A stale L1 line can only show "empty" for a slot that has been claimed meanwhile; the CAS then
// fails and returns the true entry.
// RESOLVE (single-GPU fused backward): the thread that wins a slot is the only one for its FID, so it also
// resolves the FID in the table right away — lane-level probe, expiry-timestamp bump
// (ref: entry.SetTimestamp(update_time), cuckoo_embedding_hash_table.cc:243), row index parked in the set
// entry; absent FIDs are queued (set slot) for claim_miss_kernel.  No separate resolve pass over the uniques.
struct ClaimResolve {
  const TableDev* t;
  uint32_t update_ts;
  uint32_t* miss_ctr;
  uint32_t* miss_slots;
};

template <bool RESOLVE>
__global__ void __launch_bounds__(kThreads)
fid_claim_kernel(const int64_t* __restrict__ fids, int64_t n, Entry* set, uint32_t R, int N, uint32_t epoch,
                 uint32_t* __restrict__ slot_of, uint32_t* __restrict__ owner_cnt /* [N], [256] = overflow */,
                 ClaimResolve cr) {
  __shared__ uint32_t cnt[256];
  for (int d = threadIdx.x; d < 256; d += blockDim.x) cnt[d] = 0;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t key = __ldg(fids + i);
    const uint32_t owner = N == 1 ? 0u : (uint32_t)((uint64_t)key % (uint64_t)N);
    const uint32_t base = owner * R;
    uint32_t idx = __umulhi((uint32_t)(mix64((uint64_t)key) >> 24), R);
    uint32_t found = 0xFFFFFFFFu;
    bool won = false;
    for (uint32_t probes = 0; probes < R; ++probes) {
      Entry* p = set + base + idx;
      Entry e = ld_entry(p);  // L1-cacheable
      while (e.ts != epoch) {  // empty as far as we can see: claim it (CAS against what we saw)
        Entry ne;
        ne.key = key;
        ne.row = kEmptyRow;
        ne.ts = epoch;
        const Entry old = cas_entry_old(p, e, ne);
        if (old.key == e.key && old.row == e.row && old.ts == e.ts) {
          won = true;
          e = ne;
        } else {
          e = old;  // somebody else changed it: the true entry (claimed this epoch, or a different stale one)
        }
      }
      if (e.key == key) {
        found = base + idx;
        break;
      }
      idx = idx + 1 == R ? 0 : idx + 1;
    }
    if (found == 0xFFFFFFFFu) {  // region full (owner skew): the host retries with larger regions
      owner_cnt[256] = 1;
      found = base;
      won = false;
    }
    slot_of[i] = found;
    if (won) {
      atomicAdd(&cnt[owner], 1u);
      if (RESOLVE) {
        Entry* slot = nullptr;
        const uint32_t row = probe_lane_slot(cr.t, key, &slot);
        if (row != kEmptyRow) {
          slot->ts = cr.update_ts;
          set[found].row = row;
        } else {
          cr.miss_slots[atomicAdd(cr.miss_ctr, 1u)] = found;
        }
      }
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < N; d += blockDim.x)
    if (cnt[d]) atomicAdd(owner_cnt + d, cnt[d]);
}

// FIDs the claim found absent from the table (each exactly once): take a row (free list first, then the bump
// allocator), publish {fid, row, ts} with the lock-free cuckoo insert and park row | fresh in the set entry.
__global__ void __launch_bounds__(kThreads)
claim_miss_kernel(const TableDev* __restrict__ t, Entry* set, const uint32_t* __restrict__ miss_ctr,
                  const uint32_t* __restrict__ miss_slots, uint32_t update_ts) {
  const int64_t n = (int64_t)*miss_ctr;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t sl = miss_slots[q];
    const uint32_t ticket = atomicAdd(t->ctrs + kCtrMiss, 1u);
    const uint32_t fc = t->ctrs[kCtrFree];  // stable during this kernel (finalize updates it)
    const uint32_t row = ticket < fc ? t->free_list[fc - 1 - ticket] : t->ctrs[kCtrBump] + (ticket - fc);
    if (row >= t->row_cap) {
      atomicOr(t->ctrs + kCtrError, 2u);
      continue;  // the set entry keeps kEmptyRow: the run is skipped by the apply pass
    }
    Entry e;
    e.key = ld_entry_cg(set + sl).key;
    e.row = row;
    e.ts = update_ts;
    cuckoo_insert(t, e);
    set[sl].row = row | kFreshBit;
  }
}

static const PeerOut no_peer = {};  // n == 0: the reduce kernels write their local ugrad buffer

void run_pool_backward(mono_mtable* mt, int k, const int64_t* fids_dev, int64_t n_fids,
                       const int32_t* row_offsets, int64_t n_rows, int pooling,
                       const float* pooled_grad, int64_t grad_stride, int grad_col,
                       const float* lr_host, int64_t update_time, cudaStream_t s) {
  if (n_fids <= 0) return;
  if (n_fids > ((int64_t)1 << 30)) throw ArgError("more than 2^30 fids in one call");
  if (pooling != MONO_POOL_SUM && pooling != MONO_POOL_MEAN) throw ArgError("pool_backward: SUM or MEAN");
  HostTable& ht = mt->tables[k];
  const int D = ht.dim;
  if ((D & 3) || D > 128) throw ArgError("pool_backward needs dim % 4 == 0 and dim <= 128");
  if ((grad_stride & 3) || (grad_col & 3) || (reinterpret_cast<uintptr_t>(pooled_grad) & 15))
    throw ArgError("pool_backward needs 16-byte aligned gradient rows");
  const int64_t M = n_fids;
  ensure_capacity(mt, k, (uint64_t)M, s);
  upload_tables(mt, s);
  CallSeg sg;
  sg.id_begin = 0;
  sg.id_end = M;
  sg.val_off = 0;
  sg.table = k;
  sg.lr_off = 0;
  CallBlob cb = stage_call(mt, &sg, 1, lr_host, ht.slices, s);
  const int G = pick_group(D);

  // ---- scratch layout ----
  uint32_t cap = 1024;
  while (cap < 2 * (uint64_t)M) cap <<= 1;
  int bits = 0;
  while ((1u << bits) < cap) ++bits;
  const int nblk = (int)((M + kSortTile - 1) / kSortTile);
  const size_t n_long_max = (size_t)M / kShortRun + 2;
  const size_t max_pieces = (size_t)M / kSubRun + n_long_max + 2;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_k0 = take(4 * (size_t)M), o_v0 = take(4 * (size_t)M);
  const size_t o_k1 = take(4 * (size_t)M), o_v1 = take(4 * (size_t)M);
  const size_t o_blk = take(4 * (size_t)256 * nblk);
  const size_t o_ctr = take(4096 + 4 * 256 * 4);  // counters + digit totals per pass
  const size_t o_brun = take(4 * (size_t)nblk);
  const size_t o_rs = take(4 * ((size_t)M + 1)), o_rfp = take(4 * (size_t)M), o_ridx = take(4 * (size_t)M);
  const size_t o_miss = take(4 * (size_t)M);
  const size_t o_occ = take(row_offsets ? 4 * (size_t)M : 0);
  const size_t o_ll = take(4 * n_long_max), o_llen = take(4 * n_long_max), o_lsb = take(4 * (n_long_max + 1));
  const size_t o_part = take(sizeof(float) * max_pieces * D);
  const size_t o_pd = take(sizeof(uint2) * max_pieces);
  const size_t o_ug = take(sizeof(float) * (size_t)M * D);
  const size_t o_med = take(4 * ((size_t)M / (kFastRun + 1) + 2));
  char* ws = (char*)mt->ws_a.get(off, s);
  uint32_t epoch = 0;
  Entry* set = (Entry*)mt->claim_set.get(sizeof(Entry) * (size_t)cap, s, &epoch);
  uint32_t *k0 = (uint32_t*)(ws + o_k0), *v0 = (uint32_t*)(ws + o_v0);
  uint32_t *k1 = (uint32_t*)(ws + o_k1), *v1 = (uint32_t*)(ws + o_v1);
  int32_t* blk_cnt = (int32_t*)(ws + o_blk);
  uint32_t* ctr = (uint32_t*)(ws + o_ctr);  // [0] n_runs [4] n_long [8] miss_ctr
  int32_t* dtot = (int32_t*)(ws + o_ctr + 4096);
  MONO_CUDA(cudaMemsetAsync(ctr, 0, 4096 + 4 * 256 * 4, s));

  // 1 claim: k0[i] = set slot of occurrence i; the winner of a slot resolves its FID in the table (row parked in
  //   the set entry, expiry timestamp bumped) or queues it as absent
  ClaimResolve cr;
  cr.t = mt->d_tables + k;
  cr.update_ts = (uint32_t)update_time;
  cr.miss_ctr = ctr + 8;
  cr.miss_slots = (uint32_t*)(ws + o_miss);
  fid_claim_kernel<true><<<resident_grid(fid_claim_kernel<true>, M, kThreads), kThreads, 0, s>>>(
      fids_dev, M, set, cap, 1, epoch, k0, ctr + 512, cr);
  MONO_CHECK_LAUNCH();
  // 2 absent FIDs: allocate a row + lock-free insert (few in steady state: a small grid; count stays on the device)
  claim_miss_kernel<<<resident_grid(claim_miss_kernel, std::min<int64_t>(M, 148 * 2 * kThreads), kThreads), kThreads, 0, s>>>(
      mt->d_tables + k, set, ctr + 8, cr.miss_slots, (uint32_t)update_time);
  MONO_CHECK_LAUNCH();
  upsert_finalize_kernel<<<1, 64, 0, s>>>(mt->d_tables, cb.table_ids, cb.ntab, ctr + 8, (uint32_t)update_time);
  MONO_CHECK_LAUNCH();
  // 3 stable LSD radix sort of (slot, position)  +  4 ordered run list (with each run's resolved row)
  SortWs sw;
  sw.k0 = k0; sw.v0 = v0; sw.k1 = k1; sw.v1 = v1;
  sw.blk_cnt = blk_cnt; sw.dtot = dtot;
  sw.blk_runs = (uint32_t*)(ws + o_brun);
  sw.run_start = (uint32_t*)(ws + o_rs);
  sw.run_first_pos = (uint32_t*)(ws + o_rfp);
  sw.n_runs = ctr;
  sw.claim_set = set;
  sw.rowidx = (uint32_t*)(ws + o_ridx);
  const uint32_t* skeys = nullptr;
  const uint32_t* perm = nullptr;
  sort_and_runs(sw, M, bits, 0, &skeys, &perm, s);
  uint32_t* run_start = sw.run_start;
  uint32_t* run_first_pos = sw.run_first_pos;
  // 5 reduce + update
  BwdArgs a;
  a.td = ht.dev;  // descriptor is current: ensure_capacity / upload_tables ran above
  a.t = mt->d_tables + k;
  a.fids = fids_dev;
  a.skeys = skeys;
  a.perm = perm;
  a.n = M;
  a.n_runs = ctr;
  a.run_start = run_start;
  a.run_first_pos = run_first_pos;
  a.rowidx = sw.rowidx;
  a.occ_row = nullptr;
  a.row_offsets = row_offsets;
  a.pooling = pooling;
  a.pooled_grad = pooled_grad;
  a.grad_stride = grad_stride;
  a.grad_col = grad_col;
  a.lr = cb.lr;
  a.n_long = ctr + 4;
  a.long_list = (uint32_t*)(ws + o_ll);
  a.long_len = (uint32_t*)(ws + o_llen);
  a.long_sub_base = (uint32_t*)(ws + o_lsb);
  a.partial = (float*)(ws + o_part);
  a.piece_desc = (uint2*)(ws + o_pd);
  a.ugrad = (float*)(ws + o_ug);
  a.med_list = (uint32_t*)(ws + o_med);
  a.scratch = a.ugrad;
  if (row_offsets) {
    uint32_t* occ = (uint32_t*)(ws + o_occ);
    occ_row_kernel<<<resident_grid(occ_row_kernel, n_rows, kThreads), kThreads, 0, s>>>(row_offsets, n_rows, occ);
    MONO_CHECK_LAUNCH();
    a.occ_row = occ;
  }
  launch_reduce(a, no_peer, G, M, max_pieces, s);
#define BWD2(GG, OO)                                                                                             \
  runs_apply_kernel<GG, OO><<<resident_grid(runs_apply_kernel<GG, OO>, M, kThreads / GG), kThreads, 0, s>>>(a);  \
  MONO_CHECK_LAUNCH()
#define BWD(GG)                                                         \
  switch (opt_sel) {                                                    \
    case MONO_OPT_SGD: BWD2(GG, MONO_OPT_SGD); break;                   \
    case MONO_OPT_ADAGRAD: BWD2(GG, MONO_OPT_ADAGRAD); break;           \
    case MONO_OPT_FTRL: BWD2(GG, MONO_OPT_FTRL); break;                 \
    case MONO_OPT_ADAM: BWD2(GG, MONO_OPT_ADAM); break;                 \
    default: BWD2(GG, -1); break;                                       \
  }
  const int opt_sel = ht.segs.size() == 1 ? ht.segs[0].opt_type : -1;
  switch (G) {
    case 4: BWD(4); break;
    case 8: BWD(8); break;
    case 16: BWD(16); break;
    default: BWD(32); break;
  }
#undef BWD2
#undef BWD
  ht.issued_total += (uint64_t)M;
  ht.max_update_ts = std::max<int64_t>(ht.max_update_ts, update_time);
  request_snapshot(mt, k, s);
}


// run j of the sorted offsets -> destination row: out_rows + (sorted key << shift)
template <int G>
__global__ void __launch_bounds__(kThreads) runs_emit_kernel(BwdArgs a, const uint32_t* __restrict__ skeys,
                                                             int shift, float* __restrict__ out_rows) {
  const int gl = Group<G>::gl(), c = gl * 4;
  const int64_t nr = *a.n_runs;
  const int D = a.td.dim;
  const int64_t gstride = (int64_t)gridDim.x * (kThreads / G);
  for (int64_t j = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G; j < nr; j += gstride) {
    if (c >= D) continue;
    const float4 g = __ldcs(reinterpret_cast<const float4*>(a.ugrad + (size_t)j * D + c));
    *reinterpret_cast<float4*>(out_rows + ((size_t)skeys[a.run_start[j]] << shift) + c) = g;
  }
}

// Deterministic replacement of the float-atomic scatter of pooled-row gradients
// (ref: FusedGatherGradKernel, map_id_to_embedding.cu.cc:75-118; ScatterGrad,
// fused_embedding_to_layout.h:286-347): out_rows[offs[m] : +dim] = sum over the occurrences m with that
// offset of pooled_grad[row(m)] (or /n for MEAN), summed in occurrence order.  Rows of out_rows that
// no occurrence points at are left untouched.
void run_scatter_rows(int device, const int32_t* offs_dev, int64_t M, int dim, const int32_t* row_offsets,
                      int64_t n_rows, int pooling, const float* pooled_grad, int64_t grad_stride,
                      int grad_col, float* out_rows, int64_t total_floats, cudaStream_t s) {
  MONO_CUDA(cudaSetDevice(device));
  if (M <= 0) return;
  if (M >= ((int64_t)1 << 31) || total_floats >= ((int64_t)1 << 31)) throw ArgError("scatter_rows: sizes exceed 2^31");
  if (pooling != MONO_POOL_SUM && pooling != MONO_POOL_MEAN) throw ArgError("scatter_rows: SUM or MEAN");
  if ((dim & 3) || dim > 128) throw ArgError("scatter_rows needs dim % 4 == 0 and dim <= 128");
  if ((grad_stride & 3) || (grad_col & 3) || (reinterpret_cast<uintptr_t>(pooled_grad) & 15) ||
      (reinterpret_cast<uintptr_t>(out_rows) & 15))
    throw ArgError("scatter_rows needs 16-byte aligned rows");
  const int D = dim;
  // offsets are multiples of 4 floats (16-byte aligned rows; every dim in the fused buffer is a multiple
  // of 4): drop the two zero bits from the sort key
  const int shift = 2;
  int bits = 1;
  while (((int64_t)1 << bits) < ((total_floats >> shift) + 1)) ++bits;
  const int passes = std::max(1, (bits + 7) / 8);
  const int nblk = (int)((M + kSortTile - 1) / kSortTile);
  const size_t n_long_max = (size_t)M / kShortRun + 2;
  const size_t max_pieces = (size_t)M / kSubRun + n_long_max + 2;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_v0 = take(4 * (size_t)M), o_k1 = take(4 * (size_t)M), o_v1 = take(4 * (size_t)M);
  const size_t o_k0b = take(4 * (size_t)M);
  const size_t o_blk = take(4 * (size_t)256 * nblk);
  const size_t o_ctr = take(4096 + 4 * 256 * 4);
  const size_t o_brun = take(4 * (size_t)nblk);
  const size_t o_rs = take(4 * ((size_t)M + 1)), o_rfp = take(4 * (size_t)M);
  const size_t o_occ = take(row_offsets ? 4 * (size_t)M : 0);
  const size_t o_ll = take(4 * n_long_max), o_llen = take(4 * n_long_max), o_lsb = take(4 * (n_long_max + 1));
  const size_t o_part = take(sizeof(float) * max_pieces * D);
  const size_t o_pd = take(sizeof(uint2) * max_pieces);
  const size_t o_ug = take(sizeof(float) * (size_t)M * D);
  const size_t o_med = take(4 * ((size_t)M / (kFastRun + 1) + 2));
  char* ws = nullptr;
  MONO_CUDA(cudaMallocAsync((void**)&ws, off, s));
  uint32_t* ctr = (uint32_t*)(ws + o_ctr);
  MONO_CUDA(cudaMemsetAsync(ctr, 0, 4096 + 4 * 256 * 4, s));
  SortWs sw;
  // private copy of the keys: the sort ping-pongs between k0 and k1 and must not touch the caller's offsets
  sw.k0 = (uint32_t*)(ws + o_k0b);
  MONO_CUDA(cudaMemcpyAsync(sw.k0, offs_dev, 4 * (size_t)M, cudaMemcpyDeviceToDevice, s));
  sw.v0 = (uint32_t*)(ws + o_v0);
  sw.k1 = (uint32_t*)(ws + o_k1);
  sw.v1 = (uint32_t*)(ws + o_v1);
  sw.blk_cnt = (int32_t*)(ws + o_blk);
  sw.dtot = (int32_t*)(ws + o_ctr + 4096);
  sw.blk_runs = (uint32_t*)(ws + o_brun);
  sw.run_start = (uint32_t*)(ws + o_rs);
  sw.run_first_pos = (uint32_t*)(ws + o_rfp);
  sw.n_runs = ctr;
  const uint32_t* skeys = nullptr;
  const uint32_t* perm = nullptr;
  (void)passes;
  sort_and_runs(sw, M, bits, shift, &skeys, &perm, s);
  BwdArgs a;
  std::memset(&a, 0, sizeof(a));
  a.td.dim = D;
  a.skeys = skeys;
  a.perm = perm;
  a.n = M;
  a.n_runs = ctr;
  a.run_start = sw.run_start;
  a.run_first_pos = sw.run_first_pos;
  a.row_offsets = row_offsets;
  a.pooling = pooling;
  a.pooled_grad = pooled_grad;
  a.grad_stride = grad_stride;
  a.grad_col = grad_col;
  a.n_long = ctr + 4;
  a.long_list = (uint32_t*)(ws + o_ll);
  a.long_len = (uint32_t*)(ws + o_llen);
  a.long_sub_base = (uint32_t*)(ws + o_lsb);
  a.partial = (float*)(ws + o_part);
  a.piece_desc = (uint2*)(ws + o_pd);
  a.ugrad = (float*)(ws + o_ug);
  a.med_list = (uint32_t*)(ws + o_med);
  if (row_offsets) {
    uint32_t* occ = (uint32_t*)(ws + o_occ);
    occ_row_kernel<<<resident_grid(occ_row_kernel, n_rows, kThreads), kThreads, 0, s>>>(row_offsets, n_rows, occ);
    MONO_CHECK_LAUNCH();
    a.occ_row = occ;
  }
  const int G = pick_group(D);
  launch_reduce(a, no_peer, G, M, max_pieces, s);
#define EMIT(GG)                                                                                              \
  runs_emit_kernel<GG><<<resident_grid(runs_emit_kernel<GG>, M, kThreads / GG), kThreads, 0, s>>>(a, skeys, shift, out_rows); \
  MONO_CHECK_LAUNCH()
  switch (G) {
    case 4: EMIT(4); break;
    case 8: EMIT(8); break;
    case 16: EMIT(16); break;
    default: EMIT(32); break;
  }
#undef EMIT
  MONO_CUDA(cudaFreeAsync(ws, s));
}


// ==========================================================================================
// Owner grouping: ONE grouping of a batch's FID occurrences shared by the forward (dedup + bucket by
// owner for the exchange) and the backward (deterministic per-FID gradient reduction) of the sharded
// step.  Functionally FusedReorderByIndices (ref: fused_reorder_by_indices.cc:38-123) for a single id
// list, except that the order of the distinct FIDs inside a shard is the engine's (scratch-set slot
// order), not first-occurrence order; mono_reorder_by_indices is the bit-exact op.
//
// The scratch set is split into N regions of R slots and a FID lives in region owner(fid) = fid mod N
// (linear probing wraps inside the region).  Sorting the occurrences by slot therefore yields the runs
// (one per distinct FID) already bucketed by owner: no separate partition pass, and a run's index IS its
// position in the bucketed unique list.  The per-owner distinct counts fall out of the claim kernel (one
// count per successful insert), i.e. after the FIRST kernel: they are copied to the host on a side
// stream while the sort still runs, so the host can size and enqueue the exchange without idling the GPU.
// ==========================================================================================
// bucketed unique list and the per-occurrence row offsets: run j (slot order == owner-bucketed order)
__global__ void __launch_bounds__(kThreads)
group_emit_kernel(const int64_t* __restrict__ fids, const uint32_t* __restrict__ perm,
                  const uint32_t* __restrict__ run_of_sorted, const uint32_t* __restrict__ run_first_pos,
                  const uint32_t* __restrict__ n_runs, int64_t n, int dim, int32_t* __restrict__ occ_offset,
                  int64_t* __restrict__ uniq_out) {
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = t0; i < n; i += stride) occ_offset[perm[i]] = (int32_t)(run_of_sorted[i] * (uint32_t)dim);
  const int64_t nr = *n_runs;
  for (int64_t j = t0; j < nr; j += stride) uniq_out[j] = fids[run_first_pos[j]];
}

void grouping_build(mono_grouping* g, const int64_t* fids_dev, int64_t M, int N, int dim,
                    int64_t* uniq_out, int32_t* occ_offset_out, int32_t* shard_counts_host,
                    int64_t* n_unique_host, cudaStream_t s) {
  MONO_CUDA(cudaSetDevice(g->device));
  if (N <= 0 || N > 256) throw ArgError("grouping: num_shards must be in [1, 256]");
  if (M < 0 || M >= ((int64_t)1 << 29)) throw ArgError("grouping: bad occurrence count");
  if ((dim & 3) || dim <= 0 || dim > 128) throw ArgError("grouping needs dim % 4 == 0 and dim <= 128");
  g->M = M;
  g->dim = dim;
  g->skeys = nullptr;
  if (M == 0) {
    for (int n = 0; n < N; ++n) shard_counts_host[n] = 0;
    if (n_unique_host) *n_unique_host = 0;
    return;
  }
  if (!g->h_counts) {
    MONO_CUDA(cudaHostAlloc((void**)&g->h_counts, 4 * 260, cudaHostAllocDefault));
    MONO_CUDA(cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking));
    MONO_CUDA(cudaEventCreateWithFlags(&g->ev_claimed, cudaEventDisableTiming));
    MONO_CUDA(cudaEventCreateWithFlags(&g->ev_copied, cudaEventDisableTiming));
  }
  uint64_t cap0 = 1024;
  while (cap0 < 2 * (uint64_t)M) cap0 <<= 1;
  for (int attempt = 0; attempt < 2; ++attempt) {
    // attempt 0: the regions share 2M slots (load <= 0.5 for hash-balanced owners);
    // attempt 1 (a region overflowed: heavily skewed owners): every region can hold all M FIDs.
    const uint64_t cap = attempt == 0 ? cap0 : cap0 * (uint64_t)N;
    if (cap > ((uint64_t)1 << 31)) throw ArgError("grouping: FID owners too skewed for this batch size");
    const uint32_t R = (uint32_t)(cap / (uint64_t)N);
    int bits = 0;
    while (((uint64_t)1 << bits) < cap) ++bits;
    const int nblk = (int)((M + kSortTile - 1) / kSortTile);
    const size_t n_long_max = (size_t)M / kShortRun + 2;
    const size_t max_pieces = (size_t)M / kSubRun + n_long_max + 2;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_k0 = take(4 * (size_t)M), o_v0 = take(4 * (size_t)M), o_k1 = take(4 * (size_t)M), o_v1 = take(4 * (size_t)M);
    const size_t o_blk = take(4 * (size_t)256 * nblk);
    const size_t o_ctr = take(4096 + 4 * 256 * 4 + 4 * 260);
    const size_t o_brun = take(4 * (size_t)nblk);
    const size_t o_rs = take(4 * ((size_t)M + 1)), o_rfp = take(4 * (size_t)M), o_ros = take(4 * (size_t)M);
    const size_t o_tail = off;
    // reduce() scratch
    take(4 * (size_t)M);                                                             // occ_row
    take(4 * n_long_max); take(4 * n_long_max); take(4 * (n_long_max + 1));          // long run lists
    take(sizeof(float) * max_pieces * dim);                                          // partial sums
    take(sizeof(uint2) * max_pieces);                                                // piece descriptors
    take(4 * ((size_t)M / (kFastRun + 1) + 2));                                      // medium run list
    char* ws = (char*)g->ws.get(off, s);
    g->tail = ws + o_tail;
    g->tail_bytes = off - o_tail;
    uint32_t epoch = 0;
    Entry* set = (Entry*)g->claim_set.get(sizeof(Entry) * (size_t)cap, s, &epoch);
    uint32_t* ctr = (uint32_t*)(ws + o_ctr);
    int32_t* dtot = (int32_t*)(ws + o_ctr + 4096);
    uint32_t* owner_cnt = (uint32_t*)(ws + o_ctr + 4096 + 4 * 256 * 4);
    MONO_CUDA(cudaMemsetAsync(ctr, 0, 4096 + 4 * 256 * 4 + 4 * 260, s));
    SortWs sw;
    sw.k0 = (uint32_t*)(ws + o_k0); sw.v0 = (uint32_t*)(ws + o_v0);
    sw.k1 = (uint32_t*)(ws + o_k1); sw.v1 = (uint32_t*)(ws + o_v1);
    sw.blk_cnt = (int32_t*)(ws + o_blk);
    sw.dtot = dtot;
    sw.blk_runs = (uint32_t*)(ws + o_brun);
    sw.run_start = (uint32_t*)(ws + o_rs);
    sw.run_first_pos = (uint32_t*)(ws + o_rfp);
    sw.n_runs = ctr;
    sw.run_of_sorted = (uint32_t*)(ws + o_ros);
    fid_claim_kernel<false><<<resident_grid(fid_claim_kernel<false>, M, kThreads), kThreads, 0, s>>>(
        fids_dev, M, set, R, N, epoch, sw.k0, owner_cnt, ClaimResolve{});
    MONO_CHECK_LAUNCH();
    // counts -> host on the side stream, while the sort below keeps the GPU busy
    MONO_CUDA(cudaEventRecord(g->ev_claimed, s));
    MONO_CUDA(cudaStreamWaitEvent(g->side, g->ev_claimed, 0));
    MONO_CUDA(cudaMemcpyAsync(g->h_counts, owner_cnt, 4 * 257, cudaMemcpyDeviceToHost, g->side));
    MONO_CUDA(cudaEventRecord(g->ev_copied, g->side));
    const uint32_t* skeys = nullptr;
    const uint32_t* perm = nullptr;
    sort_and_runs(sw, M, bits, 0, &skeys, &perm, s);
    group_emit_kernel<<<resident_grid(group_emit_kernel, M, kThreads), kThreads, 0, s>>>(
        fids_dev, perm, sw.run_of_sorted, sw.run_first_pos, ctr, M, dim, occ_offset_out, uniq_out);
    MONO_CHECK_LAUNCH();
    MONO_CUDA(cudaEventSynchronize(g->ev_copied));
    if (g->h_counts[256] != 0) continue;  // region overflow: redo with full-size regions
    int64_t total = 0;
    for (int n = 0; n < N; ++n) {
      shard_counts_host[n] = (int32_t)g->h_counts[n];
      total += g->h_counts[n];
    }
    if (n_unique_host) *n_unique_host = total;
    g->skeys = skeys;
    g->perm = perm;
    g->run_start = sw.run_start;
    g->run_first_pos = sw.run_first_pos;
    g->ctr = ctr;
    return;
  }
  throw ArgError("grouping: scratch set overflow (internal)");
}

static void grouping_reduce_impl(mono_grouping* g, const float* pooled_grad, int64_t grad_stride, int grad_col,
                                 const int32_t* row_offsets, int64_t n_rows, int pooling, float* out_rows,
                                 const PeerOut& po, cudaStream_t s) {
  MONO_CUDA(cudaSetDevice(g->device));
  const int64_t M = g->M;
  const int D = g->dim;
  if (M <= 0) return;
  if (!g->skeys) throw ArgError("grouping_reduce before grouping_build");
  if (pooling != MONO_POOL_SUM && pooling != MONO_POOL_MEAN) throw ArgError("grouping_reduce: SUM or MEAN");
  if ((grad_stride & 3) || (grad_col & 3) || (reinterpret_cast<uintptr_t>(pooled_grad) & 15) ||
      (po.n == 0 && (reinterpret_cast<uintptr_t>(out_rows) & 15)))
    throw ArgError("grouping_reduce needs 16-byte aligned rows");
  const size_t n_long_max = (size_t)M / kShortRun + 2;
  const size_t max_pieces = (size_t)M / kSubRun + n_long_max + 2;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_occ = take(4 * (size_t)M);
  const size_t o_ll = take(4 * n_long_max), o_llen = take(4 * n_long_max), o_lsb = take(4 * (n_long_max + 1));
  const size_t o_part = take(sizeof(float) * max_pieces * D);
  const size_t o_pd = take(sizeof(uint2) * max_pieces);
  const size_t o_med = take(4 * ((size_t)M / (kFastRun + 1) + 2));
  if (off > g->tail_bytes) throw ArgError("grouping scratch too small (internal)");
  char* ws = g->tail;
  BwdArgs a;
  std::memset(&a, 0, sizeof(a));
  a.td.dim = D;
  a.skeys = g->skeys;
  a.perm = g->perm;
  a.n = M;
  a.n_runs = g->ctr;
  a.run_start = g->run_start;
  a.run_first_pos = g->run_first_pos;
  a.row_offsets = row_offsets;
  a.pooling = pooling;
  a.pooled_grad = pooled_grad;
  a.grad_stride = grad_stride;
  a.grad_col = grad_col;
  a.n_long = g->ctr + 4;
  a.long_list = (uint32_t*)(ws + o_ll);
  a.long_len = (uint32_t*)(ws + o_llen);
  a.long_sub_base = (uint32_t*)(ws + o_lsb);
  a.partial = (float*)(ws + o_part);
  a.piece_desc = (uint2*)(ws + o_pd);
  a.ugrad = out_rows;  // runs are already in the bucketed order: the sums are written in place
  a.med_list = (uint32_t*)(ws + o_med);
  MONO_CUDA(cudaMemsetAsync(g->ctr + 4, 0, 12, s));  // n_long, n_med, work counter
  if (row_offsets) {
    uint32_t* occ = (uint32_t*)(ws + o_occ);
    occ_row_kernel<<<resident_grid(occ_row_kernel, n_rows, kThreads), kThreads, 0, s>>>(row_offsets, n_rows, occ);
    MONO_CHECK_LAUNCH();
    a.occ_row = occ;
  }
  launch_reduce(a, po, pick_group(D), M, max_pieces, s);
}

// out_rows[u] = summed gradient of the u-th distinct FID of the bucketed list (ref: ScatterGrad)
void grouping_reduce(mono_grouping* g, const float* pooled_grad, int64_t grad_stride, int grad_col,
                     const int32_t* row_offsets, int64_t n_rows, int pooling, float* out_rows, cudaStream_t s) {
  grouping_reduce_impl(g, pooled_grad, grad_stride, grad_col, row_offsets, n_rows, pooling, out_rows, no_peer, s);
}

// same, fused with the gradient exchange of the sharded backward: the row of the u-th distinct FID is stored
// into the window of the rank that owns it (replaces the gradient all-to-all, ref:
// distributed_ps_sync.py:531-573); the NVLink stores overlap the reduction, run by run.
void grouping_reduce_push(mono_grouping* g, const float* pooled_grad, int64_t grad_stride, int grad_col,
                          const int32_t* row_offsets, int64_t n_rows, int pooling, const PeerOut& po,
                          cudaStream_t s) {
  if (po.n <= 0) throw ArgError("grouping_reduce_push without a peer window");
  grouping_reduce_impl(g, pooled_grad, grad_stride, grad_col, row_offsets, n_rows, pooling, nullptr, po, s);
}

}  // namespace mono
