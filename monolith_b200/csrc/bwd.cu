// bwd.cu — the fused sparse backward: pooled-gradient scatter + sparse optimizer + expiry bump, deterministic,
// without float atomics; and the owner grouping of the sharded step that shares its machinery.
//
// Replaces (for one table) the reference's ScatterGrad / BackwardBatchKernel (fused_embedding_to_layout.h:286-347,
// .cu.cc:337-381: float atomicAdd per occurrence into a per-unique-FID gradient buffer) followed by
// MultiHashTableOptimize (multi_hash_table_update_op.cc:47-89; embedding_hash_table_tf_bridge.cc:258-341).
// Pipeline, all on the caller's stream, nothing returns to the host:
//   1 claim   : every occurrence finds / claims the scratch-set slot of its FID (same slot <=> same FID); the
//               winner of a slot resolves the FID in the table (row index, expiry-timestamp bump) or queues it
//   2 insert  : absent FIDs get a row and are published with the lock-free cuckoo insert
//   3 sort    : stable LSD radix sort of (slot, position), 11 bits per pass (2 passes for 2 M occurrences):
//               the occurrences of a FID become one contiguous run in position order
//   4 runs    : run heads -> ordered run list (start, first position, resolved row)
//   5 reduce + update : seg_reduce_kernel streams the sorted occurrences in 32-occurrence pieces; runs of <= 64
//               occurrences are summed in occurrence order (the CPU reference's order: bit-exact) and the optimizer is
//               applied from registers; hot FIDs are cut into run-aligned 32-occurrence blocks combined by a fixed
//               32-ary tree (deterministic, pairwise-accurate)
#include <algorithm>
#include <cstring>

#include "rowops.cuh"

namespace mono {

// ==========================================================================================
// Stable LSD radix sort of (key, position), 11 bits per pass.
// Items travel as uint2 {key, position} (ONE 8-byte scattered store per item and pass); the first pass reads the
// bare keys (position = index).  Per pass: tile histograms (shared-memory atomics) -> column scan over the tiles
// -> stable scatter.  The scatter kernel ranks an item among the equal digits in front of it inside its tile
// with ballots (11 votes give the lane mask of equal digits; lanes are consecutive positions, so the count of
// lower lanes is the stable rank; one MATCH.ANY instead of the 11 votes measured 2.6 us slower over the two passes) and per-warp digit counters in shared memory; tiles are ordered by the scanned
// histograms, so no atomic ever decides an output position.
// ==========================================================================================
constexpr int kRadixBits = 11;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kRadixIPT = 16;                         // items per thread
constexpr int kRadixTile = kThreads * kRadixIPT;      // 4096 items per block iteration
constexpr int kRadixWarps = kThreads / 32;

template <bool FIRST>
__device__ __forceinline__ uint2 radix_load(const void* __restrict__ in, int64_t i, int pre_shift) {
  if (FIRST) return make_uint2(static_cast<const uint32_t*>(in)[i] >> pre_shift, (uint32_t)i);
  return static_cast<const uint2*>(in)[i];
}

template <bool FIRST>
__global__ void __launch_bounds__(kThreads)
radix_hist_kernel(const void* __restrict__ in, int64_t n, int shift, int pre_shift, uint32_t* __restrict__ cnt /*[nblk][2048]*/,
                  int nblk) {
  __shared__ uint32_t h[kRadixBins];
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    for (int d = threadIdx.x; d < kRadixBins; d += kThreads) h[d] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blk * kRadixTile;
#pragma unroll
    for (int r = 0; r < kRadixIPT; ++r) {
      const int64_t i = base + r * kThreads + threadIdx.x;
      if (i < n) atomicAdd(&h[(radix_load<FIRST>(in, i, pre_shift).x >> shift) & (kRadixBins - 1)], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < kRadixBins; d += kThreads) cnt[(size_t)blk * kRadixBins + d] = h[d];
    __syncthreads();
  }
}

// exclusive prefix over the tiles, per digit column (in place), and the digit totals.
// Block = 32 digit columns x 32 tile chunks.
__global__ void __launch_bounds__(1024)
radix_colscan_kernel(uint32_t* __restrict__ cnt, int nblk, uint32_t* __restrict__ dtot) {
  __shared__ uint32_t sums[32][33];
  const int dx = threadIdx.x & 31, ch = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + dx;
  const int per = (nblk + 31) / 32;
  const int t0 = min(nblk, ch * per), t1 = min(nblk, t0 + per);
  uint32_t s = 0;
  for (int t = t0; t < t1; ++t) s += cnt[(size_t)t * kRadixBins + d];
  sums[ch][dx] = s;
  __syncthreads();
  if (ch == 0) {
    uint32_t run = 0;
    for (int q = 0; q < 32; ++q) {
      const uint32_t v = sums[q][dx];
      sums[q][dx] = run;
      run += v;
    }
    dtot[d] = run;
  }
  __syncthreads();
  uint32_t run = sums[ch][dx];
  for (int t = t0; t < t1; ++t) {
    const uint32_t v = cnt[(size_t)t * kRadixBins + d];
    cnt[(size_t)t * kRadixBins + d] = run;
    run += v;
  }
}

template <bool FIRST>
__global__ void __launch_bounds__(kThreads, 4)
radix_scatter_kernel(const void* __restrict__ in, int64_t n, int shift, int pre_shift,
                     const uint32_t* __restrict__ goff /*[nblk][2048] scanned*/, const uint32_t* __restrict__ dtot,
                     int nblk, uint2* __restrict__ out) {
  __shared__ uint16_t wc[kRadixWarps][kRadixBins];   // per-warp digit counters, then exclusive warp offsets
  __shared__ uint32_t bbase[kRadixBins];             // global offset of (tile, digit)
  __shared__ uint32_t wsum[kRadixWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  // exclusive scan of the 2048 digit totals; thread t keeps the bases of its digits q * 256 + t in registers
  uint32_t dbase[kRadixBins / kThreads];
  {
    uint32_t carry = 0;
#pragma unroll
    for (int q = 0; q < kRadixBins / kThreads; ++q) {
      const uint32_t v = dtot[q * kThreads + threadIdx.x];
      uint32_t x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) wsum[w] = x;
      __syncthreads();
      uint32_t off = 0, tot = 0;
#pragma unroll
      for (int ww = 0; ww < kRadixWarps; ++ww) {
        const uint32_t t = wsum[ww];
        if (ww < w) off += t;
        tot += t;
      }
      dbase[q] = carry + off + x - v;
      carry += tot;
      __syncthreads();
    }
  }
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    uint32_t* wz = reinterpret_cast<uint32_t*>(&wc[0][0]);
    for (int q = threadIdx.x; q < kRadixWarps * kRadixBins / 2; q += kThreads) wz[q] = 0;
    __syncthreads();
    // ---- phase 1: stable rank of every item among the equal digits of its warp's 512-item slice ----
    const int64_t wbeg = (int64_t)blk * kRadixTile + (int64_t)w * (kRadixTile / kRadixWarps);
    uint32_t dgv[kRadixIPT];   // all 16 digits first: 16 independent loads in flight (0xFFFFFFFF = past the end)
#pragma unroll
    for (int r = 0; r < kRadixIPT; ++r) {
      const int64_t i = wbeg + r * 32 + lane;
      dgv[r] = i < n ? (radix_load<FIRST>(in, i, pre_shift).x >> shift) & (kRadixBins - 1) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int r = 0; r < kRadixIPT; ++r) {  // afterwards dgv[r] = digit | rank << 16 (rank < 1024)
      const bool valid = dgv[r] != 0xFFFFFFFFu;
      const uint32_t dg = valid ? dgv[r] : 0u;
      uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
      for (int b = 0; b < kRadixBits; ++b) {
        const uint32_t bal = __ballot_sync(0xffffffffu, (dg >> b) & 1u);
        peers &= ((dg >> b) & 1u) ? bal : ~bal;
      }
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (valid && lane == leader) {
        old = wc[w][dg];
        wc[w][dg] = (uint16_t)(old + __popc(peers));
      }
      old = __shfl_sync(0xffffffffu, old, max(leader, 0));
      if (valid) dgv[r] = dg | ((old + __popc(peers & lt)) << 16);
      __syncwarp();
    }
    __syncthreads();
    // ---- warp offsets per digit and the tile's global digit offsets ----
#pragma unroll
    for (int q = 0; q < kRadixBins / kThreads; ++q) {
      const int d = q * kThreads + threadIdx.x;
      uint32_t run = 0;
#pragma unroll
      for (int ww = 0; ww < kRadixWarps; ++ww) {
        const uint32_t v = wc[ww][d];
        wc[ww][d] = (uint16_t)run;
        run += v;
      }
      bbase[d] = goff[(size_t)blk * kRadixBins + d] + dbase[q];
    }
    __syncthreads();
    // ---- phase 2: scatter (digit and rank are in registers; the items are read again, 8 loads in flight) ----
#pragma unroll
    for (int h = 0; h < kRadixIPT; h += 8) {
      uint2 itv[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int64_t i = wbeg + (h + r) * 32 + lane;
        itv[r] = make_uint2(0u, 0u);
        if (dgv[h + r] != 0xFFFFFFFFu) itv[r] = radix_load<FIRST>(in, i, pre_shift);
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint32_t v = dgv[h + r];
        if (v != 0xFFFFFFFFu) {
          const uint32_t dg = v & (kRadixBins - 1);
          out[bbase[dg] + wc[w][dg] + (v >> 16)] = itv[r];
        }
      }
    }
    __syncthreads();
  }
}

// ---- ordered run compaction: run j = j-th distinct key of the sorted array ----
__device__ __forceinline__ bool is_run_start(const uint2* __restrict__ sorted, int64_t i, int64_t n) {
  return i < n && (i == 0 || sorted[i].x != sorted[i - 1].x);
}

constexpr int kRunTile = kThreads * 8;

__global__ void __launch_bounds__(kThreads)
runs_count_kernel(const uint2* __restrict__ sorted, int64_t n, int nblk, uint32_t* __restrict__ blk_runs) {
  __shared__ uint32_t wc[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    uint32_t cnt = 0;
    for (int c = 0; c < kRunTile; c += kThreads)
      cnt += __popc(__ballot_sync(0xffffffffu, is_run_start(sorted, (int64_t)blk * kRunTile + c + threadIdx.x, n)));
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

// run list: start (sorted index), first position, resolved row (when the claim parked it in the set), and per
// 32-occurrence piece the number of run heads in front of it (what seg_reduce_kernel indexes runs with)
__global__ void __launch_bounds__(kThreads)
runs_write_kernel(const uint2* __restrict__ sorted, int64_t n, int nblk, const uint32_t* __restrict__ blk_runs,
                  uint32_t* __restrict__ run_start, uint32_t* __restrict__ run_first_pos,
                  uint32_t* __restrict__ piece_run_base, uint32_t* __restrict__ run_of_sorted /* optional */,
                  const Entry* __restrict__ claim_set /* optional */, uint32_t* __restrict__ rowidx) {
  __shared__ uint32_t wc[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int NW = kThreads / 32;
  constexpr int kPerWarp = kRunTile / NW;
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t wbeg = (int64_t)blk * kRunTile + (int64_t)w * kPerWarp;
    uint32_t cnt = 0;
    for (int c = 0; c < kPerWarp; c += 32)
      cnt += __popc(__ballot_sync(0xffffffffu, is_run_start(sorted, wbeg + c + lane, n)));
    if (lane == 0) wc[w] = cnt;
    __syncthreads();
    uint32_t base = blk_runs[blk];
    for (int ww = 0; ww < w; ++ww) base += wc[ww];
    for (int c = 0; c < kPerWarp; c += 32) {
      const int64_t i = wbeg + c + lane;
      const bool st = is_run_start(sorted, i, n);
      const uint32_t bal = __ballot_sync(0xffffffffu, st);
      if (lane == 0 && wbeg + c < n) piece_run_base[(wbeg + c) >> 5] = base;
      if (st) {
        const uint32_t j = base + __popc(bal & ((1u << lane) - 1u));
        const uint2 it = sorted[i];
        run_start[j] = (uint32_t)i;
        run_first_pos[j] = it.y;
        if (claim_set) rowidx[j] = claim_set[it.x].row;
      }
      if (run_of_sorted && i < n) run_of_sorted[i] = base + __popc(bal & (0xffffffffu >> (31 - lane))) - 1;
      base += __popc(bal);
    }
    __syncthreads();
  }
}

struct SortWs {
  uint32_t* k0;                 // input keys (position = index)
  uint2 *a, *b;                 // ping-pong item buffers [M]
  uint32_t* cnt;                // [nblk_radix][2048]
  uint32_t* dtot;               // [passes][2048]
  uint32_t* blk_runs;           // [nblk_runs]
  uint32_t* run_start;          // [M + 1]
  uint32_t* run_first_pos;      // [M]
  uint32_t* piece_run_base;     // [ceil(M / 32)]
  uint32_t* n_runs;             // device counter
  uint32_t* run_of_sorted = nullptr;  // optional [M]: run index of every sorted element
  const Entry* claim_set = nullptr;   // optional: the claim set (keys are its slots) holding resolved rows ...
  uint32_t* rowidx = nullptr;         // ... gathered per run into rowidx[j]
  static int radix_blocks(int64_t M) { return (int)((M + kRadixTile - 1) / kRadixTile); }
  static int run_blocks(int64_t M) { return (int)((M + kRunTile - 1) / kRunTile); }
  static int passes(int bits) { return std::max(1, (bits + kRadixBits - 1) / kRadixBits); }
};

// stable LSD radix sort of (key >> pre_shift, position) over `bits` key bits, then the ordered run list (one run
// per distinct key).  Everything stays on the stream; counts stay on the device.
static const uint2* sort_and_runs(const SortWs& w, int64_t M, int bits, int pre_shift, cudaStream_t s) {
  const int passes = SortWs::passes(bits);
  const int nb = SortWs::radix_blocks(M);
  const void* in = w.k0;
  uint2* out = w.a;
  for (int p = 0; p < passes; ++p) {
    uint32_t* dt = w.dtot + kRadixBins * p;
    const int ps = p == 0 ? pre_shift : 0;
    if (p == 0) {
      radix_hist_kernel<true><<<resident_grid(radix_hist_kernel<true>, nb, 1), kThreads, 0, s>>>(in, M, 0, ps, w.cnt, nb);
      MONO_CHECK_LAUNCH();
    } else {
      radix_hist_kernel<false><<<resident_grid(radix_hist_kernel<false>, nb, 1), kThreads, 0, s>>>(in, M, kRadixBits * p, 0, w.cnt, nb);
      MONO_CHECK_LAUNCH();
    }
    radix_colscan_kernel<<<kRadixBins / 32, 1024, 0, s>>>(w.cnt, nb, dt);
    MONO_CHECK_LAUNCH();
    if (p == 0) {
      radix_scatter_kernel<true><<<resident_grid(radix_scatter_kernel<true>, nb, 1), kThreads, 0, s>>>(in, M, 0, ps, w.cnt, dt, nb, out);
      MONO_CHECK_LAUNCH();
    } else {
      radix_scatter_kernel<false><<<resident_grid(radix_scatter_kernel<false>, nb, 1), kThreads, 0, s>>>(in, M, kRadixBits * p, 0, w.cnt, dt, nb, out);
      MONO_CHECK_LAUNCH();
    }
    in = out;
    out = (out == w.a) ? w.b : w.a;
  }
  const uint2* sorted = static_cast<const uint2*>(in);
  const int nr = SortWs::run_blocks(M);
  runs_count_kernel<<<resident_grid(runs_count_kernel, nr, 1), kThreads, 0, s>>>(sorted, M, nr, w.blk_runs);
  MONO_CHECK_LAUNCH();
  runs_scan_kernel<<<1, 1024, 0, s>>>(w.blk_runs, nr, w.n_runs, w.run_start, M);
  MONO_CHECK_LAUNCH();
  runs_write_kernel<<<resident_grid(runs_write_kernel, nr, 1), kThreads, 0, s>>>(
      sorted, M, nr, w.blk_runs, w.run_start, w.run_first_pos, w.piece_run_base, w.run_of_sorted, w.claim_set, w.rowidx);
  MONO_CHECK_LAUNCH();
  return sorted;
}

// ==========================================================================================
// gradient source and the per-row optimizer step
// ==========================================================================================
constexpr int kShortRun = 64;

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
  const uint2* sorted;   // {scratch-set slot (sort key), occurrence position}, sorted by key, stable
  int64_t n;             // occurrences
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
  float* ugrad;                 // [runs][D] summed gradient of every run (= unique FID), run order (MODE_STORE)
  float* scratch;               // alias of ugrad for the generic (multi-segment) apply path
};

// Where the gradient rows come from: copied out of the kernel parameters once per thread so that the
// inner loops do not re-read the constant bank (ncu showed LDCU stalls inside the unrolled loads).
struct GradSrc {
  const float* base;            // pooled_grad + grad_col + lane column
  int64_t stride;
  const uint32_t* occ_row;
  const int32_t* row_offsets;
  bool mean;
};
__device__ __forceinline__ GradSrc make_grad_src(const BwdArgs& a, int c) {
  GradSrc g;
  g.base = a.pooled_grad + a.grad_col + c;
  g.stride = a.grad_stride;
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

// destination of run j's summed row: ugrad[j], or (po.n != 0: the sharded backward's fused gradient
// exchange) row j of an owner-bucketed list whose part r lives in rank r's peer window.
__device__ __forceinline__ float* run_dst(float* ugrad, int n_parts, const int64_t* __restrict__ s_start,
                                          char* const* __restrict__ s_base, int64_t j, int D) {
  if (n_parts == 0) return ugrad + (size_t)j * D;
  const int r = peer_part(s_start, n_parts, j);
  return reinterpret_cast<float*>(s_base[r]) + (j - s_start[r]) * D;
}


// ==========================================================================================
// Gradient reduction over the sorted occurrence list (replaces the per-run-length-class kernels of round 1).
//
// The sorted list is cut into PIECES of 32 consecutive sorted occurrences; one lane group (G lanes, one 16-byte
// vector of the row per lane) owns one piece, a warp owns 32/G consecutive pieces.  Work is defined in UNITS:
//   * a SHORT run (<= kShortRun occurrences of one FID) is one unit, summed sequentially in occurrence
//     order = the CPU reference's order (bit-exact);
//   * a LONG run (> kShortRun) is cut into 32-occurrence BLOCKS aligned to the START OF THE RUN; each block is
//     one unit, summed sequentially; block sums are then combined by a fixed 32-ary tree over the block index
//     (tree_level_kernel), so the association depends on the run length only: run-to-run bit-stable whatever
//     scratch-set slot the FID got, and pairwise-accurate (checked against fp64 in tests/test_gpu_parity.py).
// A group owns the units that START inside its piece and follows its last unit past the end of the piece (at most
// 63 occurrences); occurrences in front of the first unit start belong to the previous group's last unit.  Every
// group therefore streams 32 + <64 gradient rows whatever the run-length distribution is: hot FIDs of a Zipf
// batch (half of the occurrences) are plain 32-row units like everything else — no run-length classes, no
// work lists, no atomics, no device-side scheduling.
//
// Where results go:
//   SHORT unit  MODE_APPLY: the optimizer is applied right here from registers (row index of the run parked by
//               the claim kernel; w / optimizer-state rows prefetched together with the gradient rows): the
//               per-unique gradient buffer is never written.  MODE_STORE: summed row -> run_dst (local ugrad,
//               or the owner's peer window for the sharded backward).
//   LONG block  -> part[slot], slot = 2 * piece + (block 0 ? 1 : 0): at most one block of the run that covers the
//               piece start and at most one long-run head can START in a 32-occurrence piece, and block k + i of a
//               run starts exactly i pieces after block k.  The tree levels and the final pass therefore find
//               every operand by arithmetic on the slot index: meta[slot] = {run, block, #blocks, valid}.
// ==========================================================================================
constexpr int kPiece = 32;
constexpr int kTreeFan = 128;   // members summed by one tree node (in block order)
constexpr int kFinishMax = 512; // blocks / nodes the final pass may still have to add up per run

struct SegMeta {  // one per partial slot
  uint32_t run, blk, nblk, valid;
};

struct SegArgs {
  BwdArgs b;                        // gradient source, table, run list (run_start / rowidx / run_first_pos)
  const uint32_t* piece_run_base;   // [ceil(M/32)]: number of run heads in front of piece p
  float* part;                      // [2 * n_pieces][D]
  SegMeta* meta;                    // [2 * n_pieces]
  int64_t n_pieces;
};

template <int EPL>
__device__ __forceinline__ uint32_t pick(const uint32_t (&v)[EPL], int comp) {
  uint32_t r = v[0];
#pragma unroll
  for (int c = 1; c < EPL; ++c)
    if (comp == c) r = v[c];
  return r;
}

// element `e` (0..31) of a 32-element chunk held EPL-per-lane by the G lanes of the group
template <int G>
__device__ __forceinline__ uint32_t chunk_elem(const uint32_t (&v)[32 / G], int e) {
  constexpr int EPL = 32 / G;
  const uint32_t mine = pick<EPL>(v, e % EPL);  // e is group-uniform: every lane selects the same component
  return __shfl_sync(Group<G>::mask(), mine, Group<G>::base() + e / EPL);
}

// 32 consecutive sorted items {key, position}, EPL per lane (lane gl holds items gl * EPL ..): vector loads
template <int G>
__device__ __forceinline__ void load_items(const uint2* __restrict__ src, int64_t first, int64_t n,
                                           uint32_t (&key)[32 / G], uint32_t (&pos)[32 / G]) {
  constexpr int EPL = 32 / G;
  const int64_t i0 = first + Group<G>::gl() * EPL;
  if (EPL >= 2 && i0 + EPL <= n) {
#pragma unroll
    for (int q = 0; q < EPL; q += 2) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + i0 + q);
      key[q] = v.x; pos[q] = v.y;
      key[q + (EPL >= 2 ? 1 : 0)] = v.z; pos[q + (EPL >= 2 ? 1 : 0)] = v.w;
    }
  } else {
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
      uint2 v = make_uint2(0u, 0u);
      if (i0 + q < n) v = src[i0 + q];
      key[q] = v.x;
      pos[q] = v.y;
    }
  }
}

enum { MODE_STORE = 0, MODE_APPLY = 1 };

// positions of the 4 consecutive elements e .. e+3 (e a multiple of 4, < 32) of a chunk held EPL-per-lane
template <int G>
__device__ __forceinline__ void fetch4(const uint32_t (&v)[32 / G], int e, uint32_t (&out)[4]) {
  constexpr int EPL = 32 / G;
  const uint32_t gmask = Group<G>::mask();
  const int gb = Group<G>::base();
  if (EPL == 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) out[u] = __shfl_sync(gmask, v[u % EPL], gb + e / 4);
  } else if (EPL == 8) {
    const bool hi = (e & 4) != 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) out[u] = __shfl_sync(gmask, hi ? v[(4 + u) % EPL] : v[u % EPL], gb + e / 8);
  } else if (EPL == 2) {
#pragma unroll
    for (int u = 0; u < 4; ++u) out[u] = __shfl_sync(gmask, v[u % EPL], gb + e / 2 + u / 2);
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) out[u] = __shfl_sync(gmask, v[0], gb + e + u);
  }
}

// The reduction kernel (MODE_STORE): unit sums go to run_dst (short runs) / part (long-run blocks).
// Memory-level parallelism is what bounds it (a dependent random HBM round trip costs ~2.5 us under load, so ~100 KB
// must be in flight per SM): the gradient rows of a group's stream are requested kStage = 16 at a time with cp.async
// (LDGSTS) into shared memory — every lane copies ITS 16-byte column slice of each row into a private slot, so nothing
// but the thread's own cp.async.wait orders the data and no register holds a row in flight — then the adds run in
// occurrence order out of shared memory.  3 blocks x 64 KB = 192 KB of gradient rows in flight per SM.

// G = lanes per group, VPL = 16-byte vectors of the row per lane (lane gl owns columns (gl + v * G) * 4 ..): VPL = 2 halves
// the warp instructions per occurrence (the kernel issues ~50 of them per occurrence at VPL = 1 and was issue-bound,
// profiles/r2_ncu_kernels_summary.txt) at the same bytes in flight (kStage / VPL rows of VPL vectors per thread).
// PLAIN = SUM pooling over identity occurrence rows (no CSR indirection, no MEAN divisor): compiled out.
template <int G, int VPL, bool PLAIN, bool AHEAD>
__global__ void __launch_bounds__(kThreads, 3)
seg_reduce_kernel(SegArgs sa, const PeerOut po) {
  __shared__ int64_t s_start[kMaxPeers + 1];
  __shared__ char* s_base[kMaxPeers];
  extern __shared__ float4 stage_raw[];        // [kStage][kThreads]: [slot][thread], conflict-free 16-byte accesses
  float4 (*stage)[kThreads] = reinterpret_cast<float4 (*)[kThreads]>(stage_raw);
  peer_starts(po, s_start, s_base);
  const int n_parts = po.n;
  constexpr int EPL = 32 / G;          // elements of a piece per lane
  constexpr int RPI = 32 / G;          // pieces per warp iteration
  constexpr int RS = kStage / VPL;     // gradient rows requested per stage
  const BwdArgs& a = sa.b;
  const int gl = Group<G>::gl(), grp = (threadIdx.x & 31) / G;
  const uint32_t gmask = Group<G>::mask();
  const int c = gl * 4;
  const int D = a.td.dim;
  bool in[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) in[v] = c + v * G * 4 < D;
  const int64_t M = a.n;
  const GradSrc gs = make_grad_src(a, c);
  const uint32_t* __restrict__ run_start = a.run_start;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * RPI;
  for (int64_t p0 = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * RPI; p0 < sa.n_pieces; p0 += wstride) {
    const int64_t p = p0 + grp;
    if (p >= sa.n_pieces) continue;  // whole groups leave together: the shuffles below use the group mask only
    const int64_t base = p * kPiece;
    // ---- lane-parallel loads: the piece's sort keys and positions, the run index in front of the piece ----
    uint32_t pm0[EPL], sk[EPL];
    load_items<G>(a.sorted, base, M, sk, pm0);
    // AHEAD: the positions of the next two pieces too (a group may stream up to 63 occurrences past its piece), so the
    // continuation never waits for a dependent position load in front of its gradient-row request
    uint32_t pm1[AHEAD ? EPL : 1], pm2[AHEAD ? EPL : 1];
    if (AHEAD) {
      uint32_t skx[EPL];
      load_items<G>(a.sorted, base + kPiece, M, skx, reinterpret_cast<uint32_t (&)[EPL]>(pm1));
      load_items<G>(a.sorted, base + 2 * kPiece, M, skx, reinterpret_cast<uint32_t (&)[EPL]>(pm2));
    }
    const uint32_t j0 = sa.piece_run_base[p];
    uint32_t sk_prev = __shfl_up_sync(gmask, sk[EPL - 1], 1, G);
    if (gl == 0) sk_prev = base > 0 ? a.sorted[base - 1].x : ~sk[0];
    uint32_t hb = 0;
#pragma unroll
    for (int q = 0; q < EPL; ++q) {
      const int64_t i = base + gl * EPL + q;
      const uint32_t before = q == 0 ? sk_prev : sk[q - 1];
      if (i < M && (i == 0 || sk[q] != before)) hb |= 1u << q;
    }
    const uint32_t H = __reduce_or_sync(gmask, hb << (gl * EPL));  // run heads inside the piece
    const int nheads = __popc(H);
    const int nvalid = (int)min((int64_t)kPiece, M - base);
    // run boundaries around the piece (three independent loads)
    const uint32_t rs_prev = j0 > 0 ? run_start[j0 - 1] : 0u;       // start of the run covering the piece start
    const uint32_t rs_0 = run_start[j0];                            // first head at / after the piece start
    const uint32_t rs_end = run_start[j0 + nheads];                 // end of the last run that starts in the piece
    // ---- units of this piece ----
    const int e_first = H ? __ffs(H) - 1 : nvalid;                  // leading occurrences belong to run j0 - 1
    uint32_t S = H;                                                  // unit starts
    int lead_o = -1;                                                 // start of the leading long-run block, if any
    uint32_t lead_k = 0, lead_nb = 0;
    if (e_first > 0 && rs_0 - rs_prev > (uint32_t)kShortRun) {
      const uint32_t into = (uint32_t)base - rs_prev;                // occurrences of that run in front of the piece
      const int o = (int)((kPiece - (into & (kPiece - 1))) & (kPiece - 1));
      if (o < e_first) {
        lead_o = o;
        lead_k = (into + o) / kPiece;
        lead_nb = (rs_0 - rs_prev + kPiece - 1) / kPiece;
        S |= 1u << o;
      }
    }
    const int e_last = nheads ? 31 - __clz(H) : -1;
    const uint32_t len_last = nheads ? rs_end - ((uint32_t)base + e_last) : 0u;
    const bool last_long = len_last > (uint32_t)kShortRun;
    int e_stop;  // end of the stream (exclusive), relative to the piece start; may exceed 32
    if (nheads) e_stop = last_long ? e_last + kPiece : (int)(rs_end - (uint32_t)base);
    else e_stop = lead_o >= 0 ? min(lead_o + kPiece, (int)(rs_0 - (uint32_t)base)) : 0;
    // slot bookkeeping: both slots of the piece are (in)validated by their owner, every call
    if (gl == 0) {
      SegMeta m0, m1;
      m0.run = j0 - 1; m0.blk = lead_k; m0.nblk = lead_nb; m0.valid = lead_o >= 0 ? 1u : 0u;
      m1.run = j0 + nheads - 1; m1.blk = 0; m1.nblk = (len_last + kPiece - 1) / kPiece; m1.valid = last_long ? 1u : 0u;
      *reinterpret_cast<uint4*>(sa.meta + 2 * p) = *reinterpret_cast<uint4*>(&m0);
      *reinterpret_cast<uint4*>(sa.meta + 2 * p + 1) = *reinterpret_cast<uint4*>(&m1);
    }
    if (S == 0) continue;
    const int e_begin = __ffs(S) - 1;

    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e0 = e_begin & ~(RS - 1); e0 < e_stop; e0 += RS) {
      // ---- request: up to RS gradient rows, this lane's 16-byte slices of each ----
      // (the 4-element bodies are NOT unrolled 4x more: the kernel was instruction-fetch bound when they were)
#pragma unroll 1
      for (int b4 = 0; b4 < RS; b4 += 4) {
        const int eb = e0 + b4;
        if (eb >= e_stop) break;
        uint32_t m[4];
        if (eb < kPiece) {
          fetch4<G>(pm0, eb, m);
        } else if (AHEAD) {
          if (eb < 2 * kPiece) fetch4<G>(reinterpret_cast<const uint32_t (&)[EPL]>(pm1), eb - kPiece, m);
          else fetch4<G>(reinterpret_cast<const uint32_t (&)[EPL]>(pm2), eb - 2 * kPiece, m);
        } else {  // continuation past the piece: the positions straight from memory (same address in every lane)
#pragma unroll
          for (int u = 0; u < 4; ++u) m[u] = base + eb + u < M ? a.sorted[base + eb + u].y : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = eb + u;
          if (e >= e_begin && e < e_stop) {
            const uint32_t r = (!PLAIN && gs.occ_row) ? gs.occ_row[m[u]] : m[u];
            const float* src = gs.base + (size_t)r * gs.stride;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
              if (in[v]) cp_async16(&stage[(b4 + u) * VPL + v][threadIdx.x], src + v * G * 4);
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      // ---- consume in occurrence order ----
#pragma unroll 1
      for (int b4 = 0; b4 < RS; b4 += 4) {
        const int eb = e0 + b4;
        if (eb >= e_stop) break;
        // unit starts at eb + u, and at eb + u + 1 (=> eb + u ends a unit); no starts past the piece
        const uint32_t st4 = eb < kPiece ? (S >> eb) & 0xFu : 0u;
        const uint32_t nx4 = eb + 1 < kPiece ? (S >> (eb + 1)) & 0xFu : 0u;
        float fn[4] = {1.f, 1.f, 1.f, 1.f};
        if (!PLAIN && gs.mean) {  // MEAN pooling: the divisor of each row (positions fetched again: rare path)
          uint32_t m[4];
          if (eb < kPiece) {
            fetch4<G>(pm0, eb, m);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) m[u] = base + eb + u < M ? a.sorted[base + eb + u].y : 0u;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t r = gs.occ_row ? gs.occ_row[m[u]] : m[u];
            fn[u] = (float)(gs.row_offsets[r + 1] - gs.row_offsets[r]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = eb + u;
          if (e < e_begin || e >= e_stop) continue;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (in[v]) x = stage[(b4 + u) * VPL + v][threadIdx.x];
            if (!PLAIN && gs.mean) {
              x.x = __fdiv_rn(x.x, fn[u]); x.y = __fdiv_rn(x.y, fn[u]); x.z = __fdiv_rn(x.z, fn[u]); x.w = __fdiv_rn(x.w, fn[u]);
            }
            if ((st4 >> u) & 1u) acc[v] = x; else add4(acc[v], x);
          }
          if (!(e + 1 == e_stop || ((nx4 >> u) & 1u))) continue;
          // ---- a unit ends here ----
          const int hr = e < kPiece ? __popc(H & (0xFFFFFFFFu >> (31 - e))) : nheads;  // heads at or before e
          float* dst;
          if (hr == 0) dst = sa.part + (size_t)(2 * p) * D;                          // leading block of a long run
          else if (last_long && hr == nheads) dst = sa.part + (size_t)(2 * p + 1) * D;  // block 0 of a long run
          else dst = run_dst(a.ugrad, n_parts, s_start, s_base, (int64_t)j0 + hr - 1, D);  // a short run
#pragma unroll
          for (int v = 0; v < VPL; ++v)
            if (in[v]) *reinterpret_cast<float4*>(dst + c + v * G * 4) = acc[v];
        }
      }
    }
  }
}

// One level of the fixed tree over the block index of the long runs: the slot of block k, k a multiple of
// kTreeFan * stride, receives (in place) the sum of the slots of blocks k, k + stride, ..., k + (kTreeFan - 1) * stride.
// A WARP owns a node: its 32 / G lane groups each sum a contiguous quarter (kTreeFan * G / 32 members, in block
// order, 16 requests in flight per lane), then the group sums are added in group order through shuffles — the
// association is a function of (run length, row width) only, so results are run-to-run bit-stable.  (One lane group
// per node, 8 dependent stages of 16 loads, took 28 us on the C2 batch: the hottest FIDs' chains were the kernel.)
template <int G>
__global__ void __launch_bounds__(kThreads)
tree_level_kernel(float* __restrict__ part, const SegMeta* __restrict__ meta, int64_t n_slots, uint32_t stride, int D) {
  extern __shared__ float4 stage_raw[];  // [kStage][kThreads]
  float4 (*stage)[kThreads] = reinterpret_cast<float4 (*)[kThreads]>(stage_raw);
  constexpr int NG = 32 / G;                 // lane groups per warp = slots scanned per warp iteration
  constexpr int PER = kTreeFan / NG;         // members per lane group
  static_assert(PER % 16 == 0 || PER == kTreeFan, "members per group must fill whole stages");
  const int gl = Group<G>::gl(), c = gl * 4, grp = (threadIdx.x & 31) / G;
  const bool in = c < D;
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * (kThreads / 32) * 32;
  for (int64_t q0 = ((int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 32; q0 < n_slots; q0 += wstride) {
    // every lane inspects one slot; node heads are then processed one at a time by the whole warp
    SegMeta mine;
    mine.run = mine.blk = mine.nblk = mine.valid = 0;
    if (q0 + lane < n_slots) mine = meta[q0 + lane];
    const bool head = mine.valid && mine.blk % (kTreeFan * stride) == 0 && (uint64_t)mine.blk + stride < mine.nblk;
    uint32_t heads = __ballot_sync(0xFFFFFFFFu, head);
    while (heads) {
      const int src = __ffs(heads) - 1;       // the lane that saw this head
      heads &= heads - 1;
      const uint32_t blk = __shfl_sync(0xFFFFFFFFu, mine.blk, src);
      const uint32_t nblk = __shfl_sync(0xFFFFFFFFu, mine.nblk, src);
      const int64_t q = q0 + src;
      const int64_t piece = q >> 1;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // member i of the node lives in slot q (i = 0) or 2 * (piece + i * stride): block k + i*stride starts i*stride pieces on
      for (int i0 = grp * PER; i0 < (grp + 1) * PER; i0 += 16) {
        if ((uint64_t)blk + (uint64_t)i0 * stride >= nblk) break;
#pragma unroll
        for (int u = 0; u < 16; ++u) {  // 16 UNCONDITIONAL requests (clamped to the node's own slot): all in flight together
          const int i = i0 + u;
          const bool ok = i > 0 && (uint64_t)blk + (uint64_t)i * stride < nblk;
          const size_t slot = ok ? (size_t)(2 * (piece + (int64_t)i * stride)) : (size_t)q;
          cp_async16(&stage[u][threadIdx.x], part + slot * D + (in ? c : 0));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int i = i0 + u;
          if ((uint64_t)blk + (uint64_t)i * stride < nblk) add4(acc, stage[u][threadIdx.x]);
        }
      }
      // group sums -> group 0, in group order
      float4 tot = acc;
#pragma unroll
      for (int g = 1; g < NG; ++g) {
        float4 o;
        o.x = __shfl_sync(0xFFFFFFFFu, acc.x, g * G + gl);
        o.y = __shfl_sync(0xFFFFFFFFu, acc.y, g * G + gl);
        o.z = __shfl_sync(0xFFFFFFFFu, acc.z, g * G + gl);
        o.w = __shfl_sync(0xFFFFFFFFu, acc.w, g * G + gl);
        if ((uint64_t)blk + (uint64_t)(g * PER) * stride < nblk) add4(tot, o);
      }
      __syncwarp();  // every group has read its members (slot q among them) before the node is overwritten
      if (grp == 0 && in) *reinterpret_cast<float4*>(part + (size_t)q * D + c) = tot;
    }
  }
}

// Long runs, last step: the head slot of a long run combines what the tree left (blocks 0, top, 2 * top, ... — at most
// 32 of them, in order) and applies the optimizer (MODE_APPLY) or stores the row (MODE_STORE).
template <int G, int MODE, int OPT>
__global__ void __launch_bounds__(kThreads)
long_finish_kernel(SegArgs sa, uint32_t top, const PeerOut po) {
  __shared__ int64_t s_start[kMaxPeers + 1];
  __shared__ char* s_base[kMaxPeers];
  extern __shared__ float4 stage_raw[];  // [kStage][kThreads]
  float4 (*stage)[kThreads] = reinterpret_cast<float4 (*)[kThreads]>(stage_raw);
  peer_starts(po, s_start, s_base);
  const BwdArgs& a = sa.b;
  const int gl = Group<G>::gl(), c = gl * 4;
  const int D = a.td.dim;
  const int64_t gstride = (int64_t)gridDim.x * (kThreads / G);
  for (int64_t p = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G; p < sa.n_pieces; p += gstride) {
    const SegMeta m = sa.meta[2 * p + 1];
    if (!m.valid) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) acc = *reinterpret_cast<const float4*>(sa.part + (size_t)(2 * p + 1) * D + c);
    for (uint32_t k0 = top; k0 < m.nblk; k0 += 16 * top) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {  // 16 requests in flight (clamped to the head slot)
        const uint64_t k = (uint64_t)k0 + (uint64_t)u * top;
        const size_t slot = k < m.nblk ? (size_t)(2 * (p + (int64_t)k)) : (size_t)(2 * p + 1);
        cp_async16(&stage[u][threadIdx.x], sa.part + slot * D + (c < D ? c : 0));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if ((uint64_t)k0 + (uint64_t)u * top < m.nblk) add4(acc, stage[u][threadIdx.x]);
    }
    const uint32_t j = m.run;
    if (MODE == MODE_APPLY) {
      const uint32_t ri = a.rowidx[j];
      if (ri == kEmptyRow) continue;
      const RowPre pre = bwd_prefetch<G, OPT>(a, ri, c);
      bwd_apply<G, OPT>(a, j, ri, acc, c, pre);
    } else {
      if (c < D) *reinterpret_cast<float4*>(run_dst(a.ugrad, po.n, s_start, s_base, j, D) + c) = acc;
    }
  }
}

// Apply: group per run; rowidx[j], ugrad[j] and the row's w / state are all independent loads.
// (A variant that handled two runs per group at a time was measured slower: 62 registers, 43 % occupancy.)
// prefetch != 0: the NEXT run's row index is read one iteration ahead and its weight / state rows are pulled into L2
// meanwhile, which takes the random row fetch (~2.5 us from HBM) off the per-iteration chain.
template <int G, int OPT>
__global__ void __launch_bounds__(kThreads) runs_apply_kernel(BwdArgs a, int prefetch) {
  const int gl = Group<G>::gl();
  const int c = gl * 4;
  const int64_t nr = *a.n_runs;
  const int D = a.td.dim;
  const int64_t gstride = (int64_t)gridDim.x * (kThreads / G);
  int64_t j = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G;
  uint32_t ri_next = j < nr ? a.rowidx[j] : kEmptyRow;
  for (; j < nr; j += gstride) {
    const uint32_t ri = ri_next;
    if (j + gstride < nr) {
      ri_next = a.rowidx[j + gstride];
      if (prefetch && OPT >= 0 && ri_next != kEmptyRow && !(ri_next & kFreshBit) && c < D) {
        const uint32_t row = ri_next & ~kFreshBit;
        prefetch_l2(a.td.emb + (size_t)row * a.td.emb_stride + c);
        if (OPT != MONO_OPT_SGD) {
          const float* s_row = a.td.state + (size_t)row * a.td.state_stride;
          prefetch_l2(s_row + c);
          if (OPT == MONO_OPT_FTRL || OPT == MONO_OPT_ADAM) prefetch_l2(s_row + D + c);
        }
      }
    }
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) g4 = __ldcs(reinterpret_cast<const float4*>(a.ugrad + (size_t)j * D + c));
    if (ri == kEmptyRow) continue;
    const RowPre pre = bwd_prefetch<G, OPT>(a, ri, c);
    bwd_apply<G, OPT>(a, (uint32_t)j, ri, g4, c, pre);
  }
}

// run j of the sorted offsets -> destination row: out_rows + (sorted key << shift)
template <int G>
__global__ void __launch_bounds__(kThreads) runs_emit_kernel(BwdArgs a, int shift, float* __restrict__ out_rows) {
  const int gl = Group<G>::gl(), c = gl * 4;
  const int64_t nr = *a.n_runs;
  const int D = a.td.dim;
  const int64_t gstride = (int64_t)gridDim.x * (kThreads / G);
  for (int64_t j = (int64_t)blockIdx.x * (kThreads / G) + threadIdx.x / G; j < nr; j += gstride) {
    if (c >= D) continue;
    const float4 g = __ldcs(reinterpret_cast<const float4*>(a.ugrad + (size_t)j * D + c));
    *reinterpret_cast<float4*>(out_rows + ((size_t)a.sorted[a.run_start[j]].x << shift) + c) = g;
  }
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

template <bool RESOLVE, bool DUAL = false>
__global__ void __launch_bounds__(kThreads)
fid_claim_kernel(const int64_t* __restrict__ fids, int64_t n, Entry* set, uint32_t R, int N, uint32_t epoch,
                 uint32_t* __restrict__ slot_of, uint32_t* __restrict__ owner_cnt /* [N], [256] = overflow */,
                 ClaimResolve cr, int prefetch) {
  __shared__ uint32_t cnt[256];
  for (int d = threadIdx.x; d < 256; d += blockDim.x) cnt[d] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  // the chain per occurrence is FID -> set entry (-> CAS -> table bucket for a winner), ~2.5 us per dependent round
  // trip: the NEXT occurrence's FID is requested one iteration ahead, which takes it off the chain
  // ... and so are (prefetch != 0) its scratch-set slot and, for RESOLVE, its first table bucket: both are pulled into
  // L2 while this iteration's chain runs, so the next iteration's two dependent reads are L2 hits.
  int64_t key_next = i < n ? __ldg(fids + i) : 0;
  const Entry* pf_buckets = nullptr;
  uint32_t pf_nb = 0;
  if (RESOLVE && prefetch) {
    pf_buckets = cr.t->buckets;
    pf_nb = cr.t->num_buckets;
  }
  for (; i < n; i += stride) {
    const int64_t key = key_next;
    if (i + stride < n) {
      key_next = __ldg(fids + i + stride);
      if (prefetch) {
        const uint32_t o_n = N == 1 ? 0u : (uint32_t)((uint64_t)key_next % (uint64_t)N);
        prefetch_l2(set + o_n * R + __umulhi((uint32_t)(mix64((uint64_t)key_next) >> 24), R));
        if (RESOLVE) {
          uint32_t b1, b2;
          bucket_pair(key_next, pf_nb, b1, b2);
          prefetch_l2(pf_buckets + (size_t)b1 * kBucketSlots);
        }
      }
    }
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
        const uint32_t row = probe_lane_slot<DUAL>(cr.t, key, &slot);
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
    // admission filter: the backward hands every distinct FID once per step (count 1), as the reference's
    // deduplicated optimize does; a FID below its slot's threshold gets no row (its run is skipped by the apply)
    if (should_be_filtered(t, ld_entry_cg(set + sl).key, 1u)) continue;
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


// ==========================================================================================
// host side
// ==========================================================================================
static const PeerOut no_peer = {};  // n == 0: MODE_STORE writes the local ugrad buffer

struct ReduceScratch {  // sizes of the reduction scratch for M occurrences of dim D
  int64_t n_pieces;
  size_t part_bytes, meta_bytes, prb_bytes;
  ReduceScratch(int64_t M, int D) {
    n_pieces = (M + kPiece - 1) / kPiece;
    part_bytes = sizeof(float) * 2 * (size_t)n_pieces * D;
    meta_bytes = sizeof(SegMeta) * 2 * (size_t)n_pieces;
    prb_bytes = 4 * (size_t)n_pieces;
  }
};

constexpr size_t kStageBytes = sizeof(float4) * kStage * kThreads;
template <int GL, int VPL, bool PLAIN, bool AHEAD>
static void launch_seg2(const SegArgs& sa, const PeerOut& po, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    MONO_CUDA(cudaFuncSetAttribute(seg_reduce_kernel<GL, VPL, PLAIN, AHEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)kStageBytes));
    attr_set = true;
  }
  seg_reduce_kernel<GL, VPL, PLAIN, AHEAD>
      <<<resident_grid(seg_reduce_kernel<GL, VPL, PLAIN, AHEAD>, sa.n_pieces, (kThreads / 32) * (32 / GL), kThreads,
                       kStageBytes),
         kThreads, kStageBytes, s>>>(sa, po);
  MONO_CHECK_LAUNCH();
}
template <int GL, int VPL, bool PLAIN>
static void launch_seg(const SegArgs& sa, const PeerOut& po, cudaStream_t s) {
  if (knob(KNOB_SEG_AHEAD)) launch_seg2<GL, VPL, PLAIN, true>(sa, po, s);
  else launch_seg2<GL, VPL, PLAIN, false>(sa, po, s);
}

// per-run gradient sums -> run_dst (a.ugrad, or the owners' peer windows)
static void launch_reduce(const SegArgs& sa, int G, int /*unused*/, const PeerOut& po, cudaStream_t s) {
  const int64_t np = sa.n_pieces;
  if (np <= 0) return;
  int levels = 0;
  uint64_t top = 1;
  while ((uint64_t)np > top * kFinishMax) {  // the final pass adds up to kFinishMax nodes per run, 16 loads in flight
    top *= kTreeFan;
    ++levels;
  }
  const int D = sa.b.td.dim;
  const bool plain = sa.b.occ_row == nullptr && !(sa.b.pooling == MONO_POOL_MEAN && sa.b.row_offsets != nullptr);
  const bool two = knob(KNOB_SEG_VPL) == 2;
#define SEG_GO(GG, MODE, OO)                                                                                        \
  do {                                                                                                              \
    static bool attr_set = false;                                                                                   \
    if (!attr_set) {                                                                                                \
      MONO_CUDA(cudaFuncSetAttribute(tree_level_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                     (int)kStageBytes));                                                            \
      MONO_CUDA(cudaFuncSetAttribute(long_finish_kernel<GG, MODE, OO>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)kStageBytes));                                                            \
      attr_set = true;                                                                                              \
    }                                                                                                               \
    if (two && GG >= 8) {                                                                                           \
      if (plain) launch_seg<(GG >= 8 ? GG / 2 : GG), (GG >= 8 ? 2 : 1), true>(sa, po, s);                           \
      else launch_seg<(GG >= 8 ? GG / 2 : GG), (GG >= 8 ? 2 : 1), false>(sa, po, s);                                \
    } else {                                                                                                        \
      if (plain) launch_seg<GG, 1, true>(sa, po, s);                                                                \
      else launch_seg<GG, 1, false>(sa, po, s);                                                                     \
    }                                                                                                               \
    uint32_t stride = 1;                                                                                            \
    for (int l = 0; l < levels; ++l, stride *= kTreeFan) {                                                          \
      tree_level_kernel<GG><<<resident_grid(tree_level_kernel<GG>, 2 * np, kThreads, kThreads, kStageBytes),        \
                              kThreads, kStageBytes, s>>>(sa.part, sa.meta, 2 * np, stride, D);                     \
      MONO_CHECK_LAUNCH();                                                                                          \
    }                                                                                                               \
    long_finish_kernel<GG, MODE, OO>                                                                                \
        <<<resident_grid(long_finish_kernel<GG, MODE, OO>, np, kThreads / GG, kThreads, kStageBytes), kThreads,     \
           kStageBytes, s>>>(sa, (uint32_t)top, po);                                                                \
    MONO_CHECK_LAUNCH();                                                                                            \
  } while (0)
#define SEG_G(GG) SEG_GO(GG, MODE_STORE, 0)
  switch (G) {
    case 4: SEG_G(4); break;
    case 8: SEG_G(8); break;
    case 16: SEG_G(16); break;
    default: SEG_G(32); break;
  }
#undef SEG_G
#undef SEG_GO
}

static int bits_for(uint64_t cap) {
  int bits = 0;
  while (((uint64_t)1 << bits) < cap) ++bits;
  return bits;
}

void run_pool_backward(mono_mtable* mt, int k, const int64_t* fids_dev, int64_t n_fids,
                       const int32_t* row_offsets, int64_t n_rows, int pooling,
                       const float* pooled_grad, int64_t grad_stride, int grad_col,
                       const float* lr_host, int64_t update_time, cudaStream_t s) {
  mt->note_insert(s);
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
  // the register-resident apply exists for the four vectorised optimizers; anything else: store + generic apply_row
  const int opt_sel = (ht.segs.size() == 1 && ht.segs[0].opt_type <= MONO_OPT_ADAM) ? ht.segs[0].opt_type : -1;

  // ---- scratch layout ----
  uint32_t cap = 1024;
  while ((uint64_t)cap < 2 * (uint64_t)M) cap <<= 1;
  const int bits = bits_for(cap);
  const int passes = SortWs::passes(bits);
  const ReduceScratch rs(M, D);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_k0 = take(4 * (size_t)M), o_a = take(8 * (size_t)M), o_b = take(8 * (size_t)M);
  const size_t o_cnt = take(4 * (size_t)kRadixBins * SortWs::radix_blocks(M));
  const size_t o_ctr = take(4096 + 4 * (size_t)kRadixBins * passes);  // counters + digit totals per pass
  const size_t o_brun = take(4 * (size_t)SortWs::run_blocks(M));
  const size_t o_rs = take(4 * ((size_t)M + 1)), o_rfp = take(4 * (size_t)M), o_ridx = take(4 * (size_t)M);
  const size_t o_miss = take(4 * (size_t)M);
  const size_t o_occ = take(row_offsets ? 4 * (size_t)M : 0);
  const size_t o_prb = take(rs.prb_bytes), o_part = take(rs.part_bytes), o_meta = take(rs.meta_bytes);
  const size_t o_ug = take(sizeof(float) * (size_t)M * D);
  char* ws = (char*)mt->ws_a.get(off, s);
  uint32_t epoch = 0;
  Entry* set = (Entry*)mt->claim_set.get(sizeof(Entry) * (size_t)cap, s, &epoch);
  uint32_t* ctr = (uint32_t*)(ws + o_ctr);  // [0] n_runs [8] miss_ctr [512..] owner counts (unused here)
  MONO_CUDA(cudaMemsetAsync(ctr, 0, 4096, s));

  // 1 claim: k0[i] = set slot of occurrence i; the winner of a slot resolves its FID in the table (row parked in
  //   the set entry, expiry timestamp bumped) or queues it as absent
  uint32_t* k0 = (uint32_t*)(ws + o_k0);
  ClaimResolve cr;
  cr.t = mt->d_tables + k;
  cr.update_ts = (uint32_t)update_time;
  cr.miss_ctr = ctr + 8;
  cr.miss_slots = (uint32_t*)(ws + o_miss);
  if (knob(KNOB_CLAIM_DUAL))
    fid_claim_kernel<true, true><<<resident_grid(fid_claim_kernel<true, true>, M, kThreads), kThreads, 0, s>>>(
        fids_dev, M, set, cap, 1, epoch, k0, ctr + 512, cr, knob(KNOB_CLAIM_PF));
  else
    fid_claim_kernel<true><<<resident_grid(fid_claim_kernel<true>, M, kThreads), kThreads, 0, s>>>(
        fids_dev, M, set, cap, 1, epoch, k0, ctr + 512, cr, knob(KNOB_CLAIM_PF));
  MONO_CHECK_LAUNCH();
  // 2 absent FIDs: allocate a row + lock-free insert (few in steady state; the count stays on the device)
  claim_miss_kernel<<<resident_grid(claim_miss_kernel, std::min<int64_t>(M, 148 * 2 * kThreads), kThreads), kThreads, 0, s>>>(
      mt->d_tables + k, set, ctr + 8, cr.miss_slots, (uint32_t)update_time);
  MONO_CHECK_LAUNCH();
  launch_upsert_finalize(mt, cb, ctr + 8, (uint32_t)update_time, s);
  // 3 stable LSD radix sort of (slot, position)  +  4 ordered run list (with each run's resolved row)
  SortWs sw;
  sw.k0 = k0;
  sw.a = (uint2*)(ws + o_a);
  sw.b = (uint2*)(ws + o_b);
  sw.cnt = (uint32_t*)(ws + o_cnt);
  sw.dtot = (uint32_t*)(ws + o_ctr + 4096);
  sw.blk_runs = (uint32_t*)(ws + o_brun);
  sw.run_start = (uint32_t*)(ws + o_rs);
  sw.run_first_pos = (uint32_t*)(ws + o_rfp);
  sw.piece_run_base = (uint32_t*)(ws + o_prb);
  sw.n_runs = ctr;
  sw.claim_set = set;
  sw.rowidx = (uint32_t*)(ws + o_ridx);
  const uint2* sorted = sort_and_runs(sw, M, bits, 0, s);
  // 5 reduce + update
  SegArgs sa;
  BwdArgs& a = sa.b;
  a.td = ht.dev;  // descriptor is current: ensure_capacity / upload_tables ran above
  a.t = mt->d_tables + k;
  a.fids = fids_dev;
  a.sorted = sorted;
  a.n = M;
  a.n_runs = ctr;
  a.run_start = sw.run_start;
  a.run_first_pos = sw.run_first_pos;
  a.rowidx = sw.rowidx;
  a.occ_row = nullptr;
  a.row_offsets = row_offsets;
  a.pooling = pooling;
  a.pooled_grad = pooled_grad;
  a.grad_stride = grad_stride;
  a.grad_col = grad_col;
  a.lr = cb.lr;
  a.ugrad = (float*)(ws + o_ug);
  a.scratch = a.ugrad;
  sa.piece_run_base = sw.piece_run_base;
  sa.part = (float*)(ws + o_part);
  sa.meta = (SegMeta*)(ws + o_meta);
  sa.n_pieces = rs.n_pieces;
  if (row_offsets) {
    uint32_t* occ = (uint32_t*)(ws + o_occ);
    occ_row_kernel<<<resident_grid(occ_row_kernel, n_rows, kThreads), kThreads, 0, s>>>(row_offsets, n_rows, occ);
    MONO_CHECK_LAUNCH();
    a.occ_row = occ;
  }
  launch_reduce(sa, G, -1, no_peer, s);
  // 6 apply: one streaming pass over the runs — row index, summed gradient, w / optimizer state of every distinct FID
  const int pf = knob(KNOB_APPLY_PF);
#define BWD2(GG, OO)                                                                                             \
  runs_apply_kernel<GG, OO><<<resident_grid(runs_apply_kernel<GG, OO>, M, kThreads / GG), kThreads, 0, s>>>(a, pf);  \
  MONO_CHECK_LAUNCH()
#define BWD(GG)                                                         \
  switch (opt_sel) {                                                    \
    case MONO_OPT_SGD: BWD2(GG, MONO_OPT_SGD); break;                   \
    case MONO_OPT_ADAGRAD: BWD2(GG, MONO_OPT_ADAGRAD); break;           \
    case MONO_OPT_FTRL: BWD2(GG, MONO_OPT_FTRL); break;                 \
    case MONO_OPT_ADAM: BWD2(GG, MONO_OPT_ADAM); break;                 \
    default: BWD2(GG, -1); break;                                       \
  }
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
  const int passes = SortWs::passes(bits);
  const ReduceScratch rs(M, D);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_a = take(8 * (size_t)M), o_b = take(8 * (size_t)M);
  const size_t o_cnt = take(4 * (size_t)kRadixBins * SortWs::radix_blocks(M));
  const size_t o_ctr = take(4096 + 4 * (size_t)kRadixBins * passes);
  const size_t o_brun = take(4 * (size_t)SortWs::run_blocks(M));
  const size_t o_rs = take(4 * ((size_t)M + 1)), o_rfp = take(4 * (size_t)M);
  const size_t o_occ = take(row_offsets ? 4 * (size_t)M : 0);
  const size_t o_prb = take(rs.prb_bytes), o_part = take(rs.part_bytes), o_meta = take(rs.meta_bytes);
  const size_t o_ug = take(sizeof(float) * (size_t)M * D);
  char* ws = nullptr;
  MONO_CUDA(cudaMallocAsync((void**)&ws, off, s));
  uint32_t* ctr = (uint32_t*)(ws + o_ctr);
  MONO_CUDA(cudaMemsetAsync(ctr, 0, 4096, s));
  SortWs sw;
  sw.k0 = reinterpret_cast<uint32_t*>(const_cast<int32_t*>(offs_dev));  // read-only: the first pass copies the keys out
  sw.a = (uint2*)(ws + o_a);
  sw.b = (uint2*)(ws + o_b);
  sw.cnt = (uint32_t*)(ws + o_cnt);
  sw.dtot = (uint32_t*)(ws + o_ctr + 4096);
  sw.blk_runs = (uint32_t*)(ws + o_brun);
  sw.run_start = (uint32_t*)(ws + o_rs);
  sw.run_first_pos = (uint32_t*)(ws + o_rfp);
  sw.piece_run_base = (uint32_t*)(ws + o_prb);
  sw.n_runs = ctr;
  const uint2* sorted = sort_and_runs(sw, M, bits, shift, s);
  SegArgs sa;
  std::memset(&sa, 0, sizeof(sa));
  BwdArgs& a = sa.b;
  a.td.dim = D;
  a.sorted = sorted;
  a.n = M;
  a.n_runs = ctr;
  a.run_start = sw.run_start;
  a.run_first_pos = sw.run_first_pos;
  a.row_offsets = row_offsets;
  a.pooling = pooling;
  a.pooled_grad = pooled_grad;
  a.grad_stride = grad_stride;
  a.grad_col = grad_col;
  a.ugrad = (float*)(ws + o_ug);
  sa.piece_run_base = sw.piece_run_base;
  sa.part = (float*)(ws + o_part);
  sa.meta = (SegMeta*)(ws + o_meta);
  sa.n_pieces = rs.n_pieces;
  if (row_offsets) {
    uint32_t* occ = (uint32_t*)(ws + o_occ);
    occ_row_kernel<<<resident_grid(occ_row_kernel, n_rows, kThreads), kThreads, 0, s>>>(row_offsets, n_rows, occ);
    MONO_CHECK_LAUNCH();
    a.occ_row = occ;
  }
  const int G = pick_group(D);
  launch_reduce(sa, G, -1, no_peer, s);
#define EMIT(GG)                                                                                              \
  runs_emit_kernel<GG><<<resident_grid(runs_emit_kernel<GG>, M, kThreads / GG), kThreads, 0, s>>>(a, shift, out_rows); \
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
group_emit_kernel(const int64_t* __restrict__ fids, const uint2* __restrict__ sorted,
                  const uint32_t* __restrict__ run_of_sorted, const uint32_t* __restrict__ run_first_pos,
                  const uint32_t* __restrict__ n_runs, int64_t n, int dim, int32_t* __restrict__ occ_offset,
                  int64_t* __restrict__ uniq_out) {
  const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = t0; i < n; i += stride) occ_offset[sorted[i].y] = (int32_t)(run_of_sorted[i] * (uint32_t)dim);
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
  g->sorted = nullptr;
  if (M == 0) {
    if (!shard_counts_host) throw ArgError("grouping: the device-driven step needs a non-empty batch");
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
    const int bits = bits_for(cap);
    const int passes = SortWs::passes(bits);
    const ReduceScratch rs(M, dim);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_k0 = take(4 * (size_t)M), o_a = take(8 * (size_t)M), o_b = take(8 * (size_t)M);
    const size_t o_cnt = take(4 * (size_t)kRadixBins * SortWs::radix_blocks(M));
    const size_t o_ctr = take(4096 + 4 * (size_t)kRadixBins * passes + 4 * 260);
    const size_t o_brun = take(4 * (size_t)SortWs::run_blocks(M));
    const size_t o_rs = take(4 * ((size_t)M + 1)), o_rfp = take(4 * (size_t)M), o_ros = take(4 * (size_t)M);
    const size_t o_prb = take(rs.prb_bytes);
    // reduce() scratch
    const size_t o_occ = take(4 * (size_t)M), o_part = take(rs.part_bytes), o_meta = take(rs.meta_bytes);
    char* ws = (char*)g->ws.get(off, s);
    uint32_t epoch = 0;
    Entry* set = (Entry*)g->claim_set.get(sizeof(Entry) * (size_t)cap, s, &epoch);
    uint32_t* ctr = (uint32_t*)(ws + o_ctr);
    uint32_t* owner_cnt = (uint32_t*)(ws + o_ctr + 4096 + 4 * (size_t)kRadixBins * passes);
    MONO_CUDA(cudaMemsetAsync(ctr, 0, 4096, s));
    MONO_CUDA(cudaMemsetAsync(owner_cnt, 0, 4 * 260, s));
    SortWs sw;
    sw.k0 = (uint32_t*)(ws + o_k0);
    sw.a = (uint2*)(ws + o_a);
    sw.b = (uint2*)(ws + o_b);
    sw.cnt = (uint32_t*)(ws + o_cnt);
    sw.dtot = (uint32_t*)(ws + o_ctr + 4096);
    sw.blk_runs = (uint32_t*)(ws + o_brun);
    sw.run_start = (uint32_t*)(ws + o_rs);
    sw.run_first_pos = (uint32_t*)(ws + o_rfp);
    sw.piece_run_base = (uint32_t*)(ws + o_prb);
    sw.n_runs = ctr;
    sw.run_of_sorted = (uint32_t*)(ws + o_ros);
    ClaimResolve cr0;
    std::memset(&cr0, 0, sizeof(cr0));
    fid_claim_kernel<false><<<resident_grid(fid_claim_kernel<false>, M, kThreads), kThreads, 0, s>>>(
        fids_dev, M, set, R, N, epoch, sw.k0, owner_cnt, cr0, knob(KNOB_CLAIM_PF));
    MONO_CHECK_LAUNCH();
    // counts -> host on the side stream, while the sort below keeps the GPU busy.  Device-driven callers
    // (shard_counts_host == nullptr: xstep.cu) never read them on the host: nothing waits, the per-owner counts stay
    // in owner_cnt and a region overflow is left in owner_cnt[256] for the caller to check later.
    const bool host_counts = shard_counts_host != nullptr;
    if (host_counts) {
      MONO_CUDA(cudaEventRecord(g->ev_claimed, s));
      MONO_CUDA(cudaStreamWaitEvent(g->side, g->ev_claimed, 0));
      MONO_CUDA(cudaMemcpyAsync(g->h_counts, owner_cnt, 4 * 257, cudaMemcpyDeviceToHost, g->side));
      MONO_CUDA(cudaEventRecord(g->ev_copied, g->side));
    }
    const uint2* sorted = sort_and_runs(sw, M, bits, 0, s);
    group_emit_kernel<<<resident_grid(group_emit_kernel, M, kThreads), kThreads, 0, s>>>(
        fids_dev, sorted, sw.run_of_sorted, sw.run_first_pos, ctr, M, dim, occ_offset_out, uniq_out);
    MONO_CHECK_LAUNCH();
    g->owner_cnt = owner_cnt;
    if (host_counts) {
      MONO_CUDA(cudaEventSynchronize(g->ev_copied));
      if (g->h_counts[256] != 0) continue;  // region overflow: redo with full-size regions
      int64_t total = 0;
      for (int n = 0; n < N; ++n) {
        shard_counts_host[n] = (int32_t)g->h_counts[n];
        total += g->h_counts[n];
      }
      if (n_unique_host) *n_unique_host = total;
    }
    g->sorted = sorted;
    g->run_start = sw.run_start;
    g->run_first_pos = sw.run_first_pos;
    g->piece_run_base = sw.piece_run_base;
    g->ctr = ctr;
    g->occ = (uint32_t*)(ws + o_occ);
    g->part = (float*)(ws + o_part);
    g->meta = ws + o_meta;
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
  if (!g->sorted) throw ArgError("grouping_reduce before grouping_build");
  if (pooling != MONO_POOL_SUM && pooling != MONO_POOL_MEAN) throw ArgError("grouping_reduce: SUM or MEAN");
  if ((grad_stride & 3) || (grad_col & 3) || (reinterpret_cast<uintptr_t>(pooled_grad) & 15) ||
      (po.n == 0 && (reinterpret_cast<uintptr_t>(out_rows) & 15)))
    throw ArgError("grouping_reduce needs 16-byte aligned rows");
  const ReduceScratch rs(M, D);
  SegArgs sa;
  std::memset(&sa, 0, sizeof(sa));
  BwdArgs& a = sa.b;
  a.td.dim = D;
  a.sorted = g->sorted;
  a.n = M;
  a.n_runs = g->ctr;
  a.run_start = g->run_start;
  a.run_first_pos = g->run_first_pos;
  a.row_offsets = row_offsets;
  a.pooling = pooling;
  a.pooled_grad = pooled_grad;
  a.grad_stride = grad_stride;
  a.grad_col = grad_col;
  a.ugrad = out_rows;  // runs are already in the bucketed order: the sums are written in place
  sa.piece_run_base = g->piece_run_base;
  sa.part = g->part;
  sa.meta = static_cast<SegMeta*>(g->meta);
  sa.n_pieces = rs.n_pieces;
  if (row_offsets) {
    occ_row_kernel<<<resident_grid(occ_row_kernel, n_rows, kThreads), kThreads, 0, s>>>(row_offsets, n_rows, g->occ);
    MONO_CHECK_LAUNCH();
    a.occ_row = g->occ;
  }
  launch_reduce(sa, pick_group(D), -1, po, s);
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
