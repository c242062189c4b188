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
// lower lanes is the stable rank) and per-warp digit counters in shared memory; tiles are ordered by the scanned
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
    uint32_t lr[kRadixIPT];
#pragma unroll
    for (int r = 0; r < kRadixIPT; ++r) {
      const int64_t i = wbeg + r * 32 + lane;
      const bool valid = i < n;
      const uint32_t dg = valid ? (radix_load<FIRST>(in, i, pre_shift).x >> shift) & (kRadixBins - 1) : 0u;
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
      lr[r] = old + __popc(peers & lt);
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
    // ---- phase 2: scatter ----
#pragma unroll
    for (int r = 0; r < kRadixIPT; ++r) {
      const int64_t i = wbeg + r * 32 + lane;
      if (i < n) {
        const uint2 it = radix_load<FIRST>(in, i, pre_shift);
        const uint32_t dg = (it.x >> shift) & (kRadixBins - 1);
        out[bbase[dg] + wc[w][dg] + lr[r]] = it;
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

