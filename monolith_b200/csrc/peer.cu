// peer.cu — NVLink peer window: the exchange steps of the sharded path without NCCL.
//
// The reference's synchronous multi-GPU path moves FIDs, embedding rows and gradient rows between the
// requesting rank and the owning rank with all-to-alls (ref: distributed_ps_sync.py:92-118, 531-573;
// SURVEY.md §8e).  On an NVSwitch node every GPU can store straight into every other GPU's HBM, so the
// engine gives each rank one `window` (cudaMalloc + CUDA IPC mapping in every peer) and lets the
// PRODUCING kernel write its output where the consumer will read it:
//   * owner lookup  -> rows land in the requester's window    (ops.cu lookup_push_kernel)
//   * grad reduce   -> rows land in the owner's window        (ops.cu runs_permute_push_kernel)
//   * FID buckets   -> peer_put_kernel (a plain partitioned copy, 8-byte items)
// Ordering between ranks is a flag barrier: after its stores a rank publishes a sequence number in every
// peer's flag page (fence.sys + st.release.sys) and waits for all peers' numbers (ld.acquire.sys).
// Everything is stream-ordered; the host never waits for a peer.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.h"

namespace mono {

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct PeerFlags {
  uint64_t* flags[kMaxPeers];  // flags[r]: head of rank r's window, uint64 slot per source rank
  int world, rank;
};

// One warp: lane r publishes `seq` to rank r and waits for rank r's `seq`.  A peer that never arrives
// (crashed process) trips the timeout: the kernel traps, so the job fails loudly instead of hanging.
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerFlags pf, uint64_t seq, uint64_t timeout_ns) {
  const int r = threadIdx.x;
  if (r >= pf.world) return;
  __threadfence_system();  // this rank's earlier stores (previous kernels included) before the flag
  st_release_sys(pf.flags[r] + pf.rank, seq);
  const uint64_t* mine = pf.flags[pf.rank] + r;
  const uint64_t t0 = global_ns();
  while (ld_acquire_sys(mine) < seq) {
    if (timeout_ns != 0 && global_ns() - t0 > timeout_ns) {
      printf("mono_peer: rank %d timed out waiting for rank %d at barrier %llu\n", pf.rank, r,
             (unsigned long long)seq);
      __trap();
    }
  }
}

void peer_barrier(mono_peer* p, cudaStream_t s) {
  if (!p->attached) throw ArgError("peer window is not attached");
  MONO_CUDA(cudaSetDevice(p->device));
  PeerFlags pf;
  std::memset(&pf, 0, sizeof(pf));
  for (int r = 0; r < p->world; ++r) pf.flags[r] = reinterpret_cast<uint64_t*>(p->base[r]);
  pf.world = p->world;
  pf.rank = p->rank;
  ++p->seq;
  // MONO_PEER_TIMEOUT_S: seconds a rank waits for a peer before the kernel traps (default 600; 0 = wait for ever).
  // The wait covers everything a peer does between two barriers (its dense tower, a checkpoint, a data stall).
  static const uint64_t timeout_ns = [] {
    const char* e = std::getenv("MONO_PEER_TIMEOUT_S");
    const double sec = e ? std::atof(e) : 600.0;
    return sec <= 0 ? 0ull : (uint64_t)(sec * 1e9);
  }();
  peer_barrier_kernel<<<1, 32, 0, s>>>(pf, p->seq, timeout_ns);
  MONO_CHECK_LAUNCH();
}

// partitioned copy: part r of the source goes to rank r's window.  VEC = bytes per thread access.
struct PutArgs {
  PeerOut po;                    // start[] in VEC-sized items
  const char* src[kMaxPeers];    // source of part r
};
template <class V>
__global__ void __launch_bounds__(kThreads) peer_put_kernel(PutArgs a) {
  const int64_t total = a.po.start[a.po.n];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = peer_part(a.po.start, a.po.n, i);
    const int64_t q = i - a.po.start[r];
    const V v = __ldcs(reinterpret_cast<const V*>(a.src[r]) + q);
    reinterpret_cast<V*>(a.po.base[r])[q] = v;
  }
}

static void check_region(mono_peer* p, int64_t off, int64_t bytes, const char* what) {
  if (off < 0 || bytes < 0 || (uint64_t)off + (uint64_t)bytes > p->bytes) throw ArgError(what);
}

void peer_put(mono_peer* p, int64_t region_off, const int64_t* dst_off, const void* src,
              const int64_t* src_off, const int64_t* nbytes, cudaStream_t s) {
  if (!p->attached) throw ArgError("peer window is not attached");
  MONO_CUDA(cudaSetDevice(p->device));
  PutArgs a;
  std::memset(&a, 0, sizeof(a));
  bool v16 = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (region_off & 15) == 0;
  for (int r = 0; r < p->world; ++r) {
    check_region(p, region_off + dst_off[r], nbytes[r], "peer_put: destination outside the window");
    if ((dst_off[r] | src_off[r] | nbytes[r]) & 7) throw ArgError("peer_put: offsets and sizes must be multiples of 8");
    if ((dst_off[r] | src_off[r] | nbytes[r]) & 15) v16 = false;
  }
  const int vb = v16 ? 16 : 8;
  a.po.n = p->world;
  a.po.start[0] = 0;
  for (int r = 0; r < p->world; ++r) {
    a.po.base[r] = p->base[r] + kPeerFlagBytes + region_off + dst_off[r];
    a.src[r] = static_cast<const char*>(src) + src_off[r];
    a.po.start[r + 1] = a.po.start[r] + nbytes[r] / vb;
  }
  const int64_t total = a.po.start[p->world];
  if (total == 0) return;
  if (v16)
    peer_put_kernel<uint4><<<resident_grid(peer_put_kernel<uint4>, total, kThreads), kThreads, 0, s>>>(a);
  else
    peer_put_kernel<uint2><<<resident_grid(peer_put_kernel<uint2>, total, kThreads), kThreads, 0, s>>>(a);
  MONO_CHECK_LAUNCH();
}

// partitioned pull: part r is read from rank r's window into local memory (the mirror of peer_put);
// UNR independent 16-byte loads per thread before the first store.  Alternative to the fused push kernels
// (MONO_PEER_BULK=pull): measured ~500 GB/s for a 54 MB pull between two GPUs, but the extra pass makes the
// step slower than pushing from inside the producing kernel.
struct GetArgs {
  PeerOut po;                    // base[] = local destinations, start[] in 16-byte items
  const char* src[kMaxPeers];    // source of part r (in rank r's window)
};
__global__ void __launch_bounds__(kThreads) peer_get_kernel(GetArgs a) {
  constexpr int UNR = 4;
  const int64_t total = a.po.start[a.po.n];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * UNR) {
    uint4 v[UNR];
    int r[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t i = i0 + u * stride;
      r[u] = 0;
      v[u] = make_uint4(0, 0, 0, 0);
      if (i < total) {
        r[u] = peer_part(a.po.start, a.po.n, i);
        v[u] = __ldcs(reinterpret_cast<const uint4*>(a.src[r[u]]) + (i - a.po.start[r[u]]));
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) reinterpret_cast<uint4*>(a.po.base[r[u]])[i - a.po.start[r[u]]] = v[u];
    }
  }
}

void peer_get(mono_peer* p, int64_t region_off, const int64_t* src_off, void* dst, const int64_t* dst_off,
              const int64_t* nbytes, cudaStream_t s) {
  if (!p->attached) throw ArgError("peer window is not attached");
  MONO_CUDA(cudaSetDevice(p->device));
  if ((reinterpret_cast<uintptr_t>(dst) & 15) || (region_off & 15)) throw ArgError("peer_get: 16-byte alignment");
  GetArgs a;
  std::memset(&a, 0, sizeof(a));
  a.po.n = p->world;
  for (int r = 0; r < p->world; ++r) {
    check_region(p, region_off + src_off[r], nbytes[r], "peer_get: source outside the window");
    if ((dst_off[r] | src_off[r] | nbytes[r]) & 15) throw ArgError("peer_get: offsets and sizes must be multiples of 16");
    a.po.base[r] = static_cast<char*>(dst) + dst_off[r];
    a.src[r] = p->base[r] + kPeerFlagBytes + region_off + src_off[r];
    a.po.start[r + 1] = a.po.start[r] + nbytes[r] / 16;
  }
  const int64_t total = a.po.start[p->world];
  if (total == 0) return;
  peer_get_kernel<<<resident_grid(peer_get_kernel, total, kThreads * 4), kThreads, 0, s>>>(a);
  MONO_CHECK_LAUNCH();
}

PeerOut peer_out(mono_peer* p, int64_t region_off, const int64_t* dst_item_off, const int64_t* counts,
                 int64_t item_bytes) {
  if (!p->attached) throw ArgError("peer window is not attached");
  PeerOut po;
  std::memset(&po, 0, sizeof(po));
  po.n = p->world;
  for (int r = 0; r < p->world; ++r) {
    if (counts[r] < 0 || dst_item_off[r] < 0) throw ArgError("peer push: negative count or offset");
    check_region(p, region_off + dst_item_off[r] * item_bytes, counts[r] * item_bytes,
                 "peer push: destination outside the window");
    po.base[r] = p->base[r] + kPeerFlagBytes + region_off + dst_item_off[r] * item_bytes;
    po.start[r + 1] = po.start[r] + counts[r];
  }
  return po;
}

mono_peer* peer_create(int device, int world, int rank, size_t bytes) {
  if (world < 1 || world > kMaxPeers) throw ArgError("peer window: world size must be in [1, 16]");
  if (rank < 0 || rank >= world) throw ArgError("peer window: bad rank");
  MONO_CUDA(cudaSetDevice(device));
  auto p = new mono_peer();
  p->device = device;
  p->world = world;
  p->rank = rank;
  p->bytes = (bytes + 255) & ~(size_t)255;
  // cudaMalloc (not the stream-ordered pool): the allocation must be exportable with cudaIpcGetMemHandle
  cudaError_t e = cudaMalloc((void**)&p->local, kPeerFlagBytes + p->bytes);
  if (e != cudaSuccess) {
    delete p;
    throw CudaError(std::string("peer window cudaMalloc: ") + cudaGetErrorString(e));
  }
  MONO_CUDA(cudaMemset(p->local, 0, kPeerFlagBytes));
  MONO_CUDA(cudaDeviceSynchronize());
  p->base[rank] = p->local;
  if (world == 1) p->attached = true;
  return p;
}

// Unmap the peers' windows.  Teardown is two-phase: every rank detaches, the caller synchronises the
// ranks, then every rank destroys (an exported allocation must outlive its importers' mappings).
void peer_detach(mono_peer* p) {
  MONO_CUDA(cudaSetDevice(p->device));
  MONO_CUDA(cudaDeviceSynchronize());
  for (int r = 0; r < p->world; ++r)
    if (r != p->rank && p->base[r]) {
      cudaIpcCloseMemHandle(p->base[r]);
      p->base[r] = nullptr;
    }
  p->attached = p->world == 1;
}

void peer_destroy(mono_peer* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < p->world; ++r)
    if (r != p->rank && p->base[r]) cudaIpcCloseMemHandle(p->base[r]);
  if (p->local) cudaFree(p->local);
  delete p;
}

void peer_handle(mono_peer* p, void* out64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  MONO_CUDA(cudaSetDevice(p->device));
  cudaIpcMemHandle_t h;
  MONO_CUDA(cudaIpcGetMemHandle(&h, p->local));
  std::memcpy(out64, &h, 64);
}

void peer_attach(mono_peer* p, const void* handles) {
  MONO_CUDA(cudaSetDevice(p->device));
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank || p->base[r]) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char*>(handles) + 64 * (size_t)r, 64);
    void* ptr = nullptr;
    MONO_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->base[r] = static_cast<char*>(ptr);
  }
  p->attached = true;
}

}  // namespace mono
