// tower.cu — the stand-in dense tower of bench.py's end-to-end step, as ONE kernel.
//
// NOT part of the reference's embedding interface (dense layers are out of scope): the reference model
// (markdown/demo/demo_model.py:47-84) puts an MLP on top of the pooled embeddings; the e2e bench needs SOMETHING that
// turns pooled rows into pooled-row gradients on the device so that the sparse forward and backward are driven by a
// real data dependency.  Through torch (cast, two skinny cuBLAS GEMMs, elementwise passes, GEMM back, cast) that took
// 0.68 ms per 524288-sample batch — more than the whole sparse step; fused it is one pass over the pooled rows.
//
//   x[B,64] fp32  ->  h = relu(bf16(x) W1)  ->  logit = h . w2  ->  loss = BCEWithLogits(logit, y) (mean over B)
//   dlogit = (sigmoid(logit) - y) / B  ->  dh = dlogit * w2 (masked by h > 0)  ->  dx = dh W1^T   -> fp32 [B,64]
//
// bf16 operands on tensor cores (mma.sync m16n8k16, fp32 accumulate), bf16 rounding at the points where the torch
// formulation rounds (GEMM outputs, the outer product).  One warp owns 16 rows: its x fragment is loaded straight
// from global memory in the A-fragment layout (16 independent 8-byte loads per lane), the forward accumulators are
// re-packed into the A fragments of the backward GEMM in registers, and dx is stored from the C-fragment layout
// (32-byte sectors).  W1 (and its transpose) live in shared memory, padded so that fragment reads are conflict-free.
// HBM-bound: 256 B read + 256 B written per sample.
#include <cuda_bf16.h>

#include "engine.h"

namespace mono {
namespace {

constexpr int kTowerIn = 64, kTowerHid = 64;
constexpr int kPad = 8;                       // bf16 elements of row padding: row stride 72 * 2 B = 36 words
constexpr int kTowerThreads = 256;
constexpr int kRowsPerBlockIter = (kTowerThreads / 32) * 16;

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kTowerThreads, 2)
tower_grad_kernel(const float* __restrict__ x, int64_t B, const float* __restrict__ labels,
                  const __nv_bfloat16* __restrict__ w1 /*[64 in][64 out]*/, const __nv_bfloat16* __restrict__ w2 /*[64]*/,
                  float* __restrict__ dx, float* __restrict__ partial /*[gridDim.x]*/) {
  __shared__ __align__(16) __nv_bfloat16 s_w1[kTowerIn][kTowerHid + kPad];    // [in][out]  : B operand of dx = dh W1^T
  __shared__ __align__(16) __nv_bfloat16 s_w1t[kTowerHid][kTowerIn + kPad];   // [out][in]  : B operand of h = x W1
  __shared__ float s_w2[kTowerHid];
  __shared__ float s_loss[kTowerThreads / 32];
  for (int i = threadIdx.x; i < kTowerIn * kTowerHid; i += kTowerThreads) {
    const int r = i / kTowerHid, c = i % kTowerHid;
    const __nv_bfloat16 v = w1[i];
    s_w1[r][c] = v;
    s_w1t[c][r] = v;
  }
  for (int i = threadIdx.x; i < kTowerHid; i += kTowerThreads) s_w2[i] = __bfloat162float(w2[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int qr = lane >> 2, qc = (lane & 3) * 2;   // row within the 8-row half, first column of the lane's pair
  const float inv_b = 1.0f / (float)B;
  float loss_acc = 0.f;
  const int64_t n_tiles = (B + 15) / 16;
  // the x fragment of the NEXT tile is requested before the current tile is computed: 16 independent 8-byte loads per
  // lane stay in flight under the two GEMMs and the stores (the kernel is a stream: nothing else hides HBM latency)
  float2 v[4][4];
  float yv[2];
  auto request = [&](int64_t t) {
    const int64_t a0 = t * 16 + qr, a1 = a0 + 8;
    const bool k0 = a0 < B, k1 = a1 < B;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c0 = ks * 16 + qc;
      v[ks][0] = k0 ? __ldcs(reinterpret_cast<const float2*>(x + a0 * kTowerIn + c0)) : make_float2(0.f, 0.f);
      v[ks][1] = k1 ? __ldcs(reinterpret_cast<const float2*>(x + a1 * kTowerIn + c0)) : make_float2(0.f, 0.f);
      v[ks][2] = k0 ? __ldcs(reinterpret_cast<const float2*>(x + a0 * kTowerIn + c0 + 8)) : make_float2(0.f, 0.f);
      v[ks][3] = k1 ? __ldcs(reinterpret_cast<const float2*>(x + a1 * kTowerIn + c0 + 8)) : make_float2(0.f, 0.f);
    }
    yv[0] = k0 ? labels[a0] : 0.f;
    yv[1] = k1 ? labels[a1] : 0.f;
  };
  const int64_t tstride = (int64_t)gridDim.x * (kTowerThreads / 32);
  int64_t tile = (int64_t)blockIdx.x * (kTowerThreads / 32) + w;
  if (tile < n_tiles) request(tile);
  for (; tile < n_tiles; tile += tstride) {
    const int64_t r0 = tile * 16 + qr, r1 = r0 + 8;
    const bool ok0 = r0 < B, ok1 = r1 < B;
    // ---- x fragment: 4 k-steps x {row r0 | r1} x {cols k, k+8} ----
    uint32_t xa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int q = 0; q < 4; ++q) xa[ks][q] = pack_bf16(v[ks][q].x, v[ks][q].y);
    const float y0 = yv[0], y1 = yv[1];
    if (tile + tstride < n_tiles) request(tile + tstride);
    // ---- h = relu(x W1): 8 n-tiles of 8 hidden units ----
    float h[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      h[nt][0] = h[nt][1] = h[nt][2] = h[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t* bp = reinterpret_cast<const uint32_t*>(&s_w1t[nt * 8 + qr][ks * 16 + qc]);
        mma_bf16(h[nt], xa[ks], bp[0], bp[4]);   // (k, k+1) and (k+8, k+9) of hidden unit nt*8 + qr
      }
    }
    // logit = h . w2 over the 64 hidden units: this lane holds units nt*8 + qc, +1 of rows r0 (h[][0..1]) and r1 (h[][2..3])
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int q = 0; q < 4; ++q) h[nt][q] = fmaxf(round_bf16(h[nt][q]), 0.f);
      const float wa = s_w2[nt * 8 + qc], wb = s_w2[nt * 8 + qc + 1];
      l0 += h[nt][0] * wa + h[nt][1] * wb;
      l1 += h[nt][2] * wa + h[nt][3] * wb;
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    l0 = round_bf16(l0);
    l1 = round_bf16(l1);
    // loss and d loss / d logit (mean over the batch)
    if ((lane & 3) == 0) {
      if (ok0) loss_acc += fmaxf(l0, 0.f) - l0 * y0 + log1pf(__expf(-fabsf(l0)));
      if (ok1) loss_acc += fmaxf(l1, 0.f) - l1 * y1 + log1pf(__expf(-fabsf(l1)));
    }
    const float d0 = round_bf16((1.f / (1.f + __expf(-l0)) - y0) * inv_b);
    const float d1 = round_bf16((1.f / (1.f + __expf(-l1)) - y1) * inv_b);
    // ---- dh in the C-fragment layout -> A fragments of the backward GEMM (k = hidden unit) ----
    uint32_t da[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int nt = 2 * ks + half;
        const float wa = s_w2[nt * 8 + qc], wb = s_w2[nt * 8 + qc + 1];
        da[ks][2 * half + 0] = pack_bf16(h[nt][0] > 0.f ? d0 * wa : 0.f, h[nt][1] > 0.f ? d0 * wb : 0.f);
        da[ks][2 * half + 1] = pack_bf16(h[nt][2] > 0.f ? d1 * wa : 0.f, h[nt][3] > 0.f ? d1 * wb : 0.f);
      }
    }
    // ---- dx = dh W1^T: 8 n-tiles of 8 input columns; stored from the C layout (one 32-byte sector per row and tile) ----
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t* bp = reinterpret_cast<const uint32_t*>(&s_w1[nt * 8 + qr][ks * 16 + qc]);
        mma_bf16(g, da[ks], bp[0], bp[4]);       // W1[in = nt*8 + qr][hidden k, k+1 | k+8, k+9]
      }
      const int c0 = nt * 8 + qc;
      if (ok0) __stcs(reinterpret_cast<float2*>(dx + r0 * kTowerIn + c0), make_float2(round_bf16(g[0]), round_bf16(g[1])));
      if (ok1) __stcs(reinterpret_cast<float2*>(dx + r1 * kTowerIn + c0), make_float2(round_bf16(g[2]), round_bf16(g[3])));
    }
  }
  // ---- loss: lanes -> warp -> block partial (fixed order) ----
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
  if (lane == 0) s_loss[w] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kTowerThreads / 32; ++i) t += s_loss[i];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) tower_loss_kernel(const float* __restrict__ partial, int n, float inv_b,
                                                         float* __restrict__ loss) {
  __shared__ float s[256];
  float t = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) t += partial[i];
  s[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = s[0] * inv_b;
}

}  // namespace

int tower_scratch_floats() { return 148 * 8; }

void tower_grad(const float* x, int64_t batch, const float* labels, const void* w1_bf16, const void* w2_bf16,
                float* dx, float* loss, float* scratch, cudaStream_t s) {
  if (batch <= 0) return;
  const int64_t tiles = (batch + 15) / 16;
  // persistent: two resident blocks per SM, every warp walks its tiles with the next one's loads in flight
  const int grid = (int)std::min<int64_t>((tiles + (kTowerThreads / 32) - 1) / (kTowerThreads / 32), 148 * 2);
  tower_grad_kernel<<<grid, kTowerThreads, 0, s>>>(x, batch, labels, (const __nv_bfloat16*)w1_bf16,
                                                   (const __nv_bfloat16*)w2_bf16, dx, scratch);
  MONO_CHECK_LAUNCH();
  tower_loss_kernel<<<1, 256, 0, s>>>(scratch, grid, 1.0f / (float)batch, loss);
  MONO_CHECK_LAUNCH();
}

}  // namespace mono
