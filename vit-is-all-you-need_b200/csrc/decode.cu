// Incremental (KV-cached) decode attention for VideoGPT.generate (train_videogpt.py:56-65).
//
// The reference re-runs the whole causal stack over the growing sequence for every generated token (O(N^2) attention
// and O(N) GEMM work per token).  Here the keys / values of every layer stay resident in a cache with the layout of
// the fused QKV projection output, [B, Nmax, 3, H, 64] bf16 (transformer.py:27 `(qkv h d)`), so a step appends one row
// per layer and attends one query against the cached rows:
//     o[b, h, :] = softmax(q . K[0..len)^T / sqrt(64)) V[0..len)        (the new token sees itself: len = pos + 1)
//
// HBM-bound: one CTA per (batch, head) streams len x 256 B of K and V once.  Scores: one key per thread (the query
// lives in registers), block-wide max / sum, then P V with a warp per key and a lane per pair of head dims (128-byte
// coalesced V rows).  fp32 math, bf16 output (the operand of the following LayerNorm-add).
#include "../../include/b200vit.h"
#include "common.cuh"

namespace b200 {

constexpr int DEC_THREADS = 128;
constexpr int DEC_WARPS = DEC_THREADS / 32;

__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_kernel(const __nv_bfloat16* __restrict__ cache, __nv_bfloat16* __restrict__ out, int Nmax, int H, int pos, int len) {
  extern __shared__ float s_scores[];           // [len]
  __shared__ float s_red[DEC_WARPS];
  __shared__ float s_acc[DEC_WARPS][64];
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row_stride = 3LL * H * 64;    // elements between consecutive positions
  const __nv_bfloat16* base = cache + (long long)b * Nmax * row_stride + (long long)h * 64;
  const __nv_bfloat16* qrow = base + (long long)pos * row_stride;          // slot 0: q
  const __nv_bfloat16* kbase = base + (long long)H * 64;                   // slot 1: k
  const __nv_bfloat16* vbase = base + 2LL * H * 64;                        // slot 2: v

  float q[64];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(qrow) + j);
    const float2 a = unpack_bf16(u.x), c = unpack_bf16(u.y), d = unpack_bf16(u.z), e = unpack_bf16(u.w);
    q[j * 8 + 0] = a.x; q[j * 8 + 1] = a.y; q[j * 8 + 2] = c.x; q[j * 8 + 3] = c.y;
    q[j * 8 + 4] = d.x; q[j * 8 + 5] = d.y; q[j * 8 + 6] = e.x; q[j * 8 + 7] = e.y;
  }
  // scores in the log2 domain: (q . k) / 8 * log2(e)
  const float scale = 0.125f * 1.4426950408889634f;
  float mx = -INFINITY;
  for (int key = threadIdx.x; key < len; key += DEC_THREADS) {
    const uint4* kr = reinterpret_cast<const uint4*>(kbase + (long long)key * row_stride);
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 u = __ldg(kr + j);
      const float2 a = unpack_bf16(u.x), c = unpack_bf16(u.y), d = unpack_bf16(u.z), e = unpack_bf16(u.w);
      dot = fmaf(q[j * 8 + 0], a.x, dot); dot = fmaf(q[j * 8 + 1], a.y, dot);
      dot = fmaf(q[j * 8 + 2], c.x, dot); dot = fmaf(q[j * 8 + 3], c.y, dot);
      dot = fmaf(q[j * 8 + 4], d.x, dot); dot = fmaf(q[j * 8 + 5], d.y, dot);
      dot = fmaf(q[j * 8 + 6], e.x, dot); dot = fmaf(q[j * 8 + 7], e.y, dot);
    }
    dot *= scale;
    s_scores[key] = dot;
    mx = fmaxf(mx, dot);
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int key = threadIdx.x; key < len; key += DEC_THREADS) {
    const float p = exp2f(s_scores[key] - mx);
    s_scores[key] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
  // P V: warp per key, lane per pair of head dims
  float a0 = 0.f, a1 = 0.f;
  for (int key = warp; key < len; key += DEC_WARPS) {
    const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(vbase + (long long)key * row_stride) + lane);
    const float2 v = unpack_bf16(u);
    const float p = s_scores[key];
    a0 = fmaf(p, v.x, a0); a1 = fmaf(p, v.y, a1);
  }
  s_acc[warp][2 * lane] = a0; s_acc[warp][2 * lane + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 32) {
    const float inv = 1.0f / sum;
    const float o0 = ((s_acc[0][2 * lane] + s_acc[1][2 * lane]) + (s_acc[2][2 * lane] + s_acc[3][2 * lane])) * inv;
    const float o1 = ((s_acc[0][2 * lane + 1] + s_acc[1][2 * lane + 1]) + (s_acc[2][2 * lane + 1] + s_acc[3][2 * lane + 1])) * inv;
    reinterpret_cast<uint32_t*>(out + ((long long)b * H + h) * 64)[lane] = pack_bf16(o0, o1);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_attn_decode(const void* kv_cache, void* out_bf16, int B, int Nmax, int H, int pos, void* stream) {
  B200_REQUIRE(kv_cache && out_bf16 && B > 0 && H > 0 && Nmax > 0 && pos >= 0 && pos < Nmax,
               "attn_decode: bad arguments (0 <= pos < Nmax)");
  const int len = pos + 1;
  const size_t smem = sizeof(float) * (size_t)len;
  B200_REQUIRE(smem <= 200 * 1024, "attn_decode: %d cached positions exceed the shared-memory score buffer", len);
  if (smem > 40 * 1024) {
    B200_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  attn_decode_kernel<<<B * H, DEC_THREADS, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)kv_cache,
                                                                        (__nv_bfloat16*)out_bf16, Nmax, H, pos, len);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"
