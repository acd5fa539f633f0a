// Incremental (KV-cached) decode attention for VideoGPT.generate (train_videogpt.py:56-65).
//
// The reference re-runs the whole causal stack over the growing sequence for every generated token (O(N^2) attention
// and O(N) GEMM work per token).  Here the keys / values of every layer stay resident in a per-head cache
//     kv_cache[2][B][H][Nmax][64]  bf16      (K planes, then V planes; one head's rows are contiguous)
// so a step appends one row per (layer, head) and attends one query against the cached rows:
//     o[b, h, :] = softmax(q . K[0..len)^T / sqrt(64)) V[0..len)        (the new token sees itself: len = pos + 1)
//
// HBM-bound: one CTA per (batch, head) streams len x 256 contiguous bytes of K and V once -- by ONE bulk TMA copy per
// plane into shared memory when the cache is short enough for two CTAs per SM (attn_decode_bulk_kernel: two memory round
// trips per CTA), else with per-lane global loads (attn_decode_kernel: ~6 dependent round trips of ~1 us each).  The layout matters more
// than anything else here: with the cache kept in the fused-QKV layout [B, Nmax, 3, H, 64] every key row was an isolated
// 128-byte piece at a 4.6 KB stride and the kernel ran at ~2 TB/s (12.9 us at 520 keys, batch 16); splitting the keys over
// 4 CTAs (plain launches + combine kernel: 16.4 us; thread-block cluster + DSMEM combine: 17.0 us) did not help.
// Scores and P V both read four consecutive rows per warp-wide load (8 lanes x 16 B per row), block-wide max / sum in
// between.  fp32 math, bf16 output (the operand of the following LayerNorm-add).  The position is read from
// DEVICE memory so that one captured CUDA graph serves every step of a generation.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace b200 {

constexpr int DEC_THREADS = 256;
constexpr int DEC_WARPS = DEC_THREADS / 32;

__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_kernel(const __nv_bfloat16* __restrict__ qkv_rows, const __nv_bfloat16* __restrict__ cache,
                   __nv_bfloat16* __restrict__ out, int B, int Nmax, int H, const int* __restrict__ pos_dev) {
  const int pos = min(max(__ldg(pos_dev), 0), Nmax - 1), len = pos + 1;
  extern __shared__ float s_scores[];           // [Nmax] (len used)
  __shared__ float s_red[DEC_WARPS];
  __shared__ float s_acc[DEC_WARPS][64];
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* qrow = qkv_rows + (long long)b * 3 * H * 64 + h * 64;            // q of the new token
  const __nv_bfloat16* kbase = cache + ((long long)b * H + h) * Nmax * 64;              // K plane of (b, h)
  const __nv_bfloat16* vbase = kbase + (long long)B * H * Nmax * 64;                    // V plane of (b, h)

  // Scores, coalesced: 8 lanes share one key row (16 bytes each), so a warp-wide load covers 4 consecutive rows = 512
  // contiguous bytes (a thread-per-row scheme touches 32 different 128-byte lines per load instruction and is bound by
  // L1 tag throughput); partial dot products are reduced over the 8 lanes with 3 shuffles.  8 loads in flight per lane.
  const int kg = lane >> 3, dl = lane & 7;
  float qf[8];
  {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(qrow) + dl);
    const float2 a = unpack_bf16(u.x), c = unpack_bf16(u.y), d = unpack_bf16(u.z), e = unpack_bf16(u.w);
    qf[0] = a.x; qf[1] = a.y; qf[2] = c.x; qf[3] = c.y; qf[4] = d.x; qf[5] = d.y; qf[6] = e.x; qf[7] = e.y;
  }
  // scores in the log2 domain: (q . k) / 8 * log2(e)
  const float scale = 0.125f * 1.4426950408889634f;
  float mx = -INFINITY;
  constexpr int QK_ILP = 8;
  // (warp-uniform trip count: the shuffles below need all 32 lanes; invalid keys are predicated inside)
  for (int base = warp * 4; base < len; base += DEC_WARPS * 4 * QK_ILP) {
    const int key = base + kg;
    uint4 u[QK_ILP];
#pragma unroll
    for (int i = 0; i < QK_ILP; ++i) {
      const int kk = key + i * DEC_WARPS * 4;
      u[i] = kk < len ? __ldg(reinterpret_cast<const uint4*>(kbase + (long long)kk * 64) + dl) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < QK_ILP; ++i) {
      const int kk = key + i * DEC_WARPS * 4;
      const float2 a = unpack_bf16(u[i].x), c = unpack_bf16(u[i].y), d = unpack_bf16(u[i].z), e = unpack_bf16(u[i].w);
      float dot = qf[0] * a.x;
      dot = fmaf(qf[1], a.y, dot); dot = fmaf(qf[2], c.x, dot); dot = fmaf(qf[3], c.y, dot);
      dot = fmaf(qf[4], d.x, dot); dot = fmaf(qf[5], d.y, dot); dot = fmaf(qf[6], e.x, dot); dot = fmaf(qf[7], e.y, dot);
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      if (kk < len) {
        dot *= scale;
        if (dl == 0) s_scores[kk] = dot;
        mx = fmaxf(mx, dot);
      }
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < DEC_WARPS; ++w) mx = fmaxf(mx, s_red[w]);
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
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < DEC_WARPS; ++w) sum += s_red[w];
  // P V: a warp-wide load covers 4 consecutive value rows (8 lanes x 16 B each, 512 contiguous bytes); every lane
  // accumulates 8 head dims of its key group, 8 loads in flight per lane; then shuffle + shared reduction
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  constexpr int PV_ILP = 8;
  for (int key = warp * 4 + kg; key < len; key += DEC_WARPS * 4 * PV_ILP) {
    uint4 u[PV_ILP];
#pragma unroll
    for (int i = 0; i < PV_ILP; ++i) {
      const int kk = key + i * DEC_WARPS * 4;
      u[i] = kk < len ? __ldg(reinterpret_cast<const uint4*>(vbase + (long long)kk * 64) + dl) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < PV_ILP; ++i) {
      const int kk = key + i * DEC_WARPS * 4;
      const float p = kk < len ? s_scores[kk] : 0.f;
      const float2 v0 = unpack_bf16(u[i].x), v1 = unpack_bf16(u[i].y), v2 = unpack_bf16(u[i].z), v3 = unpack_bf16(u[i].w);
      acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]); acc[2] = fmaf(p, v1.x, acc[2]); acc[3] = fmaf(p, v1.y, acc[3]);
      acc[4] = fmaf(p, v2.x, acc[4]); acc[5] = fmaf(p, v2.y, acc[5]); acc[6] = fmaf(p, v3.x, acc[6]); acc[7] = fmaf(p, v3.y, acc[7]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[warp][dl * 8 + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const float inv = 1.0f / sum;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) { o0 += s_acc[w][2 * lane]; o1 += s_acc[w][2 * lane + 1]; }
    reinterpret_cast<uint32_t*>(out + ((long long)b * H + h) * 64)[lane] = pack_bf16(o0 * inv, o1 * inv);
  }
}

// Variant with the K plane, then the V plane, of the (batch, head) brought into shared memory by ONE bulk TMA copy each
// (the rows are contiguous: len x 128 bytes): two memory round trips per CTA instead of ~6 dependent ones.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_bulk_kernel(const __nv_bfloat16* __restrict__ qkv_rows, const __nv_bfloat16* __restrict__ cache,
                        __nv_bfloat16* __restrict__ out, int B, int Nmax, int H, const int* __restrict__ pos_dev) {
  const int pos = min(max(__ldg(pos_dev), 0), Nmax - 1), len = pos + 1;
  extern __shared__ __align__(128) uint8_t dec_smem[];
  __nv_bfloat16* s_rows = reinterpret_cast<__nv_bfloat16*>(dec_smem);                 // [Nmax][64] bf16: K, then V
  float* s_scores = reinterpret_cast<float*>(dec_smem + (size_t)Nmax * 128);          // [Nmax]
  __shared__ float s_red[DEC_WARPS];
  __shared__ float s_acc[DEC_WARPS][64];
  __shared__ __align__(8) uint64_t s_bar;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* qrow = qkv_rows + (long long)b * 3 * H * 64 + h * 64;
  const __nv_bfloat16* kbase = cache + ((long long)b * H + h) * Nmax * 64;
  const __nv_bfloat16* vbase = kbase + (long long)B * H * Nmax * 64;
  const uint32_t bytes = (uint32_t)len * 128u;
  if (threadIdx.x == 0) { mbar_init(&s_bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&s_bar, bytes); bulk_g2s(s_rows, kbase, bytes, &s_bar); }
  const int kg = lane >> 3, dl = lane & 7;
  float qf[8];
  {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(qrow) + dl);
    const float2 a = unpack_bf16(u.x), c = unpack_bf16(u.y), d = unpack_bf16(u.z), e = unpack_bf16(u.w);
    qf[0] = a.x; qf[1] = a.y; qf[2] = c.x; qf[3] = c.y; qf[4] = d.x; qf[5] = d.y; qf[6] = e.x; qf[7] = e.y;
  }
  const float scale = 0.125f * 1.4426950408889634f;
  float mx = -INFINITY;
  mbar_wait(&s_bar, 0, 70);
  for (int base = warp * 4; base < len; base += DEC_WARPS * 4) {   // warp-uniform trip count (shuffles inside)
    const int kk = base + kg;
    const uint4 u = kk < len ? *(reinterpret_cast<const uint4*>(s_rows + (long long)kk * 64) + dl) : make_uint4(0, 0, 0, 0);
    const float2 a = unpack_bf16(u.x), c = unpack_bf16(u.y), d = unpack_bf16(u.z), e = unpack_bf16(u.w);
    float dot = qf[0] * a.x;
    dot = fmaf(qf[1], a.y, dot); dot = fmaf(qf[2], c.x, dot); dot = fmaf(qf[3], c.y, dot);
    dot = fmaf(qf[4], d.x, dot); dot = fmaf(qf[5], d.y, dot); dot = fmaf(qf[6], e.x, dot); dot = fmaf(qf[7], e.y, dot);
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    if (kk < len) {
      dot *= scale;
      if (dl == 0) s_scores[kk] = dot;
      mx = fmaxf(mx, dot);
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();                                   // every warp is done reading the K rows
  if (threadIdx.x == 0) {                            // V plane into the same buffer while the softmax is normalised
    fence_proxy_async_smem();
    mbar_expect_tx(&s_bar, bytes);
    bulk_g2s(s_rows, vbase, bytes, &s_bar);
  }
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < DEC_WARPS; ++w) mx = fmaxf(mx, s_red[w]);
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
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < DEC_WARPS; ++w) sum += s_red[w];
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  mbar_wait(&s_bar, 1, 71);
  for (int kk = warp * 4 + kg; kk < len; kk += DEC_WARPS * 4) {
    const uint4 u = *(reinterpret_cast<const uint4*>(s_rows + (long long)kk * 64) + dl);
    const float p = s_scores[kk];
    const float2 v0 = unpack_bf16(u.x), v1 = unpack_bf16(u.y), v2 = unpack_bf16(u.z), v3 = unpack_bf16(u.w);
    acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]); acc[2] = fmaf(p, v1.x, acc[2]); acc[3] = fmaf(p, v1.y, acc[3]);
    acc[4] = fmaf(p, v2.x, acc[4]); acc[5] = fmaf(p, v2.y, acc[5]); acc[6] = fmaf(p, v3.x, acc[6]); acc[7] = fmaf(p, v3.y, acc[7]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[warp][dl * 8 + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const float inv = 1.0f / sum;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) { o0 += s_acc[w][2 * lane]; o1 += s_acc[w][2 * lane + 1]; }
    reinterpret_cast<uint32_t*>(out + ((long long)b * H + h) * 64)[lane] = pack_bf16(o0 * inv, o1 * inv);
  }
}

// K and V of the new token (slots 1 and 2 of its fused q | k | v row) -> kv_cache[{0,1}][b][h][*pos][:]
__global__ void __launch_bounds__(256)
kv_append_kernel(const __nv_bfloat16* __restrict__ rows, __nv_bfloat16* __restrict__ cache, int B, int Nmax, int H,
                 const int* __restrict__ pos_dev) {
  const int pos = min(max(__ldg(pos_dev), 0), Nmax - 1);
  const long long total = (long long)B * 2 * H * 8;   // 16-byte chunks: (b, k|v, h, 8 chunks of a 128-byte row)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7);
    const int h = (int)((i >> 3) % H);
    const int kv = (int)((i / (8 * H)) & 1);
    const long long b = i / (16LL * H);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rows + (b * 3 + 1 + kv) * H * 64 + h * 64) + c);
    reinterpret_cast<uint4*>(cache + (((long long)kv * B + b) * H + h) * Nmax * 64 + (long long)pos * 64)[c] = v;
  }
}

// prefill: qkv[B, S, 3, H, 64] (the fused QKV projection of the whole prompt) -> K / V planes [2][B][H][Nmax][64], rows 0..S
__global__ void __launch_bounds__(256)
kv_fill_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ cache, int B, int S, int Nmax, int H) {
  const long long total = (long long)B * S * 2 * H * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i & 7);
    const int h = (int)((i >> 3) % H);
    const int kv = (int)((i / (8 * H)) & 1);
    const long long bs = i / (16LL * H);
    const long long b = bs / S;
    const int s = (int)(bs - b * S);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(qkv + (bs * 3 + 1 + kv) * H * 64 + h * 64) + c);
    reinterpret_cast<uint4*>(cache + (((long long)kv * B + b) * H + h) * Nmax * 64 + (long long)s * 64)[c] = v;
  }
}

__global__ void advance_counter_kernel(int* c, int by) { *c += by; }

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_attn_decode(const void* qkv_rows, const void* kv_cache, void* out_bf16, int B, int Nmax, int H, const int* pos_dev,
                        void* stream) {
  B200_REQUIRE(qkv_rows && kv_cache && out_bf16 && pos_dev && B > 0 && H > 0 && Nmax > 0, "attn_decode: bad arguments");
  const size_t smem = sizeof(float) * (size_t)Nmax;
  B200_REQUIRE(smem <= 200 * 1024, "attn_decode: %d cache positions exceed the shared-memory score buffer", Nmax);
  if (smem > 40 * 1024) {
    B200_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const size_t smem_bulk = (size_t)Nmax * 128 + sizeof(float) * (size_t)Nmax;
  // default: K / V planes through one bulk TMA copy each (9.4 us against 11.6 us for the global-load kernel at 520 keys,
  // batch 16) while two CTAs still fit an SM; longer caches (and bring-up knob 9 = 2) take the global-load kernel
  if (g_debug[9] != 2 && smem_bulk <= 100 * 1024) {
    B200_CUDA(cudaFuncSetAttribute(attn_decode_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bulk));
    attn_decode_bulk_kernel<<<B * H, DEC_THREADS, smem_bulk, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)qkv_rows, (const __nv_bfloat16*)kv_cache, (__nv_bfloat16*)out_bf16, B, Nmax, H, pos_dev);
    B200_CUDA(cudaGetLastError());
    return OK;
  }
  attn_decode_kernel<<<B * H, DEC_THREADS, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv_rows, (const __nv_bfloat16*)kv_cache,
                                                                        (__nv_bfloat16*)out_bf16, B, Nmax, H, pos_dev);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_kv_append(const void* qkv_rows, void* kv_cache, int B, int Nmax, int H, const int* pos_dev, void* stream) {
  B200_REQUIRE(qkv_rows && kv_cache && pos_dev && B > 0 && Nmax > 0 && H > 0, "kv_append: bad arguments");
  const long long total = (long long)B * 2 * H * 8;
  kv_append_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv_rows,
                                                                                (__nv_bfloat16*)kv_cache, B, Nmax, H, pos_dev);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_kv_fill(const void* qkv, void* kv_cache, int B, int S, int Nmax, int H, void* stream) {
  B200_REQUIRE(qkv && kv_cache && B > 0 && S > 0 && S <= Nmax && H > 0, "kv_fill: bad arguments (S <= Nmax)");
  const long long total = (long long)B * S * 2 * H * 8;
  const long long want = (total + 255) / 256;
  const int cap = num_sms() * 16;
  kv_fill_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)kv_cache,
                                                                                  B, S, Nmax, H);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_advance_counter(int* counter, int by, void* stream) {
  B200_REQUIRE(counter != nullptr, "advance_counter: null pointer");
  advance_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, by);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"
