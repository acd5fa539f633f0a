// Incremental (KV-cached) decode attention for VideoGPT.generate (train_videogpt.py:56-65).
//
// The reference re-runs the whole causal stack over the growing sequence for every generated token (O(N^2) attention
// and O(N) GEMM work per token).  Here the keys / values of every layer stay resident in a cache with the layout of
// the fused QKV projection output, [B, Nmax, 3, H, 64] bf16 (transformer.py:27 `(qkv h d)`), so a step appends one row
// per layer and attends one query against the cached rows:
//     o[b, h, :] = softmax(q . K[0..len)^T / sqrt(64)) V[0..len)        (the new token sees itself: len = pos + 1)
//
// HBM-bound in bytes (len x 256 B of K and V per (batch, head), once), latency-bound in practice: a cluster of four
// CTAs per (batch, head) splits the keys (see the kernel).  Scores: one key per thread (the query lives in registers),
// block-wide max / sum, then P V with four value rows per warp-wide load.  fp32 math, bf16 output (the operand of the
// following LayerNorm-add).
#include "../../include/b200vit.h"
#include <cooperative_groups.h>

#include "common.cuh"

namespace b200 {

constexpr int DEC_THREADS = 256;
constexpr int DEC_WARPS = DEC_THREADS / 32;
// CTAs per (batch, head).  The kernel is written for a thread-block cluster that splits the keys and combines the
// partials through distributed shared memory, but MEASURED on B200 (VideoGPT-B, batch 16, 520 keys): cluster of 4 x 128
// threads 17.0 us vs one CTA of 256 threads 13.7 us per launch, and a CUDA graph whose kernel nodes carry a cluster
// dimension replays at eager speed (1.54 instead of 0.40 ms per token) -- so the shipped configuration is 1.
constexpr int DEC_CLUSTER = 1;

// A thread-block cluster of DEC_CLUSTER CTAs serves one (batch, head): CTA r scores and weights keys
// [r * chunk, (r + 1) * chunk) on its own (flash-decoding split), then deposits its partial (max, sum, 64 weighted
// sums) in the leader CTA's shared memory through distributed shared memory; after one cluster barrier the leader
// rescales and writes the output row.  The kernel is a chain of dependent memory round trips (q, keys, values), so
// what the split buys is a four times shorter chain per CTA and four times the loads in flight per (batch, head).
__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_kernel(const __nv_bfloat16* __restrict__ cache, __nv_bfloat16* __restrict__ out, int Nmax, int H,
                   const int* __restrict__ pos_dev) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  cluster.barrier_arrive();   // "I have started": waited for just before the first remote shared-memory write
  // the position lives in device memory so that one captured CUDA graph serves every step of a generation
  const int pos = min(max(__ldg(pos_dev), 0), Nmax - 1), len = pos + 1;
  const int chunk = (len + DEC_CLUSTER - 1) / DEC_CLUSTER;
  const int k0 = rank * chunk, k1 = min(len, k0 + chunk);     // this CTA's keys (possibly none)
  extern __shared__ float s_scores[];                          // [ceil(Nmax / DEC_CLUSTER)]
  __shared__ float s_red[DEC_WARPS];
  __shared__ float s_acc[DEC_WARPS][64];
  __shared__ float s_part[DEC_CLUSTER][66];                    // leader only: per CTA {64 weighted sums, max, sum}
  const int bh = blockIdx.x / DEC_CLUSTER;
  const int b = bh / H, h = bh - b * H;
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
  for (int key = k0 + threadIdx.x; key < k1; key += DEC_THREADS) {
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
    s_scores[key - k0] = dot;
    mx = fmaxf(mx, dot);
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < DEC_WARPS; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  const float mref = mx == -INFINITY ? 0.f : mx;               // a CTA without keys contributes zeros
  float sum = 0.f;
  for (int key = k0 + threadIdx.x; key < k1; key += DEC_THREADS) {
    const float p = exp2f(s_scores[key - k0] - mref);
    s_scores[key - k0] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < DEC_WARPS; ++w) sum += s_red[w];
  // P V: a warp-wide load covers 4 value rows (8 lanes x 16 B each); every lane accumulates 8 head dims of its key
  // group, 4 loads in flight per lane; then shuffle + shared reduction
  const int kg = lane >> 3, dl = lane & 7;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int key = k0 + warp * 4 + kg; key < k1; key += DEC_WARPS * 4 * 4) {
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = key + i * DEC_WARPS * 4;
      u[i] = kk < k1 ? __ldg(reinterpret_cast<const uint4*>(vbase + (long long)kk * row_stride) + dl) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = key + i * DEC_WARPS * 4;
      const float p = kk < k1 ? s_scores[kk - k0] : 0.f;
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
  // deposit this CTA's partial in the leader's shared memory (distributed shared memory), then combine there
  cluster.barrier_wait();     // every CTA of the cluster is resident: its shared memory may be written
  float* leader = cluster.map_shared_rank(&s_part[0][0], 0);
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) t += s_acc[w][threadIdx.x];
    leader[rank * 66 + threadIdx.x] = t;
  } else if (threadIdx.x == 64) {
    leader[rank * 66 + 64] = mx;     // -inf when this CTA had no keys
    leader[rank * 66 + 65] = sum;
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x < 32) {
    float M = -INFINITY;
#pragma unroll
    for (int r = 0; r < DEC_CLUSTER; ++r) M = fmaxf(M, s_part[r][64]);
    float tot = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int r = 0; r < DEC_CLUSTER; ++r) {
      const float mr = s_part[r][64];
      const float w = mr == -INFINITY ? 0.f : exp2f(mr - M);
      tot = fmaf(s_part[r][65], w, tot);
      o0 = fmaf(s_part[r][2 * lane], w, o0);
      o1 = fmaf(s_part[r][2 * lane + 1], w, o1);
    }
    const float inv = 1.0f / tot;
    reinterpret_cast<uint32_t*>(out + ((long long)b * H + h) * 64)[lane] = pack_bf16(o0 * inv, o1 * inv);
  }
}

// cache[b, *pos, :] = row[b, :]   (row = the new token's fused q | k | v, `row_elems` = 3 * H * 64 bf16 values)
__global__ void __launch_bounds__(256)
kv_append_kernel(const __nv_bfloat16* __restrict__ rows, __nv_bfloat16* __restrict__ cache, int B, int Nmax, int row_elems,
                 const int* __restrict__ pos_dev) {
  const int pos = min(max(__ldg(pos_dev), 0), Nmax - 1);
  const int per_row = row_elems >> 3;   // 16-byte chunks
  const long long total = (long long)B * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per_row;
    const int c = (int)(i - b * per_row);
    reinterpret_cast<uint4*>(cache + ((long long)b * Nmax + pos) * row_elems)[c] =
        __ldg(reinterpret_cast<const uint4*>(rows + b * row_elems) + c);
  }
}

__global__ void advance_counter_kernel(int* c, int by) { *c += by; }

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_attn_decode(const void* kv_cache, void* out_bf16, int B, int Nmax, int H, const int* pos_dev, void* stream) {
  B200_REQUIRE(kv_cache && out_bf16 && pos_dev && B > 0 && H > 0 && Nmax > 0, "attn_decode: bad arguments");
  const size_t smem = sizeof(float) * (size_t)((Nmax + DEC_CLUSTER - 1) / DEC_CLUSTER);
  B200_REQUIRE(smem <= 200 * 1024, "attn_decode: %d cache positions exceed the shared-memory score buffer", Nmax);
  if (smem > 40 * 1024) {
    B200_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (DEC_CLUSTER > 1) {
    B200_CUDA(launch_kernel(attn_decode_kernel, dim3(B * H * DEC_CLUSTER), dim3(DEC_THREADS), smem, (cudaStream_t)stream,
                            DEC_CLUSTER, (const __nv_bfloat16*)kv_cache, (__nv_bfloat16*)out_bf16, Nmax, H, pos_dev));
  } else {
    attn_decode_kernel<<<B * H, DEC_THREADS, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)kv_cache,
                                                                          (__nv_bfloat16*)out_bf16, Nmax, H, pos_dev);
    B200_CUDA(cudaGetLastError());
  }
  return OK;
}

int b200vit_kv_append(const void* rows_bf16, void* kv_cache, int B, int Nmax, int row_elems, const int* pos_dev, void* stream) {
  B200_REQUIRE(rows_bf16 && kv_cache && pos_dev && B > 0 && Nmax > 0 && row_elems > 0 && row_elems % 8 == 0,
               "kv_append: bad arguments (row_elems must be a multiple of 8)");
  const long long total = (long long)B * (row_elems / 8);
  kv_append_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)rows_bf16,
                                                                                (__nv_bfloat16*)kv_cache, B, Nmax, row_elems, pos_dev);
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
