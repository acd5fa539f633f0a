// Classifier head plumbing + cross-entropy: the steps immediately after the encoder stack.
//
//   train_vit.py:51-53   ViTClassifier.forward: head(vit(x)[:, 0])       -> gather_tokens (fp32 rows -> bf16 GEMM operand),
//                                                                           tcgen05 GEMM (gemm.cu), scatter_tokens (backward)
//   train_vit.py:81,102  nn.CrossEntropyLoss()(pred, labels)             -> cross_entropy_fwd / _bwd
//   train_videogpt.py:53-54 cross entropy over [B*1024, 1024] logits     -> same kernels (one warp per row)
//
// All HBM-bound and small next to the stack; the point of fusing them is launch count and the backward fill: the
// gradient of `x[:, 0]` is a [B, N, d] tensor that is zero except for one row per image, which PyTorch produces with a
// fill + a strided copy and the stack's backward then casts to bf16 (three passes over 155 MB at ViT-B/16, B = 256);
// scatter_tokens writes the fp32 tensor and its bf16 twin in one pass.  Both take a token RANGE: the tokenizer encoder
// projects the first `latent_tokens` tokens (train_titok.py:41-42), the decoder de-patchifies the first `n_patches` (train_titok.py:71).
//
// Cross-entropy follows torch.nn.functional.cross_entropy (reduction='mean', ignore_index, no label smoothing, no class
// weights -- what every reference script uses): row loss = logsumexp(x) - x[label], mean over the non-ignored rows;
// d logits = (softmax(x) - onehot) * upstream / n_valid, written in the dtype of the logits.  fp32 math throughout.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace b200 {

// out[b * cnt + t, :] = bf16(x[b, t0 + t, :]) for t in [0, cnt): the cnt * d values of one image are contiguous in x
__global__ void __launch_bounds__(256)
gather_tokens_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, long long img_stride,
                     long long per_img /* cnt * d */, long long token_off) {
  const long long per4 = per_img >> 2;
  const long long total = (long long)B * per4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per4;
    const long long c = (i - b * per4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + b * img_stride + token_off + c);
    uint2 w;
    w.x = pack_bf16(v.x, v.y); w.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + b * per_img + c) = w;
  }
}

// dx[B, N, d] (fp32) = 0 except dx[:, t0 : t0 + cnt] = dy[B * cnt, d] ; optional bf16 twin.  One pass of pure stores.
template <bool DY_BF16>
__global__ void __launch_bounds__(256)
scatter_tokens_kernel(const void* __restrict__ dy_, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16, int B, int N,
                      int d, int t0, int cnt) {
  const int per_row = d >> 2;
  const long long total = (long long)B * N * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / per_row;
    const int c = (int)(i - row * per_row) * 4;
    const long long b = row / N;
    const int t = (int)(row - b * N);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= t0 && t < t0 + cnt) {
      const long long src = (b * cnt + (t - t0)) * d + c;
      if constexpr (DY_BF16) {
        const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(dy_) + src);
        const float2 p0 = unpack_bf16(u.x), p1 = unpack_bf16(u.y);
        v = make_float4(p0.x, p0.y, p1.x, p1.y);
      } else {
        v = *reinterpret_cast<const float4*>(static_cast<const float*>(dy_) + src);
      }
    }
    *reinterpret_cast<float4*>(dx + row * d + c) = v;
    if (dx16 != nullptr) {
      uint2 w;
      w.x = pack_bf16(v.x, v.y); w.y = pack_bf16(v.z, v.w);
      *reinterpret_cast<uint2*>(dx16 + row * d + c) = w;
    }
  }
}

template <typename T> __device__ __forceinline__ float ld_logit(const T* p);
template <> __device__ __forceinline__ float ld_logit<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_logit<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_logit(T* p, float v);
template <> __device__ __forceinline__ void st_logit<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_logit<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int CE_WARPS = 8;

// one warp per row: max, sum of exponentials (second read hits L1/L2), row loss
template <typename T>
__global__ void __launch_bounds__(CE_WARPS * 32)
ce_fwd_kernel(const T* __restrict__ logits, long long ld, const long long* __restrict__ labels, float* __restrict__ row_loss,
              float* __restrict__ lse, int R, int C, long long ignore_index) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * CE_WARPS + warp;
  if (r >= R) return;
  const T* x = logits + r * ld;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, ld_logit<T>(x + c));
  mx = warp_max(mx);
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += expf(ld_logit<T>(x + c) - mx);
  s = warp_sum(s);
  if (lane == 0) {
    const float l = mx + logf(s);
    const long long y = labels[r];
    lse[r] = l;
    row_loss[r] = (y == ignore_index || y < 0 || y >= C) ? 0.f : l - ld_logit<T>(x + y);
  }
}

// loss[0] = mean over non-ignored rows, loss[1] = 1 / n_valid (0 when every row is ignored); deterministic order
__global__ void __launch_bounds__(256)
ce_reduce_kernel(const float* __restrict__ row_loss, const long long* __restrict__ labels, float* __restrict__ loss, int R,
                 int C, long long ignore_index) {
  __shared__ float ssum[8];
  __shared__ float scnt[8];
  // a label outside [0, C) that is not ignore_index is an error (torch: device assert).  There is no host sync here, so
  // the loud equivalent is a NaN loss and, through loss[1], NaN gradients: a label / vocabulary mismatch cannot train silently
  float s = 0.f, n = 0.f;
  for (int r = threadIdx.x; r < R; r += 256) {
    const long long y = labels[r];
    if (y == ignore_index) continue;
    if (y < 0 || y >= C) { s = nanf(""); n += 1.f; continue; }
    s += row_loss[r]; n += 1.f;
  }
  s = warp_sum(s); n = warp_sum(n);
  if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = s; scnt[threadIdx.x >> 5] = n; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ts = 0.f, tn = 0.f;
    for (int i = 0; i < 8; ++i) { ts += ssum[i]; tn += scnt[i]; }
    loss[0] = tn > 0.f ? ts / tn : nanf("");   // torch: mean over zero rows is nan
    loss[1] = tn > 0.f ? (ts != ts ? ts : 1.f / tn) : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(CE_WARPS * 32)
ce_bwd_kernel(const T* __restrict__ logits, long long ld, const long long* __restrict__ labels, const float* __restrict__ lse,
              const float* __restrict__ loss, const float* __restrict__ dloss, T* __restrict__ dlogits, long long ldd, int R,
              int C, long long ignore_index) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * CE_WARPS + warp;
  if (r >= R) return;
  const long long y = labels[r];
  const bool ignored = (y == ignore_index || y < 0 || y >= C);
  const float scale = ignored ? 0.f : __ldg(dloss) * __ldg(loss + 1);
  const float l = lse[r];
  const T* x = logits + r * ld;
  T* dx = dlogits + r * ldd;
  for (int c = lane; c < C; c += 32) {
    const float p = expf(ld_logit<T>(x + c) - l);
    st_logit<T>(dx + c, (p - (c == y ? 1.f : 0.f)) * scale);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_gather_tokens_bf16(const float* x, void* out_bf16, int B, int N, int d, int t0, int cnt, void* stream) {
  B200_REQUIRE(x && out_bf16 && B > 0 && N > 0 && d > 0 && d % 4 == 0 && t0 >= 0 && cnt > 0 && t0 + cnt <= N,
               "gather_tokens: bad arguments (d must be a multiple of 4, 0 <= t0, t0 + cnt <= N)");
  const long long total = (long long)B * cnt * (d / 4);
  const long long want = (total + 255) / 256;
  const int cap = num_sms() * 16;
  const int grid = (int)(want < cap ? want : cap);
  gather_tokens_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, B, (long long)N * d,
                                                               (long long)cnt * d, (long long)t0 * d);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_scatter_tokens(const void* dy, int dy_is_bf16, float* dx, void* dx_bf16, int B, int N, int d, int t0, int cnt,
                           void* stream) {
  B200_REQUIRE(dy && dx && B > 0 && N > 0 && d > 0 && d % 4 == 0 && t0 >= 0 && cnt > 0 && t0 + cnt <= N,
               "scatter_tokens: bad arguments (d must be a multiple of 4, 0 <= t0, t0 + cnt <= N)");
  const long long total = (long long)B * N * (d / 4);
  const long long want = (total + 255) / 256;
  const int cap = num_sms() * 16;
  const int grid = (int)(want < cap ? want : cap);
  if (dy_is_bf16) scatter_tokens_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(dy, dx, (__nv_bfloat16*)dx_bf16, B, N, d, t0, cnt);
  else            scatter_tokens_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(dy, dx, (__nv_bfloat16*)dx_bf16, B, N, d, t0, cnt);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_cross_entropy_fwd(const void* logits, int logits_bf16, long long ld, const long long* labels, float* loss,
                              float* lse, float* row_loss, int R, int C, long long ignore_index, void* stream) {
  B200_REQUIRE(logits && labels && loss && lse && row_loss && R > 0 && C > 0 && ld >= C, "cross_entropy_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (R + CE_WARPS - 1) / CE_WARPS;
  if (logits_bf16) ce_fwd_kernel<__nv_bfloat16><<<grid, CE_WARPS * 32, 0, st>>>((const __nv_bfloat16*)logits, ld, labels, row_loss, lse, R, C, ignore_index);
  else             ce_fwd_kernel<float><<<grid, CE_WARPS * 32, 0, st>>>((const float*)logits, ld, labels, row_loss, lse, R, C, ignore_index);
  B200_CUDA(cudaGetLastError());
  ce_reduce_kernel<<<1, 256, 0, st>>>(row_loss, labels, loss, R, C, ignore_index);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_cross_entropy_bwd(const void* logits, int logits_bf16, long long ld, const long long* labels, const float* lse,
                              const float* loss, const float* dloss, void* dlogits, long long ldd, int R, int C,
                              long long ignore_index, void* stream) {
  B200_REQUIRE(logits && labels && lse && loss && dloss && dlogits && R > 0 && C > 0 && ld >= C && ldd >= C,
               "cross_entropy_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (R + CE_WARPS - 1) / CE_WARPS;
  if (logits_bf16) ce_bwd_kernel<__nv_bfloat16><<<grid, CE_WARPS * 32, 0, st>>>((const __nv_bfloat16*)logits, ld, labels, lse, loss, dloss, (__nv_bfloat16*)dlogits, ldd, R, C, ignore_index);
  else             ce_bwd_kernel<float><<<grid, CE_WARPS * 32, 0, st>>>((const float*)logits, ld, labels, lse, loss, dloss, (float*)dlogits, ldd, R, C, ignore_index);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// Token + positional embedding of the autoregressive models (train_videogpt.py:45-50):
//   x[b, s, :] = tok_embed[idx[b, s]] + pos_embed[pos0 + s]          (fp32 residual stream, the stack's input)
// backward: d_tok_embed[idx[b, s]] += dy[b, s]  (fp32 atomics; rows repeat across the batch),
//           d_pos_embed[pos0 + s]   = sum_b dy[b, s]
// ------------------------------------------------------------------------------------------------------------
namespace b200 {

__global__ void __launch_bounds__(256)
embed_fwd_kernel(const long long* __restrict__ idx, const float* __restrict__ tok, const float* __restrict__ pos,
                 float* __restrict__ out, long long rows, int S, int d, int pos0, const int* __restrict__ pos0_dev, int n_pos,
                 int vocab) {
  if (pos0_dev != nullptr) pos0 = min(max(__ldg(pos0_dev), 0), n_pos - S);   // decode: the position lives on the device
  const int per_row = d >> 2;
  const long long total = rows * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / per_row;
    const int c = (int)(i - r * per_row) * 4;
    long long t = idx[r];
    const bool bad = t < 0 || t >= vocab;          // torch's embedding raises a device assert; without a host sync the
    t = bad ? 0 : t;                               // loud equivalent is a NaN row (never read out of bounds)
    const int s = (int)(r % S);
    float4 a = __ldg(reinterpret_cast<const float4*>(tok + t * d + c));
    if (bad) a.x = a.y = a.z = a.w = nanf("");
    const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (long long)(pos0 + s) * d + c));
    *reinterpret_cast<float4*>(out + r * d + c) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
  }
}

__global__ void __launch_bounds__(256)
embed_bwd_tok_kernel(const long long* __restrict__ idx, const float* __restrict__ dy, float* __restrict__ dtok, long long rows,
                     int d, int vocab) {
  const long long total = rows * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d;
    const int c = (int)(i - r * d);
    const long long t = idx[r];
    if (t >= 0 && t < vocab) atomicAdd(dtok + t * d + c, dy[i]);
  }
}

// dpos[s, :] = sum_b dy[b, s, :]   (deterministic: one thread per output element walks the batch)
__global__ void __launch_bounds__(256)
embed_bwd_pos_kernel(const float* __restrict__ dy, float* __restrict__ dpos, int B, int S, int d) {
  const int per_row = d >> 2;
  const long long total = (long long)S * per_row;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long off = i * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float4 v = *reinterpret_cast<const float4*>(dy + (long long)b * S * d + off);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(dpos + off) = acc;
}

}  // namespace b200

extern "C" {

int b200vit_embed_fwd(const long long* idx, const float* tok_embed, const float* pos_embed, float* out, int B, int S, int d,
                      int pos0, const int* pos0_dev, int n_pos, int vocab, void* stream) {
  B200_REQUIRE(idx && tok_embed && pos_embed && out && B > 0 && S > 0 && d > 0 && d % 4 == 0 && pos0 >= 0 && vocab > 0 &&
                   n_pos >= S && (pos0_dev != nullptr || pos0 + S <= n_pos),
               "embed_fwd: bad arguments (d must be a multiple of 4, positions within pos_embed)");
  const long long total = (long long)B * S * (d / 4);
  const long long want = (total + 255) / 256;
  const int cap = b200::num_sms() * 16;
  b200::embed_fwd_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(idx, tok_embed, pos_embed, out,
                                                                                         (long long)B * S, S, d, pos0, pos0_dev, n_pos,
                                                                                         vocab);
  B200_CUDA(cudaGetLastError());
  return b200::OK;
}

int b200vit_embed_bwd(const long long* idx, const float* dy, float* dtok /* [vocab, d], overwritten */, float* dpos /* [S, d] */,
                      int B, int S, int d, int vocab, void* stream) {
  B200_REQUIRE(idx && dy && dtok && dpos && B > 0 && S > 0 && d > 0 && d % 4 == 0 && vocab > 0, "embed_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(dtok, 0, sizeof(float) * (size_t)vocab * d, st));
  const long long total = (long long)B * S * d;
  const long long want = (total + 255) / 256;
  const int cap = b200::num_sms() * 16;
  b200::embed_bwd_tok_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(idx, dy, dtok, (long long)B * S, d, vocab);
  B200_CUDA(cudaGetLastError());
  const long long tp = (long long)S * (d / 4);
  b200::embed_bwd_pos_kernel<<<(int)((tp + 255) / 256), 256, 0, st>>>(dy, dpos, B, S, d);
  B200_CUDA(cudaGetLastError());
  return b200::OK;
}

}  // extern "C"
