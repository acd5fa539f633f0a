// Patch embedding: nn.Conv2d(C, d, kernel = stride = p) + "b c h w -> b (h w) c" + pos_emb add + prepended
// extra tokens (train_vit.py:34-36,38-45; blocks.py:235-237,257-267) as ONE tcgen05 GEMM whose epilogue adds
// bias + positional embedding and writes straight into rows [extra, extra+P) of the [B, T, d] fp32 token
// buffer.  The fp32 NCHW image has to be converted to bf16 anyway, so the im2col gather is fused with that cast.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace b200 {

int gemm_patch_epilogue(const void* xcol, const void* w, const float* bias, const float* pos, float* out,
                        int rows, int N, int K, int P, int T, int extra, cudaStream_t st);

int gemm_depatch_epilogue(const void* rows_bf16, const void* w, const float* bias, float* img, int rows, int N, int K,
                          int P, int Wt, int log2p, cudaStream_t st);

// cols[(b, ph, pw), (c, i, j)] = x[b, c, ph*p + i, pw*p + j]   (fp32 -> bf16), 4 pixels per thread
__global__ void im2col_vec4_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ cols, int B, int C,
                                   int H, int W, int p) {
  const int gw = W / p, gh = H / p;
  const int K = C * p * p;
  const long long total = (long long)B * gh * gw * (K / 4);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int k4 = (int)(t % (K / 4));
    const long long row = t / (K / 4);
    const int k = k4 * 4;
    const int j = k % p, i = (k / p) % p, c = k / (p * p);
    const int pw = (int)(row % gw), ph = (int)((row / gw) % gh);
    const long long b = row / ((long long)gw * gh);
    const float4 v = *reinterpret_cast<const float4*>(x + ((b * C + c) * H + (ph * p + i)) * (long long)W + pw * p + j);
    uint2 w;
    w.x = pack_bf16(v.x, v.y); w.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(cols + row * K + k) = w;
  }
}

__global__ void im2col_scalar_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ cols, int B, int C,
                                     int H, int W, int p) {
  const int gw = W / p, gh = H / p;
  const int K = C * p * p;
  const long long total = (long long)B * gh * gw * K;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(t % K);
    const long long row = t / K;
    const int j = k % p, i = (k / p) % p, c = k / (p * p);
    const int pw = (int)(row % gw), ph = (int)((row / gw) % gh);
    const long long b = row / ((long long)gw * gh);
    cols[t] = __float2bfloat16_rn(x[((b * C + c) * H + (ph * p + i)) * (long long)W + pw * p + j]);
  }
}

// dx[b, c, ph*p+i, pw*p+j] = dcols[(b,ph,pw), (c,i,j)]   (bf16 -> fp32); patches do not overlap -> pure permutation
__global__ void col2im_kernel(const __nv_bfloat16* __restrict__ dcols, float* __restrict__ dx, int B, int C, int H,
                              int W, int p) {
  const int gw = W / p, gh = H / p;
  const int K = C * p * p;
  const long long total = (long long)B * C * H * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int xw = (int)(t % W), yh = (int)((t / W) % H), c = (int)((t / ((long long)W * H)) % C);
    const long long b = t / ((long long)W * H * C);
    const int pw = xw / p, j = xw % p, ph = yh / p, i = yh % p;
    const long long row = (b * gh + ph) * gw + pw;
    dx[t] = __bfloat162float(dcols[row * K + (c * p + i) * p + j]);
  }
}

// out[b, e, :] = extra_emb[e, :] for the prepended tokens
__global__ void broadcast_extra_kernel(const float* __restrict__ extra_emb, float* __restrict__ out, int B, int T,
                                       int extra, int d) {
  const long long total = (long long)B * extra * (d / 4);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(t % (d / 4));
    const int e = (int)((t / (d / 4)) % extra);
    const long long b = t / ((long long)(d / 4) * extra);
    reinterpret_cast<float4*>(out + (b * T + e) * d)[c4] = __ldg(reinterpret_cast<const float4*>(extra_emb + (long long)e * d) + c4);
  }
}

// Rows of the [B, T, d] token buffer that do not come out of the GEMM (blocks.py:261-267, 346-352):
//   head rows e in [0, extra):      out[b, e]             = (e == 0 ? head0 : head1)[0, :] + head_pos[e, :]
//                                   (class_embedding / mask_token + positional_embedding rows)
//   tail rows l in [0, tail):       out[b, extra + P + l] = tail_a[l, :] + tail_b[l, :]   (latent_tokens + their positions)
// head_pos / tail_b may be null.  (train_vit.ViT's extra_emb rows keep broadcast_extra_kernel above.)
__global__ void assemble_rows_kernel(const float* __restrict__ head0, const float* __restrict__ head1,
                                     const float* __restrict__ head_pos, const float* __restrict__ tail_a,
                                     const float* __restrict__ tail_b, float* __restrict__ out, int B, int T, int extra,
                                     int P, int tail, int d) {
  const int nvec = d / 4, per = extra + tail;
  const long long total = (long long)B * per * nvec;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(t % nvec);
    const int e = (int)((t / nvec) % per);
    const long long b = t / ((long long)nvec * per);
    float4 v, w = make_float4(0.f, 0.f, 0.f, 0.f);
    int row;
    if (e < extra) {
      v = __ldg(reinterpret_cast<const float4*>(e == 0 ? head0 : head1) + c4);
      if (head_pos != nullptr) w = __ldg(reinterpret_cast<const float4*>(head_pos + (long long)e * d) + c4);
      row = e;
    } else {
      const int l = e - extra;
      v = __ldg(reinterpret_cast<const float4*>(tail_a + (long long)l * d) + c4);
      if (tail_b != nullptr) w = __ldg(reinterpret_cast<const float4*>(tail_b + (long long)l * d) + c4);
      row = extra + P + l;
    }
    reinterpret_cast<float4*>(out + (b * T + row) * d)[c4] = make_float4(v.x + w.x, v.y + w.y, v.z + w.z, v.w + w.w);
  }
}

// out[b, j, :] = x[b, t0 + j, :] (fp32 copy of a token range: the ln_post input of the blocks.py encoder / decoder)
__global__ void gather_rows_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int B, long long in_stride,
                                       long long out_stride, long long in_off) {
  const long long per = out_stride / 4;
  const long long total = (long long)B * per;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long b = t / per, c = t - b * per;
    reinterpret_cast<float4*>(out + b * out_stride)[c] = __ldg(reinterpret_cast<const float4*>(x + b * in_stride + in_off) + c);
  }
}

// Backward helpers on the fp32 token gradient dtok[B, T, d]:
//   dsum[t, :]  = sum_b dtok[b, t, :]          (-> extra_emb.grad rows [0,extra), pos_emb.grad rows [extra,T))
//   dpe[(b,p),:] = bf16(dtok[b, extra + p, :]) (compact operand of the conv wgrad GEMM)
// (P = rows of each image that came out of the GEMM: T - extra for train_vit.ViT, T - extra - tail for the blocks.py
// encoder / decoder whose latent tokens follow the patches)
__global__ void patch_bwd_reduce_kernel(const float* __restrict__ dtok, float* __restrict__ dsum,
                                        __nv_bfloat16* __restrict__ dpe, int B, int T, int extra, int P, int d) {
  const int nvec = d / 4;
  const long long total = (long long)T * nvec;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(t % nvec);
    const int tok = (int)(t / nvec);
    float4 acc = make_float4(0, 0, 0, 0);
    for (int b = 0; b < B; ++b) {
      const float4 v = reinterpret_cast<const float4*>(dtok + ((long long)b * T + tok) * d)[c4];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      if (dpe != nullptr && tok >= extra && tok < extra + P) {
        uint2 w;
        w.x = pack_bf16(v.x, v.y); w.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(dpe + ((long long)b * P + (tok - extra)) * d + c4 * 4) = w;
      }
    }
    reinterpret_cast<float4*>(dsum + (long long)tok * d)[c4] = acc;
  }
}

__global__ void colsum_f32_kernel(const float* __restrict__ a, float* __restrict__ out, int rows, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += a[(long long)r * n + c];
  out[c] = s;
}

// n = 4 * n4 + tail elements: vector body plus a scalar tail (any length; the base pointers are 16-byte aligned)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4, int tail) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[i];
    uint2 w;
    w.x = pack_bf16(v.x, v.y); w.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = w;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) out[n4 * 4 + threadIdx.x] = __float2bfloat16_rn(in[n4 * 4 + threadIdx.x]);
}

// out(f32) = scale * in(bf16): gradients coming back from a bf16-compressed all-reduce (ddp.py)
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n4, int tail,
                                     float scale) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint2 u = reinterpret_cast<const uint2*>(in)[i];
    const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
    reinterpret_cast<float4*>(out)[i] = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail) out[n4 * 4 + threadIdx.x] = __bfloat162float(in[n4 * 4 + threadIdx.x]) * scale;
}

static int grid_for(long long work, int threads) {
  long long g = (work + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_cast_f32_bf16(const float* in, void* out, long long n, void* stream) {
  B200_REQUIRE(in && out && n > 0, "cast_f32_bf16: bad arguments (n=%lld)", n);
  B200_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 7) == 0, "cast_f32_bf16: pointers must be 16- / 8-byte aligned");
  cast_f32_bf16_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, n / 4, (int)(n % 4));
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_cast_bf16_f32(const void* in, float* out, long long n, float scale, void* stream) {
  B200_REQUIRE(in && out && n > 0, "cast_bf16_f32: bad arguments (n=%lld)", n);
  B200_REQUIRE(((uintptr_t)out & 15) == 0 && ((uintptr_t)in & 7) == 0, "cast_bf16_f32: pointers must be 16- / 8-byte aligned");
  cast_bf16_f32_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, out, n / 4, (int)(n % 4), scale);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_im2col_bf16(const float* x, void* cols, int B, int C, int H, int W, int p, void* stream) {
  B200_REQUIRE(x && cols && B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0, "im2col: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * C * H * W;
  if (p % 4 == 0) im2col_vec4_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(x, (__nv_bfloat16*)cols, B, C, H, W, p);
  else            im2col_scalar_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, (__nv_bfloat16*)cols, B, C, H, W, p);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_col2im_f32(const void* dcols, float* dx, int B, int C, int H, int W, int p, void* stream) {
  B200_REQUIRE(dcols && dx && B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0, "col2im: bad arguments");
  const long long n = (long long)B * C * H * W;
  col2im_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dcols, dx, B, C, H, W, p);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_patch_embed_fwd(const float* x, const void* w_bf16, const float* bias, const float* pos_emb,
                            const float* extra_emb, float* tokens, void* cols, int B, int C, int H, int W, int p,
                            int d, int extra, void* stream) {
  B200_REQUIRE(x && w_bf16 && pos_emb && tokens && cols, "patch_embed_fwd: null pointer");
  B200_REQUIRE(extra == 0 || extra_emb != nullptr, "patch_embed_fwd: extra_emb is null");
  B200_REQUIRE(d % 8 == 0 && (C * p * p) % 8 == 0, "patch_embed_fwd: d=%d and C*p*p=%d must be multiples of 8", d, C * p * p);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = b200vit_im2col_bf16(x, cols, B, C, H, W, p, stream);
  if (rc != OK) return rc;
  const int P = (H / p) * (W / p), T = P + extra, K = C * p * p;
  if (extra > 0) {
    broadcast_extra_kernel<<<grid_for((long long)B * extra * (d / 4), 256), 256, 0, st>>>(extra_emb, tokens, B, T, extra, d);
    B200_CUDA(cudaGetLastError());
  }
  return gemm_patch_epilogue(cols, w_bf16, bias, pos_emb, tokens, B * P, d, K, P, T, extra, st);
}

int b200vit_depatchify_fwd(const void* rows_bf16, const void* w_cmajor_bf16, const float* bias_cmajor, float* img, int B,
                           int Ht, int Wt, int p, int C, int d, void* stream) {
  B200_REQUIRE(rows_bf16 && w_cmajor_bf16 && img && B > 0 && Ht > 0 && Wt > 0 && C > 0, "depatchify_fwd: bad arguments");
  int log2p = 0;
  while ((1 << log2p) < p) ++log2p;
  B200_REQUIRE(p >= 4 && (1 << log2p) == p, "depatchify_fwd: patch size %d must be a power of two >= 4", p);
  B200_REQUIRE(d % 8 == 0, "depatchify_fwd: d=%d must be a multiple of 8", d);
  return gemm_depatch_epilogue(rows_bf16, w_cmajor_bf16, bias_cmajor, img, B * Ht * Wt, C * p * p, d, Ht * Wt, Wt, log2p,
                               (cudaStream_t)stream);
}

int b200vit_patch_embed_bwd_reduce(const float* dtokens, float* dsum, void* dpe_bf16, int B, int T, int extra, int d,
                                   void* stream) {
  B200_REQUIRE(dtokens && dsum && d % 4 == 0 && T > extra, "patch_embed_bwd_reduce: bad arguments");
  patch_bwd_reduce_kernel<<<grid_for((long long)T * (d / 4), 128), 128, 0, (cudaStream_t)stream>>>(
      dtokens, dsum, (__nv_bfloat16*)dpe_bf16, B, T, extra, T - extra, d);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_tokens_assemble_fwd(const void* cols_bf16, const void* w_bf16, const float* bias, const float* pos,
                                const float* head0, const float* head1, const float* head_pos, const float* tail_a,
                                const float* tail_b, float* tokens, int B, int P, int K, int d, int extra, int tail,
                                void* stream) {
  B200_REQUIRE(cols_bf16 && w_bf16 && pos && tokens && B > 0 && P > 0, "tokens_assemble_fwd: bad arguments");
  B200_REQUIRE(d % 8 == 0 && K % 8 == 0, "tokens_assemble_fwd: d=%d and K=%d must be multiples of 8", d, K);
  B200_REQUIRE(extra >= 0 && tail >= 0 && (extra == 0 || head0) && (extra <= 1 || head1) && (tail == 0 || tail_a),
               "tokens_assemble_fwd: head / tail rows requested without their sources");
  cudaStream_t st = (cudaStream_t)stream;
  const int T = extra + P + tail;
  if (extra + tail > 0) {
    assemble_rows_kernel<<<grid_for((long long)B * (extra + tail) * (d / 4), 256), 256, 0, st>>>(
        head0, head1, head_pos, tail_a, tail_b, tokens, B, T, extra, P, tail, d);
    B200_CUDA(cudaGetLastError());
  }
  return gemm_patch_epilogue(cols_bf16, w_bf16, bias, pos, tokens, B * P, d, K, P, T, extra, st);
}

int b200vit_tokens_assemble_bwd_reduce(const float* dtokens, float* dsum, void* dpe_bf16, int B, int T, int extra, int tail,
                                       int d, void* stream) {
  B200_REQUIRE(dtokens && dsum && d % 4 == 0 && extra >= 0 && tail >= 0 && T > extra + tail,
               "tokens_assemble_bwd_reduce: bad arguments");
  patch_bwd_reduce_kernel<<<grid_for((long long)T * (d / 4), 128), 128, 0, (cudaStream_t)stream>>>(
      dtokens, dsum, (__nv_bfloat16*)dpe_bf16, B, T, extra, T - extra - tail, d);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_gather_tokens_f32(const float* x, float* out, int B, int N, int d, int t0, int cnt, void* stream) {
  B200_REQUIRE(x && out && B > 0 && N > 0 && d > 0 && d % 4 == 0 && t0 >= 0 && cnt > 0 && t0 + cnt <= N,
               "gather_tokens_f32: bad arguments (d must be a multiple of 4, 0 <= t0, t0 + cnt <= N)");
  gather_rows_f32_kernel<<<grid_for((long long)B * cnt * (d / 4), 256), 256, 0, (cudaStream_t)stream>>>(
      x, out, B, (long long)N * d, (long long)cnt * d, (long long)t0 * d);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_colsum_f32(const float* a, float* out, int rows, int n, void* stream) {
  B200_REQUIRE(a && out && rows > 0 && n > 0, "colsum_f32: bad arguments");
  colsum_f32_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, out, rows, n);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"
