// Folding an affine LayerNorm into the Linear that follows it (blocks.py:66,69: ln_1 -> attn.in_proj, ln_2 -> mlp.c_fc):
//
//     Linear(gamma * xhat + beta) = xhat (W diag(gamma))^T + (b + W beta)
//
// so the GEMM operand is the affine-free xhat -- the same bf16 tensor the affine-free blocks of transformer.py save for
// backward -- and blocks.ResidualAttentionBlock runs on exactly the kernels of transformer.TransformerLayer.  Per optimizer
// step (not per row of activations):
//     fold:    W'[n,k] = bf16(W[n,k] * gamma[k]),  b'[n] = b[n] + sum_k W[n,k] * beta[k]
//     unfold:  dW[n,k] = dW'[n,k] * gamma[k] + db'[n] * beta[k],  dgamma[k] = sum_n dW'[n,k] * W[n,k],  dbeta[k] = sum_n W[n,k] * db'[n]
// (dW' and db' come out of the ordinary wgrad GEMM on xhat).  All fp32 except the bf16 operand.  Weight-sized work: N*K
// elements, against M*K activations per GEMM with M >> N.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace b200 {

// one warp per output row n
__global__ void __launch_bounds__(256)
affine_fold_kernel(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma,
                   const float* __restrict__ beta, __nv_bfloat16* __restrict__ W16, float* __restrict__ bias_out, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  const float* w = W + (long long)n * K;
  __nv_bfloat16* o = W16 + (long long)n * K;
  float acc = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 v = *reinterpret_cast<const float4*>(w + k);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + k));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + k));
    acc = fmaf(v.x, b.x, acc); acc = fmaf(v.y, b.y, acc); acc = fmaf(v.z, b.z, acc); acc = fmaf(v.w, b.w, acc);
    uint2 u;
    u.x = pack_bf16(v.x * g.x, v.y * g.y); u.y = pack_bf16(v.z * g.z, v.w * g.w);
    *reinterpret_cast<uint2*>(o + k) = u;
  }
  acc = warp_sum(acc);
  if (lane == 0) bias_out[n] = (bias != nullptr ? bias[n] : 0.f) + acc;
}

// grid (K / 32, splits): each CTA owns 32 columns x a slab of rows; 8 row groups x 32 columns per CTA; scales dW in place
// and writes its partial column sums to part[split][2][K]; unfold_finish adds the partials in a fixed order.
constexpr int UNF_ROWGROUPS = 8;
__global__ void __launch_bounds__(32 * UNF_ROWGROUPS)
affine_unfold_kernel(float* __restrict__ dW, const float* __restrict__ W, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ dbias, float* __restrict__ part, int N, int K,
                     int rows_per_cta) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + tx;
  const int n0 = blockIdx.y * rows_per_cta, n1 = min(N, n0 + rows_per_cta);
  float sg = 0.f, sb = 0.f;
  if (k < K) {
    const float g = __ldg(gamma + k), be = __ldg(beta + k);
    for (int n = n0 + ty; n < n1; n += UNF_ROWGROUPS) {
      const long long i = (long long)n * K + k;
      const float dw = dW[i], w = W[i], dbn = __ldg(dbias + n);
      sg = fmaf(dw, w, sg);
      sb = fmaf(w, dbn, sb);
      dW[i] = fmaf(dbn, be, dw * g);      // through W' = W diag(gamma) and through b' = b + W beta
    }
  }
  __shared__ float red[2][UNF_ROWGROUPS][32];
  red[0][ty][tx] = sg; red[1][ty][tx] = sb;
  __syncthreads();
  if (ty < 2 && k < K) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < UNF_ROWGROUPS; ++i) t += red[ty][i][tx];
    part[((long long)blockIdx.y * 2 + ty) * K + k] = t;
  }
}

__global__ void __launch_bounds__(256)
affine_unfold_finish_kernel(const float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta, int K,
                            int splits, int accumulate) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= K) return;
  float g = 0.f, b = 0.f;
  for (int s = 0; s < splits; ++s) { g += part[((long long)s * 2 + 0) * K + k]; b += part[((long long)s * 2 + 1) * K + k]; }
  if (accumulate) { dgamma[k] += g; dbeta[k] += b; }
  else { dgamma[k] = g; dbeta[k] = b; }
}

constexpr int UNF_SPLITS = 16;

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_affine_fold(const float* W, const float* bias, const float* gamma, const float* beta, void* W_bf16,
                        float* bias_out, int N, int K, void* stream) {
  B200_REQUIRE(W && gamma && beta && W_bf16 && bias_out && N > 0 && K > 0 && K % 4 == 0,
               "affine_fold: bad arguments (K must be a multiple of 4)");
  affine_fold_kernel<<<(N + 7) / 8, 256, 0, (cudaStream_t)stream>>>(W, bias, gamma, beta, (__nv_bfloat16*)W_bf16, bias_out, N, K);
  B200_CUDA(cudaGetLastError());
  return OK;
}

size_t b200vit_affine_unfold_workspace_size(int K) { return sizeof(float) * 2 * (size_t)UNF_SPLITS * (size_t)(K > 0 ? K : 0); }

int b200vit_affine_unfold_grads(float* dW, const float* W, const float* gamma, const float* beta, const float* dbias,
                                float* dgamma, float* dbeta, int N, int K, int accumulate, void* workspace,
                                size_t workspace_bytes, void* stream) {
  B200_REQUIRE(dW && W && gamma && beta && dbias && dgamma && dbeta && N > 0 && K > 0, "affine_unfold_grads: bad arguments");
  B200_REQUIRE(workspace && workspace_bytes >= b200vit_affine_unfold_workspace_size(K), "affine_unfold_grads: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int splits = UNF_SPLITS;
  int rows_per_cta = (N + splits - 1) / splits;
  if (rows_per_cta < UNF_ROWGROUPS) rows_per_cta = UNF_ROWGROUPS;
  splits = (N + rows_per_cta - 1) / rows_per_cta;
  affine_unfold_kernel<<<dim3((K + 31) / 32, splits), 32 * UNF_ROWGROUPS, 0, st>>>(dW, W, gamma, beta, dbias, (float*)workspace, N, K,
                                                                                 rows_per_cta);
  B200_CUDA(cudaGetLastError());
  affine_unfold_finish_kernel<<<(K + 255) / 256, 256, 0, st>>>((const float*)workspace, dgamma, dbeta, K, splits, accumulate);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"
