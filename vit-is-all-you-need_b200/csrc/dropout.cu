// Element-wise dropout helpers (HBM-bound, 16-byte accesses) and the mask dumps used by the parity tests.
#include "../../include/b200vit.h"
#include "common.cuh"
#include "dropout.cuh"

namespace b200 {

// out(bf16) = x(f32) * keep / (1 - p): the gradient of nn.Dropout applied while casting for the next GEMM
__global__ void __launch_bounds__(256) dropout_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           long long M, int d, uint32_t seed, uint32_t thr, float r) {
  const int vec_per_row = d >> 2;
  const long long nvec = M * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vec_per_row;
    const int c = (int)(i - row * vec_per_row) * 4;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const uint32_t h0 = drop_hash_rows(seed, (uint32_t)row, (uint32_t)(c >> 1));
    const uint32_t h1 = drop_hash_rows(seed, (uint32_t)row, (uint32_t)(c >> 1) + 1);
    uint2 w;
    w.x = pack_bf16(drop_keep(h0, 0, thr) ? v.x * r : 0.f, drop_keep(h0, 1, thr) ? v.y * r : 0.f);
    w.y = pack_bf16(drop_keep(h1, 0, thr) ? v.z * r : 0.f, drop_keep(h1, 1, thr) ? v.w * r : 0.f);
    reinterpret_cast<uint2*>(out)[i] = w;
  }
}

__global__ void dropout_mask_rows_kernel(uint8_t* out, long long M, int d, uint32_t seed, uint32_t thr) {
  const long long n = M * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / d;
    const int c = (int)(i - row * d);
    out[i] = drop_keep(drop_hash_rows(seed, (uint32_t)row, (uint32_t)(c >> 1)), c & 1, thr) ? 1 : 0;
  }
}

__global__ void dropout_mask_attn_kernel(uint8_t* out, int BH, int N, uint32_t seed, uint32_t thr) {
  const long long n = (long long)BH * N * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % N);
    const int q = (int)((i / N) % N);
    const int bh = (int)(i / ((long long)N * N));
    out[i] = drop_keep(drop_hash_attn(seed, (uint32_t)bh, (uint32_t)q, (uint32_t)(k >> 1)), k & 1, thr) ? 1 : 0;
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_dropout_cast_bf16(const float* x, void* out_bf16, long long M, int d, float p, unsigned int seed,
                              void* stream) {
  B200_REQUIRE(x && out_bf16 && M > 0 && d > 0 && d % 4 == 0, "dropout_cast: bad arguments (d must be a multiple of 4)");
  B200_REQUIRE(p >= 0.f && p < 1.f, "dropout_cast: p must be in [0, 1)");
  dropout_cast_kernel<<<num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, M, d, seed,
                                                                        drop_threshold(p), 1.0f / (1.0f - p));
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_dropout_mask_rows(unsigned char* out, long long M, int d, float p, unsigned int seed, void* stream) {
  B200_REQUIRE(out && M > 0 && d > 0, "dropout_mask_rows: bad arguments");
  dropout_mask_rows_kernel<<<num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(out, M, d, seed, drop_threshold(p));
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_dropout_mask_attn(unsigned char* out, int B, int H, int N, float p, unsigned int seed, void* stream) {
  B200_REQUIRE(out && B > 0 && H > 0 && N > 0, "dropout_mask_attn: bad arguments");
  dropout_mask_attn_kernel<<<num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(out, B * H, N, seed, drop_threshold(p));
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"
