// Counter-based dropout masks shared by the GEMM epilogue, the element-wise kernels and the attention kernels.
// A mask bit is a pure function of (seed, coordinates), so forward and backward regenerate it instead of storing
// it (nn.Dropout after mlp[2], transformer.py:40; dropout_p of F.scaled_dot_product_attention, transformer.py:28).
// One 32-bit hash serves TWO neighbouring elements (16 bits each): keep <=> bits >= round(p * 65536).
#pragma once
#include <stdint.h>

namespace b200 {

__host__ __device__ __forceinline__ uint32_t drop_mix32(uint32_t h) {  // murmur3 finaliser
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
// [M, d] activations: pairs along the column index
__host__ __device__ __forceinline__ uint32_t drop_hash_rows(uint32_t seed, uint32_t row, uint32_t colpair) {
  return drop_mix32(seed ^ (row * 0x9E3779B9u + colpair * 0x85EBCA77u));
}
// attention probabilities of (batch*heads + head, query, key): pairs along the key index
__host__ __device__ __forceinline__ uint32_t drop_hash_attn(uint32_t seed, uint32_t bh, uint32_t q, uint32_t kpair) {
  return drop_mix32((seed ^ (bh * 0x9E3779B9u)) + q * 0x85EBCA77u + kpair * 0xC2B2AE3Du);
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t h, uint32_t odd, uint32_t thr) {
  return (odd ? (h >> 16) : (h & 0xffffu)) >= thr;
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  const float t = p * 65536.0f + 0.5f;
  return t < 0.f ? 0u : (t > 65535.f ? 65535u : (uint32_t)t);
}

}  // namespace b200
