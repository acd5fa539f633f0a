// Fused multi-tensor AdamW (decoupled weight decay) + bf16 operand refresh, one launch for a whole parameter group.
// Replaces torch.optim.AdamW at train_vit.py:82,105 / train_titok.py:134,160 / train_videogpt.py:107,134.
//
// HBM-bound streaming kernel: per parameter element it reads p, g, m, v (16 B) and writes p, m, v (12 B) plus, for
// the GEMM weights, the bf16 copy the next forward consumes (2 B) -- the separate fp32 -> bf16 cast pass over the
// weights disappears.  A device table describes the tensors (pointers + sizes) and a chunk list maps every CTA to
// 4096 consecutive elements of one tensor, so ~150 tensors of very different sizes are processed by one grid of
// equal-work CTAs (every load of a chunk is issued before the first use).
//
// Arithmetic follows torch/optim/adam.py::_single_tensor_adam with decoupled_weight_decay (torch 2.11), in fp32 and
// in its operation order:
//     p  *= 1 - lr * wd
//     m   = m + (g - m) * (1 - beta1)                       (Tensor.lerp_)
//     v   = v * beta2 + (1 - beta2) * g * g                 (mul_ + addcmul_)
//     p  -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)    (addcdiv_),  bc_i = 1 - beta_i^step
// GradScaler support (the reference scripts call scaler.step(optim)): g is divided by *grad_scale when given, and the
// whole step -- including the step counter -- is skipped when *found_inf != 0.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace b200 {

struct AdamWTensor {   // mirrored by b200vit/optim.py (6 x 8 bytes)
  float* p; const float* g; float* m; float* v; __nv_bfloat16* w16; long long n;
};
struct AdamWChunk { int tensor; int index; };  // chunk `index` (of ADAMW_CHUNK elements) of tensor `tensor`

constexpr int ADAMW_THREADS = 256;
constexpr int ADAMW_VEC = 4;                                     // float4 per access
constexpr int ADAMW_ILP = 4;                                     // accesses per thread
constexpr int ADAMW_CHUNK = ADAMW_THREADS * ADAMW_VEC * ADAMW_ILP;  // 4096 elements per CTA

struct AdamWArgs {
  const AdamWTensor* tensors; const AdamWChunk* chunks;
  float decay;        // 1 - lr * wd   (computed in double on the host, like the Python scalar in torch)
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps;
  double lr, beta1_d, beta2_d;   // for the device-side bias corrections (device step counter)
  float step_size, rsqrt_bc2_inv;  // host-side step: lr / bc1 and sqrt(bc2); ignored when step_dev != nullptr
  const float* step_dev;           // device step counter (already incremented for this step) or nullptr
  const float* grad_scale; const float* found_inf;
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, const AdamWArgs& a, float step_size,
                                           float sqrt_bc2, float inv_scale) {
  g *= inv_scale;
  p *= a.decay;
  m = fmaf(g - m, a.one_minus_beta1, m);
  v = fmaf(a.one_minus_beta2 * g, g, v * a.beta2);
  const float denom = __fsqrt_rn(v) / sqrt_bc2 + a.eps;
  p = fmaf(-step_size, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(ADAMW_THREADS) adamw_kernel(const AdamWArgs a) {
  if (a.found_inf != nullptr && __ldg(a.found_inf) != 0.f) return;   // GradScaler: skip the step
  const AdamWChunk ck = a.chunks[blockIdx.x];
  const AdamWTensor t = a.tensors[ck.tensor];
  float step_size = a.step_size, sqrt_bc2 = a.rsqrt_bc2_inv;
  if (a.step_dev != nullptr) {
    __shared__ float sh[2];
    if (threadIdx.x == 0) {
      const double s = (double)__ldg(a.step_dev);
      sh[0] = (float)(a.lr / (1.0 - pow(a.beta1_d, s)));
      sh[1] = (float)sqrt(1.0 - pow(a.beta2_d, s));
    }
    __syncthreads();
    step_size = sh[0]; sqrt_bc2 = sh[1];
  }
  const float inv_scale = a.grad_scale != nullptr ? 1.0f / __ldg(a.grad_scale) : 1.0f;
  const long long base = (long long)ck.index * ADAMW_CHUNK;
  const long long left = t.n - base;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                        reinterpret_cast<uintptr_t>(t.v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(t.w16) & 7) == 0;
  if (vec_ok && left >= ADAMW_CHUNK) {
    float4 p[ADAMW_ILP], g[ADAMW_ILP], m[ADAMW_ILP], v[ADAMW_ILP];
#pragma unroll
    for (int i = 0; i < ADAMW_ILP; ++i) {
      const long long o = base + (long long)(i * ADAMW_THREADS + threadIdx.x) * ADAMW_VEC;
      p[i] = ld4(t.p + o); g[i] = ld4(t.g + o); m[i] = ld4(t.m + o); v[i] = ld4(t.v + o);
    }
#pragma unroll
    for (int i = 0; i < ADAMW_ILP; ++i) {
      const long long o = base + (long long)(i * ADAMW_THREADS + threadIdx.x) * ADAMW_VEC;
      adamw_elem(p[i].x, g[i].x, m[i].x, v[i].x, a, step_size, sqrt_bc2, inv_scale);
      adamw_elem(p[i].y, g[i].y, m[i].y, v[i].y, a, step_size, sqrt_bc2, inv_scale);
      adamw_elem(p[i].z, g[i].z, m[i].z, v[i].z, a, step_size, sqrt_bc2, inv_scale);
      adamw_elem(p[i].w, g[i].w, m[i].w, v[i].w, a, step_size, sqrt_bc2, inv_scale);
      *reinterpret_cast<float4*>(t.p + o) = p[i];
      *reinterpret_cast<float4*>(t.m + o) = m[i];
      *reinterpret_cast<float4*>(t.v + o) = v[i];
      if (t.w16 != nullptr) {
        uint2 w;
        w.x = pack_bf16(p[i].x, p[i].y); w.y = pack_bf16(p[i].z, p[i].w);
        *reinterpret_cast<uint2*>(t.w16 + o) = w;
      }
    }
  } else {  // ragged tail chunk or unaligned tensor: scalar accesses
    const long long end = left < ADAMW_CHUNK ? t.n : base + ADAMW_CHUNK;
    for (long long o = base + threadIdx.x; o < end; o += ADAMW_THREADS) {
      float p = t.p[o], m = t.m[o], v = t.v[o];
      adamw_elem(p, t.g[o], m, v, a, step_size, sqrt_bc2, inv_scale);
      t.p[o] = p; t.m[o] = m; t.v[o] = v;
      if (t.w16 != nullptr) t.w16[o] = __float2bfloat16_rn(p);
    }
  }
}

// step += 1 unless the GradScaler found an inf (torch: _foreach_add_(steps, 1) ... _foreach_sub_(steps, found_inf))
__global__ void adamw_advance_step_kernel(float* step, const float* found_inf) {
  if (found_inf == nullptr || *found_inf == 0.f) *step += 1.0f;
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_adamw_chunk_elems(void) { return ADAMW_CHUNK; }

int b200vit_adamw_step(const void* tensors, const void* chunks, int n_chunks, double lr, double beta1, double beta2,
                       double eps, double weight_decay, long long step, float* step_dev, const float* grad_scale,
                       const float* found_inf, void* stream) {
  B200_REQUIRE(tensors && chunks && n_chunks > 0, "adamw_step: empty tensor / chunk table");
  B200_REQUIRE(step_dev != nullptr || step >= 1, "adamw_step: step must be >= 1 (or pass a device step counter)");
  B200_REQUIRE(found_inf == nullptr || step_dev != nullptr,
               "adamw_step: found_inf needs the device step counter (a skipped step must not advance it)");
  cudaStream_t st = (cudaStream_t)stream;
  AdamWArgs a;
  a.tensors = (const AdamWTensor*)tensors;
  a.chunks = (const AdamWChunk*)chunks;
  // hyper-parameters arrive as doubles (Python floats) and are combined in double before the single rounding to
  // fp32, like the Python scalars of torch/optim/adam.py: (float)(1 - (double)0.999f) would be off by 1.3e-5 relative
  a.decay = (float)(1.0 - lr * weight_decay);
  a.beta1 = (float)beta1; a.beta2 = (float)beta2;
  a.one_minus_beta1 = (float)(1.0 - beta1);
  a.one_minus_beta2 = (float)(1.0 - beta2);
  a.eps = (float)eps; a.lr = lr; a.beta1_d = beta1; a.beta2_d = beta2;
  a.step_size = 0.f; a.rsqrt_bc2_inv = 1.f;
  if (step_dev == nullptr) {
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    a.step_size = (float)(lr / bc1);
    a.rsqrt_bc2_inv = (float)sqrt(bc2);
  } else {
    adamw_advance_step_kernel<<<1, 1, 0, st>>>(step_dev, found_inf);
    B200_CUDA(cudaGetLastError());
  }
  a.step_dev = step_dev; a.grad_scale = grad_scale; a.found_inf = found_inf;
  launch_kernel(adamw_kernel, dim3(n_chunks), dim3(ADAMW_THREADS), 0, st, 1, a);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"
