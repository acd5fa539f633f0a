// Skinny forward GEMM for the single-token decode step (M = batch <= 64 rows): y[M, N] = x[M, K] w[N, K]^T + bias.
//
// At M = 16 the tcgen05 tile kernel (gemm.cu) would keep 9 of 148 SMs busy (N / 256 tiles of which 7/8 is padding) and
// run at the latency of one CTA streaming its weight slab; the operation is weight-bandwidth-bound (VideoGPT-B: 170 MB
// of bf16 weights per generated token, 2.7 GFLOP), so this kernel is organised around streaming w once at full
// memory-level parallelism instead:
//   * one CTA per 8 output columns (N / 8 CTAs: 288 for the QKV projection), 4 (K >= 2048: 8) warps splitting K, each
//     lane issuing 16-byte loads of its weight row 4 deep and one iteration ahead, partial sums reduced through smem;
//   * mma.sync.m16n8k16 bf16 (M = 16 rows is exactly one MMA tile; up to 4 row tiles re-use the weight fragment),
//     with the K index permuted identically for both operands so that every operand fragment is one 16-byte load
//     (the dot product does not care in which order k is visited);
//   * fused epilogues of the decode step: bias -> bf16 (QKV), bias + GELU -> bf16 (+ GELU' if asked), bias + residual
//     -> fp32 (mlp[2]), bias -> fp32 (vocabulary projection).
// This is the one place where mma.sync is the right tool: there is a single 16-row tile, the tensor pipe is idle
// either way, and what is measured is HBM / L2 bandwidth.
#include "../../include/b200vit.h"
#include "common.cuh"
#include "gemm_tcgen05.cuh"

namespace b200 {

constexpr int SK_KCHUNK = 32;           // k values consumed per lane-row per step (4 lanes x 8 bf16)
constexpr int SK_UNROLL = 4;            // chunks per loop iteration: 4 independent 16-byte weight loads per lane

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {   // weights are read once per launch: keep them out of L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

struct SkinnyArgs {
  const __nv_bfloat16* x; const __nv_bfloat16* w; const float* bias;
  void* out; void* out2; const float* resid;
  int M, N, K;
};

template <int MT, int KIND, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) gemm_skinny_kernel(const SkinnyArgs a) {
  constexpr int SK_WARPS = WARPS;
  __shared__ float red[SK_WARPS][MT][32][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 8;
  const int kq = a.K / SK_WARPS;                 // this warp's K range
  const int k_begin = warp * kq, k_end = k_begin + kq;
  const int n = n0 + g;
  const bool n_ok = n < a.N;
  const __nv_bfloat16* wrow = a.w + (long long)(n_ok ? n : 0) * a.K + t * 8;
  float acc[MT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m) { acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f; }

  // software pipeline: the weight loads of iteration i + 1 are in flight while iteration i multiplies (a warp has only
  // K / WARPS / 128 iterations, each a full DRAM round trip if exposed)
  constexpr int STEP = SK_KCHUNK * SK_UNROLL;
  auto load_w = [&](int k, uint4 (&dst)[SK_UNROLL]) {
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) {
      const int kk = k + u * SK_KCHUNK;
      dst[u] = (n_ok && kk < k_end) ? ld_nc_v4(wrow + kk) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 wv[SK_UNROLL], wn[SK_UNROLL];
  load_w(k_begin, wv);
  for (int k = k_begin; k < k_end; k += STEP) {
    if (k + STEP < k_end) load_w(k + STEP, wn);
    if constexpr (MT == 1) {
      uint4 x0[SK_UNROLL], x1[SK_UNROLL];   // all activation fragments of the iteration before the first MMA
#pragma unroll
      for (int u = 0; u < SK_UNROLL; ++u) {
        const int kk = k + u * SK_KCHUNK;
        const bool ok = kk < k_end;
        x0[u] = (ok && g < a.M) ? __ldg(reinterpret_cast<const uint4*>(a.x + (long long)g * a.K + kk + t * 8)) : make_uint4(0, 0, 0, 0);
        x1[u] = (ok && g + 8 < a.M) ? __ldg(reinterpret_cast<const uint4*>(a.x + (long long)(g + 8) * a.K + kk + t * 8)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < SK_UNROLL; ++u) {
        const uint32_t b0[2] = {wv[u].x, wv[u].y}, b1[2] = {wv[u].z, wv[u].w};
        const uint32_t a0[4] = {x0[u].x, x1[u].x, x0[u].y, x1[u].y}, a1[4] = {x0[u].z, x1[u].z, x0[u].w, x1[u].w};
        mma_bf16_16816(acc[0], a0, b0);
        mma_bf16_16816(acc[0], a1, b1);
      }
    } else {
#pragma unroll
      for (int u = 0; u < SK_UNROLL; ++u) {
        const int kk = k + u * SK_KCHUNK;
        if (kk >= k_end) break;       // warp-uniform
        const uint32_t b0[2] = {wv[u].x, wv[u].y}, b1[2] = {wv[u].z, wv[u].w};
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const int r0 = m * 16 + g, r1 = r0 + 8;
          const uint4 x0 = r0 < a.M ? __ldg(reinterpret_cast<const uint4*>(a.x + (long long)r0 * a.K + kk + t * 8)) : make_uint4(0, 0, 0, 0);
          const uint4 x1 = r1 < a.M ? __ldg(reinterpret_cast<const uint4*>(a.x + (long long)r1 * a.K + kk + t * 8)) : make_uint4(0, 0, 0, 0);
          const uint32_t a0[4] = {x0.x, x1.x, x0.y, x1.y}, a1[4] = {x0.z, x1.z, x0.w, x1.w};
          mma_bf16_16816(acc[m], a0, b0);
          mma_bf16_16816(acc[m], a1, b1);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < SK_UNROLL; ++u) wv[u] = wn[u];
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
#pragma unroll
    for (int j = 0; j < 4; ++j) red[warp][m][lane][j] = acc[m][j];
  }
  __syncthreads();
  if (warp != 0) return;
  // C fragment: (g, t) holds rows g / g + 8, columns 2t / 2t + 1
  const int c0 = n0 + 2 * t;
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sacc = 0.f;
#pragma unroll
      for (int w = 0; w < SK_WARPS; ++w) sacc += red[w][m][lane][j];
      v[j] = sacc;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = m * 16 + g + half * 8;
      if (r >= a.M) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = c0 + j;
        if (c >= a.N) continue;
        float y = v[half * 2 + j] + (a.bias != nullptr ? __ldg(a.bias + c) : 0.f);
        const long long o = (long long)r * a.N + c;
        if constexpr (KIND == EPI_BF16) {
          static_cast<__nv_bfloat16*>(a.out)[o] = __float2bfloat16_rn(y);
        } else if constexpr (KIND == EPI_GELU_BF16) {
          float gl, gp;
          gelu_and_grad(y, gl, gp);
          static_cast<__nv_bfloat16*>(a.out)[o] = __float2bfloat16_rn(gl);
          if (a.out2 != nullptr) static_cast<__nv_bfloat16*>(a.out2)[o] = __float2bfloat16_rn(gp);
        } else if constexpr (KIND == EPI_RESID_F32) {
          static_cast<float*>(a.out)[o] = y + a.resid[o];
        } else {
          static_cast<float*>(a.out)[o] = y;
        }
      }
    }
  }
}

template <int KIND>
static int launch_skinny(const SkinnyArgs& a, cudaStream_t st) {
  const int grid = (a.N + 7) / 8;
  const int mt = (a.M + 15) / 16;
  // long contractions (mlp[2], K = 4d) split K over 8 warps: half the exposed DRAM round trips per warp and twice the
  // loads in flight per SM for the few (N / 8) CTAs of a narrow output
  const bool wide = a.K >= 2048 && a.K % (8 * SK_KCHUNK) == 0;
  if (wide) {
    switch (mt) {
      case 1: gemm_skinny_kernel<1, KIND, 8><<<grid, 256, 0, st>>>(a); break;
      case 2: gemm_skinny_kernel<2, KIND, 8><<<grid, 256, 0, st>>>(a); break;
      case 3: gemm_skinny_kernel<3, KIND, 8><<<grid, 256, 0, st>>>(a); break;
      default: gemm_skinny_kernel<4, KIND, 8><<<grid, 256, 0, st>>>(a); break;
    }
  } else {
    switch (mt) {
      case 1: gemm_skinny_kernel<1, KIND, 4><<<grid, 128, 0, st>>>(a); break;
      case 2: gemm_skinny_kernel<2, KIND, 4><<<grid, 128, 0, st>>>(a); break;
      case 3: gemm_skinny_kernel<3, KIND, 4><<<grid, 128, 0, st>>>(a); break;
      default: gemm_skinny_kernel<4, KIND, 4><<<grid, 128, 0, st>>>(a); break;
    }
  }
  return check_cuda(cudaGetLastError(), "gemm_skinny launch");
}

// Returns 1 when the shape is taken by the skinny kernel (and it was launched), 0 when the caller should use the tile
// kernel, a negative error code on failure.
int gemm_skinny_try(int kind, const void* x, const void* w, const float* bias, void* out, void* out2, const float* resid,
                    int M, int N, int K, cudaStream_t st) {
  if (g_debug[6] == 1) return 0;                           // bring-up knob: force the tile kernel
  if (M > 64 || K % (4 * SK_KCHUNK) != 0 || N % 2 != 0) return 0;
  SkinnyArgs a{(const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias, out, out2, resid, M, N, K};
  int rc;
  switch (kind) {
    case EPI_BF16: rc = launch_skinny<EPI_BF16>(a, st); break;
    case EPI_GELU_BF16: rc = launch_skinny<EPI_GELU_BF16>(a, st); break;
    case EPI_RESID_F32: rc = launch_skinny<EPI_RESID_F32>(a, st); break;
    case EPI_F32: rc = launch_skinny<EPI_F32>(a, st); break;
    default: return 0;
  }
  return rc == OK ? 1 : rc;
}

}  // namespace b200
