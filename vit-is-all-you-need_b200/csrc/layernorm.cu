// LayerNorm forward/backward for the fp32 residual stream (HBM-bound; one warp per row, 128-bit accesses,
// warp-shuffle statistics).  Fusions:
//   fwd: optional residual add (x_out = x + add_bf16) before normalising -> "x + attn(...)" then LN
//        (transformer.py:43-44), emits the bf16 GEMM operand and saves mean / rstd;
//   bwd: dx = dres + LN'(dy) (the residual-gradient add of the pre-norm block), optional bf16 copy of dx for
//        the next GEMM, optional affine grads (nn.LayerNorm in blocks.py:43,48).
// Matches F.layer_norm: biased variance, eps inside the sqrt.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace b200 {

constexpr int LN_THREADS = 256;
constexpr int LN_WARPS = LN_THREADS / 32;
constexpr int LN_MAXCH = 8;  // float4 chunks per lane -> d <= 1024

struct LnFwdArgs {
  const float* x; const __nv_bfloat16* add; float* x_out;
  const float* gamma; const float* beta;
  __nv_bfloat16* y; float* y_f32; float* mean; float* rstd;
  int M, d; float eps;
};

__global__ void __launch_bounds__(LN_THREADS) ln_fwd_kernel(const LnFwdArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = a.d >> 2;
  for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < a.M; row += (long long)gridDim.x * LN_WARPS) {
    const float4* xr = reinterpret_cast<const float4*>(a.x + row * a.d);
    float4 v[LN_MAXCH];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        v[i] = xr[c];
        if (a.add != nullptr) {
          const uint2 u = *reinterpret_cast<const uint2*>(a.add + row * a.d + c * 4);
          const float2 p0 = unpack_bf16(u.x), p1 = unpack_bf16(u.y);
          v[i].x += p0.x; v[i].y += p0.y; v[i].z += p1.x; v[i].w += p1.y;
        }
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    sum = warp_sum(sum);
    const float mean = sum / (float)a.d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
        sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
    }
    sq = warp_sum(sq);
    const float rstd = rsqrtf(sq / (float)a.d + a.eps);
    if (lane == 0) {
      if (a.mean) a.mean[row] = mean;
      if (a.rstd) a.rstd[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < LN_MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        if (a.x_out != nullptr) reinterpret_cast<float4*>(a.x_out + row * a.d)[c] = v[i];
        float4 o;
        o.x = (v[i].x - mean) * rstd; o.y = (v[i].y - mean) * rstd;
        o.z = (v[i].z - mean) * rstd; o.w = (v[i].w - mean) * rstd;
        if (a.gamma != nullptr) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + c);
          o.x *= g.x; o.y *= g.y; o.z *= g.z; o.w *= g.w;
        }
        if (a.beta != nullptr) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta) + c);
          o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        if (a.y != nullptr) {
          uint2 w;
          w.x = pack_bf16(o.x, o.y); w.y = pack_bf16(o.z, o.w);
          *reinterpret_cast<uint2*>(a.y + row * a.d + c * 4) = w;
        }
        if (a.y_f32 != nullptr) reinterpret_cast<float4*>(a.y_f32 + row * a.d)[c] = o;
      }
    }
  }
}

struct LnBwdArgs {
  const __nv_bfloat16* dy; const float* dy_f32; const float* x; const float* mean; const float* rstd;
  const float* gamma; const float* dres;
  float* dx; __nv_bfloat16* dx_bf16; float* dgamma; float* dbeta;
  int M, d;
  const __nv_bfloat16* xhat;  // when set (affine-free LN only): the saved bf16 normalised rows replace x / mean
};

template <bool AFFINE>
__global__ void __launch_bounds__(LN_THREADS) ln_bwd_kernel(const LnBwdArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = a.d >> 2;
  constexpr bool affine_grads = AFFINE;
  float4 accg[LN_MAXCH], accb[LN_MAXCH];
#pragma unroll
  for (int i = 0; i < LN_MAXCH; ++i) { accg[i] = make_float4(0, 0, 0, 0); accb[i] = make_float4(0, 0, 0, 0); }

  for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < a.M; row += (long long)gridDim.x * LN_WARPS) {
    const float mean = a.xhat != nullptr ? 0.f : a.mean[row], rstd = a.rstd[row];
    float4 xh[LN_MAXCH], g[LN_MAXCH];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float4 xv;
        if (a.xhat != nullptr) {
          const uint2 u = *reinterpret_cast<const uint2*>(a.xhat + row * a.d + c * 4);
          const float2 p0 = unpack_bf16(u.x), p1 = unpack_bf16(u.y);
          xv = make_float4(p0.x, p0.y, p1.x, p1.y);
        } else {
          xv = reinterpret_cast<const float4*>(a.x + row * a.d)[c];
        }
        float4 dyv;
        if (a.dy != nullptr) {
          const uint2 u = *reinterpret_cast<const uint2*>(a.dy + row * a.d + c * 4);
          const float2 p0 = unpack_bf16(u.x), p1 = unpack_bf16(u.y);
          dyv = make_float4(p0.x, p0.y, p1.x, p1.y);
        } else {
          dyv = reinterpret_cast<const float4*>(a.dy_f32 + row * a.d)[c];
        }
        if (a.xhat != nullptr) {
          xh[i] = xv;
        } else {
          xh[i].x = (xv.x - mean) * rstd; xh[i].y = (xv.y - mean) * rstd;
          xh[i].z = (xv.z - mean) * rstd; xh[i].w = (xv.w - mean) * rstd;
        }
        if (affine_grads) {
          accg[i].x += dyv.x * xh[i].x; accg[i].y += dyv.y * xh[i].y;
          accg[i].z += dyv.z * xh[i].z; accg[i].w += dyv.w * xh[i].w;
          accb[i].x += dyv.x; accb[i].y += dyv.y; accb[i].z += dyv.z; accb[i].w += dyv.w;
        }
        g[i] = dyv;
        if (a.gamma != nullptr) {
          const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma) + c);
          g[i].x *= gm.x; g[i].y *= gm.y; g[i].z *= gm.z; g[i].w *= gm.w;
        }
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      }
    }
    s1 = warp_sum(s1) / (float)a.d;
    s2 = warp_sum(s2) / (float)a.d;
#pragma unroll
    for (int i = 0; i < LN_MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        float4 o;
        o.x = (g[i].x - s1 - xh[i].x * s2) * rstd; o.y = (g[i].y - s1 - xh[i].y * s2) * rstd;
        o.z = (g[i].z - s1 - xh[i].z * s2) * rstd; o.w = (g[i].w - s1 - xh[i].w * s2) * rstd;
        if (a.dres != nullptr) {
          const float4 r = reinterpret_cast<const float4*>(a.dres + row * a.d)[c];
          o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        reinterpret_cast<float4*>(a.dx + row * a.d)[c] = o;
        if (a.dx_bf16 != nullptr) {
          uint2 w;
          w.x = pack_bf16(o.x, o.y); w.y = pack_bf16(o.z, o.w);
          *reinterpret_cast<uint2*>(a.dx_bf16 + row * a.d + c * 4) = w;
        }
      }
    }
  }

  if (affine_grads) {
    // cross-warp reduction in shared memory, then one atomic per column per CTA
    __shared__ float4 red[LN_WARPS][32];
#pragma unroll
    for (int i = 0; i < LN_MAXCH; ++i) {
      const int c = lane + 32 * i;
      if (32 * i >= nvec) continue;  // uniform across the CTA
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
        red[warp][lane] = pass == 0 ? accg[i] : accb[i];
        __syncthreads();
        if (warp == 0 && c < nvec) {
          float4 t = red[0][lane];
          for (int w = 1; w < LN_WARPS; ++w) {
            t.x += red[w][lane].x; t.y += red[w][lane].y; t.z += red[w][lane].z; t.w += red[w][lane].w;
          }
          float* dst = (pass == 0 ? a.dgamma : a.dbeta) + c * 4;
          atomicAdd(dst + 0, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
        }
      }
    }
  }
}


// ----------------------------------------------------------------------------------------------------------
// Specialised row kernels for d = CH * 128 (512 / 768 / 1024: every shipped width).  One warp per row, one row
// per warp, every global load of the row issued before the first use, no per-chunk branches: ~170 issue slots
// per row instead of ~600 in the generic kernels above, and registers low enough for >= 20 resident warps per
// SM, so the kernels run at HBM speed instead of being issue/latency bound.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream4(const float* p) {  // read-once data: do not keep it in L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ld_stream2(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

constexpr int LNF_THREADS = 128;
constexpr int LNF_WARPS = LNF_THREADS / 32;

template <int CH, bool ADD, bool AFFINE>
__global__ void __launch_bounds__(LNF_THREADS) ln_fwd_fast_kernel(const LnFwdArgs a) {
  constexpr int D = CH * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * LNF_WARPS + warp;
  pdl_wait();
  pdl_trigger();
  if (row >= a.M) return;
  const float* xr = a.x + row * D + lane * 4;
  float4 v[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) v[i] = ld_stream4(xr + i * 128);
  if constexpr (ADD) {
    const __nv_bfloat16* ar = a.add + row * D + lane * 4;
    uint2 u[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) u[i] = ld_stream2(ar + i * 128);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const float2 p0 = unpack_bf16(u[i].x), p1 = unpack_bf16(u[i].y);
      v[i].x += p0.x; v[i].y += p0.y; v[i].z += p1.x; v[i].w += p1.y;
    }
    if (a.x_out != nullptr) {
      float* xo = a.x_out + row * D + lane * 4;
#pragma unroll
      for (int i = 0; i < CH; ++i) *reinterpret_cast<float4*>(xo + i * 128) = v[i];
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(sum) * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
    sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + a.eps);
  if (lane == 0) {
    if (a.mean) a.mean[row] = mean;
    if (a.rstd) a.rstd[row] = rstd;
  }
  __nv_bfloat16* yr = a.y + row * D + lane * 4;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    float4 o;
    o.x = (v[i].x - mean) * rstd; o.y = (v[i].y - mean) * rstd;
    o.z = (v[i].z - mean) * rstd; o.w = (v[i].w - mean) * rstd;
    if constexpr (AFFINE) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + lane + 32 * i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta) + lane + 32 * i);
      o.x = fmaf(o.x, g.x, b.x); o.y = fmaf(o.y, g.y, b.y); o.z = fmaf(o.z, g.z, b.z); o.w = fmaf(o.w, g.w, b.w);
    }
    uint2 w;
    w.x = pack_bf16(o.x, o.y); w.y = pack_bf16(o.z, o.w);
    *reinterpret_cast<uint2*>(yr + i * 128) = w;
  }
}

// dx = dres + LN'(dy), affine-free or with gamma (no parameter gradients here), bf16 dy, optional bf16 copy.
// XHAT: x-hat comes from the saved bf16 output of the (affine-free) forward instead of being rebuilt from the fp32
// row and its mean -- 14 instead of 16 bytes per element, and the fp32 residual rows need not be kept for backward.
template <int CH, bool GAMMA, bool XHAT>
__global__ void __launch_bounds__(LNF_THREADS) ln_bwd_fast_kernel(const LnBwdArgs a) {
  constexpr int D = CH * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * LNF_WARPS + warp;
  pdl_wait();
  pdl_trigger();
  if (row >= a.M) return;
  const float rstd = __ldg(a.rstd + row);
  const __nv_bfloat16* dyr = a.dy + row * D + lane * 4;
  float4 xh[CH], r[CH];
  uint2 u[CH];
  [[maybe_unused]] uint2 xq[CH];
  [[maybe_unused]] float nm = 0.f;
  if constexpr (XHAT) {
    const __nv_bfloat16* hr = a.xhat + row * D + lane * 4;
#pragma unroll
    for (int i = 0; i < CH; ++i) xq[i] = ld_stream2(hr + i * 128);
  } else {
    nm = -__ldg(a.mean + row) * rstd;
    const float* xr = a.x + row * D + lane * 4;
#pragma unroll
    for (int i = 0; i < CH; ++i) xh[i] = ld_stream4(xr + i * 128);
  }
#pragma unroll
  for (int i = 0; i < CH; ++i) u[i] = ld_stream2(dyr + i * 128);
  if (a.dres != nullptr) {
    const float* rr = a.dres + row * D + lane * 4;
#pragma unroll
    for (int i = 0; i < CH; ++i) r[i] = ld_stream4(rr + i * 128);
  } else {
#pragma unroll
    for (int i = 0; i < CH; ++i) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4 g[CH];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const float2 p0 = unpack_bf16(u[i].x), p1 = unpack_bf16(u[i].y);
    g[i] = make_float4(p0.x, p0.y, p1.x, p1.y);
    if constexpr (GAMMA) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(a.gamma) + lane + 32 * i);
      g[i].x *= gm.x; g[i].y *= gm.y; g[i].z *= gm.z; g[i].w *= gm.w;
    }
    if constexpr (XHAT) {
      const float2 h0 = unpack_bf16(xq[i].x), h1 = unpack_bf16(xq[i].y);
      xh[i] = make_float4(h0.x, h0.y, h1.x, h1.y);
    } else {
      xh[i].x = fmaf(xh[i].x, rstd, nm); xh[i].y = fmaf(xh[i].y, rstd, nm);
      xh[i].z = fmaf(xh[i].z, rstd, nm); xh[i].w = fmaf(xh[i].w, rstd, nm);
    }
    s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
    s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
  }
  s1 = warp_sum(s1) * (1.0f / D);
  s2 = warp_sum(s2) * (1.0f / D);
  float* dxr = a.dx + row * D + lane * 4;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    float4 o;
    o.x = fmaf((g[i].x - s1) - xh[i].x * s2, rstd, r[i].x);
    o.y = fmaf((g[i].y - s1) - xh[i].y * s2, rstd, r[i].y);
    o.z = fmaf((g[i].z - s1) - xh[i].z * s2, rstd, r[i].z);
    o.w = fmaf((g[i].w - s1) - xh[i].w * s2, rstd, r[i].w);
    *reinterpret_cast<float4*>(dxr + i * 128) = o;
    if (a.dx_bf16 != nullptr) {
      uint2 w;
      w.x = pack_bf16(o.x, o.y); w.y = pack_bf16(o.z, o.w);
      *reinterpret_cast<uint2*>(a.dx_bf16 + row * D + lane * 4 + i * 128) = w;
    }
  }
}

template <int CH>
static bool launch_ln_fwd_fast(const LnFwdArgs& a, cudaStream_t st) {
  const int grid = (a.M + LNF_WARPS - 1) / LNF_WARPS;
  const bool affine = a.gamma != nullptr && a.beta != nullptr;
  if (a.add != nullptr) {
    if (affine) launch_kernel(ln_fwd_fast_kernel<CH, true, true>, dim3(grid), dim3(LNF_THREADS), 0, st, 1, a);
    else        launch_kernel(ln_fwd_fast_kernel<CH, true, false>, dim3(grid), dim3(LNF_THREADS), 0, st, 1, a);
  } else {
    if (affine) launch_kernel(ln_fwd_fast_kernel<CH, false, true>, dim3(grid), dim3(LNF_THREADS), 0, st, 1, a);
    else        launch_kernel(ln_fwd_fast_kernel<CH, false, false>, dim3(grid), dim3(LNF_THREADS), 0, st, 1, a);
  }
  return true;
}
// Returns true when a specialised kernel was launched.
static bool try_ln_fwd_fast(const LnFwdArgs& a, cudaStream_t st) {
  if (a.y == nullptr || a.y_f32 != nullptr || (a.gamma == nullptr) != (a.beta == nullptr)) return false;
  if (a.add == nullptr && a.x_out != nullptr) return false;
  switch (a.d) {
    case 512: return launch_ln_fwd_fast<4>(a, st);
    case 768: return launch_ln_fwd_fast<6>(a, st);
    case 1024: return launch_ln_fwd_fast<8>(a, st);
    default: return false;
  }
}
template <int CH>
static bool launch_ln_bwd_fast(const LnBwdArgs& a, cudaStream_t st) {
  const int grid = (a.M + LNF_WARPS - 1) / LNF_WARPS;
  if (a.xhat != nullptr)  launch_kernel(ln_bwd_fast_kernel<CH, false, true>, dim3(grid), dim3(LNF_THREADS), 0, st, 1, a);
  else if (a.gamma != nullptr) launch_kernel(ln_bwd_fast_kernel<CH, true, false>, dim3(grid), dim3(LNF_THREADS), 0, st, 1, a);
  else                    launch_kernel(ln_bwd_fast_kernel<CH, false, false>, dim3(grid), dim3(LNF_THREADS), 0, st, 1, a);
  return true;
}
static bool try_ln_bwd_fast(const LnBwdArgs& a, cudaStream_t st) {
  if (a.dy == nullptr || a.dgamma != nullptr) return false;
  switch (a.d) {
    case 512: return launch_ln_bwd_fast<4>(a, st);
    case 768: return launch_ln_bwd_fast<6>(a, st);
    case 1024: return launch_ln_bwd_fast<8>(a, st);
    default: return false;
  }
}

// Column sums of a bf16 matrix [M, N] -> fp32 [N] (bias gradients).  Each CTA owns 64 columns x a slab of
// rows; 8-byte loads, shared-memory reduction over the row groups, one atomic per column per CTA.
constexpr int CS_THREADS = 256;
__global__ void __launch_bounds__(CS_THREADS)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ a, float* __restrict__ out, int M, int N, int rows_per_cta) {
  const int tx = threadIdx.x & 15;   // 16 threads x 4 columns = 64 columns
  const int ty = threadIdx.x >> 4;   // 16 row groups
  const int col = blockIdx.x * 64 + tx * 4;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float4 acc = make_float4(0, 0, 0, 0);
  if (col < N) {
    for (int r = r0 + ty; r < r1; r += 16) {
      const uint2 u = *reinterpret_cast<const uint2*>(a + (long long)r * N + col);
      const float2 p0 = unpack_bf16(u.x), p1 = unpack_bf16(u.y);
      acc.x += p0.x; acc.y += p0.y; acc.z += p1.x; acc.w += p1.y;
    }
  }
  __shared__ float4 red[16][16];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < N) {
    float4 t = red[0][tx];
    for (int i = 1; i < 16; ++i) { t.x += red[i][tx].x; t.y += red[i][tx].y; t.z += red[i][tx].z; t.w += red[i][tx].w; }
    atomicAdd(out + col + 0, t.x); atomicAdd(out + col + 1, t.y);
    atomicAdd(out + col + 2, t.z); atomicAdd(out + col + 3, t.w);
  }
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_layernorm_fwd(const float* x, const void* add_bf16, float* x_out, const float* gamma,
                          const float* beta, void* y_bf16, float* y_f32, float* mean, float* rstd, int M,
                          int d, float eps, void* stream) {
  B200_REQUIRE(x && (y_bf16 || y_f32), "layernorm_fwd: null pointer");
  B200_REQUIRE(M > 0 && d > 0 && d % 4 == 0 && d <= 128 * LN_MAXCH, "layernorm_fwd: d=%d must be a multiple of 4 and <= %d", d, 128 * LN_MAXCH);
  LnFwdArgs a{x, (const __nv_bfloat16*)add_bf16, x_out, gamma, beta, (__nv_bfloat16*)y_bf16, y_f32, mean, rstd, M, d, eps};
  if (try_ln_fwd_fast(a, (cudaStream_t)stream)) {
    B200_CUDA(cudaGetLastError());
    return OK;
  }
  const int blocks = (M + LN_WARPS - 1) / LN_WARPS;
  const int grid = blocks < num_sms() * 16 ? blocks : num_sms() * 16;
  ln_fwd_kernel<<<grid, LN_THREADS, 0, (cudaStream_t)stream>>>(a);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* mean,
                          const float* rstd, const float* gamma, const float* dres, float* dx,
                          void* dx_bf16, float* dgamma, float* dbeta, int M, int d, void* stream) {
  B200_REQUIRE((dy_bf16 || dy_f32) && x && mean && rstd && dx, "layernorm_bwd: null pointer");
  B200_REQUIRE(M > 0 && d > 0 && d % 4 == 0 && d <= 128 * LN_MAXCH, "layernorm_bwd: d=%d must be a multiple of 4 and <= %d", d, 128 * LN_MAXCH);
  B200_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma and dbeta go together");
  cudaStream_t st = (cudaStream_t)stream;
  if (dgamma) {
    B200_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * d, st));
    B200_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * d, st));
  }
  LnBwdArgs a{(const __nv_bfloat16*)dy_bf16, dy_f32, x, mean, rstd, gamma, dres, dx, (__nv_bfloat16*)dx_bf16, dgamma, dbeta, M, d, nullptr};
  if (try_ln_bwd_fast(a, st)) {
    B200_CUDA(cudaGetLastError());
    return OK;
  }
  const int blocks = (M + LN_WARPS - 1) / LN_WARPS;
  const int cap = dgamma ? num_sms() * 4 : num_sms() * 16;
  const int grid = blocks < cap ? blocks : cap;
  if (dgamma) ln_bwd_kernel<true><<<grid, LN_THREADS, 0, st>>>(a);
  else        ln_bwd_kernel<false><<<grid, LN_THREADS, 0, st>>>(a);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_layernorm_bwd_xhat(const void* dy_bf16, const void* xhat_bf16, const float* rstd, const float* dres,
                               float* dx, void* dx_bf16, int M, int d, void* stream) {
  B200_REQUIRE(dy_bf16 && xhat_bf16 && rstd && dx, "layernorm_bwd_xhat: null pointer");
  B200_REQUIRE(M > 0 && d > 0 && d % 4 == 0 && d <= 128 * LN_MAXCH, "layernorm_bwd_xhat: d=%d must be a multiple of 4 and <= %d", d, 128 * LN_MAXCH);
  cudaStream_t st = (cudaStream_t)stream;
  LnBwdArgs a{(const __nv_bfloat16*)dy_bf16, nullptr, nullptr, nullptr, rstd, nullptr, dres, dx, (__nv_bfloat16*)dx_bf16,
              nullptr, nullptr, M, d, (const __nv_bfloat16*)xhat_bf16};
  if (try_ln_bwd_fast(a, st)) {
    B200_CUDA(cudaGetLastError());
    return OK;
  }
  const int blocks = (M + LN_WARPS - 1) / LN_WARPS;
  const int grid = blocks < num_sms() * 16 ? blocks : num_sms() * 16;
  ln_bwd_kernel<false><<<grid, LN_THREADS, 0, st>>>(a);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_colsum_bf16(const void* a, float* out, int M, int N, int accumulate, void* stream) {
  B200_REQUIRE(a && out && M > 0 && N > 0 && N % 4 == 0, "colsum: bad arguments (N must be a multiple of 4)");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) B200_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  const int gx = (N + 63) / 64;
  int gy = (num_sms() * 8 + gx - 1) / gx;
  int rows_per_cta = (M + gy - 1) / gy;
  if (rows_per_cta < 64) rows_per_cta = 64;
  gy = (M + rows_per_cta - 1) / rows_per_cta;
  colsum_bf16_kernel<<<dim3(gx, gy), CS_THREADS, 0, st>>>((const __nv_bfloat16*)a, out, M, N, rows_per_cta);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"
