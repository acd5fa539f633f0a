// Fused VQ nearest-codebook lookup (fp32 CUDA cores on purpose: indices must be bit-exact).
//
// Replaces, in one pass with no [rows, K] distance matrix in HBM:
//   train_titok.py:50-59  Quantizer.forward        F.normalize x2, torch.cdist, argmin, codebook gather,
//                                                   MSE losses, straight-through output
//   blocks.py:428-505     VectorQuantizer.forward  (l2-norm, expanded distance, argmin, normalised gather)
// Arithmetic contract == oracle/vq_oracle.c (sequential fmaf over the latent dim, IEEE sqrt/div,
// dist = fma(-2, dot, xx + ee), first-minimum tie rule).
//
// Layout: the (normalised) codebook is staged in shared memory (row stride chosen conflict-free for 128-bit
// loads); every warp owns ROWS rows whose normalised latents live in registers; lanes scan interleaved codes
// and keep a private (best distance, best index); a shuffle arg-min merges the lanes.
#include "../../include/b200vit.h"
#include "common.cuh"

namespace b200 {

struct VqLayout {
  long long R;             // rows (latent vectors)
  int D;                   // latent dim
  int K;                   // codes
  long long inner;         // x[(r / inner) * outer_stride + j * elem_stride + r % inner]
  long long elem_stride;
  long long outer_stride;
  int flags;               // bit0: l2-normalise; bit1: gather normalised rows
};

__host__ __device__ inline int vq_dp(int D) { return (D + 3) & ~3; }
__host__ __device__ inline int vq_stride(int DP) { return (DP % 8 == 4) ? DP : DP + 4; }

// ---- codebook preparation: normalised rows (zero padded to DP) + squared norms -------------------
__global__ void vq_prep_kernel(const float* __restrict__ codebook, int K, int D, int DP, int l2,
                               float* __restrict__ ehat, float* __restrict__ ee) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const float* c = codebook + (long long)k * D;
  float den = 1.0f;
  if (l2) {
    float ss = 0.0f;
    for (int j = 0; j < D; ++j) ss = __fmaf_rn(c[j], c[j], ss);
    den = fmaxf(__fsqrt_rn(ss), 1e-12f);
  }
  float s = 0.0f;
  for (int j = 0; j < DP; ++j) {
    float v = 0.0f;
    if (j < D) v = l2 ? __fdiv_rn(c[j], den) : c[j];
    ehat[(long long)k * DP + j] = v;
    if (j < D) s = __fmaf_rn(v, v, s);
  }
  ee[k] = s;
}

constexpr int VQ_THREADS = 256;
constexpr int VQ_WARPS = VQ_THREADS / 32;
constexpr int VQ_MAX_WARPS = 16;

// MAXT = largest block the instantiation is launched with (register cap); the block size itself is a launch parameter
template <int DP, int ROWS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
vq_fwd_kernel(const float* __restrict__ x, const float* __restrict__ codebook,
              const float* __restrict__ ehat, const float* __restrict__ ee, VqLayout L, int KC,
              long long* __restrict__ idx_out, float* __restrict__ q_out, double* __restrict__ partial) {
  extern __shared__ float4 vq_smem4[];
  float* s_codes = reinterpret_cast<float*>(vq_smem4);
  constexpr int STRIDE = (DP % 8 == 4) ? DP : DP + 4;
  float* s_ee = s_codes + (size_t)KC * STRIDE;
  __shared__ double s_part[VQ_MAX_WARPS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const long long row0 = ((long long)blockIdx.x * nwarps + warp) * ROWS;
  const bool l2 = (L.flags & 1) != 0;
  const bool gather_norm = (L.flags & 2) != 0;

  // Every lane keeps the full normalised latent of the warp's ROWS rows (redundant, but D is tiny).
  float xh[ROWS][DP];
  float xx[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const long long row = row0 + r;
    const bool valid = row < L.R;
    const float* xr = x + (valid ? (row / L.inner) * L.outer_stride + (row % L.inner) : 0);
    float raw[DP];
#pragma unroll
    for (int j = 0; j < DP; ++j) raw[j] = (valid && j < L.D) ? __ldg(xr + (long long)j * L.elem_stride) : 0.0f;
    float den = 1.0f;
    if (l2) {
      float ss = 0.0f;
#pragma unroll
      for (int j = 0; j < DP; ++j) ss = __fmaf_rn(raw[j], raw[j], ss);
      den = fmaxf(__fsqrt_rn(ss), 1e-12f);
    }
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      xh[r][j] = l2 ? __fdiv_rn(raw[j], den) : raw[j];
      s = __fmaf_rn(xh[r][j], xh[r][j], s);
    }
    xx[r] = s;
  }

  float best[ROWS];
  int besti[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) { best[r] = INFINITY; besti[r] = 0x7fffffff; }

  for (int k0 = 0; k0 < L.K; k0 += KC) {
    const int kc = min(KC, L.K - k0);
    __syncthreads();  // previous chunk fully consumed
    // stage the chunk: 16-byte copies, padded row stride
    for (int i = threadIdx.x; i < kc * (DP / 4); i += nthreads) {
      const int k = i / (DP / 4), v = i - k * (DP / 4);
      const float4 t = __ldg(reinterpret_cast<const float4*>(ehat + (long long)(k0 + k) * DP) + v);
      *reinterpret_cast<float4*>(s_codes + (size_t)k * STRIDE + v * 4) = t;
    }
    for (int i = threadIdx.x; i < kc; i += nthreads) s_ee[i] = __ldg(ee + k0 + i);
    __syncthreads();

    for (int k = lane; k < kc; k += 32) {
      float e[DP];
#pragma unroll
      for (int v = 0; v < DP / 4; ++v) {
        const float4 t = *reinterpret_cast<const float4*>(s_codes + (size_t)k * STRIDE + v * 4);
        e[v * 4 + 0] = t.x; e[v * 4 + 1] = t.y; e[v * 4 + 2] = t.z; e[v * 4 + 3] = t.w;
      }
      const float eek = s_ee[k];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        float dot = 0.0f;
#pragma unroll
        for (int j = 0; j < DP; ++j) dot = __fmaf_rn(xh[r][j], e[j], dot);
        const float dist = __fmaf_rn(-2.0f, dot, __fadd_rn(xx[r], eek));
        if (dist < best[r]) { best[r] = dist; besti[r] = k0 + k; }
      }
    }
  }

  // lanes -> warp arg-min, lowest index wins ties (== first minimum of a sequential scan)
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best[r], o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti[r], o);
      if (od < best[r] || (od == best[r] && oi < besti[r])) { best[r] = od; besti[r] = oi; }
    }
  }

  // epilogue: lane j writes latent element j (and j + 32 when D > 32)
  float acc = 0.0f;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const long long row = row0 + r;
    if (row >= L.R) continue;
    const int bi = besti[r];
    if (lane == 0) idx_out[row] = bi;
    float* qr = q_out + (row / L.inner) * L.outer_stride + (row % L.inner);
#pragma unroll
    for (int jb = 0; jb < DP; jb += 32) {
      const int j = jb + lane;
      float xv = 0.0f;
#pragma unroll
      for (int jj = 0; jj < 32 && jb + jj < DP; ++jj)
        if (jj == lane) xv = xh[r][jb + jj];
      if (j < L.D) {
        const float c = gather_norm ? __ldg(ehat + (long long)bi * DP + j) : __ldg(codebook + (long long)bi * L.D + j);
        const float diff = __fsub_rn(c, xv);
        qr[(long long)j * L.elem_stride] = __fadd_rn(xv, diff);
        acc = __fmaf_rn(diff, diff, acc);
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) s_part[warp] = (double)acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < nwarps; ++w) t += s_part[w];
    partial[blockIdx.x] = t;
  }
}

// losses[0] = mse = sum (c - xh)^2 / Nel ; losses[1] = commitment_cost * mse ; losses[2] = (1 + cc) * mse
__global__ void vq_finalize_kernel(const double* __restrict__ partial, int n, double nel, float cc,
                                   float* __restrict__ losses) {
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) t += partial[i];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0) {
    const double mse = t / nel;
    losses[0] = (float)mse;
    losses[1] = (float)(cc * mse);
    losses[2] = (float)((1.0 + cc) * mse);
  }
}

// ---- backward: straight-through + commitment/codebook losses, normalise-backward, scatter-add -----
__global__ void vq_bwd_kernel(const float* __restrict__ x, const float* __restrict__ codebook,
                              const long long* __restrict__ idx, const float* __restrict__ g,
                              const float* __restrict__ coef /*[2]: a_commit, a_code*/, VqLayout L,
                              float* __restrict__ dx, float* __restrict__ dC) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= L.R) return;
  const bool l2 = (L.flags & 1) != 0;
  const bool gather_norm = (L.flags & 2) != 0;
  const int D = L.D;
  const long long off = (row / L.inner) * L.outer_stride + (row % L.inner);
  const float inv_nel = 1.0f / ((float)L.R * (float)D);
  const float a_commit = coef[0] * 2.0f * inv_nel, a_code = coef[1] * 2.0f * inv_nel;
  const float* c = codebook + idx[row] * D;

  float den = 1.0f, cden = 1.0f;
  if (l2) {
    float ss = 0.0f;
    for (int j = 0; j < D; ++j) { const float v = x[off + j * L.elem_stride]; ss = __fmaf_rn(v, v, ss); }
    den = fmaxf(sqrtf(ss), 1e-12f);
  }
  if (gather_norm) {
    float ss = 0.0f;
    for (int j = 0; j < D; ++j) ss = __fmaf_rn(c[j], c[j], ss);
    cden = fmaxf(sqrtf(ss), 1e-12f);
  }
  float proj = 0.0f, cproj = 0.0f;
  for (int j = 0; j < D; ++j) {
    const float xh = x[off + j * L.elem_stride] / den;
    const float ch = c[j] / cden;
    const float dxh = g[off + j * L.elem_stride] + a_commit * (xh - ch);
    const float dc = a_code * (ch - xh);
    proj += xh * dxh;
    cproj += ch * dc;
  }
  for (int j = 0; j < D; ++j) {
    const float xh = x[off + j * L.elem_stride] / den;
    const float ch = c[j] / cden;
    const float dxh = g[off + j * L.elem_stride] + a_commit * (xh - ch);
    const float dc = a_code * (ch - xh);
    dx[off + j * L.elem_stride] = l2 ? (dxh - xh * proj) / den : dxh;
    atomicAdd(dC + idx[row] * D + j, gather_norm ? (dc - ch * cproj) / cden : dc);
  }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static int vq_rows_per_cta(int DP) { return VQ_WARPS * (DP <= 16 ? 4 : (DP <= 32 ? 2 : 1)); }

template <int DP, int ROWS, int MAXT = VQ_THREADS>
static int launch_vq_fwd(const float* x, const float* codebook, const float* ehat, const float* ee,
                         const VqLayout& L, long long* idx, float* q, double* partial, int grid,
                         cudaStream_t st, int threads = VQ_THREADS) {
  constexpr int STRIDE = (DP % 8 == 4) ? DP : DP + 4;
  const size_t per_code = (size_t)STRIDE * 4 + 4;
  // 4096 codes x (12 floats + norm) = 213 KB: the whole TiTok codebook in ONE chunk (a 200 KB budget split it into
  // 3936 + 160 codes, the second chunk leaving most lanes idle between two extra barriers)
  const size_t budget = 216 * 1024;
  int KC = (int)(budget / per_code);
  if (KC > L.K) KC = L.K;
  KC = (KC + 3) & ~3;  // keep s_ee 16-byte aligned after the code rows
  const size_t smem = (size_t)KC * per_code + 16;
  auto kern = vq_fwd_kernel<DP, ROWS, MAXT>;
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
  kern<<<grid, threads, smem, st>>>(x, codebook, ehat, ee, L, KC, idx, q, partial);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200vit_vq_workspace_size(long long R, int D, int K) {
  const int DP = vq_dp(D);
  const int rpc_min = DP <= 16 ? 16 : vq_rows_per_cta(DP);   // the many-warp variant may use as few as 4 warps x 4 rows per CTA
  const long long grid = (R + rpc_min - 1) / rpc_min;
  return align256((size_t)K * DP * 4) + align256((size_t)K * 4) + align256((size_t)grid * 8);
}

int b200vit_vq_fwd(const float* x, const float* codebook, long long R, int D, int K, long long inner,
                   long long elem_stride, long long outer_stride, int flags, float commitment_cost,
                   long long* indices, float* quantized, float* losses, void* workspace,
                   size_t workspace_bytes, void* stream) {
  B200_REQUIRE(x && codebook && indices && quantized && losses && workspace, "vq_fwd: null pointer");
  B200_REQUIRE(R > 0 && D > 0 && K > 0, "vq_fwd: bad sizes R=%lld D=%d K=%d", R, D, K);
  if (D > 64) {
    set_error("vq_fwd: latent dim %d > 64 is not implemented", D);
    return ERR_UNSUPPORTED;
  }
  B200_REQUIRE(workspace_bytes >= b200vit_vq_workspace_size(R, D, K), "vq_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int DP = vq_dp(D);
  char* ws = (char*)workspace;
  float* ehat = (float*)ws;
  float* ee = (float*)(ws + align256((size_t)K * DP * 4));
  double* partial = (double*)(ws + align256((size_t)K * DP * 4) + align256((size_t)K * 4));
  VqLayout L{R, D, K, inner, elem_stride, outer_stride, flags};
  vq_prep_kernel<<<(K + 127) / 128, 128, 0, st>>>(codebook, K, D, DP, flags & 1, ehat, ee);
  B200_CUDA(cudaGetLastError());
  int rpc = vq_rows_per_cta(DP);
  // When the codebook fills an SM's shared memory (K = 4096: 213 KB, one CTA per SM whatever the register count) and
  // 4 rows per warp would need more than one wave of CTAs, give each warp 8 rows: half the CTAs, half the codebook
  // staging traffic, twice the independent FMA chains per lane (66 -> 57 us at rows 8192, K 4096).  With smaller
  // codebooks two CTAs share an SM at 4 rows per warp and the 182-register 8-row variant measured slower
  // (K 2048, rows 65536: 166 -> 200 us), so it is not used there.  Same arithmetic per (row, code): bit-exact indices.
  const int stride = (DP % 8 == 4) ? DP : DP + 4;
  const bool one_cta_per_sm = (size_t)K * (stride * 4 + 4) > 113 * 1024;
  // Round 2 (ncu: 2 warps per scheduler issue 0.47 instructions per clock, long-/short-scoreboard stalls, 128 of 148 SMs
  // busy): when the codebook pins one CTA per SM, run up to 16 warps of 4 rows each -- 4 warps per scheduler hide the
  // shared-memory and FMA-chain latencies -- and size the CTA (4..16 warps) so that one wave covers all SMs.  Same
  // arithmetic per (row, code) in the same order: bit-exact indices.
  const bool many_warps = (DP == 12 || DP == 16) && one_cta_per_sm;
  int threads = VQ_THREADS;
  bool rows8 = false;
  if (many_warps) {
    const long long per_wave = (long long)num_sms() * 4;                    // rows per wave with ONE warp per CTA
    const long long waves16 = (R + per_wave * VQ_MAX_WARPS - 1) / (per_wave * VQ_MAX_WARPS);
    long long w = (R + per_wave * waves16 - 1) / (per_wave * waves16);      // warps per CTA that fill the same number of waves
    if (w < 4) w = 4;
    if (w > VQ_MAX_WARPS) w = VQ_MAX_WARPS;
    threads = (int)w * 32;
    rpc = (int)w * 4;
  } else {
    rows8 = DP <= 16 && DP >= 12 && one_cta_per_sm && (R + rpc - 1) / rpc > num_sms();
    if (rows8) rpc *= 2;
  }
  const int grid = (int)((R + rpc - 1) / rpc);
  int rc;
  if (many_warps) {
    if (DP == 12) rc = launch_vq_fwd<12, 4, 512>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st, threads);
    else          rc = launch_vq_fwd<16, 4, 512>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st, threads);
    if (rc != OK) return rc;
    vq_finalize_kernel<<<1, 32, 0, st>>>(partial, grid, (double)R * D, commitment_cost, losses);
    B200_CUDA(cudaGetLastError());
    return OK;
  }
  switch (DP) {
    case 4:  rc = launch_vq_fwd<4, 4>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st); break;
    case 8:  rc = launch_vq_fwd<8, 4>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st); break;
    case 12: rc = rows8 ? launch_vq_fwd<12, 8>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st)
                        : launch_vq_fwd<12, 4>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st); break;
    case 16: rc = rows8 ? launch_vq_fwd<16, 8>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st)
                        : launch_vq_fwd<16, 4>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st); break;
    case 20: case 24: case 28: case 32:
      // pad to 32 is not possible without changing DP; dispatch exact sizes used in practice
      if (DP == 32) { rc = launch_vq_fwd<32, 2>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st); break; }
      if (DP == 24) { rc = launch_vq_fwd<24, 2>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st); break; }
      if (DP == 20) { rc = launch_vq_fwd<20, 2>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st); break; }
      rc = launch_vq_fwd<28, 2>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st); break;
    case 64: rc = launch_vq_fwd<64, 1>(x, codebook, ehat, ee, L, indices, quantized, partial, grid, st); break;
    default:
      set_error("vq_fwd: latent dim %d (padded %d) is not instantiated", D, DP);
      return ERR_UNSUPPORTED;
  }
  if (rc != OK) return rc;
  vq_finalize_kernel<<<1, 32, 0, st>>>(partial, grid, (double)R * D, commitment_cost, losses);
  B200_CUDA(cudaGetLastError());
  return OK;
}

int b200vit_vq_bwd(const float* x, const float* codebook, const long long* indices, const float* grad_q,
                   const float* coef, long long R, int D, int K, long long inner, long long elem_stride,
                   long long outer_stride, int flags, float* dx, float* dcodebook, void* stream) {
  B200_REQUIRE(x && codebook && indices && grad_q && coef && dx && dcodebook, "vq_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  VqLayout L{R, D, K, inner, elem_stride, outer_stride, flags};
  B200_CUDA(cudaMemsetAsync(dcodebook, 0, sizeof(float) * (size_t)K * D, st));
  vq_bwd_kernel<<<(int)((R + 127) / 128), 128, 0, st>>>(x, codebook, indices, grad_q, coef, L, dx, dcodebook);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"
