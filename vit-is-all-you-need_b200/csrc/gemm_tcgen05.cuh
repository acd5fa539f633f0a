// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma
// (fp32 accumulators in TMEM, double buffered) -> tcgen05.ld epilogue with fused element-wise work.
//
//   D[M,N] = sum_k A(m,k) * B(n,k)
//
// Operand storage ("major-ness") is a template parameter so that the three contractions of a
// Linear layer run without any transposed copies:
//   forward  y  = x  W^T : A = x  [M,K] K-major   B = W  [N,K] K-major    (A_MN=0, B_MN=0)
//   dgrad    dx = dy W   : A = dy [M,K] K-major   B = W  stored [K,N]      (A_MN=0, B_MN=1)
//   wgrad    dW = dy^T x : A = dy stored [K,M]    B = x  stored [K,N]      (A_MN=1, B_MN=1)
// "MN-major" operands are fetched as 64(MN) x 64(K) TMA boxes and described to the tensor core
// with the canonical MN-major SWIZZLE_128B layout (LBO = distance between 64-wide MN groups,
// SBO = distance between 8-row K groups).
//
// Warp roles (256 threads, 1 CTA/SM): warp0 = TMA producer, warp1 = MMA issuer (one thread),
// warp2 = TMEM allocator, warps 4..7 = epilogue (each owns 32 TMEM lanes = 32 output rows).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;

struct GemmShape {
  int M, N, K;
  int tiles_m, tiles_n;
  int kb_total;    // ceil(K / 64)
  int kb_per;      // k-blocks per split
  int splits;      // number of non-empty K splits
  int total_work;  // tiles_m * tiles_n * splits
  // descriptor strides for MN-major operands (bytes); defaults 8192 / 1024 / 2048
  int mn_lbo, mn_sbo, mn_kadv;
};

enum EpiKind : int {
  EPI_BF16 = 0,        // out(bf16) = acc + bias
  EPI_GELU_BF16 = 1,   // out2(bf16) = u = acc + bias (optional) ; out(bf16) = gelu(u)
  EPI_RESID_F32 = 2,   // out(f32) = aux(f32) + acc + bias
  EPI_DGELU_BF16 = 3,  // out(bf16) = acc * gelu'(aux(bf16))
  EPI_F32 = 4,         // out(f32) = acc + bias
  EPI_ATOMIC_F32 = 5,  // out(f32) += acc   (split-K wgrad)
  EPI_PATCH_F32 = 6,   // out(f32)[b*T + extra + p] = acc + bias + pos[p]   (row = b*P + p)
};

struct EpiParams {
  void* out;
  long long ldo;
  void* out2;
  long long ldo2;
  const float* bias;
  const void* aux;
  long long ldaux;
  const float* pos;  // [P, N] fp32
  int P, T, extra;
};

template <int KIND>
struct Epilogue {
  EpiParams p;

  // One thread owns output row `row`, columns [col, col+32); nvalid (multiple of 8) of them exist.
  __device__ __forceinline__ void operator()(int row, int col, const uint32_t (&acc)[32],
                                             int nvalid) const {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);

    if constexpr (KIND == EPI_BF16 || KIND == EPI_GELU_BF16 || KIND == EPI_RESID_F32 ||
                  KIND == EPI_F32 || KIND == EPI_PATCH_F32) {
      if (p.bias != nullptr) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (q * 4 < nvalid) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col) + q);
            v[q * 4 + 0] += b.x; v[q * 4 + 1] += b.y; v[q * 4 + 2] += b.z; v[q * 4 + 3] += b.w;
          }
        }
      }
    }

    if constexpr (KIND == EPI_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q * 8 < nvalid) {
          uint4 w;
          w.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]); w.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
          w.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]); w.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
          reinterpret_cast<uint4*>(o)[q] = w;
        }
      }
    } else if constexpr (KIND == EPI_GELU_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col;
      __nv_bfloat16* o2 =
          p.out2 ? reinterpret_cast<__nv_bfloat16*>(p.out2) + (long long)row * p.ldo2 + col : nullptr;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q * 8 < nvalid) {
          uint4 w;
          if (o2 != nullptr) {
            w.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]); w.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
            w.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]); w.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
            reinterpret_cast<uint4*>(o2)[q] = w;
          }
          float g[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // GELU is applied to the bf16-rounded pre-activation, which is what backward sees.
            const float u = __bfloat162float(__float2bfloat16_rn(v[q * 8 + j]));
            g[j] = gelu_erf(u);
          }
          w.x = pack_bf16(g[0], g[1]); w.y = pack_bf16(g[2], g[3]);
          w.z = pack_bf16(g[4], g[5]); w.w = pack_bf16(g[6], g[7]);
          reinterpret_cast<uint4*>(o)[q] = w;
        }
      }
    } else if constexpr (KIND == EPI_RESID_F32) {
      float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col;
      const float* r = reinterpret_cast<const float*>(p.aux) + (long long)row * p.ldaux + col;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q * 4 < nvalid) {
          const float4 a = reinterpret_cast<const float4*>(r)[q];
          float4 w;
          w.x = a.x + v[q * 4 + 0]; w.y = a.y + v[q * 4 + 1];
          w.z = a.z + v[q * 4 + 2]; w.w = a.w + v[q * 4 + 3];
          reinterpret_cast<float4*>(o)[q] = w;
        }
      }
    } else if constexpr (KIND == EPI_DGELU_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)row * p.ldo + col;
      const __nv_bfloat16* u =
          reinterpret_cast<const __nv_bfloat16*>(p.aux) + (long long)row * p.ldaux + col;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q * 8 < nvalid) {
          const uint4 uu = reinterpret_cast<const uint4*>(u)[q];
          const float2 u0 = unpack_bf16(uu.x), u1 = unpack_bf16(uu.y), u2 = unpack_bf16(uu.z),
                       u3 = unpack_bf16(uu.w);
          uint4 w;
          w.x = pack_bf16(v[q * 8 + 0] * gelu_erf_grad(u0.x), v[q * 8 + 1] * gelu_erf_grad(u0.y));
          w.y = pack_bf16(v[q * 8 + 2] * gelu_erf_grad(u1.x), v[q * 8 + 3] * gelu_erf_grad(u1.y));
          w.z = pack_bf16(v[q * 8 + 4] * gelu_erf_grad(u2.x), v[q * 8 + 5] * gelu_erf_grad(u2.y));
          w.w = pack_bf16(v[q * 8 + 6] * gelu_erf_grad(u3.x), v[q * 8 + 7] * gelu_erf_grad(u3.y));
          reinterpret_cast<uint4*>(o)[q] = w;
        }
      }
    } else if constexpr (KIND == EPI_F32) {
      float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q * 4 < nvalid) {
          reinterpret_cast<float4*>(o)[q] =
              make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
      }
    } else if constexpr (KIND == EPI_ATOMIC_F32) {
      float* o = reinterpret_cast<float*>(p.out) + (long long)row * p.ldo + col;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nvalid) atomicAdd(o + j, v[j]);
      }
    } else if constexpr (KIND == EPI_PATCH_F32) {
      const int b = row / p.P;
      const int pp = row - b * p.P;
      float* o = reinterpret_cast<float*>(p.out) + ((long long)b * p.T + p.extra + pp) * p.ldo + col;
      const float* pe = p.pos + (long long)pp * p.ldaux + col;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (q * 4 < nvalid) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(pe) + q);
          reinterpret_cast<float4*>(o)[q] =
              make_float4(v[q * 4 + 0] + a.x, v[q * 4 + 1] + a.y, v[q * 4 + 2] + a.z,
                          v[q * 4 + 3] + a.w);
        }
      }
    }
  }
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;       // 32 KB (BN=256) / 16 KB (BN=128)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;  // two accumulator stages (512 or 256 columns)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <bool A_MN, bool B_MN, int BN, int KIND>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a,
                    const __grid_constant__ CUtensorMap tma_b, const GemmShape s,
                    const Epilogue<KIND> epi) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < s.total_work; w += gridDim.x) {
        const int split = w % s.splits;
        const int t = w / s.splits;
        const int n0 = (t % s.tiles_n) * BN;
        const int m0 = (t / s.tiles_n) * GEMM_BM;
        const int kb0 = split * s.kb_per;
        const int kb1 = min(s.kb_total, kb0 + s.kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 100 + stage);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int k0 = kb * GEMM_BK;
          if constexpr (A_MN) {
#pragma unroll
            for (int i = 0; i < GEMM_BM / 64; ++i)
              tma_load_2d(sa + i * 8192, &tma_a, &full_bar[stage], m0 + i * 64, k0);
          } else {
            tma_load_2d(sa, &tma_a, &full_bar[stage], k0, m0);
          }
          if constexpr (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sb + i * 8192, &tma_b, &full_bar[stage], n0 + i * 64, k0);
          } else {
            tma_load_2d(sb, &tma_b, &full_bar[stage], k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < s.total_work; w += gridDim.x) {
        const int split = w % s.splits;
        const int kb0 = split * s.kb_per;
        const int kb1 = min(s.kb_total, kb0 + s.kb_per);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1, 200 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, 300 + stage);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_smem_desc(sa + k * s.mn_kadv, s.mn_lbo, s.mn_sbo)
                                     : umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_smem_desc(sb + k * s.mn_kadv, s.mn_lbo, s.mn_sbo)
                                     : umma_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------- epilogue -----------------------------------
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < s.total_work; w += gridDim.x) {
      const int t = w / s.splits;
      const int n0 = (t % s.tiles_n) * BN;
      const int m0 = (t / s.tiles_n) * GEMM_BM;
      mbar_wait(&tmem_full[acc], acc_phase, 400 + acc);
      tc_fence_after();
      const int row = m0 + ew * 32 + lane;
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        const int nvalid = min(32, s.N - (n0 + c));
        if (nvalid <= 0) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        tmem_ld_wait();
        if (row < s.M) epi(row, n0 + c, v, nvalid);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace b200
