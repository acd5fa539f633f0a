// Fused flash-style multi-head self-attention for head_dim = 64 (every config of the reference:
// transformer.py:56-58, blocks.py:219-233), forward and backward, on tcgen05 tensor cores.
//
// Replaces F.scaled_dot_product_attention(q, k, v, attn_mask=None | -inf upper triangle) at
// transformer.py:28 (and the SDPA inside nn.MultiheadAttention, blocks.py:60) together with the
// "b n (qkv h d) -> qkv b h n d" / "b h n d -> b n (h d)" rearranges (transformer.py:27,29): Q, K, V are
// read straight out of the packed QKV GEMM output [B, N, 3, H, 64] by strided TMA boxes and O is written as
// [B, N, H*64]; the [N, N] score matrix never touches HBM (the reference materialises an additive mask
// buffer [block, block], transformer.py:22-25).
//
// Forward, per CTA = (batch, head, 128-query tile), looping over 128-key blocks with online softmax:
//   S = Q K^T   tcgen05.mma, both operands K-major from TMA tiles, fp32 S in TMEM
//   softmax     4 warps, one query row per thread (TMEM lane == row), exp2 with fp32 running max / sum
//   O_j = P V   P written as bf16 to a 128B-swizzled smem tile (K-major A), V used as an MN-major B
// Backward, per CTA = (batch, head, 128-key block), looping over 128-query tiles:
//   S = Q K^T, dP = dO V^T, P = exp2(S - lse), dS = P (dP - D) scale,
//   dV += P^T dO, dK += dS^T Q (accumulated in TMEM), dQ = dS K (fp32 atomics, converted afterwards).
#include "../../include/b200vit.h"
#include "common.cuh"
#include "dropout.cuh"

namespace b200 {

constexpr int AT_HD = 64;
constexpr int AT_BQ = 128;
constexpr int AT_BK = 128;
constexpr int AT_THREADS = 192;        // warps 0-3: softmax/compute, warp 4: TMA, warp 5: MMA + TMEM alloc
constexpr int AT_BWD_THREADS = 320;    // streaming backward: + warps 6-9, a second compute warpgroup
constexpr int AT_TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16 = 16 KB
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnParams {
  int B, N, H, d;
  long long sb, sn;     // token (b, n) lives at row b*sb + n*sn  ([B,N,..]: N,1 ; sequence-first [N,B,..]: 1,B)
  int causal;
  float scale;          // 1/sqrt(64)
  float scale_log2e;    // scale * log2(e)
  __nv_bfloat16* o;     // [B, N, d]
  float* lse;           // [B, H, N]  natural-log LSE of the scaled scores
  // backward only
  const __nv_bfloat16* o_in;   // [B, N, d]
  const __nv_bfloat16* do_in;  // [B, N, d]
  float* dq_acc;               // [B, N, d] fp32, zeroed
  const float* dstat;          // [B, H, N] fp32: D = rowsum(dO o O), written by attn_bwd_dstat_kernel (streaming bwd)
  __nv_bfloat16* dqkv;         // [B, N, 3d]
  // dropout on the attention probabilities (dropout_p of SDPA, transformer.py:28); drop_thr == 0: none
  uint32_t drop_seed, drop_thr;
  float drop_r;                // 1 / (1 - p)
  int dbg_skip_dq;             // bring-up knob 8: the streaming backward skips its dQ atomics (timing experiments only)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// true in exactly one lane of a fully converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// K-major SWIZZLE_128B tile (rows of 128 bytes): descriptor for the 16-element K slice `k16` (0..3)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_addr, int k16) {
  return umma_smem_desc(tile_addr + k16 * 32, 16, 1024);
}
// MN-major view of the same bytes: K advances by 16 rows (2048 B); `lbo` = distance between 64-wide MN groups
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int k16, uint32_t lbo) {
  return umma_smem_desc(tile_addr + k16 * 2048, lbo, 1024);
}

// Writes 32 bf16 (already packed as 16 x b32) of row `row` into a [128 x 128] bf16 operand tile laid out as two
// 64-column slabs of 128 rows x 128 bytes with the TMA 128B swizzle; `c32` in [0,4) selects the 32-column chunk.
__device__ __forceinline__ void store_operand_chunk(uint8_t* tile, int row, int c32, const uint32_t (&w)[16]) {
  uint8_t* slab = tile + (c32 >> 1) * (128 * 128) + row * 128;
  const int chunk0 = (c32 & 1) * 4;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = (chunk0 + q) ^ (row & 7);
    *reinterpret_cast<uint4*>(slab + chunk * 16) = make_uint4(w[q * 4 + 0], w[q * 4 + 1], w[q * 4 + 2], w[q * 4 + 3]);
  }
}

// --------------------------------------------------------------------------------------------------
// Forward
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct FwdSmem {
  static constexpr int Q = 0;
  static constexpr int KV = AT_TILE_BYTES;                  // 2 stages x (K | V)
  static constexpr int P = KV + 4 * AT_TILE_BYTES;          // 128 x 128 bf16 = 32 KB
  static constexpr int BAR = P + 2 * AT_TILE_BYTES;
  static constexpr int TOTAL = BAR + 128;                   // 114,816 B -> two CTAs per SM
};

template <bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o,
                const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = at_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte aligned bases
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, hh = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * AT_BQ;
  int nkv = (p.N + AT_BK - 1) / AT_BK;
  if (CAUSAL) nkv = min(nkv, (q0 + AT_BQ + AT_BK - 1) / AT_BK);

  if (warp == 4 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1); mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
    // first loads are issued before the CTA-wide sync so their latency overlaps the TMEM allocation
    mbar_expect_tx(q_full, AT_TILE_BYTES);
    tma_load_3d(smem + FwdSmem::Q, &tm_qkv, q_full, hh * AT_HD, q0, b);
    for (int j = 0; j < min(nkv, 2); ++j) {
      uint8_t* kdst = smem + FwdSmem::KV + j * 2 * AT_TILE_BYTES;
      mbar_expect_tx(&kv_full[j], 2 * AT_TILE_BYTES);
      tma_load_3d(kdst, &tm_qkv, &kv_full[j], p.d + hh * AT_HD, j * AT_BK, b);
      tma_load_3d(kdst + AT_TILE_BYTES, &tm_qkv, &kv_full[j], 2 * p.d + hh * AT_HD, j * AT_BK, b);
    }
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 128 columns
  const uint32_t tmem_O = tmem_base + 128;  // 64 columns

  if (warp == 4) {
    if (lane == 0) {
      for (int j = 2; j < nkv; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1, 10);
        uint8_t* kdst = smem + FwdSmem::KV + st * 2 * AT_TILE_BYTES;
        mbar_expect_tx(&kv_full[st], 2 * AT_TILE_BYTES);
        tma_load_3d(kdst, &tm_qkv, &kv_full[st], p.d + hh * AT_HD, j * AT_BK, b);
        tma_load_3d(kdst + AT_TILE_BYTES, &tm_qkv, &kv_full[st], 2 * p.d + hh * AT_HD, j * AT_BK, b);
      }
    }
  } else if (warp == 5) {
    // warp-uniform control flow, one elected lane issues (descriptors stay in uniform registers; see the short kernels)
    const bool elected = elect_one();
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, AT_HD, false, true);
    const uint32_t sQ = smem_u32(smem + FwdSmem::Q);
    const uint32_t sP = smem_u32(smem + FwdSmem::P);
    mbar_wait(q_full, 0, 20);
    for (int j = 0; j < nkv; ++j) {
      const int st = j & 1;
      const uint32_t sK = smem_u32(smem + FwdSmem::KV + st * 2 * AT_TILE_BYTES);
      const uint32_t sV = sK + AT_TILE_BYTES;
      const int nkeys = min(AT_BK, p.N - j * AT_BK);     // valid keys of this block
      const int ncols = ((nkeys + 31) >> 5) << 5;        // S columns the softmax warps will read
      const uint32_t idesc_s = umma_idesc_bf16(128, ncols, false, false);
      const uint64_t dQ = desc_kmajor(sQ, 0), dK = desc_kmajor(sK, 0);
      mbar_wait(&kv_full[st], (j >> 1) & 1, 21);
      tc_fence_after();
      if (elected) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, dQ + 2 * k, dK + 2 * k, idesc_s, k > 0);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_ready, j & 1, 22);
      tc_fence_after();
      const int ksteps = (nkeys + 15) >> 4;              // P columns beyond the valid keys are zero / unused
      uint64_t dV = desc_mnmajor(sV, 0, 8192);
      uint32_t acc = 0u;
#pragma unroll 2
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t dP = desc_kmajor(sP + (k >> 2) * AT_TILE_BYTES, k & 3);
        if (elected) umma_bf16(tmem_O, dP, dV, idesc_o, acc);
        dV += 128; acc = 1u;
      }
      if (elected) { umma_commit(o_full); umma_commit(&kv_empty[st]); }
      __syncwarp();
    }
  } else {
    // ---- softmax warps: thread <-> query row ----
    const int r = threadIdx.x;  // 0..127 == TMEM lane
    const int q = q0 + r;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const bool warp_has_rows = q0 + warp * 32 < p.N;  // warp-uniform: rows beyond the sequence do no math
    float m = -INFINITY, l = 0.f;
    float oacc[AT_HD];
#pragma unroll
    for (int i = 0; i < AT_HD; ++i) oacc[i] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      if (!warp_has_rows) {  // stay in phase with the pipeline, contribute nothing
        mbar_wait(s_full, j & 1, 33);
        mbar_arrive(p_ready);
        continue;
      }
      const int nch = (min(AT_BK, p.N - j * AT_BK) + 31) >> 5;  // 32-key chunks holding at least one valid key
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1, 30);
        tc_fence_after();
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld32(tmem_O + lane_off + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) oacc[c * 32 + i] += __uint_as_float(v[i]);
        }
      }
      mbar_wait(s_full, j & 1, 31);
      tc_fence_after();
      const bool need_mask = ((j + 1) * AT_BK > p.N) || (CAUSAL && (j + 1) * AT_BK > q0 + 1);
      auto masked_score = [&](uint32_t bits, int key) -> float {
        const bool ok = key < p.N && (!CAUSAL || key <= q);
        return ok ? __uint_as_float(bits) : -INFINITY;
      };
      // probabilities of one 32-key chunk against the reference m_use -> bf16 operand chunk in shared memory
      auto exp_chunk = [&](const uint32_t (&v)[32], int c, float m_use, float& rowsum) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(v[i]), p.scale_log2e, -m_use));
          float p1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), p.scale_log2e, -m_use));
          if (need_mask) {
            const int key = j * AT_BK + c * 32 + i;
            if (!(key < p.N && (!CAUSAL || key <= q))) p0 = 0.f;
            if (!(key + 1 < p.N && (!CAUSAL || key + 1 <= q))) p1 = 0.f;
          }
          rowsum += p0 + p1;
          if (p.drop_thr != 0) {  // the row sum (softmax denominator) is taken before dropout
            const int key = j * AT_BK + c * 32 + i;
            const uint32_t h = drop_hash_attn(p.drop_seed, (uint32_t)(b * p.H + hh), (uint32_t)q, (uint32_t)(key >> 1));
            if (!drop_keep(h, 0, p.drop_thr)) p0 = 0.f;
            if (!drop_keep(h, 1, p.drop_thr)) p1 = 0.f;
          }
          w[i >> 1] = pack_bf16(p0, p1);
        }
        store_operand_chunk(smem + FwdSmem::P, r, c, w);
      };
      auto chunk_max = [&](const uint32_t (&v)[32], int c) -> float {
        float mx = -INFINITY;
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, masked_score(v[i], j * AT_BK + c * 32 + i));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        return mx * p.scale_log2e;
      };
      // ONE pass over S (round 2; the kernel is bound by the TMEM read port, 64 B/clk/SM, and the separate row-maximum
      // pass read every score twice): the reference of this key block is max(running reference, maximum of the block's
      // FIRST 32-key chunk); later chunks are exponentiated against it as long as no row of the warp exceeds it by more
      // than 2^TAU (P <= 256 then, and O / l does not depend on the reference).  S is NOT overwritten here (P goes to
      // shared memory), so in the rare other case the block is simply redone with the exact two-pass scheme below.
      constexpr float TAU = 8.0f;
      float m_new = m, m_use = 0.f, alpha = 1.f, rowsum = 0.f;
      bool redo = false;
      {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_off, v);
        tmem_ld_wait();
        m_new = fmaxf(m, chunk_max(v, 0));
        m_use = (m_new == -INFINITY) ? 0.f : m_new;
        exp_chunk(v, 0, m_use, rowsum);
#pragma unroll 1
        for (int c = 1; c < nch; ++c) {
          tmem_ld32(tmem_S + lane_off + c * 32, v);
          tmem_ld_wait();
          if (__any_sync(0xffffffffu, chunk_max(v, c) > m_use + TAU)) { redo = true; break; }
          exp_chunk(v, c, m_use, rowsum);
        }
      }
      if (redo) {   // exact two-pass softmax of this block (warp-uniform branch)
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_S + lane_off + c * 32, v);
          tmem_ld_wait();
          mx = fmaxf(mx, chunk_max(v, c));
        }
        m_new = fmaxf(m, mx);
        m_use = (m_new == -INFINITY) ? 0.f : m_new;
        rowsum = 0.f;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_S + lane_off + c * 32, v);
          tmem_ld_wait();
          exp_chunk(v, c, m_use, rowsum);
        }
      }
      alpha = ex2_approx(m - m_use);
      l *= alpha;
#pragma unroll
      for (int i = 0; i < AT_HD; ++i) oacc[i] *= alpha;
      l += rowsum;
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // last PV
    if (warp_has_rows) {
    mbar_wait(o_full, (nkv - 1) & 1, 32);
    tc_fence_after();
    {
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tmem_ld32(tmem_O + lane_off + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) oacc[c * 32 + i] += __uint_as_float(v[i]);
      }
    }
    {
      // the P tile is dead after the last PV MMA: stage O there (swizzled) and leave through one TMA store per warp;
      // rows beyond the sequence are clipped by the tensor map
      const float inv = p.drop_r / l;   // kept probabilities are scaled by 1 / (1 - p)
      uint8_t* rowp = smem + FwdSmem::P + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint4 w;
        w.x = pack_bf16(oacc[c * 8 + 0] * inv, oacc[c * 8 + 1] * inv);
        w.y = pack_bf16(oacc[c * 8 + 2] * inv, oacc[c * 8 + 3] * inv);
        w.z = pack_bf16(oacc[c * 8 + 4] * inv, oacc[c * 8 + 5] * inv);
        w.w = pack_bf16(oacc[c * 8 + 6] * inv, oacc[c * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(rowp + ((c ^ (r & 7)) << 4)) = w;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tm_o, smem + FwdSmem::P + warp * 4096, hh * AT_HD, q0 + warp * 32, b);
        bulk_commit();
      }
      if (q < p.N && p.lse != nullptr) p.lse[((long long)b * p.H + hh) * p.N + q] = m * LN2 + logf(l);
      if (lane == 0) bulk_wait_all();   // shared memory must stay valid until the store has read it
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// --------------------------------------------------------------------------------------------------
// Forward for short sequences (N <= 256, no causal mask: the ViT token counts 65 / 197 and every other
// encoder of the reference whose sequence fits two 128-row tiles).  Persistent, one CTA per SM, a work unit is
// one (batch, head):
//   * Q, K, V of the whole head are staged once per unit (TMA, 2-deep ring: the next unit's tiles land while
//     this one computes);
//   * the score rows of a 128-query tile fit TMEM completely (<= 256 fp32 columns), so softmax is exact and
//     single-shot: no running max, no rescaling of O, no second key block;
//   * P is written back to TMEM as packed bf16 over the dead S columns and consumed by the tensor core directly
//     (tcgen05.mma with the A operand in TMEM), so it never touches shared memory;
//   * two warpgroups own the two query tiles (TMEM columns [0,256) and [256,512)), so one tile's softmax overlaps
//     the other tile's MMAs.
// TMEM columns of tile t (base 256 t): S fp32 [0, ncols), P bf16 [0, ncols/2) (overwrites S behind the reads),
// O fp32 [128, 192) (S is dead by then).
// --------------------------------------------------------------------------------------------------
// One warp: 32 rows x 64 fp32 TMEM columns (scaled per row) -> bf16 -> swizzled staging rows -> one TMA store
// (box 64 x 32; rows beyond the sequence are clipped by the tensor map).
__device__ __forceinline__ void bwd2_store_tile(uint32_t tsrc, float sc, uint8_t* stage_rows, bool write_ok,
                                                const CUtensorMap* tm_out, int col, int row, int b, int lane) {
  if (lane == 0) bulk_wait_read();  // this warp's previous store has finished reading its staging rows
  __syncwarp();
  uint32_t v0[32], v1[32];
  tmem_ld32(tsrc, v0);
  tmem_ld32(tsrc + 32, v1);
  tmem_ld_wait();
  if (write_ok) {
    uint8_t* rowp = stage_rows + lane * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 w;
      w.x = pack_bf16(__uint_as_float(v0[c * 8 + 0]) * sc, __uint_as_float(v0[c * 8 + 1]) * sc);
      w.y = pack_bf16(__uint_as_float(v0[c * 8 + 2]) * sc, __uint_as_float(v0[c * 8 + 3]) * sc);
      w.z = pack_bf16(__uint_as_float(v0[c * 8 + 4]) * sc, __uint_as_float(v0[c * 8 + 5]) * sc);
      w.w = pack_bf16(__uint_as_float(v0[c * 8 + 6]) * sc, __uint_as_float(v0[c * 8 + 7]) * sc);
      *reinterpret_cast<uint4*>(rowp + ((c ^ (lane & 7)) << 4)) = w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 w;
      w.x = pack_bf16(__uint_as_float(v1[c * 8 + 0]) * sc, __uint_as_float(v1[c * 8 + 1]) * sc);
      w.y = pack_bf16(__uint_as_float(v1[c * 8 + 2]) * sc, __uint_as_float(v1[c * 8 + 3]) * sc);
      w.z = pack_bf16(__uint_as_float(v1[c * 8 + 4]) * sc, __uint_as_float(v1[c * 8 + 5]) * sc);
      w.w = pack_bf16(__uint_as_float(v1[c * 8 + 6]) * sc, __uint_as_float(v1[c * 8 + 7]) * sc);
      *reinterpret_cast<uint4*>(rowp + (((4 + c) ^ (lane & 7)) << 4)) = w;
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_3d(tm_out, stage_rows, col, row, b);
    bulk_commit();
  }
}

constexpr int AS_THREADS = 384;  // warp 0: TMA, warp 1: MMA issue, warp 2: TMEM alloc, warps 4-7 / 8-11: softmax of query tile 0 / 1
struct FwdShortSmem {
  static constexpr int STAGE = 6 * AT_TILE_BYTES;          // Q0 Q1 | K (256 rows) | V (256 rows) = 96 KB
  static constexpr int Q = 0, K = 2 * AT_TILE_BYTES, V = 4 * AT_TILE_BYTES;
  static constexpr int OST = 2 * STAGE;                    // O staging: one [128 x 128 B] tile per query tile
  static constexpr int BAR = OST + 2 * AT_TILE_BYTES;
  static constexpr int TOTAL = BAR + 256;
};

__device__ __forceinline__ void tmem_st16_packed(uint32_t taddr, const uint32_t (&v)[16]) { tmem_st16(taddr, v); }

template <bool DROP>
__global__ void __launch_bounds__(AS_THREADS, 1)
attn_fwd_short_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o,
                      const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = at_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdShortSmem::BAR);
  uint64_t* full = bars + 0;      // [2] stage loaded (TMA tx)
  uint64_t* empty = bars + 2;     // [2] stage consumed (MMA commit)
  uint64_t* s_full = bars + 4;    // [2] S of tile t complete (MMA commit)
  uint64_t* p_ready = bars + 6;   // [2] P of tile t in TMEM (4 warp arrivals)
  uint64_t* o_full = bars + 8;    // [2] O of tile t complete (MMA commit)
  uint64_t* o_free = bars + 10;   // [2] O of tile t read out (4 warp arrivals) -> TMEM columns reusable
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = p.B * p.H;
  const int ntiles = (p.N + 127) >> 7;                 // 1 or 2 query tiles
  const int ncols = ((p.N + 15) >> 4) << 4;            // S columns = keys padded to the MMA N granularity
  const int kv_boxes = (p.N + 127) >> 7;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1); mbar_init(&empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4);
      mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // the set-up above overlapped the previous kernel's tail
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int n = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++n) {
        const int st = n & 1;
        const int b = u / p.H, hh = u - b * p.H;
        mbar_wait(&empty[st], ((n >> 1) & 1) ^ 1, 40);
        uint8_t* base = smem + st * FwdShortSmem::STAGE;
        mbar_expect_tx(&full[st], (ntiles + 2 * kv_boxes) * AT_TILE_BYTES);
        for (int t = 0; t < kv_boxes; ++t) {  // K first: the S MMAs need Q and K, V only later
          tma_load_3d(base + FwdShortSmem::K + t * AT_TILE_BYTES, &tm_qkv, &full[st], p.d + hh * AT_HD, t * 128, b);
        }
        for (int t = 0; t < ntiles; ++t)
          tma_load_3d(base + FwdShortSmem::Q + t * AT_TILE_BYTES, &tm_qkv, &full[st], hh * AT_HD, t * 128, b);
        for (int t = 0; t < kv_boxes; ++t)
          tma_load_3d(base + FwdShortSmem::V + t * AT_TILE_BYTES, &tm_qkv, &full[st], 2 * p.d + hh * AT_HD, t * 128, b);
      }
    }
  } else if (warp == 1) {
    // whole warp runs the control flow (descriptor arithmetic stays in uniform registers), one elected lane issues
    const bool elected = elect_one();
    const uint32_t idesc_s = umma_idesc_bf16(128, ncols, false, false);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, AT_HD, false, true);
    const int ksteps = ncols >> 4;
    // Issue order in steady state: S0(n), PV1(n-1), S1(n), PV0(n).  The two warpgroups then run half a period
    // apart, so one tile's softmax (MUFU-bound) overlaps the other tile's PV MMA, O read-out and hand-offs.
    // (One issuing thread per tile was measured slower: the warpgroups fall into lock-step and share the MUFU.)
    auto issue_s = [&](int n, int t) {
      const uint32_t sbase = smem_u32(smem + (n & 1) * FwdShortSmem::STAGE);
      const uint64_t dQ = desc_kmajor(sbase + FwdShortSmem::Q + t * AT_TILE_BYTES, 0);
      const uint64_t dK = desc_kmajor(sbase + FwdShortSmem::K, 0);
      mbar_wait(&o_free[t], (n & 1) ^ 1, 42);  // previous unit's O (same TMEM columns) has been read
      tc_fence_after();
      if (elected) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + t * 256, dQ + 2 * k, dK + 2 * k, idesc_s, k > 0);
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int n, int t) {
      uint64_t dV = desc_mnmajor(smem_u32(smem + (n & 1) * FwdShortSmem::STAGE) + FwdShortSmem::V, 0, 8192);
      uint32_t aP = tmem_base + t * 256;
      mbar_wait(&p_ready[t], n & 1, 43);
      tc_fence_after();
      uint32_t acc = 0u;
#pragma unroll 2
      for (int k = 0; k < ksteps; ++k) {
        if (elected) umma_bf16_ts(tmem_base + t * 256 + 128, aP, dV, idesc_o, acc);
        aP += 8; dV += 128; acc = 1u;     // 16 keys: 8 packed TMEM columns of P, 16 rows x 128 B of V
      }
      if (elected) umma_commit(&o_full[t]);
      __syncwarp();
    };
    int n = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++n) {
      mbar_wait(&full[n & 1], (n >> 1) & 1, 41);
      issue_s(n, 0);
      if (ntiles == 2) {
        if (n > 0) {
          issue_pv(n - 1, 1);
          if (elected) umma_commit(&empty[(n - 1) & 1]);  // stage n-1 fully consumed
          __syncwarp();
        }
        issue_s(n, 1);
        issue_pv(n, 0);
      } else {
        issue_pv(n, 0);
        if (elected) umma_commit(&empty[n & 1]);
        __syncwarp();
      }
    }
    if (ntiles == 2 && n > 0) {
      issue_pv(n - 1, 1);
      if (elected) umma_commit(&empty[(n - 1) & 1]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int t = (warp - 4) >> 2;        // query tile of this warpgroup
    const int wq = warp & 3;              // TMEM lane quarter == warp % 4
    if (t < ntiles) {
      const int r = wq * 32 + lane;       // row within the tile == TMEM lane
      const int q = t * 128 + r;
      const bool warp_has_rows = t * 128 + wq * 32 < p.N;
      const uint32_t tS = tmem_base + t * 256 + (static_cast<uint32_t>(wq * 32) << 16);
      const int nch = ncols >> 5;         // full 32-column chunks
      const bool tail16 = (ncols & 16) != 0;
      int n = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++n) {
        const int b = u / p.H, hh = u - b * p.H;
        const uint32_t bh = (uint32_t)u;   // == b * H + hh
        (void)bh;
        mbar_wait(&s_full[t], n & 1, 44);
        tc_fence_after();
        float l = 0.f, m2 = 0.f;
        if (warp_has_rows) {
          // SINGLE pass over S (the kernel is bound by TMEM read bandwidth, ~64 B/clk/SM: a separate row-maximum pass
          // costs as much as the exponentials).  The reference maximum is the maximum of the FIRST 32-key chunk; later
          // chunks are exponentiated against it as long as no row of the warp exceeds it by more than 2^TAU (P then
          // stays <= 256, exact in bf16's range, and O / l is invariant to the choice of the reference).  In the rare
          // other case the chunks already written are rescaled in place and the reference is raised (lazy rescaling).
          constexpr float TAU = 8.0f;
          uint32_t va[32], vb[32];
          auto chunk_max = [&](const uint32_t (&x)[32], int c) -> float {
            float mx = -INFINITY;
            if ((c + 1) * 32 <= p.N) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(x[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, c * 32 + i < p.N ? __uint_as_float(x[i]) : -INFINITY);
            }
            return mx * p.scale_log2e;
          };
          // raise the reference to cover `cm` (per row) and rescale the `nw` 16-column P pieces already in TMEM
          auto raise_reference = [&](float cm, int nw) {
            const float m_new = fmaxf(m2, cm);
            const float alpha = ex2_approx(m2 - m_new);   // 1 for the rows whose maximum did not move
            l *= alpha;
            m2 = m_new;
            tmem_st_wait();
            for (int cc = 0; cc < nw; ++cc) {
              uint32_t w[16];
              tmem_ld16(tS + cc * 16, w);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float2 f = unpack_bf16(w[i]);
                w[i] = pack_bf16(f.x * alpha, f.y * alpha);
              }
              tmem_st16(tS + cc * 16, w);
            }
          };
          // pass 2: probabilities -> packed bf16 written over the S columns already consumed
          auto exp_chunk = [&](const uint32_t (&x)[32], int c) {
            uint32_t w[16];
            if ((c + 1) * 32 <= p.N) {  // warp-uniform: the hot path carries no masking code
#pragma unroll
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(x[i]), p.scale_log2e, -m2));
                float p1 = ex2_approx(fmaf(__uint_as_float(x[i + 1]), p.scale_log2e, -m2));
                l += p0 + p1;
                if constexpr (DROP) {  // the denominator is taken before dropout; 1 / (1 - p) is applied to O
                  const uint32_t h = drop_hash_attn(p.drop_seed, bh, (uint32_t)q, (uint32_t)((c * 32 + i) >> 1));
                  if (!drop_keep(h, 0, p.drop_thr)) p0 = 0.f;
                  if (!drop_keep(h, 1, p.drop_thr)) p1 = 0.f;
                }
                w[i >> 1] = pack_bf16(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(x[i]), p.scale_log2e, -m2));
                float p1 = ex2_approx(fmaf(__uint_as_float(x[i + 1]), p.scale_log2e, -m2));
                if (c * 32 + i >= p.N) p0 = 0.f;
                if (c * 32 + i + 1 >= p.N) p1 = 0.f;
                l += p0 + p1;
                if constexpr (DROP) {
                  const uint32_t h = drop_hash_attn(p.drop_seed, bh, (uint32_t)q, (uint32_t)((c * 32 + i) >> 1));
                  if (!drop_keep(h, 0, p.drop_thr)) p0 = 0.f;
                  if (!drop_keep(h, 1, p.drop_thr)) p1 = 0.f;
                }
                w[i >> 1] = pack_bf16(p0, p1);
              }
            }
            tmem_st16(tS + c * 16, w);
          };
          auto step = [&](const uint32_t (&x)[32], int c) {
            const float cm = chunk_max(x, c);
            if (c == 0) m2 = cm;
            else if (__any_sync(0xffffffffu, cm > m2 + TAU)) raise_reference(cm, c * 2);
            exp_chunk(x, c);
          };
          if (nch > 0) tmem_ld32(tS, va);
          for (int c = 0; c < nch; c += 2) {
            tmem_ld_wait();
            if (c + 1 < nch) tmem_ld32(tS + (c + 1) * 32, vb);
            step(va, c);
            if (c + 1 < nch) {
              tmem_ld_wait();
              if (c + 2 < nch) tmem_ld32(tS + (c + 2) * 32, va);
              step(vb, c + 1);
            }
          }
          if (tail16) {
            uint32_t x[16], w[16];
            tmem_ld16(tS + nch * 32, x);
            tmem_ld_wait();
            {
              float mx = -INFINITY;
#pragma unroll
              for (int i = 0; i < 16; ++i) mx = fmaxf(mx, nch * 32 + i < p.N ? __uint_as_float(x[i]) : -INFINITY);
              const float cm = mx * p.scale_log2e;
              if (nch == 0) m2 = cm;
              else if (__any_sync(0xffffffffu, cm > m2 + TAU)) raise_reference(cm, nch * 2);
            }
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              float p0 = ex2_approx(fmaf(__uint_as_float(x[i]), p.scale_log2e, -m2));
              float p1 = ex2_approx(fmaf(__uint_as_float(x[i + 1]), p.scale_log2e, -m2));
              if (nch * 32 + i >= p.N) p0 = 0.f;
              if (nch * 32 + i + 1 >= p.N) p1 = 0.f;
              l += p0 + p1;
              if constexpr (DROP) {
                const uint32_t h = drop_hash_attn(p.drop_seed, bh, (uint32_t)q, (uint32_t)((nch * 32 + i) >> 1));
                if (!drop_keep(h, 0, p.drop_thr)) p0 = 0.f;
                if (!drop_keep(h, 1, p.drop_thr)) p1 = 0.f;
              }
              w[i >> 1] = pack_bf16(p0, p1);
            }
#pragma unroll
            for (int i = 8; i < 16; ++i) w[i] = 0u;
            // 8 packed columns; the x16 store also clears 8 columns beyond ncols/2, which nothing reads
            tmem_st16(tS + nch * 16, w);
          }
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
        // ---- O = P V is being computed; then normalise and store ----
        mbar_wait(&o_full[t], n & 1, 45);
        tc_fence_after();
        if (warp_has_rows) {
          // normalise, stage as bf16 (swizzled) and leave through one TMA store per warp
          bwd2_store_tile(tS + 128, (DROP ? p.drop_r : 1.0f) / l, smem + FwdShortSmem::OST + t * AT_TILE_BYTES + wq * 4096, true, &tm_o,
                          hh * AT_HD, t * 128 + wq * 32, b, lane);
          if (q < p.N && p.lse != nullptr) p.lse[((long long)b * p.H + hh) * p.N + q] = m2 * LN2 + logf(l);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[t]);
      }
      if (lane == 0) bulk_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// --------------------------------------------------------------------------------------------------
// Backward
// --------------------------------------------------------------------------------------------------
struct BwdSmem {
  static constexpr int K = 0;
  static constexpr int V = AT_TILE_BYTES;
  static constexpr int QDO = 2 * AT_TILE_BYTES;             // 2 stages x (Q | dO)
  static constexpr int P = QDO + 4 * AT_TILE_BYTES;         // 32 KB
  static constexpr int DS = P + 2 * AT_TILE_BYTES;          // 32 KB
  static constexpr int BAR = DS + 2 * AT_TILE_BYTES;
  static constexpr int TOTAL = BAR + 128 + 1024;
};

template <bool CAUSAL>
__global__ void __launch_bounds__(AT_BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const AttnParams p) {
  extern __shared__ uint8_t at_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(at_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;    // [2]
  uint64_t* q_empty = bars + 3;   // [2]
  uint64_t* sdp_full = bars + 5;
  uint64_t* pds_ready = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dq_free = bars + 8;
  uint64_t* dkv_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jb = blockIdx.x, hh = blockIdx.y, b = blockIdx.z;
  const int k0 = jb * AT_BK;
  const int nq = (p.N + AT_BQ - 1) / AT_BQ;
  const int i0 = CAUSAL ? jb : 0;  // first query tile that sees this key block
  const int ntiles = nq - i0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(kv_full, 1);
    mbar_init(&q_full[0], 1); mbar_init(&q_full[1], 1);
    mbar_init(&q_empty[0], 1); mbar_init(&q_empty[1], 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_ready, 256);   // both compute warpgroups
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 256);
    mbar_init(dkv_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dV = tmem_base + 256,
                 tmem_dK = tmem_base + 320, tmem_dQ = tmem_base + 384;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * AT_TILE_BYTES);
      tma_load_3d(smem + BwdSmem::K, &tm_qkv, kv_full, p.d + hh * AT_HD, k0, b);
      tma_load_3d(smem + BwdSmem::V, &tm_qkv, kv_full, 2 * p.d + hh * AT_HD, k0, b);
      for (int it = 0; it < ntiles; ++it) {
        const int st = it & 1;
        mbar_wait(&q_empty[st], ((it >> 1) & 1) ^ 1, 40);
        uint8_t* dst = smem + BwdSmem::QDO + st * 2 * AT_TILE_BYTES;
        mbar_expect_tx(&q_full[st], 2 * AT_TILE_BYTES);
        tma_load_3d(dst, &tm_qkv, &q_full[st], hh * AT_HD, (i0 + it) * AT_BQ, b);
        tma_load_3d(dst + AT_TILE_BYTES, &tm_do, &q_full[st], hh * AT_HD, (i0 + it) * AT_BQ, b);
      }
    }
  } else if (warp == 5) {
    const bool elected = elect_one();   // warp-uniform control flow, one elected lane issues
    // ragged last key block (N = 288: 32 of 128 keys): S / dP are only computed for the 32-key chunks that hold a valid
    // key and dQ only contracts over the valid keys; ragged last query tile: dV / dK only contract over the valid
    // query rows.  Columns / rows beyond are never read (dK / dV rows of invalid keys are not written out).
    const int nkeys = min(AT_BK, p.N - k0);
    const int ncols = ((nkeys + 31) >> 5) << 5;
    const int ksteps_dq = (nkeys + 15) >> 4;
    const uint32_t idesc_s = umma_idesc_bf16(128, ncols, false, false);       // S, dP: K-major x K-major
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, AT_HD, true, true);     // dV, dK: MN-major x MN-major
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, AT_HD, false, true);    // dQ: K-major x MN-major
    const uint32_t sK = smem_u32(smem + BwdSmem::K), sV = smem_u32(smem + BwdSmem::V);
    const uint32_t sP = smem_u32(smem + BwdSmem::P), sDS = smem_u32(smem + BwdSmem::DS);
    mbar_wait(kv_full, 0, 50);
    for (int it = 0; it < ntiles; ++it) {
      const int st = it & 1;
      const uint32_t sQ = smem_u32(smem + BwdSmem::QDO + st * 2 * AT_TILE_BYTES);
      const uint32_t sDO = sQ + AT_TILE_BYTES;
      mbar_wait(&q_full[st], (it >> 1) & 1, 51);
      tc_fence_after();
      if (elected) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // independent accumulators interleaved
          umma_bf16(tmem_S, desc_kmajor(sQ, k), desc_kmajor(sK, k), idesc_s, k > 0);
          umma_bf16(tmem_dP, desc_kmajor(sDO, k), desc_kmajor(sV, k), idesc_s, k > 0);
        }
        umma_commit(sdp_full);
      }
      __syncwarp();
      mbar_wait(pds_ready, it & 1, 52);
      tc_fence_after();
      // dV[key, d] += P^T dO ; dK[key, d] += dS^T Q   (contraction over the 128 query rows)
      const int qsteps = (min(AT_BQ, p.N - (i0 + it) * AT_BQ) + 15) >> 4;   // 16-row slabs holding a valid query
      if (elected) {
#pragma unroll 2
        for (int k = 0; k < qsteps; ++k) {
          umma_bf16(tmem_dV, desc_mnmajor(sP, k, 16384), desc_mnmajor(sDO, k, 8192), idesc_t, (it > 0 || k > 0) ? 1u : 0u);
          umma_bf16(tmem_dK, desc_mnmajor(sDS, k, 16384), desc_mnmajor(sQ, k, 8192), idesc_t, (it > 0 || k > 0) ? 1u : 0u);
        }
      }
      __syncwarp();
      if (it > 0) { mbar_wait(dq_free, (it - 1) & 1, 53); tc_fence_after(); }
      // dQ[q, d] = dS K   (contraction over the 128 keys)
      if (elected) {
#pragma unroll 2
        for (int k = 0; k < ksteps_dq; ++k)
          umma_bf16(tmem_dQ, desc_kmajor(sDS + (k >> 2) * AT_TILE_BYTES, k & 3), desc_mnmajor(sK, k, 8192), idesc_q, k > 0);
        umma_commit(dq_full);
        umma_commit(&q_empty[st]);
      }
      __syncwarp();
    }
    if (elected) umma_commit(dkv_full);
    __syncwarp();
  } else {
    // two compute warpgroups (warps 0-3 and 6-9) split the 128 key columns of every (key block, query tile) pair:
    // group g works on the 32-column chunks 2g, 2g+1, reads out dQ chunk g and, at the end, dK (g = 0) / dV (g = 1)
    const int grp = warp < 4 ? 0 : 1;
    const int quarter = warp & 3;                 // TMEM lane quarter == warp % 4
    const int r = quarter * 32 + lane;            // query row within the tile
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const int nch = (min(AT_BK, p.N - k0) + 31) >> 5;   // 32-key chunks holding at least one valid key
    for (int it = 0; it < ntiles; ++it) {
      const int q = (i0 + it) * AT_BQ + r;
      const bool qv = q < p.N;
      float lse2 = 0.f, Dq = 0.f;
      if (qv) {  // two scalar loads per row, issued before the S / dP MMAs are waited for (D comes from the pre-pass)
        lse2 = __ldg(p.lse + ((long long)b * p.H + hh) * p.N + q) * LOG2E;
        Dq = __ldg(p.dstat + ((long long)b * p.H + hh) * p.N + q);
      }
      mbar_wait(sdp_full, it & 1, 60);
      tc_fence_after();
#pragma unroll 1
      for (int c = 2 * grp; c < min(2 * grp + 2, nch); ++c) {
        uint32_t s[32], dp[32];
        tmem_ld32(tmem_S + lane_off + c * 32, s);
        tmem_ld32(tmem_dP + lane_off + c * 32, dp);
        tmem_ld_wait();
        uint32_t wp[16], wd[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float pv[2], dv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int key = k0 + c * 32 + i + e;
            const bool ok = qv && key < p.N && (!CAUSAL || key <= q);
            const float pe = ok ? ex2_approx(fmaf(__uint_as_float(s[i + e]), p.scale_log2e, -lse2)) : 0.f;
            float pk = pe, dpe = __uint_as_float(dp[i + e]);
            if (p.drop_thr != 0) {  // dV uses the dropped probabilities, dP = mask / (1 - p) * (dO V^T)
              const bool keep = drop_keep(drop_hash_attn(p.drop_seed, (uint32_t)(b * p.H + hh), (uint32_t)q, (uint32_t)(key >> 1)),
                                          key & 1, p.drop_thr);
              pk = keep ? pe : 0.f;
              dpe = keep ? dpe * p.drop_r : 0.f;
            }
            pv[e] = pk;
            dv[e] = pe * (dpe - Dq) * p.scale;
          }
          wp[i >> 1] = pack_bf16(pv[0], pv[1]);
          wd[i >> 1] = pack_bf16(dv[0], dv[1]);
        }
        store_operand_chunk(smem + BwdSmem::P, r, c, wp);
        store_operand_chunk(smem + BwdSmem::DS, r, c, wd);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_ready);

      mbar_wait(dq_full, it & 1, 61);
      tc_fence_after();
      {
        const int c = grp;
        uint32_t v[32];
        tmem_ld32(tmem_dQ + lane_off + c * 32, v);
        tmem_ld_wait();
        if (qv && !p.dbg_skip_dq) {
          float* dst = p.dq_acc + ((long long)b * p.sb + q * p.sn) * p.d + hh * AT_HD + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            atomicAdd(reinterpret_cast<float4*>(dst + i),
                      make_float4(__uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
        }
      }
      tc_fence_before();
      mbar_arrive(dq_free);
    }
    // dK / dV of this key block
    mbar_wait(dkv_full, 0, 62);
    tc_fence_after();
    const int key = k0 + r;
    {
      const int which = grp;  // 0: dK, 1: dV
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32((which == 0 ? tmem_dK : tmem_dV) + lane_off + c * 32, v);
        tmem_ld_wait();
        if (key < p.N) {
          const float sc = which == 1 ? p.drop_r : 1.0f;   // dV = (mask P / (1 - p))^T dO
          __nv_bfloat16* dst = p.dqkv + ((long long)b * p.sb + key * p.sn) * 3 * p.d + (which + 1) * p.d + hh * AT_HD + c * 32;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(v[i * 8 + 0]) * sc, __uint_as_float(v[i * 8 + 1]) * sc);
            w.y = pack_bf16(__uint_as_float(v[i * 8 + 2]) * sc, __uint_as_float(v[i * 8 + 3]) * sc);
            w.z = pack_bf16(__uint_as_float(v[i * 8 + 4]) * sc, __uint_as_float(v[i * 8 + 5]) * sc);
            w.w = pack_bf16(__uint_as_float(v[i * 8 + 6]) * sc, __uint_as_float(v[i * 8 + 7]) * sc);
            reinterpret_cast<uint4*>(dst)[i] = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// --------------------------------------------------------------------------------------------------
// Backward for short sequences (N <= 256: ViT-Ti/B/L token counts 65 / 197): one CTA per (batch, head) keeps
// K, V, Q and dO of the whole head resident in shared memory and ALL gradient accumulators in TMEM
// (dV_j, dK_j per key block, dQ_0 and dQ_1 across key blocks), so nothing is accumulated through HBM:
// no fp32 dQ buffer, no atomics, no conversion pass, and every operand tile is loaded exactly once.
// --------------------------------------------------------------------------------------------------
struct BwdSmallSmem {
  static constexpr int K = 0;                               // 2 tiles
  static constexpr int V = 2 * AT_TILE_BYTES;               // 2 tiles
  static constexpr int Q = 4 * AT_TILE_BYTES;               // 2 tiles
  static constexpr int DO = 6 * AT_TILE_BYTES;              // 2 tiles
  static constexpr int P = 8 * AT_TILE_BYTES;               // 32 KB
  static constexpr int DS = P + 2 * AT_TILE_BYTES;          // 32 KB
  static constexpr int BAR = DS + 2 * AT_TILE_BYTES;
  static constexpr int TOTAL = BAR + 128;
};

template <bool CAUSAL>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_bwd_small_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                      const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = at_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmallSmem::BAR);
  uint64_t* kv_full = bars + 0;    // [2]
  uint64_t* q_full = bars + 2;     // [2]
  uint64_t* sdp_full = bars + 4;
  uint64_t* pds_ready = bars + 5;
  uint64_t* mma_done = bars + 6;
  uint64_t* dkv_free = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hh = blockIdx.x, b = blockIdx.y;
  const int nt = (p.N + 127) / 128;  // 1 or 2 tiles along both queries and keys

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
    mbar_init(&q_full[0], 1); mbar_init(&q_full[1], 1);
    mbar_init(sdp_full, 1);
    mbar_init(pds_ready, 128);
    mbar_init(mma_done, 1);
    mbar_init(dkv_free, 128);
    fence_barrier_init();
    for (int t = 0; t < nt; ++t) {  // order: what the first iteration needs comes first
      mbar_expect_tx(&kv_full[t], 2 * AT_TILE_BYTES);
      tma_load_3d(smem + BwdSmallSmem::K + t * AT_TILE_BYTES, &tm_qkv, &kv_full[t], p.d + hh * AT_HD, t * 128, b);
      tma_load_3d(smem + BwdSmallSmem::V + t * AT_TILE_BYTES, &tm_qkv, &kv_full[t], 2 * p.d + hh * AT_HD, t * 128, b);
      mbar_expect_tx(&q_full[t], 2 * AT_TILE_BYTES);
      tma_load_3d(smem + BwdSmallSmem::Q + t * AT_TILE_BYTES, &tm_qkv, &q_full[t], hh * AT_HD, t * 128, b);
      tma_load_3d(smem + BwdSmallSmem::DO + t * AT_TILE_BYTES, &tm_do, &q_full[t], hh * AT_HD, t * 128, b);
    }
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_dP = tmem_base + 128, tmem_dV = tmem_base + 256,
                 tmem_dK = tmem_base + 320, tmem_dQ = tmem_base + 384;  // dQ_i at tmem_dQ + 64 * i

  if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, AT_HD, true, true);     // dV, dK
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, AT_HD, false, true);    // dQ
      const uint32_t sP = smem_u32(smem + BwdSmallSmem::P), sDS = smem_u32(smem + BwdSmallSmem::DS);
      int t = 0;
      for (int j = 0; j < nt; ++j) {
        const uint32_t sK = smem_u32(smem + BwdSmallSmem::K + j * AT_TILE_BYTES);
        const uint32_t sV = smem_u32(smem + BwdSmallSmem::V + j * AT_TILE_BYTES);
        const int nkeys = min(128, p.N - j * 128);
        const int ncols = ((nkeys + 31) >> 5) << 5;
        const uint32_t idesc_s = umma_idesc_bf16(128, ncols, false, false);
        const int kkeys = (nkeys + 15) >> 4;
        mbar_wait(&kv_full[j], 0, 70);
        if (j > 0) { mbar_wait(dkv_free, (j - 1) & 1, 71); }
        const int first_i = CAUSAL ? j : 0;
        for (int i = first_i; i < nt; ++i, ++t) {
          const uint32_t sQ = smem_u32(smem + BwdSmallSmem::Q + i * AT_TILE_BYTES);
          const uint32_t sDO = smem_u32(smem + BwdSmallSmem::DO + i * AT_TILE_BYTES);
          const int krows = (min(128, p.N - i * 128) + 15) >> 4;
          mbar_wait(&q_full[i], 0, 72);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_S, desc_kmajor(sQ, k), desc_kmajor(sK, k), idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_dP, desc_kmajor(sDO, k), desc_kmajor(sV, k), idesc_s, k > 0);
          umma_commit(sdp_full);
          mbar_wait(pds_ready, t & 1, 73);
          tc_fence_after();
          for (int k = 0; k < krows; ++k)
            umma_bf16(tmem_dV, desc_mnmajor(sP, k, 16384), desc_mnmajor(sDO, k, 8192), idesc_t, (i > first_i || k > 0) ? 1u : 0u);
          for (int k = 0; k < krows; ++k)
            umma_bf16(tmem_dK, desc_mnmajor(sDS, k, 16384), desc_mnmajor(sQ, k, 8192), idesc_t, (i > first_i || k > 0) ? 1u : 0u);
          for (int k = 0; k < kkeys; ++k)
            umma_bf16(tmem_dQ + 64 * i, desc_kmajor(sDS + (k >> 2) * AT_TILE_BYTES, k & 3), desc_mnmajor(sK, k, 8192), idesc_q,
                      (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(mma_done);
        }
      }
    }
  } else if (warp < 4) {
    const int r = threadIdx.x;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    float lse2[2] = {0.f, 0.f}, Dq[2] = {0.f, 0.f};
    int t = 0;
    for (int j = 0; j < nt; ++j) {
      const int k0 = j * 128;
      const int nch = (min(128, p.N - k0) + 31) >> 5;
      const int first_i = CAUSAL ? j : 0;
      for (int i = first_i; i < nt; ++i, ++t) {
        const int q = i * 128 + r;
        const bool qv = q < p.N;
        const bool warp_has_rows = i * 128 + warp * 32 < p.N;
        if (j == 0 && qv) {  // first visit of query tile i: row statistics
          lse2[i] = p.lse[((long long)b * p.H + hh) * p.N + q] * LOG2E;
          const uint4* orow = reinterpret_cast<const uint4*>(p.o_in + ((long long)b * p.sb + q * p.sn) * p.d + hh * AT_HD);
          const uint4* drow = reinterpret_cast<const uint4*>(p.do_in + ((long long)b * p.sb + q * p.sn) * p.d + hh * AT_HD);
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 a = orow[c], g = drow[c];
            const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
            const float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
            acc += a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y + a3.x * g3.x + a3.y * g3.y;
          }
          Dq[i] = acc;
        }
        const float my_lse = lse2[i], my_D = Dq[i];
        mbar_wait(sdp_full, t & 1, 80);
        if (t > 0) mbar_wait(mma_done, (t - 1) & 1, 81);  // previous dV/dK/dQ MMAs no longer read P / dS
        tc_fence_after();
        if (warp_has_rows) {
#pragma unroll 1
          for (int c = 0; c < nch; ++c) {
            uint32_t s[32], dp[32];
            tmem_ld32(tmem_S + lane_off + c * 32, s);
            tmem_ld32(tmem_dP + lane_off + c * 32, dp);
            tmem_ld_wait();
            uint32_t wp[16], wd[16];
#pragma unroll
            for (int e2 = 0; e2 < 32; e2 += 2) {
              float pv[2], dv[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int key = k0 + c * 32 + e2 + e;
                const bool ok = qv && key < p.N && (!CAUSAL || key <= q);
                const float pe = ok ? ex2_approx(fmaf(__uint_as_float(s[e2 + e]), p.scale_log2e, -my_lse)) : 0.f;
                float pk = pe, dpe = __uint_as_float(dp[e2 + e]);
                if (p.drop_thr != 0) {
                  const bool keep = drop_keep(drop_hash_attn(p.drop_seed, (uint32_t)(b * p.H + hh), (uint32_t)q, (uint32_t)(key >> 1)),
                                              key & 1, p.drop_thr);
                  pk = keep ? pe : 0.f;
                  dpe = keep ? dpe * p.drop_r : 0.f;
                }
                pv[e] = pk;
                dv[e] = pe * (dpe - my_D) * p.scale;
              }
              wp[e2 >> 1] = pack_bf16(pv[0], pv[1]);
              wd[e2 >> 1] = pack_bf16(dv[0], dv[1]);
            }
            store_operand_chunk(smem + BwdSmallSmem::P, r, c, wp);
            store_operand_chunk(smem + BwdSmallSmem::DS, r, c, wd);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(pds_ready);
      }
      // dK_j / dV_j are complete once the last MMA batch of this key block has finished
      mbar_wait(mma_done, (t - 1) & 1, 82);
      tc_fence_after();
      const int key = k0 + r;
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {  // 0: dK, 1: dV
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32((which == 0 ? tmem_dK : tmem_dV) + lane_off + c * 32, v);
          tmem_ld_wait();
          if (key < p.N) {
            const float sc = which == 1 ? p.drop_r : 1.0f;   // dV = (mask P / (1 - p))^T dO
            __nv_bfloat16* dst = p.dqkv + ((long long)b * p.sb + key * p.sn) * 3 * p.d + (which + 1) * p.d + hh * AT_HD + c * 32;
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              uint4 w;
              w.x = pack_bf16(__uint_as_float(v[i4 * 8 + 0]) * sc, __uint_as_float(v[i4 * 8 + 1]) * sc);
              w.y = pack_bf16(__uint_as_float(v[i4 * 8 + 2]) * sc, __uint_as_float(v[i4 * 8 + 3]) * sc);
              w.z = pack_bf16(__uint_as_float(v[i4 * 8 + 4]) * sc, __uint_as_float(v[i4 * 8 + 5]) * sc);
              w.w = pack_bf16(__uint_as_float(v[i4 * 8 + 6]) * sc, __uint_as_float(v[i4 * 8 + 7]) * sc);
              reinterpret_cast<uint4*>(dst)[i4] = w;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(dkv_free);
    }
    // dQ of both query tiles (all MMAs have completed: the last mma_done was waited on above)
    for (int i = 0; i < nt; ++i) {
      const int q = i * 128 + r;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_dQ + 64 * i + lane_off + c * 32, v);
        tmem_ld_wait();
        if (q < p.N) {
          __nv_bfloat16* dst = p.dqkv + ((long long)b * p.sb + q * p.sn) * 3 * p.d + hh * AT_HD + c * 32;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(v[i4 * 8 + 0]), __uint_as_float(v[i4 * 8 + 1]));
            w.y = pack_bf16(__uint_as_float(v[i4 * 8 + 2]), __uint_as_float(v[i4 * 8 + 3]));
            w.z = pack_bf16(__uint_as_float(v[i4 * 8 + 4]), __uint_as_float(v[i4 * 8 + 5]));
            w.w = pack_bf16(__uint_as_float(v[i4 * 8 + 6]), __uint_as_float(v[i4 * 8 + 7]));
            reinterpret_cast<uint4*>(dst)[i4] = w;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// --------------------------------------------------------------------------------------------------
// Backward for short sequences, second generation (N <= 208, no causal mask): persistent, one CTA per SM, a work
// unit is one (batch, head), everything of the head resident.  The score matrix is computed TRANSPOSED,
// S^T = K Q^T, so TMEM lanes are KEYS:
//   * P^T and dS^T (bf16) are written back over the dead S^T / dP^T columns and feed dV = P^T dO and
//     dK = dS^T Q directly as TMEM A operands (tcgen05.mma, A in TMEM): P never touches shared memory;
//   * dV_j / dK_j of a 128-key tile come out of ONE accumulation over all queries (no loop-carried accumulators),
//     into the columns freed by the bf16 rewrite; only dS^T additionally goes to shared memory (as the
//     MN-major A operand of dQ = dS K, issued once per unit when both key tiles are done);
//   * two warpgroups split the QUERY columns (tile 0: 128, tile 1: the rest); there are no fp32 atomics, no dQ
//     workspace and no conversion pass;
//   * MMAs that accumulate into the same TMEM tile are dependent (~90 clk each at N = 64), so independent
//     accumulators are interleaved (S / dP, dV / dK, dQ_0 / dQ_1);
//   * results leave through swizzled shared-memory staging (operand slots that are dead by then) and TMA stores;
//   * two otherwise idle warps compute the per-query statistics (-lse log2e, D = rowsum(dO o O)) of the NEXT unit
//     while this one runs (lanes are keys, so the statistics are needed as shared-memory vectors).
// The softmax scale of dS is applied when dQ / dK are read out (64 values per row instead of N).
// TMEM: S^T [0,128)+[128,128+w1), dP^T [256,384)+[384,384+w1); P^T over the first half of each S^T part, dS^T over
// the first half of each dP^T part; dV_j [64,128), dK_j [320,384); dQ_0 [128,192), dQ_1 [192,256).
// Shared memory (packed rows, ncols = ceil16(N)): K | V | Q | dO ([ncols x 128 B] each) | dS^T slabs (64 queries
// wide, [ncols keys x 128 B] each, 2 per query tile) | staging tile X | statistic vectors | barriers.
// --------------------------------------------------------------------------------------------------
constexpr int AB2_THREADS = 384;  // warp 0: TMA, warp 1: MMA, warps 2-3: statistics (2: TMEM alloc), warps 4-7 / 8-11: query tile 0 / 1

struct Bwd2Layout {
  int OP;       // bytes of one packed operand / dS^T slab = ncols * 128
  int sK, sV, sQ, sDO, sDS, sX, vec, bar, total;
  __host__ __device__ Bwd2Layout(int ncols, int nq) {
    OP = ncols * 128;
    sK = 0; sV = OP; sQ = 2 * OP; sDO = 3 * OP; sDS = 4 * OP;
    sX = sDS + 2 * nq * OP;        // [128 x 128 B] output staging tile (1024-aligned: OP is a multiple of 2048)
    vec = sX + AT_TILE_BYTES;
    bar = vec + 2 * 256 * 4;       // (-lse log2e, D) x 256 queries
    total = bar + 256;
  }
  // The M = 128 A-operand tiles (K_j, V_j) always span 128 rows: for tiny sequences they reach past the packed
  // operands, so the allocation must at least cover them (the surplus rows only feed lanes that are never stored).
  __host__ __device__ int alloc_bytes(int nq) const {
    const int need = sV + nq * 16384 + 1024;
    return total > need ? total : need;
  }
};

template <bool DROP>
__global__ void __launch_bounds__(AB2_THREADS, 1)
attn_bwd_short_kernel(const __grid_constant__ CUtensorMap tm_qkv_a, const __grid_constant__ CUtensorMap tm_qkv_b,
                      const __grid_constant__ CUtensorMap tm_do_a, const __grid_constant__ CUtensorMap tm_do_b,
                      const __grid_constant__ CUtensorMap tm_out, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  uint8_t* smem = at_smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int ncols = ((p.N + 15) >> 4) << 4;
  const int nt = (p.N + 127) >> 7;               // query tiles == key tiles (1 or 2)
  const int w0 = ncols < 128 ? ncols : 128, w1 = ncols - w0;
  const Bwd2Layout L(ncols, nt);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar);
  // The operands are released one by one as the unit finishes with them, so the next unit's loads (there is no room
  // for a second operand stage) overlap the tail of this one: Q / dO after the last dV / dK MMAs, K after the dQ
  // MMAs, V once the dV read-out staged in its rows has been read by the TMA store.
  uint64_t* full = bars + 13;       // [4] K, V, Q, dO loaded
  uint64_t* empty = bars + 17;      // [3] K, V, Q+dO free
  uint64_t* sdp_full = bars + 2;    // [2] S^T, dP^T columns of query tile i complete
  uint64_t* pds_ready = bars + 4;   // [2] P^T, dS^T of query tile i written (4 warps)
  uint64_t* dkv_full = bars + 6;    // dV_j, dK_j complete
  uint64_t* dkv_free = bars + 7;    // dV_j, dK_j read out (4 nt warps)
  uint64_t* dq_full = bars + 8;     // dQ complete
  uint64_t* dq_free = bars + 9;     // dQ read out (4 nt warps)
  uint64_t* stats_full = bars + 10; // statistic vectors of the unit written (2 warps)
  uint64_t* stats_free = bars + 11; // statistic vectors no longer read (4 nt warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  float* nlse = reinterpret_cast<float*>(smem + L.vec);
  float* Dv = nlse + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = p.B * p.H;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_qkv_a); tma_prefetch_desc(&tm_qkv_b);
    tma_prefetch_desc(&tm_do_a); tma_prefetch_desc(&tm_do_b); tma_prefetch_desc(&tm_out);
    for (int x = 0; x < 4; ++x) mbar_init(&full[x], 1);
    mbar_init(&empty[0], 1); mbar_init(&empty[1], 1 + 4); mbar_init(&empty[2], 1);
    mbar_init(&sdp_full[0], 1); mbar_init(&sdp_full[1], 1);
    mbar_init(&pds_ready[0], 4); mbar_init(&pds_ready[1], 4);
    mbar_init(dkv_full, 1); mbar_init(dkv_free, 4 * nt);
    mbar_init(dq_full, 1); mbar_init(dq_free, 4 * nt);
    mbar_init(stats_full, 2); mbar_init(stats_free, 4 * nt);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // the set-up above overlapped the previous kernel's tail
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int n = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++n) {
        const int b = u / p.H, hh = u - b * p.H;
        const uint32_t ph = (n & 1) ^ 1;
        mbar_wait(&empty[2], ph, 90);                      // Q, dO
        mbar_expect_tx(&full[2], L.OP);
        tma_load_3d(smem + L.sQ, &tm_qkv_a, &full[2], hh * AT_HD, 0, b);
        if (w1 > 0) tma_load_3d(smem + L.sQ + 16384, &tm_qkv_b, &full[2], hh * AT_HD, 128, b);
        mbar_expect_tx(&full[3], L.OP);
        tma_load_3d(smem + L.sDO, &tm_do_a, &full[3], hh * AT_HD, 0, b);
        if (w1 > 0) tma_load_3d(smem + L.sDO + 16384, &tm_do_b, &full[3], hh * AT_HD, 128, b);
        mbar_wait(&empty[0], ph, 90);                      // K
        mbar_expect_tx(&full[0], L.OP);
        tma_load_3d(smem + L.sK, &tm_qkv_a, &full[0], p.d + hh * AT_HD, 0, b);
        if (w1 > 0) tma_load_3d(smem + L.sK + 16384, &tm_qkv_b, &full[0], p.d + hh * AT_HD, 128, b);
        mbar_wait(&empty[1], ph, 90);                      // V
        mbar_expect_tx(&full[1], L.OP);
        tma_load_3d(smem + L.sV, &tm_qkv_a, &full[1], 2 * p.d + hh * AT_HD, 0, b);
        if (w1 > 0) tma_load_3d(smem + L.sV + 16384, &tm_qkv_b, &full[1], 2 * p.d + hh * AT_HD, 128, b);
        // there is no room for a second operand stage: pull the NEXT unit's tiles into L2 now, so that its loads
        // (issued when this unit releases the operands) are L2 hits
        const int u2 = u + gridDim.x;
        if (u2 < units) {
          const int b2 = u2 / p.H, h2 = u2 - b2 * p.H;
          tma_prefetch_3d(&tm_qkv_a, p.d + h2 * AT_HD, 0, b2);
          tma_prefetch_3d(&tm_qkv_a, h2 * AT_HD, 0, b2);
          tma_prefetch_3d(&tm_qkv_a, 2 * p.d + h2 * AT_HD, 0, b2);
          tma_prefetch_3d(&tm_do_a, h2 * AT_HD, 0, b2);
          if (w1 > 0) {
            tma_prefetch_3d(&tm_qkv_b, p.d + h2 * AT_HD, 128, b2);
            tma_prefetch_3d(&tm_qkv_b, h2 * AT_HD, 128, b2);
            tma_prefetch_3d(&tm_qkv_b, 2 * p.d + h2 * AT_HD, 128, b2);
            tma_prefetch_3d(&tm_do_b, h2 * AT_HD, 128, b2);
          }
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp runs the (warp-uniform) control flow and descriptor arithmetic, one elected lane issues:
    // descriptors then live in uniform registers and an MMA costs a handful of issue slots.  With everything
    // inside `if (lane == 0)` the address arithmetic is per-thread code and the N = 64 MMAs (32 clk of tensor
    // work each) were bound by the ~80 clk of scalar instructions between them.
    const bool elected = elect_one();
    constexpr uint32_t idesc_ts = umma_idesc_bf16(128, AT_HD, false, true);   // dV, dK: A in TMEM, B MN-major
    constexpr uint32_t idesc_dq = umma_idesc_bf16(128, AT_HD, true, true);    // dQ: A = dS^T slabs (MN-major)
    const uint32_t sK = smem_u32(smem + L.sK), sV = smem_u32(smem + L.sV), sQ = smem_u32(smem + L.sQ),
                   sDO = smem_u32(smem + L.sDO), sDS = smem_u32(smem + L.sDS);
    const uint32_t idesc_s0 = umma_idesc_bf16(128, w0, false, false);
    const uint32_t idesc_s1 = umma_idesc_bf16(128, w1 > 0 ? w1 : 16, false, false);
    int n = 0, cc = 0;  // unit counter, chain (key tile) counter
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++n) {
      for (int x = 0; x < 4; ++x) mbar_wait(&full[x], n & 1, 91);
      for (int j = 0; j < nt; ++j, ++cc) {
        // TMEM reuse: the previous chain's dV/dK (and, for the unit's first chain, the previous unit's dQ)
        // must have been read out before S^T / dP^T are overwritten
        if (cc > 0) mbar_wait(dkv_free, (cc - 1) & 1, 92);
        if (j == 0 && n > 0) mbar_wait(dq_free, (n - 1) & 1, 93);
        tc_fence_after();
        const uint64_t aK = desc_kmajor(sK + j * 16384, 0), aV = desc_kmajor(sV + j * 16384, 0);
        for (int i = 0; i < nt; ++i) {
          const uint32_t idesc_s = i == 0 ? idesc_s0 : idesc_s1;
          const uint64_t bQ = desc_kmajor(sQ + i * 16384, 0), bDO = desc_kmajor(sDO + i * 16384, 0);
          const uint32_t tS = tmem_base + i * 128, tdP = tmem_base + 256 + i * 128;
          if (elected) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // S^T and dP^T are independent accumulators: alternate them
              umma_bf16(tS, aK + 2 * k, bQ + 2 * k, idesc_s, k > 0);      // +32 bytes per 16-element K slice
              umma_bf16(tdP, aV + 2 * k, bDO + 2 * k, idesc_s, k > 0);
            }
            umma_commit(&sdp_full[i]);
          }
          __syncwarp();
        }
        for (int i = 0; i < nt; ++i) {
          const int ksteps = (i == 0 ? w0 : w1) >> 4;
          mbar_wait(&pds_ready[i], cc & 1, 94);
          tc_fence_after();
          uint64_t bDO = desc_mnmajor(sDO + i * 16384, 0, 8192), bQ = desc_mnmajor(sQ + i * 16384, 0, 8192);
          uint32_t aP = tmem_base + i * 128, aDS = tmem_base + 256 + i * 128;
          uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll 2
          for (int k = 0; k < ksteps; ++k) {
            if (elected) {
              umma_bf16_ts(tmem_base + 64, aP, bDO, idesc_ts, acc);
              umma_bf16_ts(tmem_base + 320, aDS, bQ, idesc_ts, acc);
            }
            aP += 8; aDS += 8; bDO += 128; bQ += 128;   // 16 queries: 8 packed TMEM columns, 16 rows x 128 B of B
            acc = 1u;
          }
          __syncwarp();
        }
        if (elected) {
          umma_commit(dkv_full);
          if (j == nt - 1) { umma_commit(&empty[2]); umma_commit(&empty[1]); }  // Q, dO (and V, MMA side) are done with
        }
        __syncwarp();
      }
      // dQ_i = dS_i K over all keys; dS^T of both key tiles is in shared memory (pds_ready waited above)
      {
        const int ksteps = ncols >> 4;
        uint64_t a0 = desc_mnmajor(sDS, 0, L.OP), a1 = desc_mnmajor(sDS + 2 * L.OP, 0, L.OP), bK = desc_mnmajor(sK, 0, 8192);
        uint32_t acc = 0u;
#pragma unroll 2
        for (int k = 0; k < ksteps; ++k) {
          if (elected) {
            umma_bf16(tmem_base + 128, a0, bK, idesc_dq, acc);
            if (nt == 2) umma_bf16(tmem_base + 192, a1, bK, idesc_dq, acc);
          }
          a0 += 128; a1 += 128; bK += 128;
          acc = 1u;
        }
        if (elected) { umma_commit(dq_full); umma_commit(&empty[0]); }   // K is done with
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // ---- statistics of the NEXT unit, computed while the current one runs: thread t owns queries t, t + 64, ... ----
    const int t = (warp - 2) * 32 + lane;
    int n = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++n) {
      const int b = u / p.H, hh = u - b * p.H;
      float a[4], dsum[4];
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        const int q = t + 64 * s4;
        a[s4] = -INFINITY; dsum[s4] = 0.f;
        if (q < p.N) {
          a[s4] = -__ldg(p.lse + ((long long)b * p.H + hh) * p.N + q) * LOG2E;
          const uint4* orow = reinterpret_cast<const uint4*>(p.o_in + ((long long)b * p.sb + q * p.sn) * p.d + hh * AT_HD);
          const uint4* drow = reinterpret_cast<const uint4*>(p.do_in + ((long long)b * p.sb + q * p.sn) * p.d + hh * AT_HD);
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 x = __ldg(orow + c), g = __ldg(drow + c);
            const float2 x0 = unpack_bf16(x.x), x1 = unpack_bf16(x.y), x2 = unpack_bf16(x.z), x3 = unpack_bf16(x.w);
            const float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
            acc += x0.x * g0.x + x0.y * g0.y + x1.x * g1.x + x1.y * g1.y + x2.x * g2.x + x2.y * g2.y + x3.x * g3.x + x3.y * g3.y;
          }
          dsum[s4] = acc;
        }
      }
      mbar_wait(stats_free, (n & 1) ^ 1, 98);   // the previous unit no longer reads the vectors
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) { nlse[t + 64 * s4] = a[s4]; Dv[t + 64 * s4] = dsum[s4]; }
      __syncwarp();
      if (lane == 0) mbar_arrive(stats_full);
    }
  } else {
    const int i = (warp - 4) >> 2;     // query tile (column range) of this warpgroup
    const int wq = warp & 3;           // TMEM lane quarter
    if (i < nt) {
      const int r = wq * 32 + lane;    // key row within the key tile == TMEM lane; query row for the dQ read-out
      const int wi = i == 0 ? w0 : w1;
      const int nch = wi >> 5;
      const bool tail16 = (wi & 16) != 0;
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tS = tmem_base + i * 128 + lane_off, tdP = tmem_base + 256 + i * 128 + lane_off;
      int n = 0, cc = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++n) {
        const int b = u / p.H, hh = u - b * p.H;
        if (lane == 0) bulk_wait_read();   // last unit's dQ store has left this warp's dS^T slab rows
        __syncwarp();
        mbar_wait(stats_full, n & 1, 99);
        for (int j = 0; j < nt; ++j, ++cc) {
          const int key_local = j * 128 + r;            // row in the packed operands / dS^T slabs
          const bool warp_has_keys = j * 128 + wq * 32 < ncols;
          const uint32_t bh = (uint32_t)u, kpair = (uint32_t)(key_local >> 1), kodd = (uint32_t)(key_local & 1);
          (void)bh; (void)kpair; (void)kodd;
          mbar_wait(&sdp_full[i], cc & 1, 95);
          tc_fence_after();
          if (warp_has_keys) {
            const bool store_ok = key_local < ncols;    // rows beyond the padded keys do not exist in the slabs
            auto do_chunk = [&](const uint32_t (&sv)[32], const uint32_t (&dv)[32], int c, int ncol) {
              // ncol = 32 or 16 valid S columns in this chunk
              uint32_t wp[16], wd[16];
              const float4* lv = reinterpret_cast<const float4*>(nlse + i * 128 + c * 32);
              const float4* dvv = reinterpret_cast<const float4*>(Dv + i * 128 + c * 32);
#pragma unroll
              for (int e4 = 0; e4 < 8; ++e4) {
                if (e4 * 4 < ncol) {
                  const float4 a = lv[e4], dd = dvv[e4];
                  const float p0 = ex2_approx(fmaf(__uint_as_float(sv[e4 * 4 + 0]), p.scale_log2e, a.x));
                  const float p1 = ex2_approx(fmaf(__uint_as_float(sv[e4 * 4 + 1]), p.scale_log2e, a.y));
                  const float p2 = ex2_approx(fmaf(__uint_as_float(sv[e4 * 4 + 2]), p.scale_log2e, a.z));
                  const float p3 = ex2_approx(fmaf(__uint_as_float(sv[e4 * 4 + 3]), p.scale_log2e, a.w));
                  float g0 = __uint_as_float(dv[e4 * 4 + 0]), g1 = __uint_as_float(dv[e4 * 4 + 1]);
                  float g2 = __uint_as_float(dv[e4 * 4 + 2]), g3 = __uint_as_float(dv[e4 * 4 + 3]);
                  float k0 = p0, k1 = p1, k2 = p2, k3 = p3;   // probabilities as used by dV (after dropout)
                  if constexpr (DROP) {  // this thread's key is fixed, the query index runs: one hash per element
                    const uint32_t q0 = (uint32_t)(i * 128 + c * 32 + e4 * 4);
                    const bool m0 = drop_keep(drop_hash_attn(p.drop_seed, bh, q0 + 0, kpair), kodd, p.drop_thr);
                    const bool m1 = drop_keep(drop_hash_attn(p.drop_seed, bh, q0 + 1, kpair), kodd, p.drop_thr);
                    const bool m2 = drop_keep(drop_hash_attn(p.drop_seed, bh, q0 + 2, kpair), kodd, p.drop_thr);
                    const bool m3 = drop_keep(drop_hash_attn(p.drop_seed, bh, q0 + 3, kpair), kodd, p.drop_thr);
                    k0 = m0 ? p0 : 0.f; k1 = m1 ? p1 : 0.f; k2 = m2 ? p2 : 0.f; k3 = m3 ? p3 : 0.f;
                    g0 = m0 ? g0 * p.drop_r : 0.f; g1 = m1 ? g1 * p.drop_r : 0.f;
                    g2 = m2 ? g2 * p.drop_r : 0.f; g3 = m3 ? g3 * p.drop_r : 0.f;
                  }
                  wp[e4 * 2 + 0] = pack_bf16(k0, k1);
                  wp[e4 * 2 + 1] = pack_bf16(k2, k3);
                  wd[e4 * 2 + 0] = pack_bf16(p0 * (g0 - dd.x), p1 * (g1 - dd.y));
                  wd[e4 * 2 + 1] = pack_bf16(p2 * (g2 - dd.z), p3 * (g3 - dd.w));
                } else {
                  wp[e4 * 2 + 0] = 0u; wp[e4 * 2 + 1] = 0u; wd[e4 * 2 + 0] = 0u; wd[e4 * 2 + 1] = 0u;
                }
              }
              tmem_st16(tS + c * 16, wp);    // over S^T columns already consumed (16 c <= 32 c)
              tmem_st16(tdP + c * 16, wd);
              if (store_ok) {
                uint8_t* row = smem + L.sDS + (2 * i + (c >> 1)) * L.OP + key_local * 128;
                const int chunk0 = (c & 1) * 4;
                const int nq4 = ncol >> 3;   // 16-byte pieces holding valid columns
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  if (q4 < nq4) {
                    const int chunk = (chunk0 + q4) ^ (key_local & 7);
                    *reinterpret_cast<uint4*>(row + chunk * 16) = make_uint4(wd[q4 * 4 + 0], wd[q4 * 4 + 1], wd[q4 * 4 + 2], wd[q4 * 4 + 3]);
                  }
                }
              }
            };
            uint32_t sv[32], dv[32];
            for (int c = 0; c < nch; ++c) {
              tmem_ld32(tS + c * 32, sv);
              tmem_ld32(tdP + c * 32, dv);
              tmem_ld_wait();
              do_chunk(sv, dv, c, 32);
            }
            if (tail16) {
              uint32_t s16[16], d16[16];
              tmem_ld16(tS + nch * 32, s16);
              tmem_ld16(tdP + nch * 32, d16);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 16; ++e) { sv[e] = s16[e]; dv[e] = d16[e]; }
#pragma unroll
              for (int e = 16; e < 32; ++e) { sv[e] = 0u; dv[e] = 0u; }
              do_chunk(sv, dv, nch, 16);
            }
            tmem_st_wait();
          }
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&pds_ready[i]);
            if (j == nt - 1) mbar_arrive(stats_free);   // last read of this unit's statistic vectors
          }
          // ---- dV_j (warpgroup 0) / dK_j (warpgroup 1, or 0 when there is a single query tile) ----
          mbar_wait(dkv_full, cc & 1, 96);
          tc_fence_after();
          if (warp_has_keys) {
            const bool wr = key_local < ncols;
            const int row0 = j * 128 + wq * 32;
            if (nt == 1 || i == 0)   // dV_j -> staged in the (dead) V_j rows
              bwd2_store_tile(tmem_base + 64 + lane_off, DROP ? p.drop_r : 1.0f, smem + L.sV + row0 * 128, wr, &tm_out,
                              2 * p.d + hh * AT_HD, row0, b, lane);
            if (nt == 1 || i == 1)   // dK_j -> staging tile X
              bwd2_store_tile(tmem_base + 320 + lane_off, p.scale, smem + L.sX + wq * 4096, wr, &tm_out,
                              p.d + hh * AT_HD, row0, b, lane);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dkv_free);
        }
        // ---- dQ of query tile i: lanes are query rows now; staged in this warp's rows of dS^T slab 2i (dead: the
        // dQ MMAs have completed), so no operand slot is held back from the next unit's loads ----
        mbar_wait(dq_full, n & 1, 97);
        tc_fence_after();
        if (i == 0) {   // the dV_j store staged in the V rows has been read -> V may be refilled
          if (lane == 0) { bulk_wait_read(); mbar_arrive(&empty[1]); }
          __syncwarp();
        }
        if (i * 128 + wq * 32 < ncols)
          bwd2_store_tile(tmem_base + 128 + 64 * i + lane_off, p.scale, smem + L.sDS + 2 * i * L.OP + wq * 4096,
                          i * 128 + r < ncols, &tm_out, hh * AT_HD, i * 128 + wq * 32, b, lane);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dq_free);
      }
      if (lane == 0) bulk_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// D[b, h, q] = sum_c dO[b, q, h, c] * O[b, q, h, c]: one warp per token row, fully coalesced; every key-block CTA of
// the streaming backward then needs two scalars per query row instead of re-reading 256 bytes of O and dO.
__global__ void __launch_bounds__(256) attn_bwd_dstat_kernel(const __nv_bfloat16* __restrict__ o,
                                                              const __nv_bfloat16* __restrict__ d_o,
                                                              float* __restrict__ dstat, int B, int N, int H,
                                                              long long sb, long long sn) {
  const int d = H * AT_HD;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);   // token index b * N + q
  if (row >= (long long)B * N) return;
  const int b = (int)(row / N), q = (int)(row - (long long)b * N);
  const long long off = ((long long)b * sb + (long long)q * sn) * d;
  for (int c0 = 0; c0 < d / 8; c0 += 32) {     // 8 elements per lane per step; 8 consecutive lanes = one head
    const int c = c0 + lane;
    const bool valid = c < d / 8;              // warp-uniform loop: every lane takes part in the shuffles
    float acc = 0.f;
    if (valid) {
      const uint4 x = *reinterpret_cast<const uint4*>(o + off + c * 8);
      const uint4 g = *reinterpret_cast<const uint4*>(d_o + off + c * 8);
      const float2 x0 = unpack_bf16(x.x), x1 = unpack_bf16(x.y), x2 = unpack_bf16(x.z), x3 = unpack_bf16(x.w);
      const float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
      acc = x0.x * g0.x + x0.y * g0.y + x1.x * g1.x + x1.y * g1.y + x2.x * g2.x + x2.y * g2.y + x3.x * g3.x + x3.y * g3.y;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (valid && (lane & 7) == 0) dstat[((long long)b * H + (c >> 3)) * N + q] = acc;
  }
}

// dq_acc fp32 [rows, d] -> bf16 into dqkv[:, 0:d] (row pitch 3d)
__global__ void attn_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dqkv,
                                       long long rows, int d) {
  const long long nvec = rows * (d / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / (d / 4);
    const int c = (int)(i - row * (d / 4)) * 4;
    const float4 v = reinterpret_cast<const float4*>(acc)[i];
    uint2 w;
    w.x = pack_bf16(v.x, v.y); w.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(dqkv + row * 3 * d + c) = w;
  }
}

static int make_tmap_bnd(CUtensorMap* tm, const void* base, int B, int N, int row_elems, int seq_first,
                         int box_rows = 128) {
  const uint64_t dims[3] = {(uint64_t)row_elems, (uint64_t)N, (uint64_t)B};
  const uint64_t sn = seq_first ? (uint64_t)B : 1, sb = seq_first ? 1 : (uint64_t)N;
  const uint64_t strides[3] = {2, sn * row_elems * 2, sb * row_elems * 2};
  const uint32_t box[3] = {64, (uint32_t)box_rows, 1};
  return make_tmap_nd_bf16(tm, base, 3, dims, strides, box, true);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200vit_flash_attn_fwd(const void* qkv, void* o, float* lse, int B, int N, int H, int causal,
                           int seq_first, void* stream) {
  return b200vit_flash_attn_fwd_dropout(qkv, o, lse, B, N, H, causal, seq_first, 0.f, 0u, stream);
}

int b200vit_flash_attn_fwd_dropout(const void* qkv, void* o, float* lse, int B, int N, int H, int causal,
                                   int seq_first, float dropout_p, unsigned int seed, void* stream) {
  B200_REQUIRE(qkv && o, "flash_attn_fwd: null pointer");
  B200_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "flash_attn_fwd: dropout_p must be in [0, 1)");
  B200_REQUIRE(B > 0 && N > 0 && H > 0, "flash_attn_fwd: bad sizes");
  const int d = H * AT_HD;
  CUtensorMap tm;
  int rc = make_tmap_bnd(&tm, qkv, B, N, 3 * d, seq_first);
  if (rc != OK) return rc;
  AttnParams p{};
  p.B = B; p.N = N; p.H = H; p.d = d; p.causal = causal;
  p.sb = seq_first ? 1 : N; p.sn = seq_first ? B : 1;
  p.scale = 0.125f; p.scale_log2e = 0.125f * LOG2E;
  p.o = (__nv_bfloat16*)o; p.lse = lse;
  p.drop_seed = seed; p.drop_thr = dropout_p > 0.f ? drop_threshold(dropout_p) : 0u; p.drop_r = 1.0f / (1.0f - dropout_p);
  dim3 grid((N + AT_BQ - 1) / AT_BQ, H, B);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tm_o;   // O leaves through 32-row TMA stores in every forward kernel
  rc = make_tmap_bnd(&tm_o, o, B, N, d, seq_first, 32);
  if (rc != OK) return rc;
  if (!causal && N <= 256 && g_debug[7] == 0) {
    // short sequences: persistent kernel, whole head resident, exact single-shot softmax, P in TMEM
    B200_CUDA(cudaFuncSetAttribute(attn_fwd_short_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdShortSmem::TOTAL));
    B200_CUDA(cudaFuncSetAttribute(attn_fwd_short_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdShortSmem::TOTAL));
    const int units = B * H;
    const int g = units < num_sms() ? units : num_sms();
    if (p.drop_thr != 0) launch_kernel(attn_fwd_short_kernel<true>, dim3(g), dim3(AS_THREADS), FwdShortSmem::TOTAL, st, 1, tm, tm_o, p);
    else                 launch_kernel(attn_fwd_short_kernel<false>, dim3(g), dim3(AS_THREADS), FwdShortSmem::TOTAL, st, 1, tm, tm_o, p);
    B200_CUDA(cudaGetLastError());
    return OK;
  }
  if (causal) {
    B200_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    attn_fwd_kernel<true><<<grid, AT_THREADS, FwdSmem::TOTAL, st>>>(tm, tm_o, p);
  } else {
    B200_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    attn_fwd_kernel<false><<<grid, AT_THREADS, FwdSmem::TOTAL, st>>>(tm, tm_o, p);
  }
  B200_CUDA(cudaGetLastError());
  return OK;
}

size_t b200vit_flash_attn_bwd_workspace_size(int B, int N, int H) {
  // streaming kernel: fp32 dQ accumulator [B, N, d] + D statistics [B, H, N]; resident kernels (N <= 256): nothing
  return N > 256 ? ((size_t)B * N * H * AT_HD + (size_t)B * H * N) * sizeof(float) : 0;
}

int b200vit_flash_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv,
                           int B, int N, int H, int causal, int seq_first, void* workspace,
                           size_t workspace_bytes, void* stream) {
  return b200vit_flash_attn_bwd_dropout(qkv, o, d_o, lse, dqkv, B, N, H, causal, seq_first, 0.f, 0u, workspace,
                                        workspace_bytes, stream);
}

int b200vit_flash_attn_bwd_dropout(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv,
                                   int B, int N, int H, int causal, int seq_first, float dropout_p,
                                   unsigned int seed, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(qkv && o && d_o && lse && dqkv, "flash_attn_bwd: null pointer");
  B200_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "flash_attn_bwd: dropout_p must be in [0, 1)");
  // only the streaming kernel (N > 256) accumulates dQ through an fp32 workspace; the resident kernels need none
  const bool needs_ws = N > 256;
  B200_REQUIRE(!needs_ws || (workspace && workspace_bytes >= b200vit_flash_attn_bwd_workspace_size(B, N, H)),
               "flash_attn_bwd: workspace missing or too small");
  const int d = H * AT_HD;
  CUtensorMap tm_qkv, tm_do;
  int rc = make_tmap_bnd(&tm_qkv, qkv, B, N, 3 * d, seq_first);
  if (rc != OK) return rc;
  rc = make_tmap_bnd(&tm_do, d_o, B, N, d, seq_first);
  if (rc != OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  AttnParams p{};
  p.B = B; p.N = N; p.H = H; p.d = d; p.causal = causal;
  p.sb = seq_first ? 1 : N; p.sn = seq_first ? B : 1;
  p.scale = 0.125f; p.scale_log2e = 0.125f * LOG2E;
  p.lse = const_cast<float*>(lse);
  p.o_in = (const __nv_bfloat16*)o; p.do_in = (const __nv_bfloat16*)d_o;
  p.dq_acc = (float*)workspace; p.dqkv = (__nv_bfloat16*)dqkv;
  p.drop_seed = seed; p.drop_thr = dropout_p > 0.f ? drop_threshold(dropout_p) : 0u; p.drop_r = 1.0f / (1.0f - dropout_p);
  p.dbg_skip_dq = g_debug[8];
  if (!causal && N <= 208 && g_debug[7] == 0) {
    // short sequences, transposed formulation: persistent, whole head resident, P^T kept in TMEM
    const int ncols = ((N + 15) / 16) * 16, nt = (N + 127) / 128;
    const int w0 = ncols < 128 ? ncols : 128, w1 = ncols - w0;
    CUtensorMap qa, qb, da, db;
    rc = make_tmap_bnd(&qa, qkv, B, N, 3 * d, seq_first, w0);
    if (rc != OK) return rc;
    rc = make_tmap_bnd(&da, d_o, B, N, d, seq_first, w0);
    if (rc != OK) return rc;
    qb = qa; db = da;
    if (w1 > 0) {
      rc = make_tmap_bnd(&qb, qkv, B, N, 3 * d, seq_first, w1);
      if (rc != OK) return rc;
      rc = make_tmap_bnd(&db, d_o, B, N, d, seq_first, w1);
      if (rc != OK) return rc;
    }
    CUtensorMap tout;
    rc = make_tmap_bnd(&tout, dqkv, B, N, 3 * d, seq_first, 32);
    if (rc != OK) return rc;
    const Bwd2Layout L(ncols, nt);
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_short_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.alloc_bytes(nt)));
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_short_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.alloc_bytes(nt)));
    const int units = B * H;
    const int g = units < num_sms() ? units : num_sms();
    if (p.drop_thr != 0) launch_kernel(attn_bwd_short_kernel<true>, dim3(g), dim3(AB2_THREADS), (size_t)L.alloc_bytes(nt), st, 1, qa, qb, da, db, tout, p);
    else                 launch_kernel(attn_bwd_short_kernel<false>, dim3(g), dim3(AB2_THREADS), (size_t)L.alloc_bytes(nt), st, 1, qa, qb, da, db, tout, p);
    B200_CUDA(cudaGetLastError());
    return OK;
  }
  if (N <= 256) {
    // short sequences: everything resident per (batch, head), no HBM accumulation
    dim3 grid_small(H, B);
    if (causal) {
      B200_CUDA(cudaFuncSetAttribute(attn_bwd_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmallSmem::TOTAL));
      attn_bwd_small_kernel<true><<<grid_small, AT_THREADS, BwdSmallSmem::TOTAL, st>>>(tm_qkv, tm_do, p);
    } else {
      B200_CUDA(cudaFuncSetAttribute(attn_bwd_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmallSmem::TOTAL));
      attn_bwd_small_kernel<false><<<grid_small, AT_THREADS, BwdSmallSmem::TOTAL, st>>>(tm_qkv, tm_do, p);
    }
    B200_CUDA(cudaGetLastError());
    return OK;
  }
  B200_CUDA(cudaMemsetAsync(workspace, 0, (size_t)B * N * d * sizeof(float), st));
  {
    float* dstat = (float*)workspace + (size_t)B * N * d;
    p.dstat = dstat;
    const long long rows = (long long)B * N;
    attn_bwd_dstat_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o, dstat,
                                                                   B, N, H, p.sb, p.sn);
    B200_CUDA(cudaGetLastError());
  }
  dim3 grid((N + AT_BK - 1) / AT_BK, H, B);
  if (causal) {
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL));
    attn_bwd_kernel<true><<<grid, AT_BWD_THREADS, BwdSmem::TOTAL, st>>>(tm_qkv, tm_do, p);
  } else {
    B200_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL));
    attn_bwd_kernel<false><<<grid, AT_BWD_THREADS, BwdSmem::TOTAL, st>>>(tm_qkv, tm_do, p);
  }
  B200_CUDA(cudaGetLastError());
  const long long rows = (long long)B * N;
  attn_dq_convert_kernel<<<num_sms() * 4, 256, 0, st>>>((const float*)workspace, (__nv_bfloat16*)dqkv, rows, d);
  B200_CUDA(cudaGetLastError());
  return OK;
}

}  // extern "C"
