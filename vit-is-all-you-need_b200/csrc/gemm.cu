// Host launchers + C ABI for the tcgen05 GEMM family (see gemm_tcgen05.cuh for the kernel).
#include "../../include/b200vit.h"
#include "gemm_tcgen05.cuh"

namespace b200 {

// Pick the K-split that minimises (waves x k-blocks per work item) on a persistent grid of `workers` CTAs / CTA pairs.
static void plan_splits(GemmShape& s, bool allow_split, int workers) {
  const int tiles = s.tiles_m * s.tiles_n;
  int best_splits = 1;
  if (allow_split && tiles < workers) {
    double best_cost = 1e30;
    const int max_splits = s.kb_total < 64 ? s.kb_total : 64;
    for (int sp = 1; sp <= max_splits; ++sp) {
      const int kb_per = (s.kb_total + sp - 1) / sp;
      const int eff = (s.kb_total + kb_per - 1) / kb_per;  // non-empty splits
      if (eff != sp) continue;
      const int waves = (tiles * sp + workers - 1) / workers;
      const double cost = (double)waves * (kb_per + 8.0);  // +8: prologue/epilogue per work item
      if (cost < best_cost) { best_cost = cost; best_splits = sp; }
    }
  }
  s.splits = best_splits;
  s.kb_per = (s.kb_total + best_splits - 1) / best_splits;
  s.total_work = tiles * s.splits;
}

template <bool A_MN, bool B_MN, int BN, int KIND, int NCTA>
static int launch(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                  const EpiParams& ep, void* out2, bool allow_split, cudaStream_t st) {
  using Cfg = GemmCfg<BN, KIND, NCTA>;
  GemmShape s;
  s.M = M; s.N = N; s.K = K;
  s.tiles_m = (M + GEMM_BM * NCTA - 1) / (GEMM_BM * NCTA);
  s.tiles_n = (N + BN - 1) / BN;
  s.kb_total = (K + GEMM_BK - 1) / GEMM_BK;
  s.mn_lbo = g_debug[0] ? g_debug[0] : 8192;
  s.mn_sbo = g_debug[1] ? g_debug[1] : 1024;
  s.mn_kadv = g_debug[2] ? g_debug[2] : 2048;
  int workers = num_sms() / NCTA;
  if (g_debug[3] > 0 && workers > g_debug[3]) workers = g_debug[3];
  plan_splits(s, allow_split, workers);

  CUtensorMap ta, tb;
  int rc;
  if (A_MN) rc = make_tmap_2d_bf16(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64, 64);
  else      rc = make_tmap_2d_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM, 64);
  if (rc != OK) return rc;
  if (B_MN) rc = make_tmap_2d_bf16(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64, 64);
  else      rc = make_tmap_2d_bf16(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, Cfg::B_ROWS, 64);
  if (rc != OK) return rc;

  // output tensor maps for the TMA-store epilogue (32-row slabs of 128 bytes)
  CUtensorMap to = ta, to2 = ta;
  if (!epi_direct_stores(KIND)) {
    if (epi_out_is_f32(KIND)) rc = make_tmap_2d_f32(&to, ep.out, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ldo, 32, 32);
    else                      rc = make_tmap_2d_bf16(&to, ep.out, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ldo, 32, 64);
    if (rc != OK) return rc;
    if (KIND == EPI_GELU_BF16) {
      rc = make_tmap_2d_bf16(&to2, out2, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ldo, 32, 64);
      if (rc != OK) return rc;
    }
    if (KIND == EPI_GELU_Q8) {
      rc = make_tmap_2d_u8(&to2, out2, (uint64_t)M, (uint64_t)N, (uint64_t)ep.ldo, 32, 128);
      if (rc != OK) return rc;
    }
  }

  auto kern = gemm_tcgen05_kernel<A_MN, B_MN, BN, KIND, NCTA>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done = true;
  }
  const int nwork = s.total_work < workers ? s.total_work : workers;
  B200_CUDA(launch_kernel(kern, dim3(nwork * NCTA), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, NCTA, ta, tb, to, to2, s, ep));
  return OK;
}

template <bool A_MN, bool B_MN, int KIND>
static int launch_bn(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K,
                     const EpiParams& ep, bool allow_split, cudaStream_t st, void* out2 = nullptr) {
  // 256-wide N tiles when N fills them; 128 otherwise (less padding waste for narrow outputs).
  const bool wide = (g_debug[4] == 256) || (g_debug[4] != 128 && (N % 256 == 0 || N >= 1024));
  // CTA pairs (256-row tiles, cta_group::2) whenever there is more than one 128-row tile of output.
  const bool pair = (g_debug[5] == 2) || (g_debug[5] != 1 && M > GEMM_BM);
  if (pair) {
    if (wide) return launch<A_MN, B_MN, 256, KIND, 2>(A, lda, B, ldb, M, N, K, ep, out2, allow_split, st);
    return launch<A_MN, B_MN, 128, KIND, 2>(A, lda, B, ldb, M, N, K, ep, out2, allow_split, st);
  }
  if (wide) return launch<A_MN, B_MN, 256, KIND, 1>(A, lda, B, ldb, M, N, K, ep, out2, allow_split, st);
  return launch<A_MN, B_MN, 128, KIND, 1>(A, lda, B, ldb, M, N, K, ep, out2, allow_split, st);
}

static int check_common(const void* a, const void* b, const void* c, int M, int N, int K) {
  B200_REQUIRE(a && b && c, "gemm: null pointer");
  B200_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: non-positive dimension M=%d N=%d K=%d", M, N, K);
  B200_REQUIRE(N % 8 == 0 && K % 8 == 0, "gemm: N=%d and K=%d must be multiples of 8", N, K);
  return OK;
}

static EpiParams make_ep(void* out, long long ldo) {
  EpiParams ep;
  ep.out = out; ep.ldo = ldo; ep.bias = nullptr;
  ep.aux = nullptr; ep.ldaux = 0; ep.pos = nullptr; ep.P = 1; ep.T = 1; ep.extra = 0; ep.colsum = nullptr; ep.drop_seed = 0; ep.drop_thr = 0; ep.drop_r = 1.0f;
  return ep;
}

// used by patchify.cu
int gemm_patch_epilogue(const void* xcol, const void* w, const float* bias, const float* pos,
                        float* out, int rows, int N, int K, int P, int T, int extra,
                        cudaStream_t st) {
  EpiParams ep = make_ep(out, N);
  ep.bias = bias; ep.pos = pos; ep.ldaux = N; ep.P = P; ep.T = T; ep.extra = extra;
  return launch_bn<false, false, EPI_PATCH_F32>(xcol, K, w, K, rows, N, K, ep, false, st);
}

// used by patchify.cu: y = rows @ w^T + bias written as the NCHW image (de-patchify, pixel shuffle in the epilogue)
int gemm_depatch_epilogue(const void* rows_bf16, const void* w, const float* bias, float* img, int rows, int N, int K,
                          int P, int Wt, int log2p, cudaStream_t st) {
  EpiParams ep = make_ep(img, N);
  ep.bias = bias; ep.P = P; ep.T = Wt; ep.extra = log2p;
  return launch_bn<false, false, EPI_DEPATCH_F32>(rows_bf16, K, w, K, rows, N, K, ep, false, st);
}

// gemm_skinny.cu: weight-streaming kernel for M <= 64 (the single-token decode step); 1 = launched, 0 = not applicable
int gemm_skinny_try(int kind, const void* x, const void* w, const float* bias, void* out, void* out2, const float* resid,
                    int M, int N, int K, cudaStream_t st);

}  // namespace b200

using namespace b200;

#define B200_TRY_SKINNY(kind, out, out2, resid)                                                            \
  do {                                                                                                     \
    const int _sk = gemm_skinny_try(kind, x, w, bias, out, out2, resid, M, N, K, (cudaStream_t)stream);    \
    if (_sk < 0) return _sk;                                                                               \
    if (_sk == 1) return OK;                                                                               \
  } while (0)

extern "C" {

// bring-up aid: how many CTA pairs of the forward GEMM kernel can be co-resident on this device
int b200vit_debug_max_clusters(void) {
  using Cfg = GemmCfg<256, EPI_BF16, 2>;
  auto kern = gemm_tcgen05_kernel<false, false, 256, EPI_BF16, 2>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) return -1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(num_sms());
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = -1;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) return -2;
  return n;
}

int b200vit_gemm_bias(const void* x, const void* w, const float* bias, void* y, int M, int N, int K,
                      void* stream) {
  int rc = check_common(x, w, y, M, N, K);
  if (rc) return rc;
  B200_TRY_SKINNY(EPI_BF16, y, nullptr, nullptr);
  EpiParams ep = make_ep(y, N);
  ep.bias = bias;
  return launch_bn<false, false, EPI_BF16>(x, K, w, K, M, N, K, ep, false, (cudaStream_t)stream);
}

int b200vit_gemm_bias_gelu(const void* x, const void* w, const float* bias, void* g, void* gprime, int M,
                           int N, int K, void* stream) {
  int rc = check_common(x, w, g, M, N, K);
  if (rc) return rc;
  B200_REQUIRE(gprime != nullptr, "gemm_bias_gelu: gprime is null");
  B200_TRY_SKINNY(EPI_GELU_BF16, g, gprime, nullptr);
  EpiParams ep = make_ep(g, N);
  ep.bias = bias;
  return launch_bn<false, false, EPI_GELU_BF16>(x, K, w, K, M, N, K, ep, false, (cudaStream_t)stream, gprime);
}

int b200vit_gemm_bias_residual(const void* x, const void* w, const float* bias, const float* resid,
                               float* out, int M, int N, int K, void* stream) {
  int rc = check_common(x, w, out, M, N, K);
  if (rc) return rc;
  B200_REQUIRE(resid != nullptr, "gemm_bias_residual: resid is null");
  B200_TRY_SKINNY(EPI_RESID_F32, out, nullptr, resid);
  EpiParams ep = make_ep(out, N);
  ep.bias = bias; ep.aux = resid; ep.ldaux = N;
  return launch_bn<false, false, EPI_RESID_F32>(x, K, w, K, M, N, K, ep, false, (cudaStream_t)stream);
}

int b200vit_gemm_bias_dropout_residual(const void* x, const void* w, const float* bias, const float* resid,
                                       float* out, int M, int N, int K, float p, unsigned int seed, void* stream) {
  if (p <= 0.f) return b200vit_gemm_bias_residual(x, w, bias, resid, out, M, N, K, stream);
  int rc = check_common(x, w, out, M, N, K);
  if (rc) return rc;
  B200_REQUIRE(resid != nullptr, "gemm_bias_dropout_residual: resid is null");
  B200_REQUIRE(p < 1.f, "gemm_bias_dropout_residual: p must be in [0, 1)");
  EpiParams ep = make_ep(out, N);
  ep.bias = bias; ep.aux = resid; ep.ldaux = N;
  ep.drop_seed = seed; ep.drop_thr = drop_threshold(p); ep.drop_r = 1.0f / (1.0f - p);
  return launch_bn<false, false, EPI_DROP_RESID_F32>(x, K, w, K, M, N, K, ep, false, (cudaStream_t)stream);
}

int b200vit_gemm_bias_f32(const void* x, const void* w, const float* bias, float* out, int M, int N,
                          int K, void* stream) {
  int rc = check_common(x, w, out, M, N, K);
  if (rc) return rc;
  B200_TRY_SKINNY(EPI_F32, out, nullptr, nullptr);
  EpiParams ep = make_ep(out, N);
  ep.bias = bias;
  return launch_bn<false, false, EPI_F32>(x, K, w, K, M, N, K, ep, false, (cudaStream_t)stream);
}

// dx[M,K] = dy[M,N] w[N,K]: contraction over N; output columns = K.
int b200vit_gemm_dgrad(const void* dy, const void* w, void* dx, int M, int N, int K, void* stream) {
  int rc = check_common(dy, w, dx, M, N, K);
  if (rc) return rc;
  EpiParams ep = make_ep(dx, K);
  return launch_bn<false, true, EPI_BF16>(dy, N, w, K, M, /*N_out=*/K, /*K_contract=*/N, ep, false,
                                          (cudaStream_t)stream);
}

int b200vit_gemm_dgrad_dgelu(const void* dy, const void* w, const void* gprime, void* dx, int M, int N,
                             int K, void* stream) {
  int rc = check_common(dy, w, dx, M, N, K);
  if (rc) return rc;
  B200_REQUIRE(gprime != nullptr, "gemm_dgrad_dgelu: gprime is null");
  EpiParams ep = make_ep(dx, K);
  ep.aux = gprime; ep.ldaux = K;
  return launch_bn<false, true, EPI_MUL_BF16>(dy, N, w, K, M, K, N, ep, false, (cudaStream_t)stream);
}

// The same two GEMMs with GELU'(u) carried as the 8-bit code of gemm_tcgen05.cuh (GP_LO / GP_STEP): half the bytes for the
// tensor that bounds both kernels.  256-wide tiles only (the u8 staging slab holds a warp's 128 columns).
int b200vit_gemm_bias_gelu_q8(const void* x, const void* w, const float* bias, void* g, void* gprime_q8, int M, int N, int K,
                              void* stream) {
  int rc = check_common(x, w, g, M, N, K);
  if (rc) return rc;
  B200_REQUIRE(gprime_q8 != nullptr && N % 16 == 0, "gemm_bias_gelu_q8: gprime is null or N=%d is not a multiple of 16", N);
  EpiParams ep = make_ep(g, N);
  ep.bias = bias;
  if (M > GEMM_BM) return launch<false, false, 256, EPI_GELU_Q8, 2>(x, K, w, K, M, N, K, ep, gprime_q8, false, (cudaStream_t)stream);
  return launch<false, false, 256, EPI_GELU_Q8, 1>(x, K, w, K, M, N, K, ep, gprime_q8, false, (cudaStream_t)stream);
}

int b200vit_gemm_dgrad_dgelu_q8(const void* dy, const void* w, const void* gprime_q8, void* dx, int M, int N, int K,
                                void* stream) {
  int rc = check_common(dy, w, dx, M, N, K);
  if (rc) return rc;
  B200_REQUIRE(gprime_q8 != nullptr && K % 16 == 0, "gemm_dgrad_dgelu_q8: gprime is null or K=%d is not a multiple of 16", K);
  EpiParams ep = make_ep(dx, K);
  ep.aux = gprime_q8; ep.ldaux = K;
  if (M > GEMM_BM) return launch<false, true, 256, EPI_MUL_Q8, 2>(dy, N, w, K, M, K, N, ep, nullptr, false, (cudaStream_t)stream);
  return launch<false, true, 256, EPI_MUL_Q8, 1>(dy, N, w, K, M, K, N, ep, nullptr, false, (cudaStream_t)stream);
}

float b200vit_gelu_grad_code_lo(void) { return GP_LO; }
float b200vit_gelu_grad_code_step(void) { return GP_STEP; }

// dw[N,K] = dy[M,N]^T x[M,K]: output rows = N, output cols = K, contraction over M (split-K).
// db[N] (optional) = column sums of dy, accumulated by two otherwise idle warps from the smem dy tiles.
int b200vit_gemm_wgrad_bias(const void* dy, const void* x, float* dw, float* db, int M, int N, int K,
                            int accumulate, void* stream) {
  int rc = check_common(dy, x, dw, M, N, K);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  EpiParams ep = make_ep(dw, K);
  ep.colsum = db;
  if (!accumulate) {
    B200_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * K, st));
    if (db) B200_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, st));
  }
  return launch_bn<true, true, EPI_ATOMIC_F32>(dy, N, x, K, /*M_out=*/N, /*N_out=*/K,
                                               /*K_contract=*/M, ep, true, st);
}

int b200vit_gemm_wgrad(const void* dy, const void* x, float* dw, int M, int N, int K, int accumulate,
                       void* stream) {
  return b200vit_gemm_wgrad_bias(dy, x, dw, nullptr, M, N, K, accumulate, stream);
}

}  // extern "C"
