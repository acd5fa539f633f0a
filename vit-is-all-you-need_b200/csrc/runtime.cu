// Host runtime shared by all entry points: error strings, device checks, TMA tensor-map encoding.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "../../include/b200vit.h"
#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
int g_debug[16] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return OK;
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return ERR_CUDA;
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
  }
  return g_num_sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode != nullptr) return OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return ERR_CUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return OK;
}

static int make_tmap_nd_typed(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, CUtensorMapDataType dtype);

int make_tmap_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, bool f32) {
  return make_tmap_nd_typed(out, base, rank, dims, strides_bytes, box, swizzle128,
                            f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
}

int make_tmap_2d_u8(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows, uint32_t box_cols) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[2] = {1, ld};
  const uint32_t box[2] = {box_cols, box_rows};
  return make_tmap_nd_typed(out, base, 2, dims, strides, box, true, CU_TENSOR_MAP_DATA_TYPE_UINT8);
}

// Encoded tensor maps are pure functions of (address, shape, strides, box, type, swizzle): cache them.  A GEMM launch needs
// 3-4 maps and cuTensorMapEncodeTiled costs ~1-2 us of host time each, which is what made the launch-bound configurations
// (ViT-Ti, single-token decode) need a CUDA graph; PyTorch's caching allocator hands the same addresses back step after step,
// so the hit rate in a training loop is ~100 %.  Forward runs on the Python main thread, backward on autograd's device
// thread: one mutex.  The cache is bounded (cleared when full) and holds no ownership of the memory it describes.
struct TmapKey {
  uint64_t w[13];
  bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t v : k.w) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); }
    return (size_t)h;
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static unsigned long long g_tmap_hits = 0, g_tmap_misses = 0;

static int encode_tmap_nd_typed(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                                const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, CUtensorMapDataType dtype);

static int make_tmap_nd_typed(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, CUtensorMapDataType dtype) {
  if (rank < 1 || rank > 5) { set_error("tensor map rank %d out of range", rank); return ERR_ARG; }
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.w[0] = (uint64_t)reinterpret_cast<uintptr_t>(base);
  key.w[1] = (uint64_t)rank | ((uint64_t)dtype << 8) | ((uint64_t)(swizzle128 ? 1 : 0) << 32);
  for (int i = 0; i < rank; ++i) {
    key.w[2 + i] = dims[i];
    key.w[7 + i] = (i > 0 ? strides_bytes[i] : 0) ^ ((uint64_t)box[i] << 48);
  }
  if (g_debug[11] != 1) {   // knob 11 = 1: bypass the cache (A/B)
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) { *out = it->second; ++g_tmap_hits; return OK; }
  }
  int rc = encode_tmap_nd_typed(out, base, rank, dims, strides_bytes, box, swizzle128, dtype);
  if (rc != OK) return rc;
  if (g_debug[11] != 1) {
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    if (g_tmap_cache.size() >= 8192) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *out);
    ++g_tmap_misses;
  }
  return OK;
}

static int encode_tmap_nd_typed(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                                const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128, CUtensorMapDataType dtype) {
  int rc = load_encode();
  if (rc != OK) return rc;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base pointer %p is not 16-byte aligned", base);
    return ERR_ARG;
  }
  for (int i = 1; i < rank; ++i) {
    if (strides_bytes[i] % 16 != 0) {
      set_error("TMA stride %llu (dim %d) is not a multiple of 16 bytes",
                (unsigned long long)strides_bytes[i], i);
      return ERR_ARG;
    }
  }
  CUresult r = g_encode(out, dtype, (cuuint32_t)rank,
                        const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (rank %d dims %llu,%llu box %u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
              rank > 1 ? box[1] : 0);
    return ERR_CUDA;
  }
  return OK;
}

int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  return make_tmap_nd(out, base, rank, dims, strides_bytes, box, swizzle128, false);
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[2] = {2, ld * 2};
  const uint32_t box[2] = {box_cols, box_rows};
  return make_tmap_nd(out, base, 2, dims, strides, box, true, false);
}

int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows, uint32_t box_cols) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[2] = {4, ld * 4};
  const uint32_t box[2] = {box_cols, box_rows};
  return make_tmap_nd(out, base, 2, dims, strides, box, true, true);
}

}  // namespace b200

extern "C" {

const char* b200vit_last_error(void) { return b200::g_err; }

int b200vit_version(void) { return B200VIT_VERSION; }

int b200vit_init(int device) {
  cudaDeviceProp prop;
  B200_CUDA(cudaSetDevice(device));
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    b200::set_error("device %d is sm_%d%d; libb200vit only contains sm_100a code (no fallback)", device,
                    prop.major, prop.minor);
    return b200::ERR_DEVICE;
  }
  b200::g_num_sms = prop.multiProcessorCount;
  return b200::load_encode();
}

/* bring-up aid: tensor-map cache statistics (hits in out[0], misses in out[1]) */
int b200vit_debug_tmap_cache_stats(unsigned long long* out) {
  if (out == nullptr) return b200::ERR_ARG;
  std::lock_guard<std::mutex> lock(b200::g_tmap_mu);
  out[0] = b200::g_tmap_hits; out[1] = b200::g_tmap_misses;
  return b200::OK;
}

int b200vit_debug_set(int key, int value) {
  if (key < 0 || key >= 16) return b200::ERR_ARG;
  b200::g_debug[key] = value;
  return b200::OK;
}

}  // extern "C"
