/* libb200vit — C ABI of the B200-native ViT-encoder hot path.
 *
 * The reference (SnakeOnex/vit-is-all-you-need) has no FFI: its boundary for this path is the Python
 * nn.Module surface (transformer.py:16-54, train_vit.py:30-45, train_titok.py:45-59, blocks.py:32-70,
 * blocks.py:405-505).  Every entry point below names the reference call site whose arithmetic it
 * replaces.  Conventions:
 *   - plain pointers + sizes, no framework types; all pointers are DEVICE pointers unless noted;
 *   - the caller owns every buffer (PyTorch's caching allocator on the Python side);
 *   - nothing allocates, synchronises or touches the default stream: work is enqueued on `stream`
 *     (a cudaStream_t passed as void*);
 *   - return 0 on success, negative on failure; b200vit_last_error() gives a thread-local message;
 *   - re-entrant: forward is called from the Python main thread, backward from autograd's device thread;
 *   - bf16 tensors are row-major and 16-byte aligned; "ld" arguments are in elements.
 */
#ifndef B200VIT_H_
#define B200VIT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VIT_VERSION 100

/* ---- runtime ---------------------------------------------------------------------------------- */
int b200vit_init(int device);            /* verifies sm_100, loads the TMA descriptor encoder        */
int b200vit_version(void);
const char* b200vit_last_error(void);
int b200vit_debug_set(int key, int value); /* bring-up knobs (descriptor sweeps); not for production */

/* ---- dense contractions: bf16 operands, fp32 accumulation in TMEM (tcgen05) --------------------- */
/* y[M,N](bf16) = x[M,K] w[N,K]^T + bias[N]            nn.Linear forward, transformer.py:21,27 (qkv)   */
int b200vit_gemm_bias(const void* x, const void* w, const float* bias, void* y, int M, int N, int K,
                      void* stream);
/* u = x w^T + bias ; g(bf16) = GELU_erf(u) ; u(bf16) stored when u != NULL   transformer.py:37-38,
 * blocks.py:51-52 */
int b200vit_gemm_bias_gelu(const void* x, const void* w, const float* bias, void* g, void* u, int M,
                           int N, int K, void* stream);
/* out[M,N](f32) = resid[M,N](f32) + x w^T + bias       transformer.py:39,44 ; blocks.py:53,67,69     */
int b200vit_gemm_bias_residual(const void* x, const void* w, const float* bias, const float* resid,
                               float* out, int M, int N, int K, void* stream);
/* out[M,N](f32) = x w^T + bias                                                                        */
int b200vit_gemm_bias_f32(const void* x, const void* w, const float* bias, float* out, int M, int N,
                          int K, void* stream);
/* dx[M,K](bf16) = dy[M,N] w[N,K]                       autograd of nn.Linear wrt input               */
int b200vit_gemm_dgrad(const void* dy, const void* w, void* dx, int M, int N, int K, void* stream);
/* dx[M,K](bf16) = (dy w) * GELU'(u[M,K])               autograd of Linear∘GELU, transformer.py:38-39 */
int b200vit_gemm_dgrad_dgelu(const void* dy, const void* w, const void* u, void* dx, int M, int N,
                             int K, void* stream);
/* dw[N,K](f32) (+)= dy[M,N]^T x[M,K]                   autograd of nn.Linear wrt weight              */
int b200vit_gemm_wgrad(const void* dy, const void* x, float* dw, int M, int N, int K, int accumulate,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VIT_H_ */
