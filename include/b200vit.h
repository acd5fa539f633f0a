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
int b200vit_debug_max_clusters(void);      /* bring-up aid: co-resident CTA pairs of the GEMM kernel */
int b200vit_debug_tmap_cache_stats(unsigned long long* hits_misses); /* TMA descriptor cache: out[0] hits, out[1] misses */

/* ---- dense contractions: bf16 operands, fp32 accumulation in TMEM (tcgen05) --------------------- */
/* y[M,N](bf16) = x[M,K] w[N,K]^T + bias[N]            nn.Linear forward, transformer.py:21,27 (qkv)   */
int b200vit_gemm_bias(const void* x, const void* w, const float* bias, void* y, int M, int N, int K,
                      void* stream);
/* u = x w^T + bias ; g(bf16) = GELU_erf(u) ; gprime(bf16) = GELU'(u) (saved for backward instead of u)
 * transformer.py:37-38, blocks.py:51-52 */
int b200vit_gemm_bias_gelu(const void* x, const void* w, const float* bias, void* g, void* gprime, int M,
                           int N, int K, void* stream);
/* out[M,N](f32) = resid[M,N](f32) + x w^T + bias       transformer.py:39,44 ; blocks.py:53,67,69     */
int b200vit_gemm_bias_residual(const void* x, const void* w, const float* bias, const float* resid,
                               float* out, int M, int N, int K, void* stream);
/* out[M,N](f32) = resid + dropout_p(x w^T + bias)      mlp[2] + nn.Dropout + residual, transformer.py:39-40,44;
 * keep mask = function of (seed, row, column) (csrc/dropout.cuh), kept values scaled by 1 / (1 - p)          */
int b200vit_gemm_bias_dropout_residual(const void* x, const void* w, const float* bias, const float* resid,
                                       float* out, int M, int N, int K, float p, unsigned int seed, void* stream);
/* out[M,N](f32) = x w^T + bias                                                                        */
int b200vit_gemm_bias_f32(const void* x, const void* w, const float* bias, float* out, int M, int N,
                          int K, void* stream);
/* dx[M,K](bf16) = dy[M,N] w[N,K]                       autograd of nn.Linear wrt input               */
int b200vit_gemm_dgrad(const void* dy, const void* w, void* dx, int M, int N, int K, void* stream);
/* dx[M,K](bf16) = (dy w) * gprime[M,K]                 autograd of Linear∘GELU, transformer.py:38-39;
 * gprime = GELU'(u) as written by b200vit_gemm_bias_gelu */
int b200vit_gemm_dgrad_dgelu(const void* dy, const void* w, const void* gprime, void* dx, int M, int N,
                             int K, void* stream);
/* The two GELU GEMMs above with GELU'(u) carried as an 8-bit fixed-point code instead of bf16:
 *   code = round((GELU'(u) - lo) / step), step = 1.27 / 255, lo = -27 step (GELU' lies in [-0.129, 1.129] for every u), decoded
 *   as lo + step * code: absolute error <= step / 2 = 0.0025.  Both kernels are bounded by the HBM bytes of this very tensor;
 *   the code halves them (fc1 forward writes [M, 4d] bytes less, its backward twin reads them less).  gprime_q8: uint8 [M, N]. */
int b200vit_gemm_bias_gelu_q8(const void* x, const void* w, const float* bias, void* g, void* gprime_q8, int M, int N, int K,
                              void* stream);
int b200vit_gemm_dgrad_dgelu_q8(const void* dy, const void* w, const void* gprime_q8, void* dx, int M, int N, int K,
                                void* stream);
float b200vit_gelu_grad_code_lo(void);
float b200vit_gelu_grad_code_step(void);
/* dw[N,K](f32) (+)= dy[M,N]^T x[M,K]                   autograd of nn.Linear wrt weight              */
int b200vit_gemm_wgrad(const void* dy, const void* x, float* dw, int M, int N, int K, int accumulate,
                       void* stream);
/* as above, plus db[N](f32) (+)= column sums of dy     autograd of nn.Linear wrt bias (transformer.py:21,37,39);
 * summed inside the wgrad kernel from the shared-memory dy tiles, so dy is not read a second time       */
int b200vit_gemm_wgrad_bias(const void* dy, const void* x, float* dw, float* db, int M, int N, int K,
                            int accumulate, void* stream);

/* ---- fused flash attention, head_dim 64 ----------------------------------------------------------------
 * qkv: [B, N, 3, H, 64] bf16 == the row-major output of the QKV Linear, "(qkv h d)" of transformer.py:27;
 * o: [B, N, H*64] bf16 == "b h n d -> b n (h d)" of transformer.py:29; lse: [B, H, N] fp32 (may be NULL).
 * Replaces F.scaled_dot_product_attention at transformer.py:28 (causal = additive -inf mask of
 * transformer.py:22-25) and the SDPA inside nn.MultiheadAttention (blocks.py:60).
 * seq_first != 0: tensors are [N, B, ...] (the LND layout of blocks.py:270) instead of [B, N, ...]. */
int b200vit_flash_attn_fwd(const void* qkv, void* o, float* lse, int B, int N, int H, int causal, int seq_first,
                           void* stream);
/* bytes of fp32 workspace the backward needs (0 for N <= 256: the resident kernels accumulate in TMEM only)        */
size_t b200vit_flash_attn_bwd_workspace_size(int B, int N, int H);
/* dqkv: [B, N, 3, H, 64] bf16 gradient of qkv given d_o [B, N, H*64] bf16 */
int b200vit_flash_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, int B,
                           int N, int H, int causal, int seq_first, void* workspace, size_t workspace_bytes,
                           void* stream);
/* The same with dropout on the attention probabilities (dropout_p of F.scaled_dot_product_attention,
 * transformer.py:28 -- applied in eval mode too, as the reference does).  The keep mask is a pure function of
 * (seed, batch*H + head, query, key) (csrc/dropout.cuh): backward regenerates it from the same seed.        */
int b200vit_flash_attn_fwd_dropout(const void* qkv, void* o, float* lse, int B, int N, int H, int causal,
                                   int seq_first, float dropout_p, unsigned int seed, void* stream);
int b200vit_flash_attn_bwd_dropout(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv,
                                   int B, int N, int H, int causal, int seq_first, float dropout_p,
                                   unsigned int seed, void* workspace, size_t workspace_bytes, void* stream);

/* ---- LayerNorm on the fp32 residual stream (F.layer_norm transformer.py:43-44; nn.LayerNorm blocks.py:43,48)
 * fwd: v = x (+ add_bf16) ; x_out = v (optional) ; y = LN(v) * gamma + beta -> bf16 and/or fp32 ; saves mean, rstd
 * bwd: dx = dres (optional) + LN'(dy) ; optional bf16 copy ; dgamma / dbeta (both or neither)               */
int b200vit_layernorm_fwd(const float* x, const void* add_bf16, float* x_out, const float* gamma, const float* beta,
                          void* y_bf16, float* y_f32, float* mean, float* rstd, int M, int d, float eps, void* stream);
int b200vit_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* mean,
                          const float* rstd, const float* gamma, const float* dres, float* dx, void* dx_bf16,
                          float* dgamma, float* dbeta, int M, int d, void* stream);
/* backward of the affine-free LayerNorm (transformer.py:43-44) from the SAVED bf16 forward output xhat = LN(x)
 * instead of the fp32 row and its mean: dx = dres (optional) + rstd * (dy - mean(dy) - xhat * mean(dy * xhat));
 * 14 instead of 16 bytes per element and the fp32 residual rows need not be kept for backward.               */
int b200vit_layernorm_bwd_xhat(const void* dy_bf16, const void* xhat_bf16, const float* rstd, const float* dres,
                               float* dx, void* dx_bf16, int M, int d, void* stream);
/* out[N](f32) (+)= column sums of a[M,N](bf16): bias gradients of nn.Linear                                  */
int b200vit_colsum_bf16(const void* a, float* out, int M, int N, int accumulate, void* stream);
int b200vit_colsum_f32(const float* a, float* out, int rows, int n, void* stream);
int b200vit_cast_f32_bf16(const float* in, void* out, long long n, void* stream);
/* out(f32)[n] = scale * in(bf16)[n]: gradients coming back from a bf16-compressed all-reduce (b200vit/ddp.py)            */
int b200vit_cast_bf16_f32(const void* in, float* out, long long n, float scale, void* stream);
/* out(bf16) = x(f32) * keep / (1 - p): backward of nn.Dropout (transformer.py:40) fused with the cast that feeds the
 * fc2 dgrad / wgrad GEMMs; same mask function as b200vit_gemm_bias_dropout_residual                          */
int b200vit_dropout_cast_bf16(const float* x, void* out_bf16, long long M, int d, float p, unsigned int seed,
                              void* stream);
/* test aids: the keep masks as bytes, [M, d] and [B, H, N, N]                                                */
int b200vit_dropout_mask_rows(unsigned char* out, long long M, int d, float p, unsigned int seed, void* stream);
int b200vit_dropout_mask_attn(unsigned char* out, int B, int H, int N, float p, unsigned int seed, void* stream);

/* ---- patch embedding (train_vit.py:34-45, blocks.py:235-237,257-267) ------------------------------------
 * tokens[B, extra+P, d](f32): rows [0,extra) = extra_emb, rows [extra, ..) = conv(x) + bias + pos_emb.
 * cols: caller-provided bf16 [B*P, C*p*p] im2col buffer (kept for the weight gradient).                     */
int b200vit_patch_embed_fwd(const float* x, const void* w_bf16, const float* bias, const float* pos_emb,
                            const float* extra_emb, float* tokens, void* cols, int B, int C, int H, int W, int p,
                            int d, int extra, void* stream);
/* dsum[T, d] = sum_b dtokens[b] ; dpe_bf16[B*P, d] = bf16(dtokens[:, extra:])                                */
int b200vit_patch_embed_bwd_reduce(const float* dtokens, float* dsum, void* dpe_bf16, int B, int T, int extra,
                                   int d, void* stream);
/* Token-sequence assembly of blocks.TiTokEncoder / TiTokDecoder (blocks.py:254-268, 337-352): ONE GEMM over bf16 operand
 * rows (im2col of the image for the encoder's patch_embed blocks.py:235,257; the latent rows for the decoder's
 * decoder_embed blocks.py:310,341) whose epilogue adds bias + pos and writes rows [extra, extra+P) of tokens[B, T, d]
 * (T = extra + P + tail, fp32), plus one pass that fills the rows no GEMM produces:
 *   head rows e < extra : (e == 0 ? head0 : head1)[d] + head_pos[e, d]  (class_embedding / mask_token + positional_embedding)
 *   tail rows l < tail  : tail_a[l, d] + tail_b[l, d]                   (latent_tokens + latent_token_positional_embedding)
 * i.e. torch.cat / expand / the two position adds of blocks.py:261-267 never run as separate passes.              */
int b200vit_tokens_assemble_fwd(const void* cols_bf16, const void* w_bf16, const float* bias, const float* pos,
                                const float* head0, const float* head1, const float* head_pos, const float* tail_a,
                                const float* tail_b, float* tokens, int B, int P, int K, int d, int extra, int tail,
                                void* stream);
/* dsum[T, d] = sum_b dtokens[b] (gradients of every broadcast row and positional table) ;
 * dpe_bf16[B*P, d] = bf16(dtokens[:, extra:extra+P]) (operand of the wgrad / dgrad GEMMs)                          */
int b200vit_tokens_assemble_bwd_reduce(const float* dtokens, float* dsum, void* dpe_bf16, int B, int T, int extra, int tail,
                                       int d, void* stream);
/* out[B, cnt, d] (f32) = x[:, t0:t0+cnt, :] -- the ln_post input of blocks.py:276,359                               */
int b200vit_gather_tokens_f32(const float* x, float* out, int B, int N, int d, int t0, int cnt, void* stream);
/* Affine LayerNorm folded into the Linear that follows it (blocks.py:66 ln_1 -> attn.in_proj, blocks.py:69 ln_2 -> mlp.c_fc):
 * Linear(gamma*xhat + beta) = xhat (W diag(gamma))^T + (b + W beta), so blocks.ResidualAttentionBlock runs on the
 * affine-free LayerNorm / GEMM kernels of transformer.TransformerLayer and keeps only the bf16 xhat for backward.
 *   fold  : W_bf16[N,K] = bf16(W * gamma) ; bias_out[N] = bias (may be NULL) + W beta
 *   unfold: dW[n,k] = dW'[n,k] gamma[k] + dbias[n] beta[k] (in place; dW' = the wgrad result on xhat, dbias = its bias
 *           gradient) ; dgamma[K] (+)= sum_n dW'[n,k] W[n,k] ; dbeta[K] (+)= sum_n W[n,k] dbias[n]   (fixed summation order) */
int b200vit_affine_fold(const float* W, const float* bias, const float* gamma, const float* beta, void* W_bf16,
                        float* bias_out, int N, int K, void* stream);
size_t b200vit_affine_unfold_workspace_size(int K);
int b200vit_affine_unfold_grads(float* dW, const float* W, const float* gamma, const float* beta, const float* dbias,
                                float* dgamma, float* dbeta, int N, int K, int accumulate, void* workspace,
                                size_t workspace_bytes, void* stream);
/* de-patchify tail of the tokenizer decoders (train_titok.py:67,72-74): 1x1 Conv2d(d -> C*p*p) over the patch tokens +
 * "b (p1 p2 c) h w -> b c (h p1) (w p2)" as ONE GEMM whose epilogue stores straight into the NCHW image.
 * rows_bf16[B*P, d]; w_cmajor_bf16[C*p*p, d] and bias_cmajor[C*p*p] hold the conv's output channels re-ordered from the
 * reference's (p1 p2 c) to (c p1 p2); img[B, C, Ht*p, Wt*p] fp32; p a power of two >= 4.  Backward: im2col_bf16 of the
 * image gradient yields exactly the (c p1 p2)-ordered rows for the wgrad / dgrad GEMMs.                        */
int b200vit_depatchify_fwd(const void* rows_bf16, const void* w_cmajor_bf16, const float* bias_cmajor, float* img, int B,
                           int Ht, int Wt, int p, int C, int d, void* stream);
int b200vit_im2col_bf16(const float* x, void* cols, int B, int C, int H, int W, int p, void* stream);
int b200vit_col2im_f32(const void* dcols, float* dx, int B, int C, int H, int W, int p, void* stream);

/* ---- VQ nearest-codebook lookup (train_titok.py:50-59 Quantizer; blocks.py:428-505 VectorQuantizer) -------
 * x element (r, j) at x[(r / inner) * outer_stride + j * elem_stride + r % inner]  ([R,D]: 1,1,D ; bchw: hw,hw,chw)
 * flags bit0: l2-normalise x and codes; bit1: gather normalised code rows (blocks.py) instead of raw rows.
 * indices[R] int64 (bit-exact vs oracle/vq_oracle.c), quantized = xh + (c - xh) in x's layout,
 * losses[3] = { mse, commitment_cost * mse, (1 + commitment_cost) * mse }.                                  */
size_t b200vit_vq_workspace_size(long long R, int D, int K);
int b200vit_vq_fwd(const float* x, const float* codebook, long long R, int D, int K, long long inner,
                   long long elem_stride, long long outer_stride, int flags, float commitment_cost,
                   long long* indices, float* quantized, float* losses, void* workspace, size_t workspace_bytes,
                   void* stream);
/* coef[2] (device): upstream scalars for d mean((c-xh)^2)/d xh and /d c ; dcodebook[K,D] is overwritten       */
int b200vit_vq_bwd(const float* x, const float* codebook, const long long* indices, const float* grad_q,
                   const float* coef, long long R, int D, int K, long long inner, long long elem_stride,
                   long long outer_stride, int flags, float* dx, float* dcodebook, void* stream);

/* ---- classifier head plumbing + cross-entropy (train_vit.py:51-53,81,102; train_videogpt.py:53-54) -----------
 * gather_tokens : out_bf16[B*cnt, d] = bf16(x[:, t0:t0+cnt, :]) of x[B, N, d] (fp32) -- the operand of the head GEMM
 *   (cnt = 1), of TiTokEncoder.proj (train_titok.py:41-42) and of the de-patchify GEMM (train_titok.py:71)
 * scatter_tokens: dx[B, N, d] (fp32) = 0 except dx[:, t0:t0+cnt] = dy[B*cnt, d] (bf16 or fp32); optional bf16 twin
 * cross_entropy_fwd: logits[R, C] (row stride ld; bf16 or fp32), labels[R] int64 -> loss[0] = mean over rows whose
 *   label != ignore_index of (logsumexp(x) - x[label]), loss[1] = 1 / n_valid, lse[R]; row_loss[R] is scratch.
 * cross_entropy_bwd: dlogits (dtype of the logits, row stride ldd) = (softmax(x) - onehot) * *dloss * loss[1].   */
int b200vit_gather_tokens_bf16(const float* x, void* out_bf16, int B, int N, int d, int t0, int cnt, void* stream);
int b200vit_scatter_tokens(const void* dy, int dy_is_bf16, float* dx, void* dx_bf16, int B, int N, int d, int t0, int cnt,
                           void* stream);
int b200vit_cross_entropy_fwd(const void* logits, int logits_bf16, long long ld, const long long* labels, float* loss,
                              float* lse, float* row_loss, int R, int C, long long ignore_index, void* stream);
int b200vit_cross_entropy_bwd(const void* logits, int logits_bf16, long long ld, const long long* labels, const float* lse,
                              const float* loss, const float* dloss, void* dlogits, long long ldd, int R, int C,
                              long long ignore_index, void* stream);

/* token + positional embedding of the autoregressive models (train_videogpt.py:45-50):
 * out[b, s, :] = tok_embed[idx[b, s]] + pos_embed[pos0 + s] (fp32);  backward: dtok[vocab, d] (overwritten, fp32 atomics
 * over repeated tokens) and dpos[S, d] = sum_b dy[b, s].                                                       */
int b200vit_embed_fwd(const long long* idx, const float* tok_embed, const float* pos_embed, float* out, int B, int S, int d,
                      int pos0, const int* pos0_dev /* optional: position from device memory (graph-captured decode) */,
                      int n_pos, int vocab, void* stream);
int b200vit_embed_bwd(const long long* idx, const float* dy, float* dtok, float* dpos, int B, int S, int d, int vocab,
                      void* stream);

/* incremental decode attention for VideoGPT.generate (train_videogpt.py:56-65).  kv_cache[2][B][H][Nmax][64] bf16: K planes
 * then V planes, one head's rows contiguous (the fused-QKV layout made every key row an isolated 128-byte piece: ~2 TB/s).
 * kv_fill copies the prompt's fused QKV projection qkv[B, S, 3, H, 64] into rows 0..S; kv_append writes the K / V of the
 * new token (slots 1, 2 of its fused row qkv_rows[B, 3*H*64]) at position *pos_dev; attn_decode:
 * out_bf16[B, H*64] = softmax(q . K[0..pos]^T / 8) V[0..pos] with q = slot 0 of qkv_rows.  The position is read from DEVICE
 * memory so that one captured CUDA graph serves every step of a generation; advance_counter bumps it.         */
int b200vit_attn_decode(const void* qkv_rows, const void* kv_cache, void* out_bf16, int B, int Nmax, int H, const int* pos_dev,
                        void* stream);
int b200vit_kv_append(const void* qkv_rows, void* kv_cache, int B, int Nmax, int H, const int* pos_dev, void* stream);
int b200vit_kv_fill(const void* qkv, void* kv_cache, int B, int S, int Nmax, int H, void* stream);
int b200vit_advance_counter(int* counter, int by, void* stream);

/* ---- fused multi-tensor AdamW + bf16 operand refresh (torch.optim.AdamW at train_vit.py:82,105, ------------
 * train_titok.py:134,160, train_videogpt.py:107,134; arithmetic of torch/optim/adam.py::_single_tensor_adam)
 * tensors: device array of { float* p; const float* g; float* m; float* v; bf16* w16 (or NULL); long long n } ;
 * chunks: device array of { int tensor; int chunk_index } -- one CTA per b200vit_adamw_chunk_elems() elements.
 * step: 1-based step number (host) -- or step_dev: device fp32 counter that the call advances (CUDA-graph capture,
 * GradScaler).  grad_scale / found_inf: optional device scalars of torch.amp.GradScaler: g /= *grad_scale; when
 * *found_inf != 0 nothing (not even the counter) changes.                                                     */
int b200vit_adamw_chunk_elems(void);
int b200vit_adamw_step(const void* tensors, const void* chunks, int n_chunks, double lr, double beta1, double beta2,
                       double eps, double weight_decay, long long step, float* step_dev, const float* grad_scale,
                       const float* found_inf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VIT_H_ */
