/* ORACLE — test infrastructure only (never linked into or called by the product path).
 *
 * Plain-C restatement of the reference's VQ nearest-codebook lookup with a FIXED fp32 operation order,
 * so that the CUDA kernel (csrc/vq.cu) can be checked BIT-EXACTLY (indices, distances, quantised values):
 *
 *   train_titok.py:50-59   Quantizer.forward        (normalise x and codebook, cdist+argmin, raw-row gather)
 *   blocks.py:428-505      VectorQuantizer.forward  (optional l2-norm, expanded distance, normalised-row gather)
 *
 * Arithmetic contract (every operation is a single correctly-rounded fp32 op, in this order):
 *   ss   = fma(x_j, x_j, ss) for j = 0..D-1 (ss starts at 0)        sum of squares
 *   den  = max(sqrt(ss), 1e-12)                                     F.normalize eps
 *   xh_j = x_j / den
 *   xx   = fma(xh_j, xh_j, xx) ; ee_k likewise on the (normalised) code
 *   dot  = fma(xh_j, e_kj, dot) for j = 0..D-1
 *   dist = fma(-2, dot, xx + ee_k)                                  == blocks.py:440-442
 *   idx  = first k with the smallest dist (strict '<' scanning k upward == torch.argmin tie rule)
 *   q_j  = xh_j + (c_j - xh_j)                                      straight-through value
 * This contract is pinned against the reference modules by tests/test_oracle_golden.py (index equality on
 * every committed fixture).  Build: oracle/build.py (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static void normalize_row(const float* x, long stride, int D, float* out, float* norm_out) {
  float ss = 0.0f;
  for (int j = 0; j < D; ++j) ss = fmaf(x[j * stride], x[j * stride], ss);
  float n = sqrtf(ss);
  float den = n > 1e-12f ? n : 1e-12f;
  for (int j = 0; j < D; ++j) out[j] = x[j * stride] / den;
  if (norm_out) *norm_out = n;
}

/* x element (r, j) lives at x[(r / inner) * outer_stride + j * elem_stride + (r % inner)].
 *   [R, D] contiguous:     inner = 1,   elem_stride = 1,   outer_stride = D
 *   [b, c, h, w] (blocks): inner = h*w, elem_stride = h*w, outer_stride = c*h*w
 * flags: bit0 = l2-normalise x and codes for the search, bit1 = gather normalised rows (blocks.py) instead
 * of raw rows (train_titok.py).
 * Outputs: idx[R], quantized (same layout as x), sumsq[0] = sum over all elements of (c - xh)^2 in double,
 * dist_out (optional, [R] best distance). */
void vq_oracle_fwd(const float* x, const float* codebook, long R, int D, long K, long inner, long elem_stride,
                   long outer_stride, int flags, int64_t* idx, float* quantized, double* sumsq,
                   float* dist_out) {
  const int l2 = flags & 1, gather_norm = (flags >> 1) & 1;
  float* eh = (float*)malloc(sizeof(float) * (size_t)K * D);
  float* ee = (float*)malloc(sizeof(float) * (size_t)K);
  for (long k = 0; k < K; ++k) {
    if (l2) normalize_row(codebook + k * D, 1, D, eh + k * D, 0);
    else for (int j = 0; j < D; ++j) eh[k * D + j] = codebook[k * D + j];
    float s = 0.0f;
    for (int j = 0; j < D; ++j) s = fmaf(eh[k * D + j], eh[k * D + j], s);
    ee[k] = s;
  }
  float* xh = (float*)malloc(sizeof(float) * (size_t)D);
  double acc = 0.0;
  for (long r = 0; r < R; ++r) {
    const float* xr = x + (r / inner) * outer_stride + (r % inner);
    if (l2) normalize_row(xr, elem_stride, D, xh, 0);
    else for (int j = 0; j < D; ++j) xh[j] = xr[j * elem_stride];
    float xx = 0.0f;
    for (int j = 0; j < D; ++j) xx = fmaf(xh[j], xh[j], xx);
    float best = INFINITY;
    long bi = 0;
    for (long k = 0; k < K; ++k) {
      float dot = 0.0f;
      for (int j = 0; j < D; ++j) dot = fmaf(xh[j], eh[k * D + j], dot);
      float dist = fmaf(-2.0f, dot, xx + ee[k]);
      if (dist < best) { best = dist; bi = k; }
    }
    idx[r] = bi;
    if (dist_out) dist_out[r] = best;
    const float* c = (gather_norm ? eh : codebook) + bi * D;
    float* qr = quantized + (r / inner) * outer_stride + (r % inner);
    for (int j = 0; j < D; ++j) {
      float diff = c[j] - xh[j];
      qr[j * elem_stride] = xh[j] + diff;
      acc += (double)diff * (double)diff;
    }
  }
  *sumsq = acc;
  free(eh); free(ee); free(xh);
}

/* Backward of (quantized, losses) wrt x and the codebook (verified against autograd through
 * the golden fixtures).  g = upstream grad of `quantized` (layout of x); a_commit / a_code are the upstream
 * scalars multiplying d/dxh of mean((c-xh)^2) and d/dc of mean((c-xh)^2) respectively
 * (Quantizer: a_commit = 0.25*dL, a_code = dL).  dC must be zeroed by the caller. */
void vq_oracle_bwd(const float* x, const float* codebook, const int64_t* idx, const float* g, long R, int D,
                   long K, long inner, long elem_stride, long outer_stride, int flags, float a_commit,
                   float a_code, float* dx, float* dC) {
  const int l2 = flags & 1, gather_norm = (flags >> 1) & 1;
  const double nel = (double)R * D;
  float* xh = (float*)malloc(sizeof(float) * (size_t)D);
  float* ch = (float*)malloc(sizeof(float) * (size_t)D);
  for (long r = 0; r < R; ++r) {
    const long off = (r / inner) * outer_stride + (r % inner);
    float n = 1.0f;
    if (l2) normalize_row(x + off, elem_stride, D, xh, &n);
    else for (int j = 0; j < D; ++j) xh[j] = x[off + j * elem_stride];
    const float* craw = codebook + idx[r] * D;
    float cn = 1.0f;
    if (gather_norm) normalize_row(craw, 1, D, ch, &cn);
    else for (int j = 0; j < D; ++j) ch[j] = craw[j];
    double proj = 0.0, cproj = 0.0;
    for (int j = 0; j < D; ++j) {
      double dxh = (double)g[off + j * elem_stride] + (double)a_commit * 2.0 * ((double)xh[j] - ch[j]) / nel;
      double dc = (double)a_code * 2.0 * ((double)ch[j] - xh[j]) / nel;
      proj += (double)xh[j] * dxh;
      cproj += (double)ch[j] * dc;
    }
    const double den = n > 1e-12f ? n : 1e-12f;
    const double cden = cn > 1e-12f ? cn : 1e-12f;
    for (int j = 0; j < D; ++j) {
      double dxh = (double)g[off + j * elem_stride] + (double)a_commit * 2.0 * ((double)xh[j] - ch[j]) / nel;
      double dc = (double)a_code * 2.0 * ((double)ch[j] - xh[j]) / nel;
      dx[off + j * elem_stride] = (float)(l2 ? (dxh - xh[j] * proj) / den : dxh);
      dC[idx[r] * D + j] += (float)(gather_norm ? (dc - ch[j] * cproj) / cden : dc);
    }
  }
  free(xh); free(ch);
}
