// Edge-aware smoothness on mean-normalised disparity, forward + backward.
//
// Reference: trainer.py:560-564 (norm_disp = disp / (mean_hw(disp) + 1e-7)) and
// get_smooth_loss, layers.py:203-216.  The mean normalisation couples every pixel of a sample, but only
// through one positive scalar r = 1 / (mean + eps): |d_x (r d)| = r |d_x d| and sign(d_x (r d)) = sign(d_x d),
// so ONE pass over the planes (all pyramid levels in one launch) works on the raw disparity and accumulates
//   sum d, sum |dx d| e, sum |dy d| e, sum g_d * d   per block, and writes g_d = dL/d(norm disp);
// the last block to finish reduces the partials (fixed order), forms the sample means, scales the loss sums by
// r and leaves the two per-sample scalars that turn g_d into the gradient w.r.t. the raw disparity
//   g_disp = g_d / (m+eps) - sum(g_d*disp) / (N (m+eps)^2)
// (applied by the disparity backward on the fly, or by sm_apply_px for stand-alone callers).
// Same host/device phase structure as bbd_strip.cuh.
#pragma once
#include "bbd_common.cuh"

namespace bbd {

using SmoothArgs = bbd_smooth_args;

constexpr int SM_NT = 256;
constexpr int SM_CHUNK = 2048;  // pixels per block

BBD_HD int sm_chunks(int h, int w) { return (h * w + SM_CHUNK - 1) / SM_CHUNK; }
BBD_HD float* sm_slot(const SmoothArgs& a, int lvl, int b, int which) {
  return a.scratch + (((size_t)lvl * a.batch + b) * 4 + which) * a.max_chunks;
}

// ---------------------------------------------------------------------------------------------------
// The pass in row-walking form (round 2).  A warp owns 30 columns (lanes 1..30; lanes 0 and 31 are the
// neighbours' columns) of a chunk of rows and walks down: every edge weight exp(-mean_c |dI|) is
// evaluated once and handed to the pixel on its other side by a shuffle (x) or kept in a register for
// the next row (y); no div/mod per pixel.  Writes g_d = dL/d(norm disp) and per-block partial sums
// (sum |dx d| e, sum |dy d| e, sum g_d * disp).  Plain arithmetic (reciprocal multiply): the forward value
// agrees with the reference to ~1e-7 relative.
// With `defer_norm` the last block to finish (ticket in the scratch buffer; the sums it forms are in a fixed
// order, so the result does not depend on which block that is) reduces the partials to the level losses and to
// the two per-sample scalars that turn g_d into the gradient w.r.t. the raw disparity,
//     g_disp = g_d * coef[0] - coef[1],   coef = (1/(m+eps), sum(g_d*disp) / (N (m+eps)^2)),
// which the disparity backward (d2d_backward_px) applies on the fly -- no third pass over the planes.
// ---------------------------------------------------------------------------------------------------
#ifndef BBD_SMR_RC
#define BBD_SMR_RC 16
#endif
constexpr int SMR_TW = 30, SMR_WARPS = 4, SMR_RC = BBD_SMR_RC;
BBD_HD int smr_nbx(int w) { return (w + SMR_TW * SMR_WARPS - 1) / (SMR_TW * SMR_WARPS); }
BBD_HD int smr_nby(int h) { return (h + SMR_RC - 1) / SMR_RC; }
BBD_HD int smr_blocks(int h, int w) { return smr_nbx(w) * smr_nby(h); }

#if defined(__CUDA_ARCH__)
BBD_HD float smr_up(float v) { return __shfl_up_sync(0xffffffffu, v, 1); }
BBD_HD float smr_down(float v) { return __shfl_down_sync(0xffffffffu, v, 1); }
BBD_HD float smr_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
#elif defined(BBD_EMU)
BBD_HD float smr_up(float v) { return simt::shfl_up(v, 1); }
BBD_HD float smr_down(float v) { return simt::shfl_down(v, 1); }
BBD_HD float smr_xor(float v, int m) { return simt::shfl_xor(v, m); }
#else
BBD_HD float smr_up(float v) { return v; }
BBD_HD float smr_down(float v) { return v; }
BBD_HD float smr_xor(float v, int) { return v; }
#endif

BBD_HD float smr_sign(float d) { return d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f); }

// one lane of one warp: out[3] = warp totals (valid in every lane)
BBD_HD void sm_rows_lane(const SmoothArgs& a, int lvl, int b, int bx, int by, int warp, int lane, float out[4]) {
  const int h = a.h[lvl], w = a.w[lvl], n = h * w;
  const float* d = a.disp[lvl] + (size_t)b * n;
  const float* img = a.img[lvl] + (size_t)b * 3 * n;
  float* g = a.gdisp[lvl] ? a.gdisp[lvl] + (size_t)b * n : nullptr;
  const float inx = 1.0f / ((float)a.batch * (float)h * (float)(w - 1));
  const float iny = 1.0f / ((float)a.batch * (float)(h - 1) * (float)w);
  const int x = (bx * SMR_WARPS + warp) * SMR_TW + lane - 1;
  const bool valid = x >= 0 && x < w;
  const bool own = lane >= 1 && lane <= SMR_TW && x < w;
  const int xc = x < 0 ? 0 : (x >= w ? w - 1 : x);
  const int y0 = by * SMR_RC, y1 = (y0 + SMR_RC < h) ? y0 + SMR_RC : h;
  const int ys = y0 > 0 ? y0 - 1 : 0;  // one row above the chunk seeds the vertical term
  float stx = 0.0f, sty = 0.0f, sgd = 0.0f, sd = 0.0f, sy_prev = 0.0f;
  float dn = d[ys * w + xc], In[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) In[c] = img[c * n + ys * w + xc];
  for (int y = ys; y < y1; ++y) {
    const float draw = dn, d0 = draw;  // raw disparity: the normalisation is a positive scale applied at the end
    float I0[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) I0[c] = In[c];
    const int yn = (y + 1 < h) ? y + 1 : y;
    dn = d[yn * w + xc];
#pragma unroll
    for (int c = 0; c < 3; ++c) In[c] = img[c * n + yn * w + xc];
    const float dr = smr_down(d0);
    float ax = 0.0f, ay = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ax += fabsf(I0[c] - smr_down(I0[c]));
      ay += fabsf(I0[c] - In[c]);
    }
    const bool has_r = valid && x < w - 1 && lane < 31, has_d = y < h - 1;
    const float ex = has_r ? expf(-ax * BBD_THIRD) : 0.0f;
    const float ey = has_d ? expf(-ay * BBD_THIRD) : 0.0f;
    const float dfx = d0 - dr, dfy = d0 - dn;
    const float sx = smr_sign(dfx) * ex * inx, sy = smr_sign(dfy) * ey * iny;
    const float sx_up = smr_up(sx);
    const float sxl = (lane > 0 && x > 0) ? sx_up : 0.0f;
    if (y >= y0 && own) {
      const float gd = sx - sxl + sy - sy_prev;
      if (g) g[y * w + x] = gd;
      stx += fabsf(dfx) * ex;
      sty += fabsf(dfy) * ey;
      sgd += gd * draw;
      sd += draw;
    }
    sy_prev = sy;
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    stx += smr_xor(stx, m);
    sty += smr_xor(sty, m);
    sgd += smr_xor(sgd, m);
    sd += smr_xor(sd, m);
  }
  out[0] = stx;
  out[1] = sty;
  out[2] = sgd;
  out[3] = sd;
}

// fixed-order reduction of the block partials of one (level, sample): thread-serial, tiny
BBD_HD void sm_finish_sample(const SmoothArgs& a, int lvl, int b, float* coef, float sums[2]) {
  const int nb = smr_blocks(a.h[lvl], a.w[lvl]);
  const float* p0 = sm_slot(a, lvl, b, 0);
  const float* p1 = sm_slot(a, lvl, b, 1);
  const float* p2 = sm_slot(a, lvl, b, 2);
  const float* p3 = sm_slot(a, lvl, b, 3);
  float tx = 0.0f, ty = 0.0f, gd = 0.0f, sd = 0.0f;
  for (int i = 0; i < nb; ++i) { tx += p1[i]; ty += p2[i]; gd += p3[i]; sd += p0[i]; }
  const int n = a.h[lvl] * a.w[lvl];
  const float den = a.normalize ? sd / (float)n + 1e-7f : 1.0f;
  const float rden = 1.0f / den;
  sums[0] = tx * rden;
  sums[1] = ty * rden;
  if (coef) {
    coef[((size_t)lvl * a.batch + b) * 2] = rden;
    coef[((size_t)lvl * a.batch + b) * 2 + 1] = a.normalize ? gd / ((float)n * den * den) : 0.0f;
  }
}
// g_disp = g_d * coef[0] - coef[1] in place (non-deferred callers)
BBD_HD void sm_apply_px(const SmoothArgs& a, const float* coef, int lvl, int b, int i) {
  float* g = a.gdisp[lvl] + (size_t)b * a.h[lvl] * a.w[lvl];
  const float* c = coef + ((size_t)lvl * a.batch + b) * 2;
  g[i] = g[i] * c[0] - c[1];
}
BBD_HD float sm_level_loss(const SmoothArgs& a, int lvl, const float* tx, const float* ty) {
  float sx = 0.0f, sy = 0.0f;
  for (int b = 0; b < a.batch; ++b) { sx += tx[b]; sy += ty[b]; }
  const float h = (float)a.h[lvl], w = (float)a.w[lvl], B = (float)a.batch;
  return sx / (B * h * (w - 1.0f)) + sy / (B * (h - 1.0f) * w);
}

}  // namespace bbd