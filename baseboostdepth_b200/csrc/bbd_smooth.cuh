// Edge-aware smoothness on mean-normalised disparity, forward + backward.
//
// Reference: trainer.py:560-564 (norm_disp = disp / (mean_hw(disp) + 1e-7)) and
// get_smooth_loss, layers.py:203-216.  All pyramid levels are handled by one launch per
// stage (grid.z = level); the three stages are separated by kernel boundaries because
// the mean normalisation couples every pixel of a sample:
//   stage 1  per-chunk sums of disp                      -> sample mean m
//   stage 2  per-pixel terms, loss partials, g_d = dL/d(norm disp), partials of sum(g_d * disp)
//   stage 3  g_disp = g_d / (m+eps) - sum(g_d*disp) / (N (m+eps)^2); level loss
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

// fixed-order block sum of one value per thread via shared memory (red: [NT + NT/16])
BBD_HD void sm_park(float* red, int tid, float v) { red[tid] = v; }
BBD_HD void sm_l1(float* red, int tid) {
  if (tid < SM_NT / 16) {
    float s = 0.0f;
    for (int i = 0; i < 16; ++i) s += red[tid * 16 + i];
    red[SM_NT + tid] = s;
  }
}
BBD_HD float sm_l2(const float* red) {
  float s = 0.0f;
  for (int i = 0; i < SM_NT / 16; ++i) s += red[SM_NT + i];
  return s;
}

BBD_HD float sm_stage1_thread(const SmoothArgs& a, int lvl, int b, int chunk, int tid) {
  const int n = a.h[lvl] * a.w[lvl];
  const float* d = a.disp[lvl] + (size_t)b * n;
  float s = 0.0f;
  for (int i = chunk * SM_CHUNK + tid; i < (chunk + 1) * SM_CHUNK && i < n; i += SM_NT) s += d[i];
  return s;
}

BBD_HD float sm_sample_mean(const SmoothArgs& a, int lvl, int b) {
  const float* p = sm_slot(a, lvl, b, 0);
  const int nc = sm_chunks(a.h[lvl], a.w[lvl]);
  float s = 0.0f;
  for (int i = 0; i < nc; ++i) s += p[i];
  return s / (float)(a.h[lvl] * a.w[lvl]);
}

// exp(-mean_c |I(i) - I(j)|)
BBD_HD float sm_edge(const float* img, int n, int i, int j) {
  float s = fabsf(sub(img[i], img[j]));
  s = add(s, fabsf(sub(img[n + i], img[n + j])));
  s = add(s, fabsf(sub(img[2 * n + i], img[2 * n + j])));
  return expf(-mul(s, BBD_THIRD));
}

// returns partial sums (tx, ty, g_d*disp) of this thread; writes g_d into gdisp
BBD_HD void sm_stage2_thread(const SmoothArgs& a, int lvl, int b, int chunk, int tid, float mean, float out[3]) {
  const int h = a.h[lvl], w = a.w[lvl], n = h * w;
  const float* d = a.disp[lvl] + (size_t)b * n;
  const float* img = a.img[lvl] + (size_t)b * 3 * n;
  float* g = a.gdisp[lvl] ? a.gdisp[lvl] + (size_t)b * n : nullptr;
  const float den = a.normalize ? add(mean, 1e-7f) : 1.0f;
  const float rden = div_(1.0f, den);  // div_const(x, den, rden) == x / den up to rare last-bit cases
  const float inx = 1.0f / ((float)a.batch * (float)h * (float)(w - 1));
  const float iny = 1.0f / ((float)a.batch * (float)(h - 1) * (float)w);
  float stx = 0.0f, sty = 0.0f, sgd = 0.0f;
  for (int i = chunk * SM_CHUNK + tid; i < (chunk + 1) * SM_CHUNK && i < n; i += SM_NT) {
    const int x = i % w, y = i / w;
    const float d0 = div_const(d[i], den, rden);
    float gd = 0.0f;
    if (x < w - 1) {
      const float diff = sub(d0, div_const(d[i + 1], den, rden));
      const float e = sm_edge(img, n, i, i + 1);
      stx += mul(fabsf(diff), e);
      gd += (diff > 0.0f ? e : (diff < 0.0f ? -e : 0.0f)) * inx;
    }
    if (x > 0) {
      const float diff = sub(div_const(d[i - 1], den, rden), d0);
      const float e = sm_edge(img, n, i - 1, i);
      gd -= (diff > 0.0f ? e : (diff < 0.0f ? -e : 0.0f)) * inx;
    }
    if (y < h - 1) {
      const float diff = sub(d0, div_const(d[i + w], den, rden));
      const float e = sm_edge(img, n, i, i + w);
      sty += mul(fabsf(diff), e);
      gd += (diff > 0.0f ? e : (diff < 0.0f ? -e : 0.0f)) * iny;
    }
    if (y > 0) {
      const float diff = sub(div_const(d[i - w], den, rden), d0);
      const float e = sm_edge(img, n, i - w, i);
      gd -= (diff > 0.0f ? e : (diff < 0.0f ? -e : 0.0f)) * iny;
    }
    if (g) g[i] = gd;
    sgd += gd * d[i];
  }
  out[0] = stx;
  out[1] = sty;
  out[2] = sgd;
}

// sum over the sample of g_d * disp (stage 2 partials), once per block
BBD_HD float sm_sample_gd_dot(const SmoothArgs& a, int lvl, int b) {
  const float* p = sm_slot(a, lvl, b, 3);
  const int nc = sm_chunks(a.h[lvl], a.w[lvl]);
  float s = 0.0f;
  for (int i = 0; i < nc; ++i) s += p[i];
  return s;
}

BBD_HD void sm_stage3_thread(const SmoothArgs& a, int lvl, int b, int chunk, int tid, float mean, float gd_dot) {
  if (!a.gdisp[lvl]) return;
  const int n = a.h[lvl] * a.w[lvl];
  float* g = a.gdisp[lvl] + (size_t)b * n;
  if (!a.normalize) return;  // g_d already is the gradient
  const float den = add(mean, 1e-7f);
  const float coupling = gd_dot / ((float)n * den * den);
  const float inv = 1.0f / den;
  for (int i = chunk * SM_CHUNK + tid; i < (chunk + 1) * SM_CHUNK && i < n; i += SM_NT) g[i] = g[i] * inv - coupling;
}

// level loss = sum_tx / (B h (w-1)) + sum_ty / (B (h-1) w); thread partial over (b, chunk) pairs
BBD_HD void sm_loss_thread(const SmoothArgs& a, int lvl, int tid, float out[2]) {
  const int nc = sm_chunks(a.h[lvl], a.w[lvl]);
  float sx = 0.0f, sy = 0.0f;
  for (int i = tid; i < a.batch * nc; i += SM_NT) {
    const int b = i / nc, c = i % nc;
    sx += sm_slot(a, lvl, b, 1)[c];
    sy += sm_slot(a, lvl, b, 2)[c];
  }
  out[0] = sx;
  out[1] = sy;
}

}  // namespace bbd
