// Per-element host/device bodies of the small operators around the fused loss:
// disparity -> depth (and back), on-demand warping, and the module-level
// BackprojectDepth / Project3D / SSIM operators.  Kernels in bbd_kernels.cu map one
// thread to one element (grid-stride); tests/emu loops over elements on the CPU.
#pragma once
#include "bbd_common.cuh"

namespace bbd {

// ---- F.interpolate(bilinear, align_corners=False) source taps (ATen UpSample.h) ----------
struct Lerp {
  int i0, i1;
  float l0, l1;
};
// scale = float(in_size) / float(out_size), computed once per level (area_pixel_compute_scale)
BBD_HD Lerp up_taps(int o, int in_size, float scale) {
  float src = sub(mul(scale, add((float)o, 0.5f)), 0.5f);
  if (src < 0.0f) src = 0.0f;
  Lerp t;
  t.i0 = (int)src;
  t.i1 = t.i0 + ((t.i0 < in_size - 1) ? 1 : 0);
  t.l1 = sub(src, (float)t.i0);
  t.l0 = sub(1.0f, t.l1);
  return t;
}

BBD_HD float d2d_up(const float* d, int w, const Lerp& ty, const Lerp& tx) {
  const float* r0 = d + ty.i0 * w;
  const float* r1 = d + ty.i1 * w;
  // rounding pattern of ATen's upsample_bilinear2d (checked bit-for-bit against torch CPU):
  // each lerp is fma(w0, a, RN(w1 * b))
  const float top = fma_(tx.l0, r0[tx.i0], mul(tx.l1, r0[tx.i1]));
  const float bot = fma_(tx.l0, r1[tx.i0], mul(tx.l1, r1[tx.i1]));
  return fma_(ty.l0, top, mul(ty.l1, bot));
}

// one full-resolution pixel: upsample + disp_to_depth (layers.py:13-22)
BBD_HD float d2d_forward_px(const bbd_d2d_args& a, int lvl, int b, int oy, int ox, float sy, float sx) {
  const int h = a.h[lvl], w = a.w[lvl];
  const float* d = a.disp[lvl] + (size_t)b * h * w;
  const float up = d2d_up(d, w, up_taps(oy, h, sy), up_taps(ox, w, sx));
  if (a.sql) return up;
  return rcp_rn(add(a.min_disp, mul(a.disp_span, up)));
}

// one low-resolution disparity pixel: gather d(loss)/d(disp) from the full-res pixels it fed.
// For an integer factor f the pixel (iy, ix) is a tap of the outputs in [f*i - f, f*i + 2f);
// the weight of output o for input i is l0 if i0 == i plus l1 if i1 == i (both can hold at the
// clamped borders), exactly the transpose of d2d_up.
BBD_HD float d2d_axis_weight(int o, int i, int in_size, float scale) {
  const Lerp t = up_taps(o, in_size, scale);
  return (t.i0 == i ? t.l0 : 0.0f) + (t.i1 == i ? t.l1 : 0.0f);
}

// ---- separable gather for the factors 2, 4, 8 ------------------------------------------------
// factor of a level if the two-pass path applies to it, else 0
BBD_HD int d2d_sep_factor(const bbd_d2d_args& a, int lvl) {
  const int h = a.h[lvl], w = a.w[lvl];
  for (int f = 2; f <= 8; f *= 2)
    if (a.height == f * h && a.width == f * w) return f;
  return 0;
}
// offset (floats) of level lvl's row-sum plane (B,H,w) inside the scratch buffer
BBD_HD size_t d2d_scratch_offset(const bbd_d2d_args& a, int lvl) {
  size_t off = 0;
  for (int l = 0; l < lvl; ++l)
    if (d2d_sep_factor(a, l)) off += (size_t)a.batch * a.height * a.w[l];
  return off;
}
// Weight of full-resolution output o = F*i + t for low-resolution input i, t in [-F/2, 3F/2): the
// transpose of d2d_up for an integer factor F.  src(o) - i = (t + 0.5)/F - 0.5 is a dyadic number,
// so up_taps computes l0/l1 exactly and the tent below reproduces them bit for bit.  At the clamped
// borders both taps of the outer half-cell fall on the border input: its weight there is 1.
template <int F>
BBD_HD float d2d_tent(int t, int i, int in_size) {
  float wgt = 1.0f - fabsf(((float)t + 0.5f) * (1.0f / (float)F) - 0.5f);
  if (i == 0 && t < F / 2) wgt = 1.0f;
  if (i == in_size - 1 && t >= F / 2) wgt = 1.0f;
  return wgt;
}
// pass 1: for one full-resolution row oy and one low-resolution column ix, the weighted sum over
// the 2F outputs that have ix as a tap (chain rule of disp_to_depth applied on the fly)
template <int F>
BBD_HD float d2d_hpass(const bbd_d2d_args& a, int lvl, int b, int oy, int ix) {
  const int w = a.w[lvl], H = a.height, W = a.width;
  const size_t plane = ((size_t)lvl * a.batch + b) * H * W + (size_t)oy * W;
  const float* gd = a.gdepth + plane;
  const float* dep = a.depth + plane;
  const float nspan = -a.disp_span;
  const int ox0 = ix * F - F / 2;
  float acc = 0.0f;
  if (ix > 0 && ix < w - 1) {
    // interior column: all 2F taps are inside the row and carry the plain tent weights (constants)
    const float* g = gd + ox0;
    const float* dp = dep + ox0;
#pragma unroll
    for (int k = 0; k < 2 * F; ++k) {
      float v = g[k];
      if (!a.sql) { const float d = dp[k]; v *= nspan * d * d; }
      acc += d2d_tent<F>(k - F / 2, 1, 3) * v;
    }
    return acc;
  }
#pragma unroll
  for (int k = 0; k < 2 * F; ++k) {
    const int ox = ox0 + k;
    if (ox < 0 || ox >= W) continue;
    float v = gd[ox];
    if (!a.sql) { const float d = dep[ox]; v *= nspan * d * d; }
    acc += d2d_tent<F>(k - F / 2, ix, w) * v;
  }
  return acc;
}
// pass 2: weighted sum of the row sums over the 2F rows that have iy as a tap
template <int F>
BBD_HD float d2d_vpass(const bbd_d2d_args& a, int lvl, int b, int iy, int ix) {
  const int h = a.h[lvl], w = a.w[lvl], H = a.height;
  const float* tmp = a.scratch + d2d_scratch_offset(a, lvl) + (size_t)b * H * w;
  const int oy0 = iy * F - F / 2;
  float acc = 0.0f;
  if (iy > 0 && iy < h - 1) {  // interior row: plain tent weights, no bounds checks
    const float* col = tmp + (size_t)oy0 * w + ix;
#pragma unroll
    for (int k = 0; k < 2 * F; ++k) acc += d2d_tent<F>(k - F / 2, 1, 3) * col[(size_t)k * w];
    return acc;
  }
#pragma unroll
  for (int k = 0; k < 2 * F; ++k) {
    const int oy = oy0 + k;
    if (oy < 0 || oy >= H) continue;
    acc += d2d_tent<F>(k - F / 2, iy, h) * tmp[(size_t)oy * w + ix];
  }
  return acc;
}

// integer factor F known at compile time: the 2F+2 candidate weights per axis live in registers
template <int F>
BBD_HD float d2d_backward_gather(const bbd_d2d_args& a, int lvl, int b, int iy, int ix, float sy, float sx) {
  constexpr int N = 2 * F + 2;
  const int h = a.h[lvl], w = a.w[lvl], H = a.height, W = a.width;
  const float* gd = a.gdepth + ((size_t)lvl * a.batch + b) * H * W;
  const float* dep = a.depth + ((size_t)lvl * a.batch + b) * H * W;
  const float nspan = -a.disp_span;
  const int oy0 = iy * F - F / 2 - 1, ox0 = ix * F - F / 2 - 1;
  float wx[N], wy[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int ox = ox0 + k, oy = oy0 + k;
    wx[k] = (ox >= 0 && ox < W) ? d2d_axis_weight(ox, ix, w, sx) : 0.0f;
    wy[k] = (oy >= 0 && oy < H) ? d2d_axis_weight(oy, iy, h, sy) : 0.0f;
  }
  float acc = 0.0f;
#pragma unroll
  for (int r = 0; r < N; ++r) {
    if (wy[r] == 0.0f) continue;
    const int base = (oy0 + r) * W + ox0;
    float row = 0.0f;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      if (wx[k] == 0.0f) continue;
      float v = gd[base + k];
      if (!a.sql) { const float d = dep[base + k]; v *= nspan * d * d; }
      row += wx[k] * v;
    }
    acc += wy[r] * row;
  }
  return acc;
}

BBD_HD float d2d_backward_px(const bbd_d2d_args& a, int lvl, int b, int iy, int ix, float sy, float sx) {
  const int h = a.h[lvl], w = a.w[lvl], H = a.height, W = a.width;
  const float* gd = a.gdepth + ((size_t)lvl * a.batch + b) * H * W;
  const float* dep = a.depth + ((size_t)lvl * a.batch + b) * H * W;
  const float nspan = -a.disp_span;
  float acc;
  if (h == H && w == W) {  // same resolution: the interpolation is the identity
    const int o = iy * W + ix;
    acc = gd[o];
    if (!a.sql) acc *= nspan * dep[o] * dep[o];
  } else if (a.scratch && H == 2 * h && W == 2 * w) {
    acc = d2d_vpass<2>(a, lvl, b, iy, ix);
  } else if (a.scratch && H == 4 * h && W == 4 * w) {
    acc = d2d_vpass<4>(a, lvl, b, iy, ix);
  } else if (a.scratch && H == 8 * h && W == 8 * w) {
    acc = d2d_vpass<8>(a, lvl, b, iy, ix);
  } else {
    // src(o) = (o + 0.5)/f - 0.5 lies in (i-1, i+1) for o in [f*i - f/2, f*i + 3f/2 - 1]; one extra
    // output on each side covers the clamped borders and rounding
    const int fy = H / h, fx = W / w;
    int oy_lo = iy * fy - fy / 2 - 1, oy_hi = iy * fy + (3 * fy) / 2 + 1;
    int ox_lo = ix * fx - fx / 2 - 1, ox_hi = ix * fx + (3 * fx) / 2 + 1;
    if (oy_lo < 0) oy_lo = 0;
    if (ox_lo < 0) ox_lo = 0;
    if (oy_hi > H) oy_hi = H;
    if (ox_hi > W) ox_hi = W;
    acc = 0.0f;
    for (int oy = oy_lo; oy < oy_hi; ++oy) {
      const float wy = d2d_axis_weight(oy, iy, h, sy);
      if (wy == 0.0f) continue;
      float row = 0.0f;
      const float* g = gd + oy * W;
      const float* dp = dep + oy * W;
      for (int ox = ox_lo; ox < ox_hi; ++ox) {
        const float wx = d2d_axis_weight(ox, ix, w, sx);
        float v = g[ox];
        if (!a.sql) v *= nspan * dp[ox] * dp[ox];
        row += wx * v;
      }
      acc += wy * row;
    }
  }
  acc *= a.gscale[lvl];
  if (a.gsmooth[lvl]) {
    float gs = a.gsmooth[lvl][((size_t)b * h + iy) * w + ix];
    if (a.gsmooth_coef) {  // finish the deferred mean-normalisation of the smoothness gradient
      const float* c = a.gsmooth_coef + ((size_t)lvl * a.batch + b) * 2;
      gs = gs * c[0] - c[1];
    }
    acc += a.gsmooth_scale[lvl] * gs;
  }
  return acc;
}

// ---- single-launch backward (integer factors 1, 2, 4, 8; width a multiple of 4) -----------------------------
// One block owns one low-resolution row (level, sample, iy).  Its threads first walk the 2F full-resolution rows
// that have iy as a tap, four columns each (16-byte loads of gdepth and depth, chain rule of disp_to_depth applied
// on the fly), and park the vertical tent sums of the whole row in shared memory; then one thread per
// low-resolution column adds its 2F neighbours with the horizontal tent weights and finishes the element.
// No scratch plane, no second launch: every full-resolution element is read by the two blocks whose windows
// contain it.  The two helpers below are that block's two phases (tests/emu runs them in a plain loop).
BBD_HD float d2d_epilogue(const bbd_d2d_args& a, int lvl, int b, int iy, int ix, float acc) {
  const int h = a.h[lvl], w = a.w[lvl];
  acc *= a.gscale[lvl];
  if (a.gsmooth[lvl]) {
    float gs = a.gsmooth[lvl][((size_t)b * h + iy) * w + ix];
    if (a.gsmooth_coef) {  // finish the deferred mean-normalisation of the smoothness gradient
      const float* c = a.gsmooth_coef + ((size_t)lvl * a.batch + b) * 2;
      gs = gs * c[0] - c[1];
    }
    acc += a.gsmooth_scale[lvl] * gs;
  }
  return acc;
}
template <int F>
BBD_HD void d2d_fused_col4(const bbd_d2d_args& a, int lvl, int b, int iy, int x4, float* v) {
  const int h = a.h[lvl], H = a.height, W = a.width;
  const size_t plane = ((size_t)lvl * a.batch + b) * H * W;
  const float nspan = -a.disp_span;
  v[0] = v[1] = v[2] = v[3] = 0.0f;
  constexpr int N = (F == 1) ? 1 : 2 * F;   // rows that have iy as a tap
  constexpr int CH = N < 4 ? N : 4;         // rows whose loads are issued together (branch-free: a row outside the
                                            // image is read at a clamped address and weighted 0)
  const int oy0 = (F == 1) ? iy : iy * F - F / 2;
#pragma unroll 1
  for (int k0 = 0; k0 < N; k0 += CH) {  // kept rolled: one batch of loads in flight at a time, bounded registers
    f4 g[CH], d[CH];
    float wy[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int oy = oy0 + k0 + k;
      const bool in = oy >= 0 && oy < H;
      const size_t o = plane + (size_t)(in ? oy : iy * F) * W + x4;
      wy[k] = in ? ((F == 1) ? 1.0f : d2d_tent<F>(k0 + k - F / 2, iy, h)) : 0.0f;
      g[k] = load4(a.gdepth + o);
      if (!a.sql) d[k] = load4(a.depth + o);
    }
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      float t[4] = {g[k].x, g[k].y, g[k].z, g[k].w};
      if (!a.sql) {
        t[0] *= nspan * d[k].x * d[k].x; t[1] *= nspan * d[k].y * d[k].y;
        t[2] *= nspan * d[k].z * d[k].z; t[3] *= nspan * d[k].w * d[k].w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += wy[k] * t[j];
    }
  }
}
template <int F>
BBD_HD float d2d_fused_out(const bbd_d2d_args& a, int lvl, int b, int iy, int ix, const float* srow) {
  const int w = a.w[lvl], W = a.width;
  float acc = 0.0f;
  if (F == 1) {
    acc = srow[ix];
  } else {
    const int ox0 = ix * F - F / 2;
#pragma unroll
    for (int k = 0; k < 2 * F; ++k) {
      const int ox = ox0 + k;
      if (ox < 0 || ox >= W) continue;
      acc += d2d_tent<F>(k - F / 2, ix, w) * srow[ox];
    }
  }
  return d2d_epilogue(a, lvl, b, iy, ix, acc);
}
// integer factor of a level for the single-launch path (1, 2, 4, 8), else 0
BBD_HD int d2d_fused_factor(const bbd_d2d_args& a, int lvl) {
  if (a.width % 4) return 0;
  if (a.h[lvl] == a.height && a.w[lvl] == a.width) return 1;
  return d2d_sep_factor(a, lvl);
}

// ---- transformation_from_parameters (layers.py:25-100) ---------------------------------------
// Value with three tangents (d/d axis-angle components); enough operators for the formulas.
struct Dual3 {
  float v, d[3];
};
BBD_HD Dual3 dconst(float v) { return Dual3{v, {0.0f, 0.0f, 0.0f}}; }
BBD_HD Dual3 dvar(float v, int i) { Dual3 r = dconst(v); r.d[i] = 1.0f; return r; }
BBD_HD Dual3 operator+(const Dual3& a, const Dual3& b) { return Dual3{add(a.v, b.v), {a.d[0] + b.d[0], a.d[1] + b.d[1], a.d[2] + b.d[2]}}; }
BBD_HD Dual3 operator-(const Dual3& a, const Dual3& b) { return Dual3{sub(a.v, b.v), {a.d[0] - b.d[0], a.d[1] - b.d[1], a.d[2] - b.d[2]}}; }
BBD_HD Dual3 operator*(const Dual3& a, const Dual3& b) {
  return Dual3{mul(a.v, b.v), {a.d[0] * b.v + a.v * b.d[0], a.d[1] * b.v + a.v * b.d[1], a.d[2] * b.v + a.v * b.d[2]}};
}
BBD_HD Dual3 operator/(const Dual3& a, const Dual3& b) {
  const float q = div_(a.v, b.v), ib = 1.0f / b.v;
  return Dual3{q, {(a.d[0] - q * b.d[0]) * ib, (a.d[1] - q * b.d[1]) * ib, (a.d[2] - q * b.d[2]) * ib}};
}

// Rotation block of rot_from_axisangle (layers.py:61-100) as duals; R[r][c]
BBD_HD void pose_rotation(const float* v, Dual3 R[3][3]) {
  const Dual3 x0 = dvar(v[0], 0), y0 = dvar(v[1], 1), z0 = dvar(v[2], 2);
  // angle = ||v||  (torch.norm: sqrt of the sum of squares)
  const float n2 = add(add(mul(v[0], v[0]), mul(v[1], v[1])), mul(v[2], v[2]));
  Dual3 angle;
  angle.v = sqrtf(n2);
  for (int i = 0; i < 3; ++i) angle.d[i] = angle.v > 0.0f ? v[i] / angle.v : 0.0f;
  const Dual3 den = angle + dconst(1e-7f);
  const Dual3 x = x0 / den, y = y0 / den, z = z0 / den;
  Dual3 ca, sa;
  ca.v = cosf(angle.v);
  sa.v = sinf(angle.v);
  for (int i = 0; i < 3; ++i) { ca.d[i] = -sa.v * angle.d[i]; sa.d[i] = ca.v * angle.d[i]; }
  const Dual3 C = dconst(1.0f) - ca;
  const Dual3 xs = x * sa, ys = y * sa, zs = z * sa;
  const Dual3 xC = x * C, yC = y * C, zC = z * C;
  const Dual3 xyC = x * yC, yzC = y * zC, zxC = z * xC;
  R[0][0] = x * xC + ca; R[0][1] = xyC - zs;    R[0][2] = zxC + ys;
  R[1][0] = xyC + zs;    R[1][1] = y * yC + ca; R[1][2] = yzC - xs;
  R[2][0] = zxC - ys;    R[2][1] = yzC + xs;    R[2][2] = z * zC + ca;
}

// forward: T (row-major 4x4).  Non-inverted: T @ R = [R | t]; inverted: R^T @ T(-t) = [R^T | R^T(-t)],
// the last column accumulated like the 4x4 matmul (k ascending, multiply then add).
BBD_HD void pose_forward_one(const float* aa, const float* tr, int invert, float* T) {
  Dual3 R[3][3];
  pose_rotation(aa, R);
  for (int i = 0; i < 16; ++i) T[i] = 0.0f;
  T[15] = 1.0f;
  if (!invert) {
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[r][c].v;
      T[r * 4 + 3] = tr[r];
    }
  } else {
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[c][r].v;
      float acc = mul(R[0][r].v, mul(tr[0], -1.0f));
      acc = add(acc, mul(R[1][r].v, mul(tr[1], -1.0f)));
      acc = add(acc, mul(R[2][r].v, mul(tr[2], -1.0f)));
      T[r * 4 + 3] = acc;
    }
  }
}

BBD_HD void pose_backward_one(const float* aa, const float* tr, int invert, const float* gT, float* gaa, float* gtr) {
  Dual3 R[3][3];
  pose_rotation(aa, R);
  float ga[3] = {0.0f, 0.0f, 0.0f}, gt[3] = {0.0f, 0.0f, 0.0f};
  if (!invert) {
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 3; ++i) ga[i] += gT[r * 4 + c] * R[r][c].d[i];
      gt[r] = gT[r * 4 + 3];
    }
  } else {
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 3; ++i) ga[i] += gT[r * 4 + c] * R[c][r].d[i];
      // T[r][3] = -sum_k R[k][r] * t[k]
      const float g = gT[r * 4 + 3];
      for (int k = 0; k < 3; ++k) {
        gt[k] -= g * R[k][r].v;
        for (int i = 0; i < 3; ++i) ga[i] -= g * tr[k] * R[k][r].d[i];
      }
    }
  }
  for (int i = 0; i < 3; ++i) { gaa[i] = ga[i]; gtr[i] = gt[i]; }
}

// ---- P = (K @ T)[:3] (layers.py:182), one output element.  For these tiny batched 4x4 products
// ATen's CPU bmm takes its naive loop (separately rounded multiply and add, k ascending) -- checked
// bit-for-bit; the large (3x4)@(4xHW) products go through the FMA chain used in project_pixel.
BBD_HD float pose_pack_elem(const float* K, const float* T, int i, int j) {
  float acc = mul(K[i * 4], T[j]);
  acc = add(acc, mul(K[i * 4 + 1], T[4 + j]));
  acc = add(acc, mul(K[i * 4 + 2], T[8 + j]));
  acc = add(acc, mul(K[i * 4 + 3], T[12 + j]));
  return acc;
}
// gT[k][j] = sum_{i<3} K[i][k] * gP[i][j]
BBD_HD float pose_pack_grad_elem(const float* K, const float* gP, int k, int j) {
  return K[k] * gP[j] + K[4 + k] * gP[4 + j] + K[8 + k] * gP[8 + j];
}

// ---- on-demand warp (trainer.py:434-442): one output pixel, three channels ----------------
BBD_HD void warp_px(int H, int W, const float* images, const float* depth, const float* inv_K, const float* P,
                    int n, int py, int px, float* warped, float* grid) {
  Cam cam;
  load_cam(cam, inv_K + (size_t)n * 16, P + (size_t)n * 12, W, H);
  Sample s;
  const size_t o = (size_t)py * W + px;
  project_pixel(cam, px, py, depth[(size_t)n * H * W + o], W, H, s);
  Taps tp;
  make_taps(s, W, H, tp);
  const float* src = images + (size_t)n * 3 * H * W;
  float* dst = warped + (size_t)n * 3 * H * W;
  for (int c = 0; c < 3; ++c) dst[(size_t)c * H * W + o] = tap_channel(src + (size_t)c * H * W, tp);
  if (grid) {
    grid[((size_t)n * 2) * H * W + o] = mul(sub(div_(s.ux, (float)(W - 1)), 0.5f), 2.0f);
    grid[((size_t)n * 2 + 1) * H * W + o] = mul(sub(div_(s.uy, (float)(H - 1)), 0.5f), 2.0f);
  }
}

// ---- F.grid_sample(bilinear, border, align_corners=True) as a standalone operator ------------
// coordinate part of ATen's grid_sampler for one axis: unnormalise + clip (+ gradient mask)
BBD_HD float gs_coord(float g, int size, float& mask) {
  const float lim = (float)(size - 1);
  float v = mul(mul(add(g, 1.0f), 0.5f), lim);
  if (!(v > 0.0f)) { v = 0.0f; mask = 0.0f; } else if (v >= lim) { v = lim; mask = 0.0f; } else { mask = 1.0f; }
  return v;
}
BBD_HD void gs_sample(int H, int W, float gx, float gy, Sample& s, Taps& t) {
  s.ix = gs_coord(gx, W, s.mx);
  s.iy = gs_coord(gy, H, s.my);
  s.x0 = (int)floorf(s.ix);
  s.y0 = (int)floorf(s.iy);
  make_taps(s, W, H, t);
}
BBD_HD void grid_sample_px(int C, int H, int W, int HoWo, const float* images, const float* grid, int n, int o, float* out) {
  Sample s; Taps t;
  gs_sample(H, W, grid[((size_t)n * 2) * HoWo + o], grid[((size_t)n * 2 + 1) * HoWo + o], s, t);
  for (int c = 0; c < C; ++c)
    out[((size_t)n * C + c) * HoWo + o] = tap_channel(images + ((size_t)n * C + c) * H * W, t);
}
BBD_HD void grid_sample_grad_px(int C, int H, int W, int HoWo, const float* images, const float* grid, const float* gout,
                                int n, int o, float* ggrid) {
  Sample s; Taps t;
  gs_sample(H, W, grid[((size_t)n * 2) * HoWo + o], grid[((size_t)n * 2 + 1) * HoWo + o], s, t);
  float gix = 0.0f, giy = 0.0f;
  for (int c = 0; c < C; ++c)
    tap_channel_grad(images + ((size_t)n * C + c) * H * W, s, t, gout[((size_t)n * C + c) * HoWo + o], gix, giy);
  // d ix / d gx = (W-1)/2 (align_corners), times the clip mask
  ggrid[((size_t)n * 2) * HoWo + o] = gix * s.mx * (0.5f * (float)(W - 1));
  ggrid[((size_t)n * 2 + 1) * HoWo + o] = giy * s.my * (0.5f * (float)(H - 1));
}

// ---- gradient w.r.t. the sampled image: the transpose of the bilinear gather, sorted by destination -------
// ATen scatters every output's upstream gradient onto its four taps with atomic adds (grid_sampler_2d_backward).
// Here the outputs are sorted by the linear index of their north-west tap (the caller sorts the keys below with a
// stable sort); a source pixel then finds the outputs that touch it in four contiguous runs of that order (those
// whose north-west tap is the pixel itself, its west, its north and its north-west neighbour) and adds them in
// run order -- a gather: no atomics, bit-reproducible.
BBD_HD int gs_dest_key(int H, int W, int HoWo, const float* grid, int n, int o) {
  Sample s; Taps t;
  gs_sample(H, W, grid[((size_t)n * 2) * HoWo + o], grid[((size_t)n * 2 + 1) * HoWo + o], s, t);
  return n * H * W + t.onw;
}
// seg_start[k] = first position of the sorted key array holding a key >= k, for k = 0 .. n_keys (position i marks
// the keys between its predecessor's and its own)
BBD_HD void gs_segment_mark(const int32_t* keys_sorted, int n_items, int n_keys, int i, int32_t* seg_start) {
  const int lo = (i == 0) ? 0 : keys_sorted[i - 1] + 1;
  const int hi = (i == n_items) ? n_keys : keys_sorted[i];
  for (int k = lo; k <= hi; ++k) seg_start[k] = i;
}
BBD_HD float gs_image_grad_px(int C, int H, int W, int HoWo, const float* grid, const float* gout, const int32_t* seg_start,
                              const int32_t* order, int n, int c, int p) {
  const int y = p / W, x = p - y * W;
  const int base = n * H * W;
  float acc = 0.0f;
  // role r: this pixel is the (NW, NE, SW, SE) tap of the outputs whose north-west tap is (p, p-1, p-W, p-W-1)
  for (int r = 0; r < 4; ++r) {
    const int dx = r & 1, dy = r >> 1;
    if ((dx && x == 0) || (dy && y == 0)) continue;
    const int key = base + p - dx - dy * W;
    for (int i = seg_start[key]; i < seg_start[key + 1]; ++i) {
      const int og = order[i];              // global output index n * HoWo + o
      const int o = og - n * HoWo;
      Sample s; Taps t;
      gs_sample(H, W, grid[((size_t)n * 2) * HoWo + o], grid[((size_t)n * 2 + 1) * HoWo + o], s, t);
      const float w = (r == 0) ? t.wnw : ((r == 1) ? t.wne : ((r == 2) ? t.wsw : t.wse));
      acc = fma_(gout[((size_t)n * C + c) * HoWo + o], w, acc);
    }
  }
  return acc;
}

// ---- BackprojectDepth (layers.py:160-167) --------------------------------------------------
BBD_HD void backproject_px(int HW, int W, const float* depth, const float* inv_K, int n, int i, float* points) {
  const float* ik = inv_K + (size_t)n * 16;
  const float x = (float)(i % W), y = (float)(i / W);
  const float d = depth[(size_t)n * HW + i];
  float* p = points + (size_t)n * 4 * HW + i;
  p[0] = mul(d, fma_(ik[2], 1.0f, fma_(ik[1], y, mul(ik[0], x))));
  p[HW] = mul(d, fma_(ik[6], 1.0f, fma_(ik[5], y, mul(ik[4], x))));
  p[2 * (size_t)HW] = mul(d, fma_(ik[10], 1.0f, fma_(ik[9], y, mul(ik[8], x))));
  p[3 * (size_t)HW] = 1.0f;
}
BBD_HD float backproject_grad_px(int HW, int W, const float* inv_K, const float* gpoints, int n, int i) {
  const float* ik = inv_K + (size_t)n * 16;
  const float x = (float)(i % W), y = (float)(i / W);
  const float* g = gpoints + (size_t)n * 4 * HW + i;
  const float rx = fma_(ik[2], 1.0f, fma_(ik[1], y, mul(ik[0], x)));
  const float ry = fma_(ik[6], 1.0f, fma_(ik[5], y, mul(ik[4], x)));
  const float rz = fma_(ik[10], 1.0f, fma_(ik[9], y, mul(ik[8], x)));
  return g[0] * rx + g[HW] * ry + g[2 * (size_t)HW] * rz;
}

// ---- Project3D (layers.py:181-195) ----------------------------------------------------------
BBD_HD void project_px(int H, int W, const float* points, const float* P, float eps, int n, int i, float* pix) {
  const int HW = H * W;
  const float* p = P + (size_t)n * 12;
  const float* q = points + (size_t)n * 4 * HW + i;
  const float X = q[0], Y = q[HW], Z = q[2 * (size_t)HW], Wc = q[3 * (size_t)HW];
  const float cx = fma_(p[3], Wc, fma_(p[2], Z, fma_(p[1], Y, mul(p[0], X))));
  const float cy = fma_(p[7], Wc, fma_(p[6], Z, fma_(p[5], Y, mul(p[4], X))));
  const float cz = fma_(p[11], Wc, fma_(p[10], Z, fma_(p[9], Y, mul(p[8], X))));
  const float zz = add(cz, eps);
  pix[((size_t)n * 2) * HW + i] = mul(sub(div_(div_(cx, zz), (float)(W - 1)), 0.5f), 2.0f);
  pix[((size_t)n * 2 + 1) * HW + i] = mul(sub(div_(div_(cy, zz), (float)(H - 1)), 0.5f), 2.0f);
}
// gradient wrt points (written) and wrt P (accumulated into gP[12] by the caller's thread)
BBD_HD void project_grad_px(int H, int W, const float* points, const float* P, float eps, const float* gpix, int n, int i,
                            float* gpoints, float gP[12]) {
  const int HW = H * W;
  const float* p = P + (size_t)n * 12;
  const float* q = points + (size_t)n * 4 * HW + i;
  const float X[4] = {q[0], q[HW], q[2 * (size_t)HW], q[3 * (size_t)HW]};
  float c[3];
  for (int r = 0; r < 3; ++r) c[r] = fma_(p[4 * r + 3], X[3], fma_(p[4 * r + 2], X[2], fma_(p[4 * r + 1], X[1], mul(p[4 * r], X[0]))));
  const float zz = add(c[2], eps);
  const float inv = 1.0f / zz;
  const float gux = gpix[((size_t)n * 2) * HW + i] * 2.0f / (float)(W - 1);
  const float guy = gpix[((size_t)n * 2 + 1) * HW + i] * 2.0f / (float)(H - 1);
  const float gc[3] = {gux * inv, guy * inv, -(gux * c[0] + guy * c[1]) * inv * inv};
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 4; ++k) gP[4 * r + k] += gc[r] * X[k];
  float* g = gpoints + (size_t)n * 4 * HW + i;
  for (int k = 0; k < 4; ++k) g[(size_t)k * HW] = p[k] * gc[0] + p[4 + k] * gc[1] + p[8 + k] * gc[2];
}

// ---- SSIM operator (layers.py:235-249), any channel count, reads global memory -------------
BBD_HD float plane_at(const float* p, int H, int W, int y, int x) { return p[(size_t)reflect1(y, H) * W + reflect1(x, W)]; }

BBD_HD void ssim_window(const float* x, const float* y, int H, int W, int py, int px, WinX& wx, WinY& wy) {
  float sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
  bool first = true;
  for (int dy = -1; dy <= 1; ++dy)
    for (int dx = -1; dx <= 1; ++dx) {
      const float a = plane_at(x, H, W, py + dy, px + dx), b = plane_at(y, H, W, py + dy, px + dx);
      if (first) { sx = a; sy = b; sxx = mul(a, a); syy = mul(b, b); sxy = mul(a, b); first = false; }
      else { sx = add(sx, a); sy = add(sy, b); sxx = add(sxx, mul(a, a)); syy = add(syy, mul(b, b)); sxy = add(sxy, mul(a, b)); }
    }
  wx.sx = sx; wx.sxx = sxx; wx.sxy = sxy;
  wy = target_stats(sy, syy);
}

BBD_HD float ssim_px(const float* x, const float* y, int H, int W, int py, int px) {
  WinX wx; WinY wy; SsimParts q;
  ssim_window(x, y, H, W, py, px, wx, wy);
  return ssim_channel(wx, wy, q);
}

// gradient at pixel (py,px) of plane x (and y): gather over the windows that contain it
BBD_HD void ssim_grad_px(const float* x, const float* y, const float* gout, int H, int W, int py, int px, float* gx, float* gy) {
  float ax = 0, bx = 0, cx = 0, ay = 0, by = 0, cy = 0;
  for (int dy = -1; dy <= 1; ++dy) {
    const int cyy = py + dy;
    if (cyy < 0 || cyy >= H) continue;
    const float my = ((py == 1 && dy == -1) || (py == H - 2 && dy == 1)) ? 2.0f : 1.0f;
    for (int dx = -1; dx <= 1; ++dx) {
      const int cxx = px + dx;
      if (cxx < 0 || cxx >= W) continue;
      const float m = my * (((px == 1 && dx == -1) || (px == W - 2 && dx == 1)) ? 2.0f : 1.0f);
      WinX wx; WinY wy; SsimParts q;
      ssim_window(x, y, H, W, cyy, cxx, wx, wy);
      ssim_channel(wx, wy, q);
      const float g = gout[(size_t)cyy * W + cxx] * m;
      float a, b, c;
      if (gx) { ssim_coefs(q, wy, g, a, b, c); ax += a; bx += b; cx += c; }
      if (gy) { ssim_coefs_y(q, wy, g, a, b, c); ay += a; by += b; cy += c; }
    }
  }
  const size_t o = (size_t)py * W + px;
  if (gx) gx[o] = ax + bx * x[o] + cx * y[o];
  if (gy) gy[o] = ay + by * y[o] + cy * x[o];
}

// ---- loss assembly (trainer.py:557-570), a handful of scalars: one thread ---------------------------
BBD_HD void loss_combine(int n, const float* reproj, const float* smooth, const float* weight, float num_scales,
                         float* per_scale, float* total) {
  float tot = 0.0f;
  for (int s = 0; s < n; ++s) {
    const float l = add(reproj[s], mul(weight[s], smooth[s]));
    per_scale[s] = l;
    tot = (s == 0) ? l : add(tot, l);
  }
  *total = div_(tot, num_scales);
}
BBD_HD void loss_combine_grad(int n, const float* g_total, const float* g_per_scale, const float* weight,
                              float num_scales, float* g_reproj, float* g_smooth) {
  const float gt = g_total ? div_(*g_total, num_scales) : 0.0f;
  for (int s = 0; s < n; ++s) {
    const float g = gt + (g_per_scale ? g_per_scale[s] : 0.0f);
    g_reproj[s] = g;
    g_smooth[s] = g * weight[s];
  }
}

// torchvision ToTensor on an 8-bit frame: float32(v) / 255, correctly rounded.
BBD_HD float u8_to_unit(uint8_t v) { return div_((float)v, 255.0f); }

}  // namespace bbd
