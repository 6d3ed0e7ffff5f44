// Fused reprojection loss, tile phases v2: lanes are columns.
//
// A block of NW warps owns a tile of 28 x TH target pixels.  The 32 lanes of a warp map to
// the 32 columns of the tile's +-2 halo region (28 interior + 4 halo), warps stride over rows.
// Every shared-memory plane has a row pitch of 32 floats, so a warp always touches 32
// consecutive words (conflict-free) and no division / modulo is needed to find a slot.
// Work is organised in phases separated by block barriers and the
// phase bodies are __host__ __device__ so tests/emu can step them on the CPU.
//
// Regions (padded image coordinates; reflection resolved when a value is produced):
//   R2: rows 0..TH+3, lanes 0..31   warped / target values        u = x0-2+lane, v = y0-2+row
//   R1: rows 0..TH+1, lanes 1..30   window centres                (R1 row i = R2 row i+1)
//   IN: rows 0..TH-1, lanes 2..29   pixels whose gradient we own  (IN row q = R2 row q+2)
#pragma once
#include "bbd_common.cuh"

#ifndef BBD_UNROLL_WARP
#define BBD_UNROLL_WARP 4
#endif
#ifndef BBD_UNROLL_STATS
#define BBD_UNROLL_STATS 1
#endif
#ifndef BBD_PACKED_STATS
#define BBD_PACKED_STATS 1
#endif
#ifndef BBD_PACKED_VERTICAL
#define BBD_PACKED_VERTICAL 1
#endif
#ifndef BBD_PACKED_PROJECT
#define BBD_PACKED_PROJECT 1
#endif
#ifndef BBD_PACKED_CHANNELS
#define BBD_PACKED_CHANNELS 1
#endif
#ifndef BBD_BWD_ROWS
#define BBD_BWD_ROWS 2  // rows a warp walks together in the backward (one exchange buffer each)
#endif
#define BBD_PRAGMA(x) _Pragma(#x)
#define BBD_UNROLL(n) BBD_PRAGMA(unroll n)

namespace bbd {

template <int TH_, int NW_>
struct StripCfg {
  static constexpr int TW = 28, TH = TH_, NW = NW_, NT = NW_ * 32;
  static constexpr int P = 32;  // row pitch (floats) of every shared plane
  static constexpr int R2H = TH + 4, R1H = TH + 2;
  static constexpr int R2N = R2H * P, R1N = R1H * P, INN = TH * P;
  static constexpr int RED_SEG = 16;
  static_assert(NT % RED_SEG == 0 && NT / RED_SEG <= RED_SEG * 2, "reduction shape");
};

template <class C>
struct StripSmem {
  float* tgt;    // [3][R2N]
  float* pred;   // [npred][3][R2N]  warped tiles: one per candidate (KEEP) or a single reused buffer
  float* tst;    // [6][R1N]  target window mean / variance term per channel
  float* stash;  // [9][R1N]  window moments of the best candidate -> gradient coefficients
  float* best;   // [R1N]
  int* bidx;     // [R1N]
  float* gd;     // [INN]
  float* red;    // [NW][12] warp partials of the block reduction (the CPU harness brings its own scratch)
  int* anywin;   // [BBD_MAX_REP]
  static constexpr size_t floats(int npred) {
    return 3 * C::R2N + (size_t)npred * 3 * C::R2N + 6 * C::R1N + 9 * C::R1N + 2 * C::R1N + C::INN + C::NW * 12 +
           BBD_MAX_REP;
  }
  BBD_HD void carve(float* base, int npred) {
    tgt = base; base += 3 * C::R2N;
    pred = base; base += (size_t)npred * 3 * C::R2N;
    tst = base; base += 6 * C::R1N;
    best = base; base += C::R1N;     // directly behind tst: together they host the backward's exchange buffers
    stash = base; base += 9 * C::R1N;
    bidx = reinterpret_cast<int*>(base); base += C::R1N;
    gd = base; base += C::INN;
    red = base; base += C::NW * 12;
    anywin = reinterpret_cast<int*>(base);
  }
};

// Per-thread constants of a tile.
struct StripCtx {
  int s, b, tile, ntiles;
  int x0, y0;
  int lane, warp;
  int u;        // padded column of this lane
  int px;       // reflected (real) column
  bool col_in;  // padded column lies inside the image (a window centre / owned pixel can live here)
};

template <class C>
BBD_HD StripCtx make_strip(int bx, int by, int bz, int tid, int batch, int H, int W) {
  StripCtx t;
  const int tiles_x = (W + C::TW - 1) / C::TW, tiles_y = (H + C::TH - 1) / C::TH;
  t.s = bz / batch;
  t.b = bz % batch;
  t.x0 = bx * C::TW;
  t.y0 = by * C::TH;
  t.tile = by * tiles_x + bx;
  t.ntiles = tiles_x * tiles_y;
  t.lane = tid & 31;
  t.warp = tid >> 5;
  t.u = t.x0 - 2 + t.lane;
  t.px = reflect1(t.u, W);
  t.col_in = t.u >= 0 && t.u < W;
  return t;
}

// 3x3 sums over a plane with pitch 32, top-left at p (row-major like ATen's avg_pool2d loop)
BBD_HD float w9(const float* p) {
  float s = p[0];
  s = add(s, p[1]); s = add(s, p[2]);
  s = add(s, p[32]); s = add(s, p[33]); s = add(s, p[34]);
  s = add(s, p[64]); s = add(s, p[65]); s = add(s, p[66]);
  return s;
}
BBD_HD float w9p(const float* p, const float* q) {
  float s = mul(p[0], q[0]);
  s = add(s, mul(p[1], q[1])); s = add(s, mul(p[2], q[2]));
  s = add(s, mul(p[32], q[32])); s = add(s, mul(p[33], q[33])); s = add(s, mul(p[34], q[34]));
  s = add(s, mul(p[64], q[64])); s = add(s, mul(p[65], q[65])); s = add(s, mul(p[66], q[66]));
  return s;
}

// the same sums for two window positions p0 / p1 (and q0 / q1) at once
BBD_HD f2 ld2(const float* p0, const float* p1, int o) { return mk2(p0[o], p1[o]); }
BBD_HD f2 w9_2(const float* p0, const float* p1) {
  f2 s = ld2(p0, p1, 0);
  s = add(s, ld2(p0, p1, 1)); s = add(s, ld2(p0, p1, 2));
  s = add(s, ld2(p0, p1, 32)); s = add(s, ld2(p0, p1, 33)); s = add(s, ld2(p0, p1, 34));
  s = add(s, ld2(p0, p1, 64)); s = add(s, ld2(p0, p1, 65)); s = add(s, ld2(p0, p1, 66));
  return s;
}
// sums of x*x and x*y over the window from one pass over the loads
BBD_HD void w9pp_2(const float* x0, const float* x1, const float* y0, const float* y1, f2& sxx, f2& sxy, f2& sx) {
  const int offs[9] = {0, 1, 2, 32, 33, 34, 64, 65, 66};
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const f2 x = ld2(x0, x1, offs[i]), y = ld2(y0, y1, offs[i]);
    if (i == 0) { sx = x; sxx = mul(x, x); sxy = mul(x, y); }
    else { sx = add(sx, x); sxx = add(sxx, mul(x, x)); sxy = add(sxy, mul(x, y)); }
  }
}

// One word global -> shared.  On the device this is an asynchronous copy (LDGSTS: no register
// staging, the thread does not wait for the data); rs_load_wait() must precede the barrier that
// publishes it.
BBD_HD void copy_word_async(float* dst, const float* src) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
  *dst = *src;
#endif
}
BBD_HD void rs_load_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// Target tile (3 channels) over R2, reflection resolved.  Asynchronous on the device: the kernel
// overlaps it with the first candidate's warp phase (which does not read the target) and calls
// rs_load_wait() before the barrier that precedes the first use of sm.tgt.  (Staging the depth
// tile the same way was measured: no gain, its loads are already hidden.)
// window sum and sum of squares of one plane at two positions (same order as w9 / w9p)
BBD_HD void w9sq_2(const float* p0, const float* p1, f2& s, f2& ss) {
  const int offs[9] = {0, 1, 2, 32, 33, 34, 64, 65, 66};
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const f2 y = ld2(p0, p1, offs[i]);
    if (i == 0) { s = y; ss = mul(y, y); }
    else { s = add(s, y); ss = add(ss, mul(y, y)); }
  }
}
// target_stats for two positions
BBD_HD void target_stats2(const f2& sy, const f2& syy, f2& mu, f2& sig) {
  mu = ninth(sy);
  sig = sub(ninth(syy), mul(mu, mu));
}

template <class C>
BBD_HD void rs_load_target(const bbd_reproj_args& a, StripSmem<C>& sm, const StripCtx& t, int tid) {
  const int H = a.height, W = a.width, HW = H * W;
  const float* img = a.target + (size_t)t.b * 3 * HW;
  for (int j = t.warp; j < C::R2H; j += C::NW) {
    const int o = reflect1(t.y0 - 2 + j, H) * W + t.px;
    const int i = j * C::P + t.lane;
    copy_word_async(sm.tgt + i, img + o);
    copy_word_async(sm.tgt + C::R2N + i, img + HW + o);
    copy_word_async(sm.tgt + 2 * C::R2N + i, img + 2 * HW + o);
  }
}

// per-scale reset of the accumulators
template <class C>
BBD_HD void rs_begin_scale(StripSmem<C>& sm, int tid) {
  if (tid < BBD_MAX_REP) sm.anywin[tid] = 0;
  for (int i = tid; i < C::INN; i += C::NT) sm.gd[i] = 0.0f;
}

// is R1 slot (row i, this lane) a window centre inside the image?
template <class C>
BBD_HD bool rs_center(const bbd_reproj_args& a, const StripCtx& t, int i, int& py) {
  py = t.y0 - 1 + i;
  return t.lane >= 1 && t.lane <= 30 && t.col_in && py >= 0 && py < a.height;
}

// Target window statistics, once per tile.  (Reading them precomputed from global memory was
// measured slower: the tile's first phase is latency-bound and this arithmetic overlaps it.)
template <class C>
BBD_HD void rs_target_stats(const bbd_reproj_args& a, StripSmem<C>& sm, const StripCtx& t) {
  if (a.no_ssim) return;
  // (packing these rows like rs_stats was measured: spills at 96 registers, kernel +10 us)
  for (int i = t.warp; i < C::R1H; i += C::NW) {
    int py;
    if (!rs_center<C>(a, t, i, py)) continue;
    const int o = i * C::P + t.lane - 1;  // top-left of the window in R2 coordinates
    const int j = i * C::P + t.lane;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* y = sm.tgt + c * C::R2N + o;
      const WinY w = target_stats(w9(y), w9p(y, y));
      sm.tst[(2 * c) * C::R1N + j] = w.mu;
      sm.tst[(2 * c + 1) * C::R1N + j] = w.sig;
    }
  }
}

BBD_HD void rs_candidate(const bbd_reproj_args& a, int b, int k, const float*& src, Cam& cam) {
  const int32_t* e = a.tab.rep + ((size_t)b * BBD_MAX_REP + k) * 4;
  src = a.frames[e[0]] + (size_t)e[1] * 3 * a.height * a.width;
  load_cam(cam, a.inv_K + (size_t)e[3] * 16, a.P + (size_t)e[2] * 12, a.width, a.height);
}

template <class C, bool KEEP>
BBD_HD void rs_warp(const bbd_reproj_args& a, StripSmem<C>& sm, const StripCtx& t, int k) {
  const int H = a.height, W = a.width, HW = H * W;
  const float* src;
  Cam cam;
  rs_candidate(a, t.b, k, src, cam);
  const float* depth = a.depth + ((size_t)t.s * a.batch + t.b) * HW;
  float* pred = sm.pred + (KEEP ? (size_t)k * 3 * C::R2N : 0);
  constexpr int ITERS = (C::R2H + C::NW - 1) / C::NW;
#if BBD_PACKED_PROJECT
  // rows in pairs: the projection arithmetic of two rows of the column runs packed
  BBD_UNROLL(BBD_UNROLL_WARP / 2)
  for (int m = 0; m < ITERS; m += 2) {
    const int j0 = t.warp + m * C::NW;
    if (C::R2H % C::NW != 0 && j0 >= C::R2H) break;
    const int j1raw = j0 + C::NW;
    const bool two = (m + 1 < ITERS) && j1raw < C::R2H;
    const int j1 = two ? j1raw : j0;
    const int py0 = reflect1(t.y0 - 2 + j0, H), py1 = reflect1(t.y0 - 2 + j1, H);
    Sample s[2];
    project_pixel2(cam, t.px, py0, py1, depth[py0 * W + t.px], depth[py1 * W + t.px], s[0], s[1]);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r == 1 && !two) break;
      Taps tp;
      make_taps(s[r], W, H, tp);
      const int i = (r ? j1 : j0) * C::P + t.lane;
      pred[i] = tap_channel(src, tp);
      pred[C::R2N + i] = tap_channel(src + HW, tp);
      pred[2 * C::R2N + i] = tap_channel(src + 2 * HW, tp);
    }
  }
#else
  BBD_UNROLL(BBD_UNROLL_WARP)
  for (int m = 0; m < ITERS; ++m) {
    const int j = t.warp + m * C::NW;
    if (C::R2H % C::NW != 0 && j >= C::R2H) break;
    const int py = reflect1(t.y0 - 2 + j, H);
    Sample s;
    project_pixel(cam, t.px, py, depth[py * W + t.px], W, H, s);
    Taps tp;
    make_taps(s, W, H, tp);
    const int i = j * C::P + t.lane;
    pred[i] = tap_channel(src, tp);
    pred[C::R2N + i] = tap_channel(src + HW, tp);
    pred[2 * C::R2N + i] = tap_channel(src + 2 * HW, tp);
  }
#endif
}

template <class C, bool KEEP>
BBD_HD void rs_stats(const bbd_reproj_args& a, StripSmem<C>& sm, const StripCtx& t, int k) {
  const float* pred = sm.pred + (KEEP ? (size_t)k * 3 * C::R2N : 0);
  constexpr int ITERS = (C::R1H + C::NW - 1) / C::NW;
  int m_begin = 0;
#if BBD_PACKED_STATS
  // rows i0 = warp and i1 = warp + NW as one packed pair (both exist: R1H >= 2 NW is asserted)
  if (!a.no_ssim) {
    static_assert(C::R1H >= 2 * C::NW, "packed statistics need two full row sets");
    m_begin = 2;
    const int i0 = t.warp, i1 = t.warp + C::NW;
    int py0, py1;
    const bool v0 = rs_center<C>(a, t, i0, py0), v1 = rs_center<C>(a, t, i1, py1);
    if (v0 || v1) {
      const int lane = (t.lane < 1) ? 1 : t.lane;  // keep the window inside the plane for idle lanes
      const int o0 = i0 * C::P + lane - 1, o1 = i1 * C::P + lane - 1;
      const int j0 = i0 * C::P + t.lane, j1 = i1 * C::P + t.lane;
      f2 ssim_sum = bc2(0.0f), l1_sum = bc2(0.0f);
      f2 wsx[3], wsxx[3], wsxy[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* x = pred + c * C::R2N;
        const float* y = sm.tgt + c * C::R2N;
        const f2 d = sub(ld2(y + o0, y + o1, C::P + 1), ld2(x + o0, x + o1, C::P + 1));
        const f2 l1 = mk2(fabsf(d.x), fabsf(d.y));
        l1_sum = (c == 0) ? l1 : add(l1_sum, l1);
        w9pp_2(x + o0, x + o1, y + o0, y + o1, wsxx[c], wsxy[c], wsx[c]);
        const f2 muy = mk2(sm.tst[(2 * c) * C::R1N + j0], sm.tst[(2 * c) * C::R1N + j1]);
        const f2 sigy = mk2(sm.tst[(2 * c + 1) * C::R1N + j0], sm.tst[(2 * c + 1) * C::R1N + j1]);
        f2 m0, m1, m2;  // the sums are dead from here on: keep the moments in their place
        const f2 v = ssim_channel2(wsx[c], wsxx[c], wsxy[c], muy, sigy, m0, m1, m2);
        wsx[c] = m0; wsxx[c] = m1; wsxy[c] = m2;
        ssim_sum = (c == 0) ? v : add(ssim_sum, v);
      }
      // 0.85 * mean_c(ssim) + 0.15 * mean_c(l1), packed (same order as photometric_mix)
      const f2 loss = add(mul(bc2(BBD_W_SSIM), mul(ssim_sum, bc2(BBD_THIRD))), mul(bc2(BBD_W_L1), mul(l1_sum, bc2(BBD_THIRD))));
      if (v0 && (k == 0 || loss.x < sm.best[j0] || loss.x != loss.x)) {
        sm.best[j0] = loss.x;
        sm.bidx[j0] = k;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          sm.stash[(3 * c) * C::R1N + j0] = wsx[c].x;
          sm.stash[(3 * c + 1) * C::R1N + j0] = wsxx[c].x;
          sm.stash[(3 * c + 2) * C::R1N + j0] = wsxy[c].x;
        }
      }
      if (v1 && (k == 0 || loss.y < sm.best[j1] || loss.y != loss.y)) {
        sm.best[j1] = loss.y;
        sm.bidx[j1] = k;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          sm.stash[(3 * c) * C::R1N + j1] = wsx[c].y;
          sm.stash[(3 * c + 1) * C::R1N + j1] = wsxx[c].y;
          sm.stash[(3 * c + 2) * C::R1N + j1] = wsxy[c].y;
        }
      }
    }
  }
#endif
  BBD_UNROLL(BBD_UNROLL_STATS)
  for (int m = m_begin; m < ITERS; ++m) {
    const int i = t.warp + m * C::NW;
    if (i >= C::R1H) break;
    int py;
    if (!rs_center<C>(a, t, i, py)) continue;
    const int o = i * C::P + t.lane - 1;
    const int ctr = o + C::P + 1;
    const int j = i * C::P + t.lane;
    float ssim_sum = 0.0f, l1_sum = 0.0f;
    WinX wx[3];
#if BBD_PACKED_STATS && BBD_PACKED_CHANNELS
    // a single row: channels 0 and 1 of the same pixel as one packed pair (identical rounding and
    // summation order), channel 2 scalar
    if (!a.no_ssim) {
      const float* x0 = pred;
      const float* x1 = pred + C::R2N;
      const float* y0 = sm.tgt;
      const float* y1 = sm.tgt + C::R2N;
      const f2 d = sub(ld2(y0, y1, ctr), ld2(x0, x1, ctr));
      l1_sum = add(fabsf(d.x), fabsf(d.y));
      f2 sx, sxx, sxy, m0, m1, m2;
      w9pp_2(x0 + o, x1 + o, y0 + o, y1 + o, sxx, sxy, sx);
      const f2 muy = mk2(sm.tst[j], sm.tst[2 * C::R1N + j]);
      const f2 sigy = mk2(sm.tst[C::R1N + j], sm.tst[3 * C::R1N + j]);
      const f2 v = ssim_channel2(sx, sxx, sxy, muy, sigy, m0, m1, m2);
      ssim_sum = add(v.x, v.y);
      wx[0].sx = m0.x; wx[0].sxx = m1.x; wx[0].sxy = m2.x;
      wx[1].sx = m0.y; wx[1].sxx = m1.y; wx[1].sxy = m2.y;
      {
        const float* x = pred + 2 * C::R2N;
        const float* y = sm.tgt + 2 * C::R2N;
        l1_sum = add(l1_sum, fabsf(sub(y[ctr], x[ctr])));
        wx[2].sx = w9(x + o);
        wx[2].sxx = w9p(x + o, x + o);
        wx[2].sxy = w9p(x + o, y + o);
        WinY wy;
        wy.mu = sm.tst[4 * C::R1N + j];
        wy.sig = sm.tst[5 * C::R1N + j];
        SsimParts q;
        const float v2 = ssim_channel(wx[2], wy, q);
        wx[2].sx = q.mux; wx[2].sxx = q.sigx; wx[2].sxy = q.sigxy;
        ssim_sum = add(ssim_sum, v2);
      }
    } else
#endif
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* x = pred + c * C::R2N;
      const float* y = sm.tgt + c * C::R2N;
      const float l1 = fabsf(sub(y[ctr], x[ctr]));
      l1_sum = (c == 0) ? l1 : add(l1_sum, l1);
      if (!a.no_ssim) {
        wx[c].sx = w9(x + o);
        wx[c].sxx = w9p(x + o, x + o);
        wx[c].sxy = w9p(x + o, y + o);
        WinY wy;
        wy.mu = sm.tst[(2 * c) * C::R1N + j];
        wy.sig = sm.tst[(2 * c + 1) * C::R1N + j];
        SsimParts q;
        const float v = ssim_channel(wx[c], wy, q);
        wx[c].sx = q.mux; wx[c].sxx = q.sigx; wx[c].sxy = q.sigxy;  // stash the moments, not the sums
        ssim_sum = (c == 0) ? v : add(ssim_sum, v);
      }
    }
    const float loss = photometric_mix(ssim_sum, l1_sum, a.no_ssim != 0);
    if (k == 0 || loss < sm.best[j] || loss != loss) {  // a NaN candidate wins, as in torch.min
      sm.best[j] = loss;
      sm.bidx[j] = k;
      if (!a.no_ssim) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          sm.stash[(3 * c) * C::R1N + j] = wx[c].sx;
          sm.stash[(3 * c + 1) * C::R1N + j] = wx[c].sxx;
          sm.stash[(3 * c + 2) * C::R1N + j] = wx[c].sxy;
        }
      }
    }
  }
}

// Winner against the identity minimum, loss partial, gradient coefficients of the winner.
template <class C>
BBD_HD float rs_select(const bbd_reproj_args& a, StripSmem<C>& sm, const StripCtx& t, int n_rep) {
  const int H = a.height, W = a.width;
  const float wgt = 1.0f / ((float)a.batch * (float)H * (float)W);
  const float g_ssim = wgt * BBD_W_SSIM * BBD_THIRD;
  float part = 0.0f;
  for (int i = t.warp; i < C::R1H; i += C::NW) {
    int py;
    const bool inside = rs_center<C>(a, t, i, py);
    const int j = i * C::P + t.lane;
    int win = -1;
    if (inside) {
      const size_t o = ((size_t)t.b * H + py) * W + t.px;
      const float idm = a.ident_min[o];
      const float bst = sm.best[j];
      const bool rep_wins = !(bst > idm);  // ties go to the lower index = the warped candidate; so does NaN
      if (rep_wins) win = sm.bidx[j];
      const bool interior = t.lane >= 2 && t.lane <= 29 && i >= 1 && i <= C::TH;
      if (interior) {
        part += (rep_wins && idm == idm) ? bst : idm;  // a NaN on either side reaches the mean (torch.min)
        if (a.winner)
          a.winner[(((size_t)t.s * a.batch + t.b) * H + py) * W + t.px] =
              (uint8_t)(rep_wins ? win : n_rep + (a.ident_arg ? a.ident_arg[o] : 0));
      }
    }
    if (a.need_grad) {
      if (win >= 0 && !a.no_ssim) {
        sm.anywin[win] = 1;
#if BBD_PACKED_STATS && BBD_PACKED_CHANNELS
        {  // channels 0 and 1 as one packed pair (the exact part keeps its rounding), channel 2 below
          const f2 mux = mk2(sm.stash[j], sm.stash[3 * C::R1N + j]);
          const f2 sigx = mk2(sm.stash[C::R1N + j], sm.stash[4 * C::R1N + j]);
          const f2 sigxy = mk2(sm.stash[2 * C::R1N + j], sm.stash[5 * C::R1N + j]);
          const f2 muy = mk2(sm.tst[j], sm.tst[2 * C::R1N + j]);
          const f2 sigy = mk2(sm.tst[C::R1N + j], sm.tst[3 * C::R1N + j]);
          SsimParts2 q2;
          ssim_from_moments2(mux, sigx, sigxy, muy, sigy, q2);
          f2 ca, cb, cc;
          ssim_coefs2(q2, muy, g_ssim, ca, cb, cc);
          sm.stash[j] = ca.x; sm.stash[C::R1N + j] = cb.x; sm.stash[2 * C::R1N + j] = cc.x;
          sm.stash[3 * C::R1N + j] = ca.y; sm.stash[4 * C::R1N + j] = cb.y; sm.stash[5 * C::R1N + j] = cc.y;
        }
        constexpr int C_BEGIN = 2;
#else
        constexpr int C_BEGIN = 0;
#endif
#pragma unroll
        for (int c = C_BEGIN; c < 3; ++c) {
          WinY wy;
          wy.mu = sm.tst[(2 * c) * C::R1N + j];
          wy.sig = sm.tst[(2 * c + 1) * C::R1N + j];
          SsimParts q;  // the stash holds the window moments of the winner (rs_stats)
          ssim_from_moments(sm.stash[(3 * c) * C::R1N + j], sm.stash[(3 * c + 1) * C::R1N + j],
                            sm.stash[(3 * c + 2) * C::R1N + j], wy, q);
          float ca, cb, cc;
          ssim_coefs(q, wy, g_ssim, ca, cb, cc);
          sm.stash[(3 * c) * C::R1N + j] = ca;
          sm.stash[(3 * c + 1) * C::R1N + j] = cb;
          sm.stash[(3 * c + 2) * C::R1N + j] = cc;
        }
      } else {
        if (win >= 0) sm.anywin[win] = 1;
#pragma unroll
        for (int c = 0; c < 9; ++c) sm.stash[c * C::R1N + j] = 0.0f;
      }
    }
    sm.bidx[j] = win;
  }
  return part;
}

// ---- separable form of the backward gather ------------------------------------------------
// The 3x3 masked sum of the coefficient planes is done as a vertical pass (every lane sums the
// three window centres of its own column) followed by a horizontal pass over the neighbouring
// lanes.  The exchange buffer is warp-private ([10][32] floats per warp, carved out of the target
// statistics planes, which are dead once the winners are selected), so the two passes are
// separated by a warp barrier only.
// does any active lane of the warp satisfy p?  (the CPU harness steps one thread at a time: p itself)
BBD_HD bool warp_any(bool p) {
#if defined(__CUDA_ARCH__)
  return __any_sync(__activemask(), p) != 0;
#else
  return p;
#endif
}

template <class C>
BBD_HD float* rs_xch(StripSmem<C>& sm, const StripCtx& t, int buf) { return sm.tst + (t.warp * BBD_BWD_ROWS + buf) * 320; }

template <class C>
BBD_HD void rs_bwd_vertical(const bbd_reproj_args& a, StripSmem<C>& sm, const StripCtx& t, int k, int q, int buf = 0) {
  static_assert(C::NW * BBD_BWD_ROWS * 320 <= 7 * C::R1N, "exchange buffers must fit the planes that are dead after the selection");
  const int py = t.y0 + q;
  if (py >= a.height || t.lane < 1 || t.lane > 30) return;
  float* x = rs_xch<C>(sm, t, buf);
  const float my0 = (py == 1) ? 2.0f : 1.0f, my2 = (py == a.height - 2) ? 2.0f : 1.0f;
  bool any = false;
#if BBD_PACKED_STATS && BBD_PACKED_VERTICAL
  // the nine coefficient planes as four packed pairs + one scalar
  f2 v2[4];
  float v8 = 0.0f;
#pragma unroll
  for (int e = 0; e < 4; ++e) v2[e] = bc2(0.0f);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int j = (q + dy) * C::P + t.lane;
    const bool mine = sm.bidx[j] == k;
    if (!warp_any(mine)) continue;
    any = any || mine;
    const float my = mine ? (dy == 0 ? my0 : (dy == 2 ? my2 : 1.0f)) : 0.0f;
    const f2 m2 = bc2(my);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      v2[e] = fma_(m2, mk2(sm.stash[(2 * e) * C::R1N + j], sm.stash[(2 * e + 1) * C::R1N + j]), v2[e]);
    v8 += my * sm.stash[8 * C::R1N + j];
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    x[(2 * e) * 32 + t.lane] = v2[e].x;
    x[(2 * e + 1) * 32 + t.lane] = v2[e].y;
  }
  x[8 * 32 + t.lane] = v8;
#else
  float v[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) v[e] = 0.0f;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int j = (q + dy) * C::P + t.lane;  // R1 row q+dy = window centre row q+1 + (dy-1)
    const bool mine = sm.bidx[j] == k;
    // skip the row only if no lane of the warp has a winner there (uniform branch); otherwise every
    // lane adds its coefficients times a 0/1 mask -- no divergence
    if (!warp_any(mine)) continue;
    any = any || mine;
    const float my = mine ? (dy == 0 ? my0 : (dy == 2 ? my2 : 1.0f)) : 0.0f;
#pragma unroll
    for (int e = 0; e < 9; ++e) v[e] += my * sm.stash[e * C::R1N + j];
  }
#pragma unroll
  for (int e = 0; e < 9; ++e) x[e * 32 + t.lane] = v[e];
#endif
  x[9 * 32 + t.lane] = any ? 1.0f : 0.0f;
}

template <class C, bool KEEP>
BBD_HD void rs_bwd_horizontal(const bbd_reproj_args& a, StripSmem<C>& sm, const StripCtx& t, int k, int q,
                              const float* src, const Cam& cam, float gP[12], int buf = 0) {
  const int H = a.height, W = a.width, HW = H * W;
  const int py = t.y0 + q;
  if (py >= H || t.lane < 2 || t.lane > 29 || t.u >= W) return;
  const float* x = rs_xch<C>(sm, t, buf);
  if (x[9 * 32 + t.lane - 1] + x[9 * 32 + t.lane] + x[9 * 32 + t.lane + 1] == 0.0f) return;
  const float wgt = 1.0f / ((float)a.batch * (float)H * (float)W);
  const float g_l1 = a.no_ssim ? wgt * BBD_THIRD : wgt * BBD_W_L1 * BBD_THIRD;
  // multiplicity of a neighbouring window: a reflected border pixel sits twice in it
  const float mx0 = (t.u == 1) ? 2.0f : 1.0f, mx2 = (t.u == W - 2) ? 2.0f : 1.0f;
  const int ctr2 = (q + 2) * C::P + t.lane;
  const bool own = sm.bidx[(q + 1) * C::P + t.lane] == k;
  // KEEP: every candidate's warped tile is still in shared memory.  Otherwise the warped value of
  // this pixel is recomputed from the four taps the gradient needs anyway (bit-identical to
  // rs_warp), so shared memory does not grow with the number of candidates (tri-min / decomp).
  const float* depth = a.depth + ((size_t)t.s * a.batch + t.b) * HW;
  Sample s;
  project_pixel(cam, t.px, py, depth[py * W + t.px], W, H, s);
  Taps tp;
  make_taps(s, W, H, tp);
  float gix = 0.0f, giy = 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* xa = x + (3 * c) * 32 + t.lane;
    const float sa = mx0 * xa[-1] + xa[0] + mx2 * xa[1];
    const float sb = mx0 * xa[31] + xa[32] + mx2 * xa[33];
    const float sc = mx0 * xa[63] + xa[64] + mx2 * xa[65];
    const float* plane = src + c * HW;
    const float xv = KEEP ? sm.pred[((size_t)k * 3 + c) * C::R2N + ctr2] : tap_channel(plane, tp);
    const float yv = sm.tgt[c * C::R2N + ctr2];
    float g = sa + sb * xv + sc * yv;
    if (own) {
      const float d = sub(yv, xv);  // l1 = |target - pred|; abs'(0) = 0
      g += (d > 0.0f) ? -g_l1 : ((d < 0.0f) ? g_l1 : 0.0f);
    }
    tap_channel_grad(plane, s, tp, g, gix, giy);
  }
  float gdep = 0.0f;
  chain_to_depth_pose(cam, s, gix, giy, gdep, gP);
  sm.gd[q * C::P + t.lane] += gdep;
}

template <class C>
BBD_HD void rs_store_gdepth(const bbd_reproj_args& a, StripSmem<C>& sm, const StripCtx& t) {
  const int H = a.height, W = a.width;
  float* out = a.gdepth + ((size_t)t.s * a.batch + t.b) * H * W;
  if (t.lane < 2 || t.lane > 29 || t.u >= W) return;
  for (int q = t.warp; q < C::TH; q += C::NW) {
    const int py = t.y0 + q;
    if (py < H) out[py * W + t.u] = sm.gd[q * C::P + t.lane];
  }
}

// ---------------------------------------------------------------------------------------
// Identity pre-pass on the same tile geometry: min over the sample's sources of
// (photometric(source, target) + noise), and the target window statistics for the main kernel.
// ---------------------------------------------------------------------------------------
template <class C>
struct IdentStripSmem {
  float* tgt;   // [3][R2N]
  float* src;   // [3][R2N]
  float* tst;   // [6][INN]
  float* best;  // [INN]
  int* arg;     // [INN]
  static constexpr size_t floats() { return 6 * C::R2N + 8 * C::INN; }
  BBD_HD void carve(float* base) {
    tgt = base; base += 3 * C::R2N;
    src = base; base += 3 * C::R2N;
    tst = base; base += 6 * C::INN;
    best = base; base += C::INN;
    arg = reinterpret_cast<int*>(base);
  }
};

template <class C>
BBD_HD void is_load(const float* img, float* dst, const StripCtx& t, int H, int W) {
  const int HW = H * W;
  for (int j = t.warp; j < C::R2H; j += C::NW) {
    const int o = reflect1(t.y0 - 2 + j, H) * W + t.px;
    const int i = j * C::P + t.lane;
    dst[i] = img[o];
    dst[C::R2N + i] = img[HW + o];
    dst[2 * C::R2N + i] = img[2 * HW + o];
  }
}

template <class C>
BBD_HD bool is_owned(const StripCtx& t, int q, int H, int W, int& py) {
  py = t.y0 + q;
  return t.lane >= 2 && t.lane <= 29 && t.u < W && py < H;
}

template <class C>
BBD_HD void is_target_stats(const bbd_ident_args& a, IdentStripSmem<C>& sm, const StripCtx& t) {
  if (a.no_ssim) return;
  const int H = a.height, W = a.width;
  int q_begin = t.warp;
#if BBD_PACKED_STATS
  {
    const int q0 = t.warp, q1 = t.warp + C::NW;
    q_begin = t.warp + 2 * C::NW;
    int py0, py1;
    const bool v0 = is_owned<C>(t, q0, H, W, py0), v1 = is_owned<C>(t, q1, H, W, py1);
    if (v0 || v1) {
      const int lane = (t.lane < 1) ? 1 : t.lane;
      const int o0 = (q0 + 1) * C::P + lane - 1, o1 = (q1 + 1) * C::P + lane - 1;
      const int j0 = q0 * C::P + t.lane, j1 = q1 * C::P + t.lane;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* y = sm.tgt + c * C::R2N;
        f2 sy, syy, mu, sig;
        w9sq_2(y + o0, y + o1, sy, syy);
        target_stats2(sy, syy, mu, sig);
        if (v0) { sm.tst[(2 * c) * C::INN + j0] = mu.x; sm.tst[(2 * c + 1) * C::INN + j0] = sig.x; }
        if (v1) { sm.tst[(2 * c) * C::INN + j1] = mu.y; sm.tst[(2 * c + 1) * C::INN + j1] = sig.y; }
      }
    }
  }
#endif
  for (int q = q_begin; q < C::TH; q += C::NW) {
    int py;
    if (!is_owned<C>(t, q, H, W, py)) continue;
    const int o = (q + 1) * C::P + t.lane - 1;  // window top-left in R2 coordinates
    const int j = q * C::P + t.lane;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* y = sm.tgt + c * C::R2N + o;
      const WinY w = target_stats(w9(y), w9p(y, y));
      sm.tst[(2 * c) * C::INN + j] = w.mu;
      sm.tst[(2 * c + 1) * C::INN + j] = w.sig;
    }
  }
}

template <class C>
BBD_HD void is_candidate(const bbd_ident_args& a, IdentStripSmem<C>& sm, const StripCtx& t, int jcand, const float* noise) {
  const int H = a.height, W = a.width;
  int q_begin = t.warp;
#if BBD_PACKED_STATS
  // rows q0 = warp and q1 = warp + NW as one packed pair (same operation order as the scalar path)
  if (!a.no_ssim) {
    static_assert(C::TH >= 2 * C::NW, "packed identity statistics need two full row sets");
    const int q0 = t.warp, q1 = t.warp + C::NW;
    q_begin = t.warp + 2 * C::NW;
    int py0, py1;
    const bool v0 = is_owned<C>(t, q0, H, W, py0), v1 = is_owned<C>(t, q1, H, W, py1);
    if (v0 || v1) {
      const int lane = (t.lane < 1) ? 1 : t.lane;  // keep the window inside the plane for idle lanes
      const int o0 = (q0 + 1) * C::P + lane - 1, o1 = (q1 + 1) * C::P + lane - 1;
      const int j0 = q0 * C::P + t.lane, j1 = q1 * C::P + t.lane;
      f2 ssim_sum = bc2(0.0f), l1_sum = bc2(0.0f);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* x = sm.src + c * C::R2N;
        const float* y = sm.tgt + c * C::R2N;
        const f2 d = sub(ld2(y + o0, y + o1, C::P + 1), ld2(x + o0, x + o1, C::P + 1));
        const f2 l1 = mk2(fabsf(d.x), fabsf(d.y));
        l1_sum = (c == 0) ? l1 : add(l1_sum, l1);
        f2 sx, sxx, sxy, m0, m1, m2;
        w9pp_2(x + o0, x + o1, y + o0, y + o1, sxx, sxy, sx);
        const f2 muy = mk2(sm.tst[(2 * c) * C::INN + j0], sm.tst[(2 * c) * C::INN + j1]);
        const f2 sigy = mk2(sm.tst[(2 * c + 1) * C::INN + j0], sm.tst[(2 * c + 1) * C::INN + j1]);
        const f2 v = ssim_channel2(sx, sxx, sxy, muy, sigy, m0, m1, m2);
        ssim_sum = (c == 0) ? v : add(ssim_sum, v);
      }
      const f2 loss = add(mul(bc2(BBD_W_SSIM), mul(ssim_sum, bc2(BBD_THIRD))), mul(bc2(BBD_W_L1), mul(l1_sum, bc2(BBD_THIRD))));
      if (v0) {
        const float val = add(loss.x, mul(noise[py0 * W + t.u], a.noise_scale));
        if (jcand == 0 || val < sm.best[j0] || val != val) { sm.best[j0] = val; sm.arg[j0] = jcand; }
      }
      if (v1) {
        const float val = add(loss.y, mul(noise[py1 * W + t.u], a.noise_scale));
        if (jcand == 0 || val < sm.best[j1] || val != val) { sm.best[j1] = val; sm.arg[j1] = jcand; }
      }
    }
  }
#endif
  for (int q = q_begin; q < C::TH; q += C::NW) {
    int py;
    if (!is_owned<C>(t, q, H, W, py)) continue;
    const int o = (q + 1) * C::P + t.lane - 1, ctr = o + C::P + 1;
    const int j = q * C::P + t.lane;
    float ssim_sum = 0.0f, l1_sum = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* x = sm.src + c * C::R2N;
      const float* y = sm.tgt + c * C::R2N;
      const float l1 = fabsf(sub(y[ctr], x[ctr]));
      l1_sum = (c == 0) ? l1 : add(l1_sum, l1);
      if (!a.no_ssim) {
        WinX wx;
        wx.sx = w9(x + o);
        wx.sxx = w9p(x + o, x + o);
        wx.sxy = w9p(x + o, y + o);
        WinY wy;
        wy.mu = sm.tst[(2 * c) * C::INN + j];
        wy.sig = sm.tst[(2 * c + 1) * C::INN + j];
        SsimParts parts;
        const float v = ssim_channel(wx, wy, parts);
        ssim_sum = (c == 0) ? v : add(ssim_sum, v);
      }
    }
    const float loss = photometric_mix(ssim_sum, l1_sum, a.no_ssim != 0);
    const float val = add(loss, mul(noise[py * W + t.u], a.noise_scale));
    if (jcand == 0 || val < sm.best[j] || val != val) {
      sm.best[j] = val;
      sm.arg[j] = jcand;
    }
  }
}

template <class C>
BBD_HD void is_store(const bbd_ident_args& a, IdentStripSmem<C>& sm, const StripCtx& t) {
  const int H = a.height, W = a.width;
  for (int q = t.warp; q < C::TH; q += C::NW) {
    int py;
    if (!is_owned<C>(t, q, H, W, py)) continue;
    const size_t o = ((size_t)t.b * H + py) * W + t.u;
    a.ident_min[o] = sm.best[q * C::P + t.lane];
    if (a.ident_arg) a.ident_arg[o] = (uint8_t)sm.arg[q * C::P + t.lane];
  }
}

// block reduction helpers (fixed order, no shuffles, no atomics)
template <class C, int K>
BBD_HD void rs_park(float* red, int tid, const float* v) {
#pragma unroll
  for (int i = 0; i < K; ++i) red[i * C::NT + tid] = v[i];
}
template <class C, int K>
BBD_HD void rs_level1(float* red, int tid) {
  constexpr int SEGS = C::NT / C::RED_SEG;
  if (tid < K * SEGS) {
    const int comp = tid / SEGS, seg = tid % SEGS;
    const float* src = red + comp * C::NT + seg * C::RED_SEG;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < C::RED_SEG; ++i) s += src[i];
    red[12 * C::NT + comp * C::RED_SEG * 2 + seg] = s;
  }
}
template <class C, int K>
BBD_HD void rs_level2(const float* red, int tid, float* out) {
  constexpr int SEGS = C::NT / C::RED_SEG;
  if (tid < K) {
    const float* src = red + 12 * C::NT + tid * C::RED_SEG * 2;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < SEGS; ++i) s += src[i];
    out[tid] = s;
  }
}

}  // namespace bbd
