// Tile phases of the fused reprojection loss and of the identity pre-pass.
//
// One thread block owns a TW x TH tile of target pixels of one (scale, sample).
// Work is organised as *phases*; inside a phase every thread strides over the
// slots of a region and touches only shared-memory cells it owns or cells that
// were completed in an earlier phase, and phases are separated by a block
// barrier.  The kernel (bbd_kernels.cu) calls phase(tid) / __syncthreads();
// the test-only CPU harness (tests/emu) calls `for tid: phase(tid)` -- the
// phase bodies are this one source.
//
// Regions of a tile (padded image coordinates, reflection resolved on load):
//   R2 = tile + 2 halo : warped/target pixel values needed by the windows
//   R1 = tile + 1 halo : window centres whose SSIM gradient reaches the tile
//   IN = tile          : pixels whose depth gradient this block owns
#pragma once
#include "bbd_common.cuh"

namespace bbd {

template <int TW_, int TH_, int NT_>
struct TileCfg {
  static constexpr int TW = TW_, TH = TH_, NT = NT_;
  static constexpr int R2W = TW + 4, R2H = TH + 4, R2N = R2W * R2H;
  static constexpr int R1W = TW + 2, R1H = TH + 2, R1N = R1W * R1H;
  static constexpr int INN = TW * TH;
  static constexpr int RED_SEG = 16;  // second-level width of the block reduction
  static_assert(NT % RED_SEG == 0 && NT / RED_SEG <= RED_SEG * 2, "reduction shape");
};

// Shared-memory carve-up (floats).  `pred` holds every candidate's warped tile so that the
// backward can revisit them after the winner is known.
template <class C>
struct ReprojSmem {
  float* tgt;    // [3][R2N]
  float* pred;   // [max_rep][3][R2N]
  float* tst;    // [6][R1N]   target mean / variance term per channel
  float* stash;  // [9][R1N]   window sums of the best candidate -> gradient coefficients
  float* best;   // [R1N]
  int* bidx;     // [R1N]      best candidate so far; after select: winner or -1
  float* gd;     // [INN]      depth gradient accumulator
  float* red;    // [12][NT] + [12][RED_SEG*2]
  int* anywin;   // [BBD_MAX_REP]
  static constexpr size_t floats(int max_rep) {
    return 3 * C::R2N + (size_t)max_rep * 3 * C::R2N + 6 * C::R1N + 9 * C::R1N + 2 * C::R1N + C::INN +
           12 * C::NT + 12 * C::RED_SEG * 2 + BBD_MAX_REP;
  }
  BBD_HD void carve(float* base, int max_rep) {
    tgt = base; base += 3 * C::R2N;
    pred = base; base += (size_t)max_rep * 3 * C::R2N;
    tst = base; base += 6 * C::R1N;
    stash = base; base += 9 * C::R1N;
    best = base; base += C::R1N;
    bidx = reinterpret_cast<int*>(base); base += C::R1N;
    gd = base; base += C::INN;
    red = base; base += 12 * C::NT + 12 * C::RED_SEG * 2;
    anywin = reinterpret_cast<int*>(base);
  }
};

struct TileId {
  int s, b;      // scale, target sample
  int x0, y0;    // image coordinates of the tile origin
  int tile;      // linear tile index inside the (scale, sample) plane
  int ntiles;
};

// Row-major 3x3 window sum, like ATen avg_pool2d's accumulation loop.
template <int STRIDE>
BBD_HD float win9(const float* p) {
  float s = p[0];
  s = add(s, p[1]); s = add(s, p[2]);
  s = add(s, p[STRIDE]); s = add(s, p[STRIDE + 1]); s = add(s, p[STRIDE + 2]);
  s = add(s, p[2 * STRIDE]); s = add(s, p[2 * STRIDE + 1]); s = add(s, p[2 * STRIDE + 2]);
  return s;
}
template <int STRIDE>
BBD_HD float win9_prod(const float* p, const float* q) {
  float s = mul(p[0], q[0]);
  s = add(s, mul(p[1], q[1])); s = add(s, mul(p[2], q[2]));
  s = add(s, mul(p[STRIDE], q[STRIDE])); s = add(s, mul(p[STRIDE + 1], q[STRIDE + 1]));
  s = add(s, mul(p[STRIDE + 2], q[STRIDE + 2]));
  s = add(s, mul(p[2 * STRIDE], q[2 * STRIDE])); s = add(s, mul(p[2 * STRIDE + 1], q[2 * STRIDE + 1]));
  s = add(s, mul(p[2 * STRIDE + 2], q[2 * STRIDE + 2]));
  return s;
}

// ---------------------------------------------------------------------------------------
// Block reduction of K per-thread values (K <= 12) in a fixed order: thread partials are
// parked in shared memory, then summed in two levels.  No shuffles, no atomics.
// ---------------------------------------------------------------------------------------
template <class C, int K>
BBD_HD void red_park(float* red, int tid, const float* v) {
  for (int i = 0; i < K; ++i) red[i * C::NT + tid] = v[i];
}
template <class C, int K>
BBD_HD void red_level1(float* red, int tid) {
  constexpr int SEGS = C::NT / C::RED_SEG;  // partial sums per component after level 1
  if (tid < K * SEGS) {
    const int comp = tid / SEGS, seg = tid % SEGS;
    const float* src = red + comp * C::NT + seg * C::RED_SEG;
    float s = 0.0f;
    for (int i = 0; i < C::RED_SEG; ++i) s += src[i];
    red[12 * C::NT + comp * C::RED_SEG * 2 + seg] = s;
  }
}
template <class C, int K>
BBD_HD void red_level2(const float* red, int tid, float* out) {
  constexpr int SEGS = C::NT / C::RED_SEG;
  if (tid < K) {
    const float* src = red + 12 * C::NT + tid * C::RED_SEG * 2;
    float s = 0.0f;
    for (int i = 0; i < SEGS; ++i) s += src[i];
    out[tid] = s;
  }
}

// ---------------------------------------------------------------------------------------
// Fused reprojection loss
// ---------------------------------------------------------------------------------------
template <class C>
BBD_HD void rp_load_target(const bbd_reproj_args& a, ReprojSmem<C>& sm, const TileId& t, int tid) {
  const int H = a.height, W = a.width;
  const float* img = a.target + (size_t)t.b * 3 * H * W;
  for (int i = tid; i < C::R2N; i += C::NT) {
    const int sx = i % C::R2W, sy = i / C::R2W;
    const int px = reflect1(t.x0 - 2 + sx, W), py = reflect1(t.y0 - 2 + sy, H);
    const size_t o = (size_t)py * W + px;
    sm.tgt[i] = img[o];
    sm.tgt[C::R2N + i] = img[(size_t)H * W + o];
    sm.tgt[2 * C::R2N + i] = img[2 * (size_t)H * W + o];
  }
  if (tid < BBD_MAX_REP) sm.anywin[tid] = 0;
  for (int i = tid; i < C::INN; i += C::NT) sm.gd[i] = 0.0f;
}

template <class C>
BBD_HD bool r1_center(const bbd_reproj_args& a, const TileId& t, int j, int& px, int& py) {
  px = t.x0 - 1 + j % C::R1W;
  py = t.y0 - 1 + j / C::R1W;
  return px >= 0 && px < a.width && py >= 0 && py < a.height;
}

template <class C>
BBD_HD void rp_target_stats(const bbd_reproj_args& a, ReprojSmem<C>& sm, const TileId& t, int tid) {
  if (a.no_ssim) return;
  for (int j = tid; j < C::R1N; j += C::NT) {
    int px, py;
    if (!r1_center<C>(a, t, j, px, py)) continue;
    const int o = (j / C::R1W) * C::R2W + (j % C::R1W);  // R2 slot of the window's top-left
    for (int c = 0; c < 3; ++c) {
      const float* y = sm.tgt + c * C::R2N + o;
      const WinY w = target_stats(win9<C::R2W>(y), win9_prod<C::R2W>(y, y));
      sm.tst[(2 * c) * C::R1N + j] = w.mu;
      sm.tst[(2 * c + 1) * C::R1N + j] = w.sig;
    }
  }
}

BBD_HD void rp_candidate(const bbd_reproj_args& a, int b, int k, const float*& src, Cam& cam) {
  const int32_t* e = a.tab.rep + ((size_t)b * BBD_MAX_REP + k) * 4;
  src = a.frames[e[0]] + (size_t)e[1] * 3 * a.height * a.width;
  load_cam(cam, a.inv_K + (size_t)e[3] * 16, a.P + (size_t)e[2] * 12, a.width, a.height);
}

template <class C>
BBD_HD void rp_warp(const bbd_reproj_args& a, ReprojSmem<C>& sm, const TileId& t, int k, int tid) {
  const int H = a.height, W = a.width;
  const float* src;
  Cam cam;
  rp_candidate(a, t.b, k, src, cam);
  const float* depth = a.depth + ((size_t)t.s * a.batch + t.b) * H * W;
  float* pred = sm.pred + (size_t)k * 3 * C::R2N;
  for (int i = tid; i < C::R2N; i += C::NT) {
    const int sx = i % C::R2W, sy = i / C::R2W;
    const int px = reflect1(t.x0 - 2 + sx, W), py = reflect1(t.y0 - 2 + sy, H);
    Sample s;
    project_pixel(cam, px, py, depth[(size_t)py * W + px], W, H, s);
    Taps tp;
    make_taps(s, W, H, tp);
    pred[i] = tap_channel(src, tp);
    pred[C::R2N + i] = tap_channel(src + (size_t)H * W, tp);
    pred[2 * C::R2N + i] = tap_channel(src + 2 * (size_t)H * W, tp);
  }
}

template <class C>
BBD_HD void rp_stats(const bbd_reproj_args& a, ReprojSmem<C>& sm, const TileId& t, int k, int tid) {
  const float* pred = sm.pred + (size_t)k * 3 * C::R2N;
  for (int j = tid; j < C::R1N; j += C::NT) {
    int px, py;
    if (!r1_center<C>(a, t, j, px, py)) continue;
    const int o = (j / C::R1W) * C::R2W + (j % C::R1W);
    const int ctr = o + C::R2W + 1;
    float ssim_sum = 0.0f, l1_sum = 0.0f;
    WinX wx[3];
    for (int c = 0; c < 3; ++c) {
      const float* x = pred + c * C::R2N;
      const float* y = sm.tgt + c * C::R2N;
      const float l1 = fabsf(sub(y[ctr], x[ctr]));
      l1_sum = (c == 0) ? l1 : add(l1_sum, l1);
      if (!a.no_ssim) {
        wx[c].sx = win9<C::R2W>(x + o);
        wx[c].sxx = win9_prod<C::R2W>(x + o, x + o);
        wx[c].sxy = win9_prod<C::R2W>(x + o, y + o);
        WinY wy;
        wy.mu = sm.tst[(2 * c) * C::R1N + j];
        wy.sig = sm.tst[(2 * c + 1) * C::R1N + j];
        SsimParts q;
        const float v = ssim_channel(wx[c], wy, q);
        ssim_sum = (c == 0) ? v : add(ssim_sum, v);
      }
    }
    const float loss = photometric_mix(ssim_sum, l1_sum, a.no_ssim != 0);
    if (k == 0 || loss < sm.best[j]) {
      sm.best[j] = loss;
      sm.bidx[j] = k;
      if (!a.no_ssim) {
        for (int c = 0; c < 3; ++c) {
          sm.stash[(3 * c) * C::R1N + j] = wx[c].sx;
          sm.stash[(3 * c + 1) * C::R1N + j] = wx[c].sxx;
          sm.stash[(3 * c + 2) * C::R1N + j] = wx[c].sxy;
        }
      }
    }
  }
}

// Winner against the identity minimum; loss partial; gradient coefficients of the winner.
// Returns this thread's partial sum of to_optimise over the tile pixels it visited.
template <class C>
BBD_HD float rp_select(const bbd_reproj_args& a, ReprojSmem<C>& sm, const TileId& t, int n_rep, int tid) {
  const int H = a.height, W = a.width;
  const float wgt = 1.0f / ((float)a.batch * (float)H * (float)W);
  const float g_ssim = wgt * BBD_W_SSIM * BBD_THIRD;
  float part = 0.0f;
  for (int j = tid; j < C::R1N; j += C::NT) {
    int px, py;
    const bool inside = r1_center<C>(a, t, j, px, py);
    int win = -1;
    if (inside) {
      const size_t o = ((size_t)t.b * H + py) * W + px;
      const float idm = a.ident_min[o];
      const float bst = sm.best[j];
      const bool rep_wins = bst <= idm;  // ties go to the lower index = the warped candidate
      if (rep_wins) win = sm.bidx[j];
      const int jx = j % C::R1W, jy = j / C::R1W;
      const bool interior = jx >= 1 && jx <= C::TW && jy >= 1 && jy <= C::TH;
      if (interior) {
        part += rep_wins ? bst : idm;
        if (a.winner)
          a.winner[(((size_t)t.s * a.batch + t.b) * H + py) * W + px] =
              (uint8_t)(rep_wins ? win : n_rep + (a.ident_arg ? a.ident_arg[o] : 0));
      }
    }
    if (a.need_grad) {
      if (win >= 0 && !a.no_ssim) {
        sm.anywin[win] = 1;
        for (int c = 0; c < 3; ++c) {
          WinX wx;
          wx.sx = sm.stash[(3 * c) * C::R1N + j];
          wx.sxx = sm.stash[(3 * c + 1) * C::R1N + j];
          wx.sxy = sm.stash[(3 * c + 2) * C::R1N + j];
          WinY wy;
          wy.mu = sm.tst[(2 * c) * C::R1N + j];
          wy.sig = sm.tst[(2 * c + 1) * C::R1N + j];
          SsimParts q;
          ssim_channel(wx, wy, q);
          float ca, cb, cc;
          ssim_coefs(q, wy, g_ssim, ca, cb, cc);
          sm.stash[(3 * c) * C::R1N + j] = ca;
          sm.stash[(3 * c + 1) * C::R1N + j] = cb;
          sm.stash[(3 * c + 2) * C::R1N + j] = cc;
        }
      } else {
        if (win >= 0) sm.anywin[win] = 1;
        for (int c = 0; c < 9; ++c) sm.stash[c * C::R1N + j] = 0.0f;
      }
    }
    sm.bidx[j] = win;
  }
  return part;
}

// Backward of candidate k over the tile: gather the SSIM coefficients of the windows this
// candidate won, add the L1 term, push through the bilinear taps into depth and pose.
template <class C>
BBD_HD void rp_backward(const bbd_reproj_args& a, ReprojSmem<C>& sm, const TileId& t, int k, int tid, float gP[12]) {
  const int H = a.height, W = a.width;
  const float wgt = 1.0f / ((float)a.batch * (float)H * (float)W);
  const float g_l1 = a.no_ssim ? wgt * BBD_THIRD : wgt * BBD_W_L1 * BBD_THIRD;
  const float* src;
  Cam cam;
  rp_candidate(a, t.b, k, src, cam);
  const float* depth = a.depth + ((size_t)t.s * a.batch + t.b) * H * W;
  const float* pred = sm.pred + (size_t)k * 3 * C::R2N;
  for (int i = 0; i < 12; ++i) gP[i] = 0.0f;
  for (int q = tid; q < C::INN; q += C::NT) {
    const int qx = q % C::TW, qy = q / C::TW;
    const int px = t.x0 + qx, py = t.y0 + qy;
    if (px >= W || py >= H) continue;
    // multiplicity of a neighbouring window: a reflected border pixel sits twice in it
    float mxw[3] = {1.0f, 1.0f, 1.0f}, myw[3] = {1.0f, 1.0f, 1.0f};
    if (px == 1) mxw[0] = 2.0f;
    if (px == W - 2) mxw[2] = 2.0f;
    if (py == 1) myw[0] = 2.0f;
    if (py == H - 2) myw[2] = 2.0f;
    float sa[3] = {0, 0, 0}, sb[3] = {0, 0, 0}, sc[3] = {0, 0, 0};
    bool any = false;
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) {
        const int j = (qy + dy) * C::R1W + (qx + dx);
        if (sm.bidx[j] != k) continue;
        any = true;
        const float m = mxw[dx] * myw[dy];
        for (int c = 0; c < 3; ++c) {
          sa[c] += m * sm.stash[(3 * c) * C::R1N + j];
          sb[c] += m * sm.stash[(3 * c + 1) * C::R1N + j];
          sc[c] += m * sm.stash[(3 * c + 2) * C::R1N + j];
        }
      }
    if (!any) continue;
    const int ctr2 = (qy + 2) * C::R2W + (qx + 2);
    const bool own = sm.bidx[(qy + 1) * C::R1W + (qx + 1)] == k;
    float gpred[3];
    for (int c = 0; c < 3; ++c) {
      const float x = pred[c * C::R2N + ctr2], y = sm.tgt[c * C::R2N + ctr2];
      float g = sa[c] + sb[c] * x + sc[c] * y;
      if (own) {
        const float d = sub(y, x);  // l1 = |target - pred|; abs'(0) = 0
        g += (d > 0.0f) ? -g_l1 : ((d < 0.0f) ? g_l1 : 0.0f);
      }
      gpred[c] = g;
    }
    Sample s;
    project_pixel(cam, px, py, depth[(size_t)py * W + px], W, H, s);
    Taps tp;
    make_taps(s, W, H, tp);
    float gix = 0.0f, giy = 0.0f;
    for (int c = 0; c < 3; ++c) tap_channel_grad(src + (size_t)c * H * W, s, tp, gpred[c], gix, giy);
    float gdep = 0.0f;
    chain_to_depth_pose(cam, s, gix, giy, gdep, gP);
    sm.gd[q] += gdep;
  }
}

template <class C>
BBD_HD void rp_store_gdepth(const bbd_reproj_args& a, ReprojSmem<C>& sm, const TileId& t, int tid) {
  const int H = a.height, W = a.width;
  float* out = a.gdepth + ((size_t)t.s * a.batch + t.b) * H * W;
  for (int q = tid; q < C::INN; q += C::NT) {
    const int px = t.x0 + q % C::TW, py = t.y0 + q / C::TW;
    if (px < W && py < H) out[(size_t)py * W + px] = sm.gd[q];
  }
}

// ---------------------------------------------------------------------------------------
// Identity pre-pass: min over the sample's sources of (photometric(source, target) + noise)
// ---------------------------------------------------------------------------------------
template <class C>
struct IdentSmem {
  float* tgt;   // [3][R1N]
  float* src;   // [3][R1N]
  float* tst;   // [6][INN]
  float* best;  // [INN]
  int* arg;     // [INN]
  static constexpr size_t floats() { return 6 * C::R1N + 6 * C::INN + 2 * C::INN; }
  BBD_HD void carve(float* base) {
    tgt = base; base += 3 * C::R1N;
    src = base; base += 3 * C::R1N;
    tst = base; base += 6 * C::INN;
    best = base; base += C::INN;
    arg = reinterpret_cast<int*>(base);
  }
};

template <class C>
BBD_HD void id_load(const bbd_ident_args& a, const float* img, float* dst, const TileId& t, int tid) {
  const int H = a.height, W = a.width;
  for (int i = tid; i < C::R1N; i += C::NT) {
    const int px = reflect1(t.x0 - 1 + i % C::R1W, W), py = reflect1(t.y0 - 1 + i / C::R1W, H);
    const size_t o = (size_t)py * W + px;
    dst[i] = img[o];
    dst[C::R1N + i] = img[(size_t)H * W + o];
    dst[2 * C::R1N + i] = img[2 * (size_t)H * W + o];
  }
}

template <class C>
BBD_HD void id_target_stats(const bbd_ident_args& a, IdentSmem<C>& sm, int tid) {
  if (a.no_ssim) return;
  for (int q = tid; q < C::INN; q += C::NT) {
    const int o = (q / C::TW) * C::R1W + (q % C::TW);
    for (int c = 0; c < 3; ++c) {
      const float* y = sm.tgt + c * C::R1N + o;
      const WinY w = target_stats(win9<C::R1W>(y), win9_prod<C::R1W>(y, y));
      sm.tst[(2 * c) * C::INN + q] = w.mu;
      sm.tst[(2 * c + 1) * C::INN + q] = w.sig;
    }
  }
}

template <class C>
BBD_HD void id_candidate(const bbd_ident_args& a, IdentSmem<C>& sm, const TileId& t, int jcand, const float* noise, int tid) {
  const int H = a.height, W = a.width;
  for (int q = tid; q < C::INN; q += C::NT) {
    const int qx = q % C::TW, qy = q / C::TW;
    const int px = t.x0 + qx, py = t.y0 + qy;
    if (px >= W || py >= H) continue;
    const int o = qy * C::R1W + qx, ctr = o + C::R1W + 1;
    float ssim_sum = 0.0f, l1_sum = 0.0f;
    for (int c = 0; c < 3; ++c) {
      const float* x = sm.src + c * C::R1N;
      const float* y = sm.tgt + c * C::R1N;
      const float l1 = fabsf(sub(y[ctr], x[ctr]));
      l1_sum = (c == 0) ? l1 : add(l1_sum, l1);
      if (!a.no_ssim) {
        WinX wx;
        wx.sx = win9<C::R1W>(x + o);
        wx.sxx = win9_prod<C::R1W>(x + o, x + o);
        wx.sxy = win9_prod<C::R1W>(x + o, y + o);
        WinY wy;
        wy.mu = sm.tst[(2 * c) * C::INN + q];
        wy.sig = sm.tst[(2 * c + 1) * C::INN + q];
        SsimParts parts;
        const float v = ssim_channel(wx, wy, parts);
        ssim_sum = (c == 0) ? v : add(ssim_sum, v);
      }
    }
    const float loss = photometric_mix(ssim_sum, l1_sum, a.no_ssim != 0);
    const float val = add(loss, mul(noise[(size_t)py * W + px], a.noise_scale));
    if (jcand == 0 || val < sm.best[q]) {
      sm.best[q] = val;
      sm.arg[q] = jcand;
    }
  }
}

template <class C>
BBD_HD void id_store(const bbd_ident_args& a, IdentSmem<C>& sm, const TileId& t, int tid) {
  const int H = a.height, W = a.width;
  for (int q = tid; q < C::INN; q += C::NT) {
    const int px = t.x0 + q % C::TW, py = t.y0 + q / C::TW;
    if (px >= W || py >= H) continue;
    const size_t o = ((size_t)t.b * H + py) * W + px;
    a.ident_min[o] = sm.best[q];
    if (a.ident_arg) a.ident_arg[o] = (uint8_t)sm.arg[q];
  }
}

// ---------------------------------------------------------------------------------------
// Whole-block drivers.  SYNC is a functor: __syncthreads() on the device; on the host the
// emulator runs the phases for all threads in turn, so it is a no-op there and the loops
// over `tid` live in the emulator (see tests/emu/bbd_emu.cpp).
// ---------------------------------------------------------------------------------------
BBD_HD TileId make_tile(int bx, int by, int bz, int batch, int height, int width, int TW, int TH) {
  TileId t;
  const int tiles_x = (width + TW - 1) / TW, tiles_y = (height + TH - 1) / TH;
  t.s = bz / batch;
  t.b = bz % batch;
  t.x0 = bx * TW;
  t.y0 = by * TH;
  t.tile = by * tiles_x + bx;
  t.ntiles = tiles_x * tiles_y;
  return t;
}

}  // namespace bbd
