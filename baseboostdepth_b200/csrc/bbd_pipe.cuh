// Fused reprojection loss, pipelined form (round 2): the row program of bbd_stream.cuh split over three
// cooperating warps of one block, so that each warp's register state fits 128 registers (15 resident warps
// per SM instead of 8) and the irregular gathers are decoupled from the regular stencil arithmetic.
//
//   warp G (gather)   : back-project + project a row (bit-exact chain, stream_chain), gather the 4 taps of every
//                       candidate (LDG.128 from the channel-interleaved source copy), bilinear value and tap
//                       gradients, the Jacobian pieces of the backward -> ring X in shared memory.  Runs one row
//                       ahead of its own arithmetic: the taps of row i+1 are in flight while row i is interpolated.
//   warp S (statistics): 3x3 window sums (horizontal neighbours read from ring X, vertical sums slide in
//                       registers), SSIM + L1 mix, per-pixel minimum against the identity plane, loss sum,
//                       winner plane, SSIM gradient coefficients of the winner -> ring CO.
//   warp B (backward) : 3x3 gather of the coefficient rows (horizontal from ring CO, vertical sliding), chain
//                       through the parked tap gradients to depth (one store per pixel) and to the 12 entries of
//                       P (register accumulators, fixed-order warp reduction at the end of the segment).
//
// The regular planes (target, depth, identity minimum) arrive through the TMA unit into a ring of row boxes
// that all three warps read.  Hand-over is by mbarriers (full / empty per ring slot); the only release point
// of a row is warp B (warp S when no gradient is wanted), which is by construction the last reader of the
// row's ring X slot and of its TMA box.
//
// Arithmetic is that of stream_unit ("fast statistics" contract, see bbd_stream.cuh); the partial-sum
// layout, the unit order and the fused finalize are identical, so the two forms are interchangeable per launch.
#pragma once
#include "bbd_stream.cuh"

#ifndef BBD_PIPE_DX
#define BBD_PIPE_DX 5  // rows of ring X (gather -> statistics, backward)
#endif
#ifndef BBD_PIPE_DC
#define BBD_PIPE_DC 3  // rows of ring CO (statistics -> backward)
#endif
#define BBD_PIPE_LA 2  // TMA look-ahead in rows

namespace bbd {

// ---- mbarrier (shared memory, 8 bytes = two floats) ---------------------------------------------------
BBD_HD void mb_init(float* bar, int count) {
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
#else
  uint32_t* w = reinterpret_cast<uint32_t*>(bar);
  w[0] = 0u;                                    // phase
  w[1] = ((uint32_t)count << 16) | (uint32_t)count;  // count | pending
#endif
}
BBD_HD void mb_init_fence() {
#if defined(__CUDA_ARCH__)
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
BBD_HD void mb_arrive(float* bar) {
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
#else
  uint32_t* w = reinterpret_cast<uint32_t*>(bar);
  uint32_t pending = (w[1] & 0xffffu) - 1u;
  if (pending == 0u) { w[0] ^= 1u; pending = w[1] >> 16; }
  w[1] = (w[1] & 0xffff0000u) | pending;
#endif
}
// returns once the phase of parity `parity` has completed
BBD_HD void mb_wait(float* bar, unsigned parity) {
#if defined(__CUDA_ARCH__)
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  for (int spin = 0; spin < (1 << 26) && !done; ++spin)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  if (!done) __trap();  // a hand-over that never happens must not hang the device
#elif defined(BBD_EMU)
  const volatile uint32_t* w = reinterpret_cast<const volatile uint32_t*>(bar);
  while ((w[0] & 1u) == (parity & 1u)) simt::yield();
#else
  (void)bar; (void)parity;
#endif
}
BBD_HD void block_sync() {
#if defined(__CUDA_ARCH__)
  __syncthreads();
#elif defined(BBD_EMU)
  simt::syncthreads();
#endif
}

template <int K, bool GRAD>
struct PipeSmem {
  static constexpr int DX = BBD_PIPE_DX, DC = BBD_PIPE_DC, LA = BBD_PIPE_LA, DT = DX + LA;
  static constexpr int TROW = 256, TBOX = 36, TDEP = 128, TIDM = 192;  // one TMA ring row (see StreamSmem)
  static constexpr int OFF_T = 0;
  static constexpr int NBAR = DT + 2 * DX + 2 * DC + 1;
  static constexpr int OFF_BAR = OFF_T + DT * TROW;
  static constexpr int OFF_CST = OFF_BAR + ((2 * NBAR + 31) / 32) * 32;
  static constexpr int CST = 32 * K;
  // ring X row: x[3] | gx[3] gy[3] jx jy ax ay ux uy (GRAD)
  static constexpr int XE = GRAD ? 15 : 3;
  static constexpr int XSLOT = XE * 32 * K;
  static constexpr int OFF_X = OFF_CST + CST;
  // ring CO row: 9 coefficients (masked by the winner) | winner of the lane (one word)
  static constexpr int COSLOT = GRAD ? (9 * K + 1) * 32 : 0;
  static constexpr int OFF_CO = OFF_X + DX * XSLOT;
  static constexpr int FLOATS = OFF_CO + DC * COSLOT;
  // barrier order inside OFF_BAR (two floats each)
  static constexpr int B_TFULL = 0, B_XFULL = DT, B_XEMPTY = DT + DX, B_CFULL = DT + 2 * DX, B_CEMPTY = DT + 2 * DX + DC, B_DONE = DT + 2 * DX + 2 * DC;
};

// What every lane of every warp of a unit knows about its place.
struct PipeCtx {
  int s, b, sb, x0, y0, y1, u, px, n_rows, unit_in_sb, upb, n_rep_raw;
  bool col_in, centre_lane, own_lane;
};
BBD_HD PipeCtx pipe_ctx(const bbd_reproj_args& a, int unit, int lane, int seg_rows) {
  typedef StreamGeo Geo;
  PipeCtx c;
  const int H = a.height, W = a.width;
  const int nstrips = Geo::strips(W), nsegs = (H + seg_rows - 1) / seg_rows;
  c.upb = nstrips * nsegs;
  c.s = unit % a.num_scales;  // scale-minor unit order, as stream_unit
  const int rest = unit / a.num_scales;
  c.b = rest / c.upb;
  const int rem = rest - c.b * c.upb;
  c.sb = c.s * a.batch + c.b;
  const int seg = rem / nstrips, strip = rem - seg * nstrips;
  c.unit_in_sb = rem;
  c.x0 = strip * Geo::TW;
  c.y0 = seg * seg_rows;
  c.y1 = (c.y0 + seg_rows < H) ? c.y0 + seg_rows : H;
  c.n_rows = c.y1 - c.y0 + 4;  // rows y0-2 .. y1+1
  c.u = c.x0 - 2 + lane;
  c.px = reflect1(c.u, W);
  c.col_in = c.u >= 0 && c.u < W;
  c.centre_lane = lane >= 1 && lane <= 30 && c.col_in;
  c.own_lane = lane >= 2 && lane <= 29 && c.col_in;
  c.n_rep_raw = a.tab.hdr[(size_t)c.b * 4];
  return c;
}
// box column of the (possibly reflected) pixel that lane `l` of the strip stands for
BBD_HD int pipe_box_col(int x0, int l, int W) {
  const int u = x0 - 2 + l;
  int li = l + 2 + reflect1(u, W) - u;
  return li < 0 ? 0 : (li > 35 ? 35 : li);
}

BBD_HD void pipe_tma_issue(const StreamTmaMaps& tm, const bbd_reproj_args& a, float* slot, float* bar, int x, int y, int s, int b) {
  tma_row_issue(tm, a, slot, bar, x, y, s, b);
#if !defined(__CUDA_ARCH__)
  mb_arrive(bar);  // emulation: the copy is synchronous
#endif
}

// ---------------------------------------------------------------------------------------------------
// warp G
// ---------------------------------------------------------------------------------------------------
template <int K, bool GRAD>
struct PipeGather {
  typedef typename SVec<K>::V V;
  typedef PipeSmem<K, GRAD> SM;
  const bbd_reproj_args& a;
  const PipeCtx& c;
  const StreamTmaMaps& tm;
  float* smem;
  int lane, li;
  const float* src[K];
  float xf, wm1, hm1, rw, rh;

  // row i: wait for its ring slot and its planes, project, start the tap loads, park the Jacobian pieces
  BBD_HD void project(int i, f4* taps, V& ex, V& ey) {
    const int H = a.height, W = a.width;
    float* bars = smem + SM::OFF_BAR;
    const int xs = i % SM::DX;
    if (i >= SM::DX) mb_wait(bars + 2 * (SM::B_XEMPTY + xs), (unsigned)(i / SM::DX - 1) & 1u);
    if (i + SM::LA < c.n_rows && elect_one(lane)) {
      const int ts = (i + SM::LA) % SM::DT;
      pipe_tma_issue(tm, a, smem + SM::OFF_T + ts * SM::TROW, bars + 2 * (SM::B_TFULL + ts), c.x0 - 4, reflect1(c.y0 - 2 + i + SM::LA, H), c.s, c.b);
    }
    const int ts = i % SM::DT;
    mb_wait(bars + 2 * (SM::B_TFULL + ts), (unsigned)(i / SM::DT) & 1u);
    const float depth = smem[SM::OFF_T + ts * SM::TROW + SM::TDEP + li];
    const int py = reflect1(c.y0 - 2 + i, H);
    V P[12], ray[3], ux, uy, rz, ixr, iyr;
    stream_chain<V>(smem + SM::OFF_CST, xf, (float)py, depth, wm1, hm1, rw, rh, P, ray, ux, uy, rz, ixr, iyr);
    V mx, my;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float ix = vget(ixr, k), iy = vget(iyr, k);
      vset(mx, k, (ix > 0.0f && ix < wm1) ? 1.0f : 0.0f);  // clip_coordinates_set_grad: the border counts as outside
      vset(my, k, (iy > 0.0f && iy < hm1) ? 1.0f : 0.0f);
      ix = fminf(fmaxf(ix, 0.0f), wm1);
      iy = fminf(fmaxf(iy, 0.0f), hm1);
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      vset(ex, k, ix - fx0);
      vset(ey, k, iy - fy0);
      const int xi = (int)fx0, yi = (int)fy0;
      const int dx = (xi + 1 < W) ? 4 : 0;  // absent taps have weight 0: read the present one again
      const int dy = (yi + 1 < H) ? 4 * W : 0;
      const float* p = src[k] + (size_t)(yi * W + xi) * 4;
      taps[0 * K + k] = load4(p);
      taps[1 * K + k] = load4(p + dx);
      taps[2 * K + k] = load4(p + dy);
      taps[3 * K + k] = load4(p + dy + dx);
    }
    if (GRAD) {
      V q[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) q[j] = fma_(P[4 * j + 2], ray[2], fma_(P[4 * j + 1], ray[1], mul(P[4 * j], ray[0])));
      const V ax = mul(mx, rz), ay = mul(my, rz);
      const V jx = mul(ax, fma_(vneg(ux), q[2], q[0])), jy = mul(ay, fma_(vneg(uy), q[2], q[1]));
      float* row = smem + SM::OFF_X + xs * SM::XSLOT + lane * K;
      sts_v(row + 9 * 32 * K, jx); sts_v(row + 10 * 32 * K, jy);
      sts_v(row + 11 * 32 * K, ax); sts_v(row + 12 * 32 * K, ay);
      sts_v(row + 13 * 32 * K, ux); sts_v(row + 14 * 32 * K, uy);
    }
  }

  // row i: bilinear value and tap gradients from the landed taps -> ring X, hand the row over
  BBD_HD void finish(int i, const f4* taps, const V& ex, const V& ey) {
    float* bars = smem + SM::OFF_BAR;
    const int xs = i % SM::DX;
    float* row = smem + SM::OFF_X + xs * SM::XSLOT + lane * K;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      V vnw, vne, vsw, vse;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        vset(vnw, k, f4c(taps[0 * K + k], ch)); vset(vne, k, f4c(taps[1 * K + k], ch));
        vset(vsw, k, f4c(taps[2 * K + k], ch)); vset(vse, k, f4c(taps[3 * K + k], ch));
      }
      const V dtop = sub(vne, vnw), dbot = sub(vse, vsw);
      const V top = fma_(ex, dtop, vnw), bot = fma_(ex, dbot, vsw);
      const V gy = sub(bot, top);
      sts_v(row + ch * 32 * K, fma_(ey, gy, top));
      if (GRAD) {
        sts_v(row + (3 + ch) * 32 * K, fma_(ey, sub(dbot, dtop), dtop));
        sts_v(row + (6 + ch) * 32 * K, gy);
      }
    }
    warp_sync();
    if (lane == 0) mb_arrive(bars + 2 * (SM::B_XFULL + xs));
  }
};

template <int K, bool GRAD>
BBD_HD void pipe_gather(const bbd_reproj_args& a, const PipeCtx& c, int lane, float* smem, const StreamTmaMaps& tm) {
  typedef typename SVec<K>::V V;
  typedef PipeSmem<K, GRAD> SM;
  const int H = a.height, W = a.width, HW = H * W;
  PipeGather<K, GRAD> g = {a, c, tm, smem, lane, pipe_box_col(c.x0, lane, W)};
  float* cst = smem + SM::OFF_CST;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    // a missing second candidate repeats the first one: it ties, never wins, gets no gradient
    const int kk = (k < c.n_rep_raw) ? k : 0;
    const int32_t* e = a.tab.rep + ((size_t)c.b * BBD_MAX_REP + kk) * 4;
    g.src[k] = a.frames_rgba[e[0]] + (size_t)e[1] * HW * 4;
    const float* Pk = a.P + (size_t)e[2] * 12;
    const float* iK = a.inv_K + (size_t)e[3] * 16;
    if (lane < 12) cst[lane * K + k] = Pk[lane];
    if (lane >= 12 && lane < 21) {
      const int i = lane - 12;
      cst[lane * K + k] = iK[(i / 3) * 4 + (i % 3)];
    }
  }
  warp_sync();
  g.xf = (float)c.px;
  g.wm1 = (float)(W - 1);
  g.hm1 = (float)(H - 1);
  g.rw = div_(1.0f, g.wm1);
  g.rh = div_(1.0f, g.hm1);
  if (elect_one(lane)) {
    float* bars = smem + SM::OFF_BAR;
    for (int q = 0; q < SM::LA && q < c.n_rows; ++q)
      pipe_tma_issue(tm, a, smem + SM::OFF_T + q * SM::TROW, bars + 2 * (SM::B_TFULL + q), c.x0 - 4, reflect1(c.y0 - 2 + q, H), c.s, c.b);
  }
  // two tap register sets: the loads of row i+1 fly while row i is interpolated
  f4 t0[4 * K], t1[4 * K];
  V ex0, ey0, ex1, ey1;
  g.project(0, t0, ex0, ey0);
  for (int i = 0; i < c.n_rows; i += 2) {
    const bool has1 = i + 1 < c.n_rows, has2 = i + 2 < c.n_rows;  // warp-uniform
    if (has1) g.project(i + 1, t1, ex1, ey1);
    g.finish(i, t0, ex0, ey0);
    if (has1) {
      if (has2) g.project(i + 2, t0, ex0, ey0);
      g.finish(i + 1, t1, ex1, ey1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// warp S
// ---------------------------------------------------------------------------------------------------
template <int K, bool GRAD>
BBD_HD void pipe_stats(const bbd_reproj_args& a, const PipeCtx& c, int lane, float* smem, int part_stride) {
  typedef typename SVec<K>::V V;
  typedef PipeSmem<K, GRAD> SM;
  const int H = a.height, W = a.width, HW = H * W;
  float* bars = smem + SM::OFF_BAR;
  const int n_rep = c.n_rep_raw < K ? c.n_rep_raw : K;
  const float wgt = 1.0f / ((float)a.batch * (float)H * (float)W);
  const bool no_ssim = a.no_ssim != 0;
  const float g_ssim = wgt * BBD_W_SSIM * BBD_THIRD;
  const float w_ssim = BBD_W_SSIM * BBD_THIRD, w_l1 = no_ssim ? BBD_THIRD : BBD_W_L1 * BBD_THIRD;
  const float ninth = 0.111111111938953399658203125f;
  const int ll = lane > 0 ? lane - 1 : 0, lr = lane < 31 ? lane + 1 : 31;  // the edge lanes stand in for themselves
  const int li = pipe_box_col(c.x0, lane, W), lil = pipe_box_col(c.x0, ll, W), lir = pipe_box_col(c.x0, lr, W);

  Slide<V> sx[3], sxx[3], sxy[3];
  Slide<float> st[3], stt[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) { sx[ch].reset(); sxx[ch].reset(); sxy[ch].reset(); st[ch].reset(); stt[ch].reset(); }
  V l1_prev = vbc<V>(0.0f);
  float loss_acc = 0.0f;

  for (int i = 0; i < c.n_rows; ++i) {
    const int xs = i % SM::DX, ts = i % SM::DT;
    mb_wait(bars + 2 * (SM::B_XFULL + xs), (unsigned)(i / SM::DX) & 1u);
    mb_wait(bars + 2 * (SM::B_TFULL + ts), (unsigned)(i / SM::DT) & 1u);
    const float* xrow = smem + SM::OFF_X + xs * SM::XSLOT;
    const float* trow = smem + SM::OFF_T + ts * SM::TROW;
    V vx[3], vxx[3], vxy[3];
    float vt[3], vtt[3];
    V l1v = vbc<V>(0.0f);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const V x = lds_v<V>(xrow + (ch * 32 + lane) * K), xl = lds_v<V>(xrow + (ch * 32 + ll) * K), xr = lds_v<V>(xrow + (ch * 32 + lr) * K);
      const float t = trow[ch * SM::TBOX + li], tl = trow[ch * SM::TBOX + lil], tr = trow[ch * SM::TBOX + lir];
      l1v = add(l1v, vabs(sub(vbc<V>(t), x)));
      vx[ch] = sx[ch].push(add(add(xl, x), xr));
      vxx[ch] = sxx[ch].push(fma_(xr, xr, fma_(xl, xl, mul(x, x))));
      vxy[ch] = sxy[ch].push(fma_(xr, vbc<V>(tr), fma_(xl, vbc<V>(tl), mul(x, vbc<V>(t)))));
      vt[ch] = st[ch].push(add(add(tl, t), tr));
      vtt[ch] = stt[ch].push(fma_(tr, tr, fma_(tl, tl, mul(t, t))));
    }
    if (i >= 2) {  // centre row rb = row i-1: its three window rows have been pushed
      const int rb = c.y0 - 3 + i;
      const bool centre = c.centre_lane && rb >= 0 && rb < H;
      const bool own_b = c.own_lane && rb >= c.y0 && rb < c.y1;
      V lossv;
      V co[9];
      if (!no_ssim) {
        V ssum = vbc<V>(0.0f);
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float muy = mul(vt[ch], ninth);
          const float sigy = fma_(-muy, muy, mul(vtt[ch], ninth));
          const float cy1 = fma_(muy, muy, BBD_C1), cy2 = add(sigy, BBD_C2);
          const V mux = mul(vx[ch], vbc<V>(ninth));
          const V sigx = fma_(vneg(mux), mux, mul(vxx[ch], vbc<V>(ninth)));
          const V sigxy = fma_(vneg(mux), vbc<V>(muy), mul(vxy[ch], vbc<V>(ninth)));
          const V n1 = fma_(mux, vbc<V>(2.0f * muy), vbc<V>(BBD_C1));
          const V n2 = fma_(vbc<V>(2.0f), sigxy, vbc<V>(BBD_C2));
          const V d1 = fma_(mux, mux, vbc<V>(cy1));
          const V d2 = add(sigx, vbc<V>(cy2));
          const V rd = vrcp_raw(mul(d1, d2));
          const V rr = mul(mul(n1, n2), rd);
          const V raw = fma_(rr, vbc<V>(-0.5f), vbc<V>(0.5f));
          ssum = add(ssum, vsat(raw));
          if (GRAD) {
            // d value / d x(q) = ca + cb * x(q) + cc * y(q) for every pixel q of the window; torch.clamp passes the
            // gradient on [0, 1] only: raw = (1 - rr) / 2 is inside exactly when |rr| <= 1 (the affine map is exact at +-1)
            V wc = mul(rd, vbc<V>(g_ssim * (-1.0f / 9.0f)));
#pragma unroll
            for (int k = 0; k < K; ++k)
              if (!(fabsf(vget(rr, k)) <= 1.0f)) vset(wc, k, 0.0f);
            const V rwc = mul(rr, wc);
            co[3 * ch + 2] = mul(wc, n1);
            co[3 * ch + 1] = vneg(mul(rwc, d1));
            co[3 * ch] = fma_(mul(wc, vbc<V>(muy)), sub(n2, n1), vneg(mul(mul(rwc, mux), sub(d2, d1))));
          }
        }
        lossv = fma_(ssum, vbc<V>(w_ssim), mul(l1_prev, vbc<V>(w_l1)));
      } else {
        lossv = mul(l1_prev, vbc<V>(w_l1));
#pragma unroll
        for (int j = 0; j < 9; ++j) co[j] = vbc<V>(0.0f);
      }
      float best = vget(lossv, 0);
      int kbest = 0;
#pragma unroll
      for (int k = 1; k < K; ++k) {
        const float lk = vget(lossv, k);
        if (lk < best || lk != lk) { best = lk; kbest = k; }  // a NaN candidate wins, as in torch.min
      }
      int win = -1;
      if (centre) {
        const size_t o = (size_t)rb * W + c.u;
        const float idm = smem[SM::OFF_T + ((i - 1) % SM::DT) * SM::TROW + SM::TIDM + li];  // centre lanes: px == u
        const bool rep_wins = (n_rep > 0) && !(best > idm);  // ties and NaN go to the warped candidate
        if (rep_wins) win = kbest;
        if (own_b) {
          loss_acc += (rep_wins && idm == idm) ? best : idm;  // a NaN on either side reaches the mean
          if (a.winner)
            a.winner[(size_t)c.sb * HW + o] = (uint8_t)(rep_wins ? kbest : c.n_rep_raw + (a.ident_arg ? a.ident_arg[(size_t)c.b * HW + o] : 0));
        }
      }
      if (GRAD) {
        const int n = i - 2;  // sequence number of this coefficient row
        const int cs = n % SM::DC;
        if (n >= SM::DC) mb_wait(bars + 2 * (SM::B_CEMPTY + cs), (unsigned)(n / SM::DC - 1) & 1u);
        float* crow = smem + SM::OFF_CO + cs * SM::COSLOT;
        V sel_k;
#pragma unroll
        for (int k = 0; k < K; ++k) vset(sel_k, k, (win == k) ? 1.0f : 0.0f);
#pragma unroll
        for (int j = 0; j < 9; ++j) sts_v(crow + (j * 32 + lane) * K, mul(co[j], sel_k));
        reinterpret_cast<int*>(crow)[9 * 32 * K + lane] = win;
        warp_sync();
        if (lane == 0) mb_arrive(bars + 2 * (SM::B_CFULL + cs));
      }
    }
    if (!GRAD && i >= 1) {  // rows up to i-1 are of no further use to anybody
      warp_sync();
      if (lane == 0) mb_arrive(bars + 2 * (SM::B_XEMPTY + (i - 1) % SM::DX));
    }
    l1_prev = l1v;
  }
  {
    float v = loss_acc;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += lane_xor(v, m);
    if (lane == 0) {
      a.loss_part[(size_t)c.sb * part_stride + c.unit_in_sb] = v;
      if (GRAD) mb_arrive(bars + 2 * SM::B_DONE);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// warp B
// ---------------------------------------------------------------------------------------------------
template <int K>
BBD_HD void pipe_backward(const bbd_reproj_args& a, const PipeCtx& c, int lane, float* smem, int part_stride) {
  typedef typename SVec<K>::V V;
  typedef PipeSmem<K, true> SM;
  const int H = a.height, W = a.width, HW = H * W;
  float* bars = smem + SM::OFF_BAR;
  const float* cst = smem + SM::OFF_CST;
  const int n_rep = c.n_rep_raw < K ? c.n_rep_raw : K;
  const float wgt = 1.0f / ((float)a.batch * (float)H * (float)W);
  const bool no_ssim = a.no_ssim != 0;
  const float g_l1 = no_ssim ? wgt * BBD_THIRD : wgt * BBD_W_L1 * BBD_THIRD;
  // a reflected border pixel sits twice in the window of its inner neighbour
  const float mxl = (c.u == 1) ? 2.0f : 1.0f, mxr = (c.u == W - 2) ? 2.0f : 1.0f;
  const int ll = lane > 0 ? lane - 1 : 0, lr = lane < 31 ? lane + 1 : 31;
  const int li = pipe_box_col(c.x0, lane, W);
  const float xf = (float)c.px;

  Slide<V> sc[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) sc[j].reset();
  V accA[3], accB[3], accC[3];  // sum gc_i*d, sum gc_i*d*y, sum gc_i  (pose gradient, factored)
#pragma unroll
  for (int j = 0; j < 3; ++j) { accA[j] = vbc<V>(0.0f); accB[j] = vbc<V>(0.0f); accC[j] = vbc<V>(0.0f); }
  int win_prev = -1;
  float* gd_base = a.gdepth + (size_t)c.sb * HW + c.u;

  // coefficient rows exist for row indices ci = 1 .. n_rows-2 (centre rows y0-1 .. y1)
  for (int ci = 1; ci <= c.n_rows - 2; ++ci) {
    const int n = ci - 1, cs = n % SM::DC;
    mb_wait(bars + 2 * (SM::B_CFULL + cs), (unsigned)(n / SM::DC) & 1u);
    const float* crow = smem + SM::OFF_CO + cs * SM::COSLOT;
    const int rc = c.y0 - 3 + ci;  // the pixel row whose 3x3 gather completes with this coefficient row
    const float m_bot = (rc == H - 2) ? 2.0f : 1.0f;      // weight of centre row rc+1 for pixel row rc
    const float m_top_next = (rc + 1 == 1) ? 2.0f : 1.0f;  // weight of centre row rc for pixel row rc+1
    V co[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const V cm = lds_v<V>(crow + (j * 32 + lane) * K), cl = lds_v<V>(crow + (j * 32 + ll) * K), cr = lds_v<V>(crow + (j * 32 + lr) * K);
      // lane 0 / 31 read themselves as their missing neighbour: they are halo lanes, their sums are never used
      const V h = fma_(cl, vbc<V>(mxl), fma_(cr, vbc<V>(mxr), cm));
      co[j] = fma_(vbc<V>(m_bot), h, sc[j].p2);  // S(rc) = m_top * h(rc-1) + h(rc) + m_bot * h(rc+1)
      sc[j].p2 = fma_(vbc<V>(m_top_next), sc[j].p1, h);
      sc[j].p1 = h;
    }
    const int win = reinterpret_cast<const int*>(crow)[9 * 32 * K + lane];
    warp_sync();
    if (lane == 0) mb_arrive(bars + 2 * (SM::B_CEMPTY + cs));

    if (rc >= c.y0) {
      const int ix = ci - 1;  // ring index of pixel row rc
      const int xs = ix % SM::DX, ts = ix % SM::DT;
      mb_wait(bars + 2 * (SM::B_XFULL + xs), (unsigned)(ix / SM::DX) & 1u);  // long since complete: orders the reads below
      mb_wait(bars + 2 * (SM::B_TFULL + ts), (unsigned)(ix / SM::DT) & 1u);
      const float* xrow = smem + SM::OFF_X + xs * SM::XSLOT + lane * K;
      const float* trow = smem + SM::OFF_T + ts * SM::TROW;
      V gix = vbc<V>(0.0f), giy = vbc<V>(0.0f);
      V gl1;
#pragma unroll
      for (int k = 0; k < K; ++k) vset(gl1, k, (win_prev == k) ? g_l1 : 0.0f);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const V xc = lds_v<V>(xrow + ch * 32 * K), gxc = lds_v<V>(xrow + (3 + ch) * 32 * K), gyc = lds_v<V>(xrow + (6 + ch) * 32 * K);
        const float tc = trow[ch * SM::TBOX + li];
        V g = fma_(co[3 * ch + 2], vbc<V>(tc), fma_(co[3 * ch + 1], xc, co[3 * ch]));
        V sg;  // l1 = |target - pred|: d/d pred = -sign(target - pred), abs'(0) = 0
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float d = tc - vget(xc, k);
          vset(sg, k, (d > 0.0f) ? -1.0f : ((d < 0.0f) ? 1.0f : 0.0f));
        }
        g = fma_(sg, gl1, g);
        gix = fma_(g, gxc, gix);
        giy = fma_(g, gyc, giy);
      }
      if (!c.own_lane) { gix = vbc<V>(0.0f); giy = vbc<V>(0.0f); }
      const V jx = lds_v<V>(xrow + 9 * 32 * K), jy = lds_v<V>(xrow + 10 * 32 * K);
      const V ax = lds_v<V>(xrow + 11 * 32 * K), ay = lds_v<V>(xrow + 12 * 32 * K);
      const V ux = lds_v<V>(xrow + 13 * 32 * K), uy = lds_v<V>(xrow + 14 * 32 * K);
      const float dc = trow[SM::TDEP + li];
      const V gd = fma_(gix, jx, mul(giy, jy));
      float gdep = vget(gd, 0);
#pragma unroll
      for (int k = 1; k < K; ++k) gdep += vget(gd, k);
      if (c.own_lane) gd_base[(size_t)rc * W] = gdep;
      // d/dP, factored: P-row i gets gc_i * (X, Y, Z, 1) with (X,Y,Z) = depth * ray, ray linear in (x, y)
      const V gc0 = mul(gix, ax), gc1 = mul(giy, ay);
      const V gc2 = vneg(fma_(gc0, ux, mul(gc1, uy)));
      const float yc = (float)rc;
      const V w0 = mul(gc0, vbc<V>(dc)), w1 = mul(gc1, vbc<V>(dc)), w2 = mul(gc2, vbc<V>(dc));
      accA[0] = add(accA[0], w0); accA[1] = add(accA[1], w1); accA[2] = add(accA[2], w2);
      accB[0] = fma_(w0, vbc<V>(yc), accB[0]); accB[1] = fma_(w1, vbc<V>(yc), accB[1]); accB[2] = fma_(w2, vbc<V>(yc), accB[2]);
      accC[0] = add(accC[0], gc0); accC[1] = add(accC[1], gc1); accC[2] = add(accC[2], gc2);
    }
    win_prev = win;
    // row ci-1 (ring X slot and TMA box) is free for the gather warp
    warp_sync();
    if (lane == 0) mb_arrive(bars + 2 * (SM::B_XEMPTY + (ci - 1) % SM::DX));
  }

  // pose-gradient partials: fixed-order warp reduction, lane 0 writes
  {
    V gP[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const V xa = mul(accA[i], vbc<V>(xf));
#pragma unroll
      for (int j = 0; j < 3; ++j)
        gP[4 * i + j] = fma_(ldc<V>(cst, 12 + 3 * j), xa, fma_(ldc<V>(cst, 12 + 3 * j + 1), accB[i], mul(ldc<V>(cst, 12 + 3 * j + 2), accA[i])));
      gP[4 * i + 3] = accC[i];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      V v = gP[i];
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) v = add(v, vlane_xor(v, m));
      gP[i] = v;
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (k >= n_rep) continue;
        float* out = a.gpose_part + (((size_t)c.sb * BBD_MAX_REP + k) * part_stride + c.unit_in_sb) * 12;
#pragma unroll
        for (int i = 0; i < 12; ++i) out[i] = vget(gP[i], k);
      }
      for (int k = n_rep; k < BBD_MAX_REP; ++k) {
        float* out = a.gpose_part + (((size_t)c.sb * BBD_MAX_REP + k) * part_stride + c.unit_in_sb) * 12;
#pragma unroll
        for (int i = 0; i < 12; ++i) out[i] = 0.0f;
      }
    }
  }
  mb_wait(bars + 2 * SM::B_DONE, 0u);  // the statistics warp has written the unit's loss partial
}

// The fused finalize of stream_unit, for the warp that finishes a unit last (same tickets, same order of adding).
template <int K, bool GRAD>
BBD_HD void pipe_finalize(const bbd_reproj_args& a, const PipeCtx& c, int lane, int part_stride) {
  if (!a.tickets) return;
  const int H = a.height, W = a.width;
  const int n_rep = c.n_rep_raw < K ? c.n_rep_raw : K;
  const size_t tiles = (size_t)part_stride;
  const int sb = c.sb, s = c.s, b = c.b, upb = c.upb;
  warp_sync();
#if defined(__CUDA_ARCH__)
  __threadfence();
  int last = 0;
  if (lane == 0) last = (atomicAdd(a.tickets + sb, 1) == upb - 1) ? 1 : 0;
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
#else
  if (c.unit_in_sb != upb - 1) return;  // emulation runs the units in order
#endif
  {
    float v = 0.0f;
    for (int i = lane; i < upb; i += 32) v += ldcg1(a.loss_part + (size_t)sb * tiles + i);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += lane_xor(v, m);
    if (lane == 0) a.pair_sum[sb] = v;
  }
  if (GRAD && a.gpose_out && lane < 12 * K) {
    const int k = lane / 12, comp = lane - 12 * k;
    if (k < n_rep) {
      const float* p = a.gpose_part + (((size_t)sb * BBD_MAX_REP + k) * tiles) * 12 + comp;
      float v4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      int i = 0;
      for (; i + 4 <= upb; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v4[j] += ldcg1(p + (size_t)(i + j) * 12);
      }
      for (; i < upb; ++i) v4[0] += ldcg1(p + (size_t)i * 12);
      const float v = (v4[0] + v4[1]) + (v4[2] + v4[3]);
      const int pose = a.tab.rep[((size_t)b * BBD_MAX_REP + k) * 4 + 2];
      a.gpose_out[((size_t)s * a.num_pose + pose) * 12 + comp] = v;
    }
  }
#if defined(__CUDA_ARCH__)
  __threadfence();
  int last_s = 0;
  if (lane == 0) {
    a.tickets[sb] = 0;
    last_s = (atomicAdd(a.tickets + a.num_scales * a.batch + s, 1) == a.batch - 1) ? 1 : 0;
  }
  last_s = __shfl_sync(0xffffffffu, last_s, 0);
  if (!last_s) return;
  __threadfence();
#else
  if (b != a.batch - 1) return;
#endif
  if (lane == 0) {
    float tot = 0.0f;
    for (int i = 0; i < a.batch; ++i) tot += ldcg1(a.pair_sum + (size_t)s * a.batch + i);
    a.loss_out[s] = tot / ((float)a.batch * (float)H * (float)W);
    a.tickets[a.num_scales * a.batch + s] = 0;
  }
}

// One block of (GRAD ? 3 : 2) warps = one unit.  `tid` is the thread index in the block.
template <int K, bool GRAD>
BBD_HD void pipe_unit(const bbd_reproj_args& a, int unit, int tid, float* smem, int part_stride, const StreamTmaMaps& tm, int seg_rows) {
  typedef PipeSmem<K, GRAD> SM;
  const int warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    float* bars = smem + SM::OFF_BAR;
    for (int i = 0; i < SM::NBAR; ++i) mb_init(bars + 2 * i, 1);
    mb_init_fence();
  }
  block_sync();
  const PipeCtx c = pipe_ctx(a, unit, lane, seg_rows);
  if (warp == 0) {
    pipe_gather<K, GRAD>(a, c, lane, smem, tm);
  } else if (warp == 1) {
    pipe_stats<K, GRAD>(a, c, lane, smem, part_stride);
    if (!GRAD) pipe_finalize<K, false>(a, c, lane, part_stride);
  } else if (GRAD) {
    pipe_backward<K>(a, c, lane, smem, part_stride);
    pipe_finalize<K, true>(a, c, lane, part_stride);
  }
}

}  // namespace bbd
