// TEST-ONLY host harness: steps the thread-block phases of the sm_100a kernels on the CPU.
//
// The device kernels in baseboostdepth_b200/csrc/bbd_kernels.cu are thin drivers around the
// __host__ __device__ phase functions of bbd_strip.cuh / bbd_smooth.cuh / bbd_ops.cuh.  This
// file instantiates the same phase functions with g++ and runs, for every block, each
// phase for tid = 0..NT-1 before moving to the next (a barrier between phases), so index
// logic, halo/reflection handling, candidate tables and the analytic gradients can be
// checked against the oracle in the GPU-less build container.  It is never loaded by the
// package; exported symbols are emu_* and take HOST pointers.
#include <stdlib.h>
#include <string.h>

#include <vector>

#define BBD_EMU 1
#include "simt.h"
#include "../../baseboostdepth_b200/csrc/bbd_ops.cuh"
#include "../../baseboostdepth_b200/csrc/bbd_smooth.cuh"
#include "../../baseboostdepth_b200/csrc/bbd_strip.cuh"
#include "../../baseboostdepth_b200/csrc/bbd_stream.cuh"
#include "../../baseboostdepth_b200/csrc/bbd_pipe.cuh"

using namespace bbd;
#ifndef BBD_TILE_H
#define BBD_TILE_H 16
#endif
using SCfg = StripCfg<BBD_TILE_H, 6>;
#define FOR_STID for (int tid = 0; tid < SCfg::NT; ++tid)

extern "C" {

// number of significands (all 2^23, three binades) for which div_const(a, d, RN(1/d)) != a / d
long emu_div_const_mismatches(float d) {
  const float y = 1.0f / d;
  long bad = 0;
  const uint32_t exps[3] = {100u, 127u, 140u};
  for (int e = 0; e < 3; ++e)
    for (uint32_t m = 0; m < (1u << 23); ++m) {
      const uint32_t bits = (exps[e] << 23) | m;
      float a;
      memcpy(&a, &bits, 4);
      if (div_const(a, d, y) != a / d) ++bad;
    }
  return bad;
}

static int tile_parts(int height, int width) {
  return ((width + SCfg::TW - 1) / SCfg::TW) * ((height + SCfg::TH - 1) / SCfg::TH);
}
int emu_reproj_tiles(int32_t height, int32_t width) {
  int a = tile_parts(height, width);
  if (StreamGeo::strips(width) * stream_max_segs(height) > a) a = StreamGeo::strips(width) * stream_max_segs(height);
  if (StreamGeoM::units(height, width) > a) a = StreamGeoM::units(height, width);
  return a;
}
// same choice as the launcher in bbd_kernels.cu
static bool use_stream(const bbd_reproj_args* a) {
  if (a->force_tile || a->min_rep < 1 || a->max_rep > BBD_MAX_REP) return false;
  if (BBD_STREAM_ASYNC && a->max_rep > 2) return false;
  bool any = false;
  for (int f = 0; f < BBD_MAX_FRAMES; ++f) {
    if (a->frames[f] && !a->frames_rgba[f]) return false;
    any = any || a->frames_rgba[f];
  }
  return any;
}
// same choices as the launcher in bbd_kernels.cu
static bool split_multi(const bbd_reproj_args* a) {
  return use_stream(a) && a->max_rep > 2 && a->need_grad && a->winner != nullptr;
}
static int parts_used(const bbd_reproj_args* a, bool grad_parts) {
  if (!use_stream(a)) return tile_parts(a->height, a->width);
  if (a->max_rep > 2 && !(grad_parts && split_multi(a))) return StreamGeoM::units(a->height, a->width);
  const int rh = a->max_rep > 2 ? stream_seg_rows_uneven(a->height, a->width, a->num_scales * a->batch, 148 * 8)
                                : stream_seg_rows(a->height, a->width, a->num_scales * a->batch, 148 * 8);
  return StreamGeo::strips(a->width) * ((a->height + rh - 1) / rh);
}

int emu_project_coords(int32_t n, int32_t H, int32_t W, const float* depth, const float* inv_K, const float* P, float* grid,
                       float* pix) {
  for (int b = 0; b < n; ++b)
    for (int i = 0; i < H * W; ++i) stream_coords_px(H, W, depth, inv_K, P, b, i / W, i % W, grid, pix);
  return 0;
}

long emu_skipped_sweeps_total(void) { return emu_skipped_sweeps(); }

int emu_pack_rgba(int32_t n, int32_t H, int32_t W, const float* planar, float* rgba) {
  const size_t HW = (size_t)H * W;
  for (size_t img = 0; img < (size_t)n; ++img)
    for (size_t j = 0; j < HW; ++j) {
      float* o = rgba + (img * HW + j) * 4;
      o[0] = planar[img * 3 * HW + j];
      o[1] = planar[img * 3 * HW + HW + j];
      o[2] = planar[img * 3 * HW + 2 * HW + j];
      o[3] = 0.0f;
    }
  return 0;
}

}  // extern "C"
template <int K, bool GRAD, bool MULTI, bool WING = false>
static void emu_stream(const bbd_reproj_args& a) {
  const int n_units = a.num_scales * a.batch * parts_used(&a, WING);
  const int stride = emu_reproj_tiles(a.height, a.width);
  // like the launcher: the TMA-staged variant when the planes can be described to the TMA unit
#ifdef BBD_EMU_NO_TMA
  const bool tma = false;
#else
  const bool tma = !BBD_STREAM_ASYNC && a.width % 4 == 0;
#endif
  std::vector<float> smem(StreamSmem<K, true, MULTI>::FLOATS + StreamSmem<K, false, MULTI>::FLOATS + PipeSmem<K, GRAD>::FLOATS);
  StreamTmaMaps none = {nullptr, nullptr, nullptr};
  const int seg_rows = (MULTI && !WING) ? BBD_STREAM_RHM
                       : (MULTI ? stream_seg_rows_uneven(a.height, a.width, a.num_scales * a.batch, 148 * 8)
                                : stream_seg_rows(a.height, a.width, a.num_scales * a.batch, 148 * 8));
  // like the launcher: the pipelined three-warp form for single-sweep launches when BBD_PIPE=1
  const char* pe = getenv("BBD_PIPE");
  const bool pipe = tma && !MULTI && (pe ? pe[0] != '0' : false);
  for (int unit = 0; unit < n_units; ++unit) {
#if !BBD_STREAM_ASYNC
    if (pipe) {
      simt::run_block(GRAD ? 96 : 64, [&](int tid) { pipe_unit<K, GRAD>(a, unit, tid, smem.data(), stride, none, seg_rows); });
      continue;
    }
    if (tma) {
      simt::run_block(32, [&](int tid) { stream_unit<K, GRAD, true, MULTI, WING>(a, unit, tid, smem.data(), stride, none, seg_rows); });
      continue;
    }
#endif
    simt::run_block(32, [&](int tid) { stream_unit<K, GRAD, false, MULTI, WING>(a, unit, tid, smem.data(), stride, none, seg_rows); });
  }
  (void)tma;
}
extern "C" {

int emu_ident_forward(const bbd_ident_args* ap) {
  const bbd_ident_args& a = *ap;
  if (!a.force_tile) {
    const int seg_rows = ident_seg_rows(a.height, a.width, a.batch, 148 * 16);
    const int n_units = a.batch * IdentGeo::strips(a.width) * ((a.height + seg_rows - 1) / seg_rows);
    for (int unit = 0; unit < n_units; ++unit)
      simt::run_block(32, [&](int tid) { ident_unit(a, const_cast<float* const*>(a.frames_rgba), unit, tid, seg_rows); });
    return 0;
  }
  std::vector<float> smem(IdentStripSmem<SCfg>::floats());
  std::vector<StripCtx> ctx(SCfg::NT);
  const int H = a.height, W = a.width;
  const int gx = (W + SCfg::TW - 1) / SCfg::TW, gy = (H + SCfg::TH - 1) / SCfg::TH;
  for (int bz = 0; bz < a.batch; ++bz)
    for (int by = 0; by < gy; ++by)
      for (int bx = 0; bx < gx; ++bx) {
        IdentStripSmem<SCfg> sm;
        sm.carve(smem.data());
        FOR_STID ctx[tid] = make_strip<SCfg>(bx, by, bz, tid, a.batch, H, W);
        const int b = ctx[0].b;
        const int32_t* hdr = a.tab.hdr + (size_t)b * 4;
        const int n_id = hdr[1];
        const float* noise = a.noise[hdr[2]] + (size_t)hdr[3] * H * W;
        FOR_STID is_load<SCfg>(a.target + (size_t)b * 3 * H * W, sm.tgt, ctx[tid], H, W);
        FOR_STID is_target_stats<SCfg>(a, sm, ctx[tid]);
        for (int j = 0; j < n_id; ++j) {
          const int32_t* e = a.tab.ident + ((size_t)b * BBD_MAX_IDENT + j) * 2;
          FOR_STID is_load<SCfg>(a.frames[e[0]] + (size_t)e[1] * 3 * H * W, sm.src, ctx[tid], H, W);
          FOR_STID is_candidate<SCfg>(a, sm, ctx[tid], j, noise);
        }
        FOR_STID is_store<SCfg>(a, sm, ctx[tid]);
      }
  return 0;
}

int emu_reproj_fused(const bbd_reproj_args* ap) {
  const bbd_reproj_args& a = *ap;
  if (use_stream(ap)) {
    if (a.max_rep == 1) { if (a.need_grad) emu_stream<1, true, false>(a); else emu_stream<1, false, false>(a); }
    else if (a.max_rep == 2) { if (a.need_grad) emu_stream<2, true, false>(a); else emu_stream<2, false, false>(a); }
#if !BBD_STREAM_ASYNC
    else if (split_multi(&a)) { emu_stream<2, false, true>(a); emu_stream<2, true, true, true>(a); }
    else { if (a.need_grad) emu_stream<2, true, true>(a); else emu_stream<2, false, true>(a); }
#endif
    return 0;
  }
  // the CPU harness runs the reuse (non-KEEP) variant for batches with more than two candidates,
  // like the launcher does
  const bool keep = StripSmem<SCfg>::floats(a.max_rep) * sizeof(float) <= 75 * 1024;
  std::vector<float> smem(StripSmem<SCfg>::floats(keep ? a.max_rep : 1));
  std::vector<float> red(12 * SCfg::NT + 12 * SCfg::RED_SEG * 2);  // host-side reduction scratch
  std::vector<float> gPs((size_t)SCfg::NT * 12), parts(SCfg::NT);
  std::vector<StripCtx> ctx(SCfg::NT);
  const int gx = (a.width + SCfg::TW - 1) / SCfg::TW, gy = (a.height + SCfg::TH - 1) / SCfg::TH;
  for (int bz = 0; bz < a.num_scales * a.batch; ++bz)
    for (int by = 0; by < gy; ++by)
      for (int bx = 0; bx < gx; ++bx) {
        StripSmem<SCfg> sm;
        sm.carve(smem.data(), keep ? a.max_rep : 1);
        sm.red = red.data();
        FOR_STID ctx[tid] = make_strip<SCfg>(bx, by, bz, tid, a.batch, a.height, a.width);
        const int b = ctx[0].b, tile = ctx[0].tile, ntiles = ctx[0].ntiles;
        const int n_rep = a.tab.hdr[(size_t)b * 4];
        FOR_STID rs_load_target<SCfg>(a, sm, ctx[tid], tid);
        FOR_STID rs_target_stats<SCfg>(a, sm, ctx[tid]);
        {
          const int s = ctx[0].s;
          FOR_STID rs_begin_scale<SCfg>(sm, tid);
          for (int k = 0; k < n_rep; ++k) {
            if (keep) {
              FOR_STID rs_warp<SCfg, true>(a, sm, ctx[tid], k);
              FOR_STID rs_stats<SCfg, true>(a, sm, ctx[tid], k);
            } else {
              FOR_STID rs_warp<SCfg, false>(a, sm, ctx[tid], k);
              FOR_STID rs_stats<SCfg, false>(a, sm, ctx[tid], k);
            }
          }
          FOR_STID parts[tid] = rs_select<SCfg>(a, sm, ctx[tid], n_rep);
          FOR_STID rs_park<SCfg, 1>(sm.red, tid, &parts[tid]);
          FOR_STID rs_level1<SCfg, 1>(sm.red, tid);
          FOR_STID rs_level2<SCfg, 1>(sm.red, tid, a.loss_part + ((size_t)s * a.batch + b) * ntiles + tile);
          if (!a.need_grad) continue;
          for (int k = 0; k < BBD_MAX_REP; ++k) {
            float* out = a.gpose_part + ((((size_t)s * a.batch + b) * BBD_MAX_REP + k) * ntiles + tile) * 12;
            if (k >= n_rep || !sm.anywin[k]) {
              for (int i = 0; i < 12; ++i) out[i] = 0.0f;
              continue;
            }
            {
              const float* src;
              Cam cam;
              rs_candidate(a, b, k, src, cam);
              for (size_t i = 0; i < gPs.size(); ++i) gPs[i] = 0.0f;
              for (int m = 0; m * SCfg::NW < SCfg::TH; m += BBD_BWD_ROWS) {   // BBD_BWD_ROWS rows per warp at a time, lanes in lockstep
                for (int r = 0; r < BBD_BWD_ROWS; ++r)
                  FOR_STID { const int q = ctx[tid].warp + (m + r) * SCfg::NW; if (q < SCfg::TH) rs_bwd_vertical<SCfg>(a, sm, ctx[tid], k, q, r); }
                for (int r = 0; r < BBD_BWD_ROWS; ++r)
                  FOR_STID { const int q = ctx[tid].warp + (m + r) * SCfg::NW; if (q < SCfg::TH) { if (keep) rs_bwd_horizontal<SCfg, true>(a, sm, ctx[tid], k, q, src, cam, &gPs[(size_t)tid * 12], r); else rs_bwd_horizontal<SCfg, false>(a, sm, ctx[tid], k, q, src, cam, &gPs[(size_t)tid * 12], r); } }
              }
            }
            FOR_STID rs_park<SCfg, 12>(sm.red, tid, &gPs[(size_t)tid * 12]);
            FOR_STID rs_level1<SCfg, 12>(sm.red, tid);
            FOR_STID rs_level2<SCfg, 12>(sm.red, tid, out);
          }
          FOR_STID rs_store_gdepth<SCfg>(a, sm, ctx[tid]);
        }
      }
  return 0;
}

const char* emu_reproj_kernel_name(const bbd_reproj_args*) { return "emulation"; }
int emu_reproj_finalizes_itself(const bbd_reproj_args* a) {
  return a && use_stream(a) && a->max_rep <= 2 && a->tickets && a->pair_sum && a->loss_out && (!a->need_grad || a->gpose_out) ? 1 : 0;
}

int emu_reproj_finalize(const bbd_reproj_args* ap, float* loss, float* gpose) {
  const bbd_reproj_args& a = *ap;
  const int ntiles = emu_reproj_tiles(a.height, a.width), used = parts_used(ap, false), used_grad = parts_used(ap, true);
  for (int s = 0; s < a.num_scales; ++s) {
    float tot = 0.0f;
    for (int i = 0; i < a.batch * used; ++i) tot += a.loss_part[(size_t)s * a.batch * ntiles + (size_t)(i / used) * ntiles + (i % used)];
    loss[s] = tot / ((float)a.batch * (float)a.height * (float)a.width);
  }
  if (!gpose) return 0;
  for (int s = 0; s < a.num_scales; ++s)
    for (int pose = 0; pose < a.num_pose; ++pose)
      for (int c = 0; c < 12; ++c) {
        float acc = 0.0f;
        for (int b = 0; b < a.batch; ++b) {
          const int n_rep = a.tab.hdr[(size_t)b * 4];
          for (int k = 0; k < n_rep; ++k) {
            if (a.tab.rep[((size_t)b * BBD_MAX_REP + k) * 4 + 2] != pose) continue;
            const float* p = a.gpose_part + (((size_t)s * a.batch + b) * BBD_MAX_REP + k) * ntiles * 12;
            for (int tI = 0; tI < used_grad; ++tI) acc += p[(size_t)tI * 12 + c];
          }
        }
        gpose[((size_t)s * a.num_pose + pose) * 12 + c] = acc;
      }
  return 0;
}

static int smooth_max_parts(int levels, const int32_t* h, const int32_t* w) {
  int mc = 1;
  for (int l = 0; l < levels; ++l) {
    const int c = sm_chunks(h[l], w[l]) > smr_blocks(h[l], w[l]) ? sm_chunks(h[l], w[l]) : smr_blocks(h[l], w[l]);
    mc = c > mc ? c : mc;
  }
  return mc;
}
size_t emu_smooth_scratch_floats(int32_t batch, int32_t levels, const int32_t* h, const int32_t* w) {
  return (size_t)levels * batch * 4 * smooth_max_parts(levels, h, w) + (size_t)levels * batch * 2 + 4;
}

int emu_smooth_fused(const bbd_smooth_args* in) {
  bbd_smooth_args a = *in;
  const int mc = smooth_max_parts(a.levels, a.h, a.w);
  a.max_chunks = mc;
  float* tail = a.scratch + (size_t)a.levels * a.batch * 4 * mc;
  float* coef = a.defer_norm ? a.coef : tail;
  // stage 2: one fiber block per warp of the device kernel's blocks
  for (int lvl = 0; lvl < a.levels; ++lvl)
    for (int b = 0; b < a.batch; ++b) {
      const int nbx = smr_nbx(a.w[lvl]), nby = smr_nby(a.h[lvl]);
      for (int by = 0; by < nby; ++by)
        for (int bx = 0; bx < nbx; ++bx) {
          float tot[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          for (int warp = 0; warp < SMR_WARPS; ++warp) {
            float out[4];
            simt::run_block(32, [&](int tid) {
              float o[4];
              sm_rows_lane(a, lvl, b, bx, by, warp, tid, o);
              if (tid == 0) { out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; out[3] = o[3]; }
            });
            for (int k = 0; k < 4; ++k) tot[k] += out[k];
          }
          for (int k = 0; k < 4; ++k) sm_slot(a, lvl, b, k < 3 ? 1 + k : 0)[by * nbx + bx] = tot[k];
        }
    }
  std::vector<float> tx(a.levels * a.batch), ty(a.levels * a.batch);
  for (int i = 0; i < a.levels * a.batch; ++i) {
    float sums[2];
    sm_finish_sample(a, i / a.batch, i % a.batch, coef, sums);
    tx[i] = sums[0];
    ty[i] = sums[1];
  }
  for (int lvl = 0; lvl < a.levels; ++lvl) a.loss[lvl] = sm_level_loss(a, lvl, tx.data() + lvl * a.batch, ty.data() + lvl * a.batch);
  if (!a.defer_norm && a.normalize)
    for (int lvl = 0; lvl < a.levels; ++lvl)
      if (a.gdisp[lvl])
        for (int b = 0; b < a.batch; ++b)
          for (int i = 0; i < a.h[lvl] * a.w[lvl]; ++i) sm_apply_px(a, coef, lvl, b, i);
  return 0;
}

int emu_disp_to_depth_forward(const bbd_d2d_args* ap) {
  const bbd_d2d_args& a = *ap;
  const int HW = a.height * a.width;
  for (int lvl = 0; lvl < a.levels; ++lvl) {
    const float sy = (float)a.h[lvl] / (float)a.height, sx = (float)a.w[lvl] / (float)a.width;
    for (int b = 0; b < a.batch; ++b)
      for (int i = 0; i < HW; ++i)
        a.depth[((size_t)lvl * a.batch + b) * HW + i] = d2d_forward_px(a, lvl, b, i / a.width, i % a.width, sy, sx);
  }
  return 0;
}

size_t emu_d2d_scratch_floats(const bbd_d2d_args* a) { return d2d_scratch_offset(*a, a->levels); }

int emu_disp_to_depth_backward_pass1(const bbd_d2d_args* ap) {
  const bbd_d2d_args& a = *ap;
  if (a.scratch)
    for (int lvl = 0; lvl < a.levels; ++lvl) {
      const int f = d2d_sep_factor(a, lvl);
      if (!f) continue;
      const int w = a.w[lvl], H = a.height;
      float* tmp = a.scratch + d2d_scratch_offset(a, lvl);
      for (int i = 0; i < a.batch * H * w; ++i) {
        const int ix = i % w, r = i / w, oy = r % H, b = r / H;
        tmp[i] = f == 2 ? d2d_hpass<2>(a, lvl, b, oy, ix) : (f == 4 ? d2d_hpass<4>(a, lvl, b, oy, ix) : d2d_hpass<8>(a, lvl, b, oy, ix));
      }
    }
  return 0;
}

int emu_disp_to_depth_backward_pass2(const bbd_d2d_args* ap, int32_t level_begin, int32_t level_end) {
  const bbd_d2d_args& a = *ap;
  for (int lvl = level_begin; lvl < level_end; ++lvl) {
    const int h = a.h[lvl], w = a.w[lvl];
    const float sy = (float)h / (float)a.height, sx = (float)w / (float)a.width;
    for (int b = 0; b < a.batch; ++b)
      for (int i = 0; i < h * w; ++i) a.gdisp[lvl][(size_t)b * h * w + i] = d2d_backward_px(a, lvl, b, i / w, i % w, sy, sx);
  }
  return 0;
}

}  // extern "C"
template <int F>
static void emu_d2d_fused_level(const bbd_d2d_args& a, int lvl) {
  std::vector<float> srow(a.width);
  for (int b = 0; b < a.batch; ++b)
    for (int iy = 0; iy < a.h[lvl]; ++iy) {
      for (int x4 = 0; x4 < a.width; x4 += 4) d2d_fused_col4<F>(a, lvl, b, iy, x4, srow.data() + x4);
      for (int ix = 0; ix < a.w[lvl]; ++ix)
        a.gdisp[lvl][((size_t)b * a.h[lvl] + iy) * a.w[lvl] + ix] = d2d_fused_out<F>(a, lvl, b, iy, ix, srow.data());
    }
}
extern "C" {
int emu_disp_to_depth_backward(const bbd_d2d_args* ap) {
  const bbd_d2d_args& a = *ap;
  bool fused = true;  // same choice as the launcher
  for (int l = 0; l < a.levels; ++l) fused = fused && d2d_fused_factor(a, l) != 0;
  if (fused) {
    for (int l = 0; l < a.levels; ++l) {
      const int f = d2d_fused_factor(a, l);
      if (f == 1) emu_d2d_fused_level<1>(a, l);
      else if (f == 2) emu_d2d_fused_level<2>(a, l);
      else if (f == 4) emu_d2d_fused_level<4>(a, l);
      else emu_d2d_fused_level<8>(a, l);
    }
    return 0;
  }
  emu_disp_to_depth_backward_pass1(ap);
  return emu_disp_to_depth_backward_pass2(ap, 0, ap->levels);
}

int emu_pose_forward(int32_t n, const float* aa, const float* tr, int32_t invert, float* T) {
  for (int i = 0; i < n; ++i) pose_forward_one(aa + (size_t)i * 3, tr + (size_t)i * 3, invert, T + (size_t)i * 16);
  return 0;
}
int emu_pose_backward(int32_t n, const float* aa, const float* tr, int32_t invert, const float* gT, float* gaa, float* gtr) {
  for (int i = 0; i < n; ++i)
    pose_backward_one(aa + (size_t)i * 3, tr + (size_t)i * 3, invert, gT + (size_t)i * 16, gaa + (size_t)i * 3, gtr + (size_t)i * 3);
  return 0;
}

int emu_pose_pack_forward(int32_t n, const float* K, const int32_t* k_row, const float* T, float* P) {
  for (int i = 0; i < n * 12; ++i)
    P[i] = pose_pack_elem(K + (size_t)k_row[i / 12] * 16, T + (size_t)(i / 12) * 16, (i % 12) / 4, i % 4);
  return 0;
}
int emu_pose_pack_backward(int32_t n, const float* K, const int32_t* k_row, const float* gP, float* gT) {
  for (int i = 0; i < n * 16; ++i)
    gT[i] = pose_pack_grad_elem(K + (size_t)k_row[i / 16] * 16, gP + (size_t)(i / 16) * 12, (i % 16) / 4, i % 4);
  return 0;
}

int emu_warp_forward(int32_t n, int32_t H, int32_t W, const float* images, const float* depth, const float* inv_K,
                     const float* P, float* warped, float* grid) {
  for (int b = 0; b < n; ++b)
    for (int i = 0; i < H * W; ++i) warp_px(H, W, images, depth, inv_K, P, b, i / W, i % W, warped, grid);
  return 0;
}

int emu_backproject_forward(int32_t n, int32_t H, int32_t W, const float* depth, const float* inv_K, float* points) {
  for (int b = 0; b < n; ++b)
    for (int i = 0; i < H * W; ++i) backproject_px(H * W, W, depth, inv_K, b, i, points);
  return 0;
}
int emu_backproject_backward(int32_t n, int32_t H, int32_t W, const float* inv_K, const float* gpoints, float* gdepth) {
  for (int b = 0; b < n; ++b)
    for (int i = 0; i < H * W; ++i) gdepth[(size_t)b * H * W + i] = backproject_grad_px(H * W, W, inv_K, gpoints, b, i);
  return 0;
}
int emu_project_forward(int32_t n, int32_t H, int32_t W, const float* points, const float* P, float eps, float* pix) {
  for (int b = 0; b < n; ++b)
    for (int i = 0; i < H * W; ++i) project_px(H, W, points, P, eps, b, i, pix);
  return 0;
}
int emu_project_chunks(int32_t, int32_t) { return 1; }
int emu_project_backward(int32_t n, int32_t H, int32_t W, const float* points, const float* P, float eps, const float* gpix,
                         float* gpoints, float* gP_part) {
  for (int b = 0; b < n; ++b) {
    float gP[12] = {0};
    for (int i = 0; i < H * W; ++i) project_grad_px(H, W, points, P, eps, gpix, b, i, gpoints, gP);
    for (int k = 0; k < 12; ++k) gP_part[(size_t)b * 12 + k] = gP[k];
  }
  return 0;
}
int emu_grid_sample_forward(int32_t n, int32_t C, int32_t H, int32_t W, int32_t Ho, int32_t Wo, const float* images,
                            const float* grid, float* out) {
  for (int b = 0; b < n; ++b)
    for (int o = 0; o < Ho * Wo; ++o) grid_sample_px(C, H, W, Ho * Wo, images, grid, b, o, out);
  return 0;
}
int emu_grid_sample_dest_keys(int32_t n, int32_t H, int32_t W, int32_t Ho, int32_t Wo, const float* grid, int32_t* keys) {
  for (int b = 0; b < n; ++b)
    for (int o = 0; o < Ho * Wo; ++o) keys[(size_t)b * Ho * Wo + o] = gs_dest_key(H, W, Ho * Wo, grid, b, o);
  return 0;
}
int emu_grid_sample_backward_image(int32_t n, int32_t C, int32_t H, int32_t W, int32_t Ho, int32_t Wo, const float* grid,
                                   const float* gout, const int32_t* keys_sorted, const int32_t* order, int32_t* seg_start,
                                   float* gimages) {
  const int n_items = n * Ho * Wo, n_keys = n * H * W;
  for (int i = 0; i <= n_items; ++i) gs_segment_mark(keys_sorted, n_items, n_keys, i, seg_start);
  for (int b = 0; b < n; ++b)
    for (int c = 0; c < C; ++c)
      for (int p = 0; p < H * W; ++p)
        gimages[((size_t)b * C + c) * H * W + p] = gs_image_grad_px(C, H, W, Ho * Wo, grid, gout, seg_start, order, b, c, p);
  return 0;
}
int emu_grid_sample_backward(int32_t n, int32_t C, int32_t H, int32_t W, int32_t Ho, int32_t Wo, const float* images,
                             const float* grid, const float* gout, float* ggrid) {
  for (int b = 0; b < n; ++b)
    for (int o = 0; o < Ho * Wo; ++o) grid_sample_grad_px(C, H, W, Ho * Wo, images, grid, gout, b, o, ggrid);
  return 0;
}

int emu_ssim_forward(int32_t n, int32_t ch, int32_t H, int32_t W, const float* x, const float* y, float* out) {
  for (size_t pl = 0; pl < (size_t)n * ch; ++pl)
    for (int i = 0; i < H * W; ++i) out[pl * H * W + i] = ssim_px(x + pl * H * W, y + pl * H * W, H, W, i / W, i % W);
  return 0;
}
int emu_ssim_backward(int32_t n, int32_t ch, int32_t H, int32_t W, const float* x, const float* y, const float* gout,
                      float* gx, float* gy) {
  for (size_t pl = 0; pl < (size_t)n * ch; ++pl)
    for (int i = 0; i < H * W; ++i)
      ssim_grad_px(x + pl * H * W, y + pl * H * W, gout + pl * H * W, H, W, i / W, i % W, gx ? gx + pl * H * W : nullptr,
                   gy ? gy + pl * H * W : nullptr);
  return 0;
}

int emu_loss_combine_forward(int32_t n, const float* reproj, const float* smooth, const float* weight, float num_scales,
                             float* per_scale, float* total) {
  loss_combine(n, reproj, smooth, weight, num_scales, per_scale, total);
  return 0;
}
int emu_loss_combine_backward(int32_t n, const float* g_total, const float* g_per_scale, const float* weight,
                              float num_scales, float* g_reproj, float* g_smooth) {
  loss_combine_grad(n, g_total, g_per_scale, weight, num_scales, g_reproj, g_smooth);
  return 0;
}

int emu_u8_to_f32(const uint8_t* src, float* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = u8_to_unit(src[i]);
  return 0;
}

}  // extern "C"
