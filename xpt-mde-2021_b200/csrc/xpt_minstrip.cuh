// xpt_minstrip.cuh -- k_min_strip: the min-over-sources losses (MonoDepth2LossMultiScale, reference
// losses.py:198-232; MoALossMultiScale :282-321) on k_fused's strip geometry.
//
// Same semantics as k_photo_min (xpt_minloss.cuh: every scale's synthesis up-sampled to full resolution by
// resize_bilinear :377-383, photometric term per pixel AND channel against the full-resolution target, tf.reduce_min
// over the sources with the gradient split equally among ties, a black synthesised pixel has loss 0 and wins), but
// built like the fused kernel instead of one thread per statistics position:
//   * one CTA (512 threads, 2 CTAs/SM) = one 64x13 full-resolution centre tile of one (scale, snippet); warps 0..14
//     own one statistics row each, a lane owns a 2-pixel strip, so the 3x3 window sums are separable (vertical sums
//     first) on 8-byte shared-memory rows and run on the packed FP32 pipe (FFMA2 / FADD2 / FMUL2);
//   * the running minimum, the tie count and the first winner of every (pixel, channel) of a strip live in REGISTERS
//     across both sweeps (k_photo_min kept them in shared memory and re-read them per source);
//   * the target's window statistics are re-formed per source from the rows the cross term loads anyway (18 registers
//     of statistics kept across the source loop pushed the running minima into local memory: -25 % time without);
//   * up-sampling taps (lo, hi, lerp per tile row / column) are tabulated once per tile;
//   * the adjoint of the up-sampling is reduced inside the tile first: dL/dS of the centre goes to shared memory, a
//     horizontal and a vertical pass fold it onto the tile's low-resolution footprint, and ONE atomic per low-resolution
//     value and tile reaches L2 (k_photo_min: 12 atomics per full-resolution pixel and source).  A level at full
//     resolution owns its pixels: plain stores.
#pragma once
#include <cassert>
#include <type_traits>
#include "xpt_fused.cuh"
#include "xpt_minloss.cuh"

// XPT_MS_CHECK 1: device-side bounds assertions of every table-driven index (a checking build for the parity tests;
// compute-sanitizer is not available on the pool)
#ifndef XPT_MS_CHECK
#define XPT_MS_CHECK 0
#endif
#if XPT_MS_CHECK
#define XPT_MS_ASSERT(c) assert(c)
#else
#define XPT_MS_ASSERT(c) ((void)0)
#endif

namespace xpt {

constexpr int kMSTPitch = 36;                   // low-resolution footprint columns of a 64-wide tile at scale >= 2: <= 34
constexpr int kMSTabs = 640;                    // x0/x1/fx [68] + y0/y1/fy [17], then the footprint's column / row ranges
constexpr int kMSRanges = 256;                  // offset of rXa[36], rXb[36], rYa[16], rYb[16] inside the table block
constexpr int kMSFlowTabs = 368;                // offset of the tap tables of the flow-warped view (CMB), same layout as the first 255

template <bool GRAD, bool PAIR>
struct MinStripSmem {
  static constexpr int sx = 0;                                  // [3][17][68] target tile
  static constexpr int sy = sx + 3 * kFRegion + 4;              // [3][17][68] up-sampled synthesis of the current source
  static constexpr int tab = sy + 3 * kFRegion + 4;             // tap tables
  static constexpr int red = tab + kMSTabs;                     // [16] loss scratch
  static constexpr int sA = red + 32;                           // GRAD: [3][15][68] x3 adjoint coefficient planes
  static constexpr int sB = sA + (GRAD ? 3 * kFStats + 4 : 0);
  static constexpr int sC = sB + (GRAD ? 3 * kFStats + 4 : 0);
  static constexpr int sL = sC + (GRAD ? 3 * kFStats + 4 : 0);  // GRAD && PAIR: [3][15][68] local (L1) derivative
  static constexpr int sG = sL + (GRAD && PAIR ? 3 * kFStats + 4 : 0);  // GRAD: [3][13][64] dL/dS of the centre
  static constexpr int sT = sG + (GRAD ? 3 * kFCentre : 0);     // GRAD: [3][13][36] after the horizontal pass
  static constexpr int kFloats = sT + (GRAD ? 3 * kFCH * kMSTPitch : 0);
  static constexpr size_t kBytes = sizeof(float) * kFloats;
};

// PAIR: the L1 and the SSIM loss of one loss set (moaL1 + moaSSIM, md2L1 + md2SSIM) in ONE launch -- the two minima are
// independent, but the up-sampled tiles, the black-pixel masks, the target tile and the up-sampling adjoint are shared;
// the gradient written is pair_c_l1 dL1/dS + pair_c_ssim dSSIM/dS.
// CMB: CombinedLossMultiScale (losses.py:235-279) instead of the minimum over sources -- per source, the static term
// counts where it is smaller than the term of the flow-warped view (a.cmb_flow, up-sampled from its own size); one sweep:
// the flow term of a source goes to the registers the minimum lives in, the static term is compared with it, and the
// gradient of the kept terms is formed and reduced right away.
template <bool GRAD, bool PAIR, bool CMB>
__global__ void __launch_bounds__(kFThreads, 2) k_min_strip(const __grid_constant__ MinLossArgs a) {
  using SM = MinStripSmem<GRAD, PAIR>;
  constexpr int NM = PAIR ? 2 : 1;              // minima tracked: [0] = the method (PAIR: L1), [1] = SSIM of the pair
  extern __shared__ __align__(16) float smem[];
  float* const sx = smem + SM::sx;
  float* const sy = smem + SM::sy;
  int* const tX0 = reinterpret_cast<int*>(smem + SM::tab);
  int* const tX1 = tX0 + kFRW;
  float* const tFX = smem + SM::tab + 2 * kFRW;
  int* const tY0 = reinterpret_cast<int*>(smem + SM::tab + 3 * kFRW);
  int* const tY1 = tY0 + kFRH;
  float* const tFY = smem + SM::tab + 3 * kFRW + 2 * kFRH;
  // adjoint of the up-sampling: the tile columns (rows) whose LOWER tap is footprint column X (row Y) form one
  // contiguous range, those whose UPPER tap is X another; packed first | last << 8, tile-relative, empty = 1 | 0 << 8
  int* const rXa = reinterpret_cast<int*>(smem + SM::tab + kMSRanges);
  int* const rXb = rXa + kMSTPitch;
  int* const rYa = rXb + kMSTPitch;
  int* const rYb = rYa + 16;
  int* const fX0 = reinterpret_cast<int*>(smem + SM::tab + kMSFlowTabs);
  int* const fX1 = fX0 + kFRW;
  float* const fFX = smem + SM::tab + kMSFlowTabs + 2 * kFRW;
  int* const fY0 = reinterpret_cast<int*>(smem + SM::tab + kMSFlowTabs + 3 * kFRW);
  int* const fY1 = fY0 + kFRH;
  float* const fFY = smem + SM::tab + kMSFlowTabs + 3 * kFRW + 2 * kFRH;
  float* const red = smem + SM::red;
  float* const sA = smem + SM::sA;
  float* const sB = smem + SM::sB;
  float* const sC = smem + SM::sC;
  float* const sG = smem + SM::sG;
  float* const sT = smem + SM::sT;
  float* const sLp = PAIR ? smem + SM::sL : sA; // plane of the local derivative (single L1 / L2 launch: sA is free)

  const int b = blockIdx.x;
  int t = blockIdx.y;
  const int l = t / a.tiles;
  t -= l * a.tiles;
  const int H = a.H, W = a.W, h = a.h[l], w = a.w[l];
  const int ty0 = (t / a.tiles_x) * kFCH, tx0 = (t % a.tiles_x) * kFCW;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool ssim = PAIR || a.method == 2, do_l1 = PAIR || a.method != 2, l2 = !PAIR && a.method == 1;
  constexpr int iS = PAIR ? 1 : 0;              // index of the SSIM minimum
  const bool identity = (h == H) && (w == W);
  const int nsrc = CMB ? a.N : a.N + a.NS;

  // ---- tap tables (resize_bilinear, half-pixel centres) of the tile's rows and columns, target tile ----------
  if (tid < kFRW) {
    int lo, hi; float f;
    up_taps(min(max(tx0 - 2 + tid, 0), W - 1), w, (float)w / (float)W, lo, hi, f);
    tX0[tid] = lo; tX1[tid] = hi; tFX[tid] = f;
  } else if (tid >= 96 && tid < 96 + kFRH) {
    int lo, hi; float f;
    up_taps(min(max(ty0 - 2 + (tid - 96), 0), H - 1), h, (float)h / (float)H, lo, hi, f);
    tY0[tid - 96] = lo; tY1[tid - 96] = hi; tFY[tid - 96] = f;
  } else if (CMB && tid >= 128 && tid < 128 + kFRW) {
    int lo, hi; float f;
    up_taps(min(max(tx0 - 2 + (tid - 128), 0), W - 1), a.cmb_w, (float)a.cmb_w / (float)W, lo, hi, f);
    fX0[tid - 128] = lo; fX1[tid - 128] = hi; fFX[tid - 128] = f;
  } else if (CMB && tid >= 224 && tid < 224 + kFRH) {
    int lo, hi; float f;
    up_taps(min(max(ty0 - 2 + (tid - 224), 0), H - 1), a.cmb_h, (float)a.cmb_h / (float)H, lo, hi, f);
    fY0[tid - 224] = lo; fY1[tid - 224] = hi; fFY[tid - 224] = f;
  }
  {
    const float* tgt = a.target + b * a.tgt_bs;
#pragma unroll
    for (int it = 0; it < kFYIters; ++it) {
      const int i = tid + it * kFThreads;
      if (i < kFRegion) {
        const int ry = i / kFP, rx = i - ry * kFP;
        const int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
          const float* p = tgt + ((long long)gy * W + gx) * 3;
          v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2);
        }
        sx[i] = v0; sx[kFRegion + i] = v1; sx[2 * kFRegion + i] = v2;
      }
    }
  }
  // (staging each source's low-resolution footprint in shared memory, next source prefetched during the terms, was
  // measured: +10 % time -- the taps already hit L1)
  if (GRAD && (ty0 == 0 || ty0 + kFCH + 1 > H))    // block-uniform: a statistics row outside the image is never written
    for (int i = tid; i < kFStats; i += kFThreads)
      if ((unsigned)(ty0 - 1 + i / kFP) >= (unsigned)H) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          sA[c * kFStats + i] = 0.f; sB[c * kFStats + i] = 0.f; sC[c * kFStats + i] = 0.f;
          if (PAIR) sLp[c * kFStats + i] = 0.f;
        }
      }
  __syncthreads();

  // ---- strip coordinates (as in k_fused) -------------------------------------------------------------------
  const int qy = wid < kFSH ? wid : lane, q0 = wid < kFSH ? lane * 2 : 2 * (kFSStrips - 1);
  const bool s_active = (wid < kFSH || lane < kFSH) && (unsigned)(ty0 - 1 + qy) < (unsigned)H;
  const int cyy = tid >> 5, c0 = (tid & 31) * 2;
  const bool g_active = cyy < kFCH && ty0 + cyy < H;

  const float gb = (GRAD && a.gbatch) ? __ldg(a.gbatch + b) : 1.f;
  const float coefL = gb * a.norm[l] * (PAIR ? a.pair_c_l1 : 1.f), coefS = gb * a.norm[l] * (PAIR ? a.pair_c_ssim : 1.f);

  float2 inv_cnt, s_centre;               // per pixel of the strip: 1/#taps (0 outside the image), counted in the loss
  {
    const int gy = ty0 - 1 + qy;
    const bool row_in = s_active && gy >= 0 && gy < H;
    const int cy = min(gy + 1, H - 1) - max(gy - 1, 0) + 1;
    float iv[2], ce[2];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const int q = q0 + o, gx = tx0 - 1 + q;
      const bool inb = row_in && q < kFSW && gx >= 0 && gx < W;
      const int cx = min(gx + 1, W - 1) - max(gx - 1, 0) + 1;
      iv[o] = inb ? box_inv(cy * cx) : 0.f;
      ce[o] = (inb && qy >= 1 && qy <= kFCH && q >= 1 && q <= kFCW) ? 1.f : 0.f;
    }
    inv_cnt = f2(iv[0], iv[1]); s_centre = f2(ce[0], ce[1]);
    asm volatile("" : "+f"(inv_cnt.x), "+f"(inv_cnt.y), "+f"(s_centre.x), "+f"(s_centre.y));
  }
  // running minimum per (channel, pixel of the strip); code = both pixels' (first winner << 8 | number of sources at
  // the minimum), 16 bits each
  float vmin[NM][3][2];
  unsigned code[NM][3];
#pragma unroll
  for (int k = 0; k < NM; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) { vmin[k][c][0] = vmin[k][c][1] = 3.0e38f; code[k][c] = 0u; }

  // ---- up-sampled region of source m into sy (zero outside the image) ---------------------------------------
  auto upsample_from = [&](const float* low, int lw, bool ident, const int* X0, const int* X1, const float* FX,
                           const int* Y0, const int* Y1, const float* FY) {
#pragma unroll
    for (int it = 0; it < kFYIters; ++it) {
      const int i = tid + it * kFThreads;
      if (i < kFRegion) {
        const int ry = i / kFP, rx = i - ry * kFP;
        const int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
        float yv[3] = {0.f, 0.f, 0.f};
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
          if (ident) {
            const float* p = low + ((size_t)gy * lw + gx) * 3;
            yv[0] = __ldg(p); yv[1] = __ldg(p + 1); yv[2] = __ldg(p + 2);
          } else {
            const int x0 = X0[rx] * 3, x1 = X1[rx] * 3;
            const float fx = FX[rx], fy = FY[ry];
            const float* r0 = low + (size_t)Y0[ry] * lw * 3;
            const float* r1 = low + (size_t)Y1[ry] * lw * 3;
            XPT_MS_ASSERT(x0 >= 0 && x1 >= x0 && x1 < lw * 3 && Y0[ry] >= 0 && Y1[ry] >= Y0[ry]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float tl = __ldg(r0 + x0 + c), tr = __ldg(r0 + x1 + c);
              const float bl = __ldg(r1 + x0 + c), br = __ldg(r1 + x1 + c);
              // explicit roundings: both sweeps must reproduce the same value bit for bit (ties compare with ==)
              const float top = __fmaf_rn(__fsub_rn(tr, tl), fx, tl);
              const float bot = __fmaf_rn(__fsub_rn(br, bl), fx, bl);
              yv[c] = __fmaf_rn(__fsub_rn(bot, top), fy, top);
            }
          }
        }
        sy[i] = yv[0]; sy[kFRegion + i] = yv[1]; sy[2 * kFRegion + i] = yv[2];
      }
    }
  };
  auto upsample = [&](int m) {
    upsample_from(m < a.N ? a.synth[l] + ((size_t)b * a.N + m) * h * w * 3
                          : a.stereo[l] + ((size_t)b * a.NS + (m - a.N)) * h * w * 3, w, identity, tX0, tX1, tFX, tY0, tY1, tFY);
  };
  auto upsample_flow = [&](int m) {
    upsample_from(a.cmb_flow + ((size_t)b * a.N + m) * a.cmb_h * a.cmb_w * 3, a.cmb_w, a.cmb_h == H && a.cmb_w == W,
                  fX0, fX1, fFX, fY0, fY1, fFY);
  };

  // ---- the strip's terms of the source in sy.  COEF = false (sweep 1): update the running minimum.  COEF = true
  // (sweep 2): route the upstream gradient to the winners and leave the adjoint coefficients in sA / sB / sC
  // (SSIM: Hh a-terms as in k_fused, summed over the 3x3 window later; L1 / L2: the local derivative in sA). --------
  // MODE 0: sweep 1 of the minimum.  1: sweep 2 of the minimum.  2 (CMB): the flow term of source m -> vmin.
  // 3 (CMB): the static term of source m, kept where it is smaller than the flow term: loss sum and, GRAD, coefficients.
  float lacc[NM];
#pragma unroll
  for (int k = 0; k < NM; ++k) lacc[k] = 0.f;
  auto strip_terms = [&](int m, auto mode_tag) {
    constexpr int MODE = decltype(mode_tag)::value;
    constexpr bool COEF = MODE == 1 || (MODE == 3 && GRAD);
    if (!s_active) return;
    const int mid = (qy + 1) * kFP + q0;
    bool bk0, bk1;              // black (invalid) synthesised pixel: tf.where(mask, 0, loss) -- value 0, no gradient
    {
      const float2 m0a = lds2(sy + mid), m0b = lds2(sy + mid + 2);
      const float2 m1a = lds2(sy + kFRegion + mid), m1b = lds2(sy + kFRegion + mid + 2);
      const float2 m2a = lds2(sy + 2 * kFRegion + mid), m2b = lds2(sy + 2 * kFRegion + mid + 2);
      bk0 = ((m0a.y + m1a.y) + m2a.y) == 0.f;
      bk1 = ((m0b.x + m1b.x) + m2b.x) == 0.f;
    }
    // one method's value at the strip's two pixels: running minimum (sweep 1) or the share of the upstream gradient
    // this source receives (sweep 2; tf.reduce_min splits it equally among the sources attaining the minimum)
    auto update = [&](int k, int c, float v0, float v1) {         // (branch-free)
      const unsigned fresh = ((unsigned)m << 8) | 1u;
      const bool lt0 = v0 < vmin[k][c][0], eq0 = v0 == vmin[k][c][0];
      const bool lt1 = v1 < vmin[k][c][1], eq1 = v1 == vmin[k][c][1];
      unsigned c0_ = code[k][c] & 0xffffu, c1_ = code[k][c] >> 16;
      c0_ = lt0 ? fresh : c0_ + (eq0 ? 1u : 0u);
      c1_ = lt1 ? fresh : c1_ + (eq1 ? 1u : 0u);
      vmin[k][c][0] = lt0 ? v0 : vmin[k][c][0];
      vmin[k][c][1] = lt1 ? v1 : vmin[k][c][1];
      code[k][c] = c0_ | (c1_ << 16);
    };
    auto share = [&](int k, int c, float v0, float v1, bool ok0, bool ok1, float coef, float& g0, float& g1) {
      if constexpr (MODE == 3) {
        // tf.cast(static_loss < flow_loss): a constant mask; the kept terms are summed over the tile's centre
        const bool keep0 = v0 < vmin[k][c][0], keep1 = v1 < vmin[k][c][1];
        if (keep0 && s_centre.x != 0.f) lacc[k] += v0;
        if (keep1 && s_centre.y != 0.f) lacc[k] += v1;
        g0 = (keep0 && ok0 && !bk0 && inv_cnt.x != 0.f) ? coef : 0.f;
        g1 = (keep1 && ok1 && !bk1 && inv_cnt.y != 0.f) ? coef : 0.f;
        return;
      }
      // (opaque per source: otherwise the compiler hoists the shares of all twelve minima out of the source loop and
      // keeps them in local memory -- 45 local loads per source)
      unsigned cd = code[k][c];
      asm volatile("" : "+r"(cd));
      const unsigned c0_ = cd & 0xffffu, c1_ = cd >> 16;
      const int n0 = c0_ & 0xff, n1 = c1_ & 0xff;
      const bool win0 = n0 == 1 ? (int)(c0_ >> 8) == m : v0 == vmin[k][c][0];
      const bool win1 = n1 == 1 ? (int)(c1_ >> 8) == m : v1 == vmin[k][c][1];
      float s0 = coef, s1 = coef;
      if (n0 > 1 || n1 > 1) {       // a real tie is rare: keep the two IEEE divisions (10 % of the kernel's instructions
        s0 = coef / (float)max(n0, 1); s1 = coef / (float)max(n1, 1);     // when evaluated unconditionally) off the common path
      }
      g0 = (win0 && ok0 && !bk0 && inv_cnt.x != 0.f) ? s0 : 0.f;
      g1 = (win1 && ok1 && !bk1 && inv_cnt.y != 0.f) ? s1 : 0.f;
    };
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* py = sy + c * kFRegion + qy * kFP + q0;
      const float* px = sx + c * kFRegion + qy * kFP + q0;
      const int so = c * kFStats + qy * kFP + q0;
      const float2 ya1 = lds2(py + kFP), yb1 = lds2(py + kFP + 2), xa1 = lds2(px + kFP), xb1 = lds2(px + kFP + 2);
      if (do_l1) {
        const float d0 = ya1.y - xa1.y, d1 = yb1.x - xb1.x;
        float v0 = l2 ? __fmul_rn(d0, d0) : fabsf(d0), v1 = l2 ? __fmul_rn(d1, d1) : fabsf(d1);
        if (bk0) v0 = 0.f;
        if (bk1) v1 = 0.f;
        if constexpr (MODE == 0) update(0, c, v0, v1);
        else if constexpr (MODE == 2) { vmin[0][c][0] = v0; vmin[0][c][1] = v1; }
        else {
          float g0, g1;
          share(0, c, v0, v1, true, true, coefL, g0, g1);
          if constexpr (COEF)
            *reinterpret_cast<float2*>(sLp + so) = f2(l2 ? g0 * 2.f * d0 : g0 * sgnf(d0), l2 ? g1 * 2.f * d1 : g1 * sgnf(d1));
        }
      }
      if (ssim) {
        const float2 ya0 = lds2(py), yb0 = lds2(py + 2), ya2 = lds2(py + 2 * kFP), yb2 = lds2(py + 2 * kFP + 2);
        const float2 xa0 = lds2(px), xb0 = lds2(px + 2), xa2 = lds2(px + 2 * kFP), xb2 = lds2(px + 2 * kFP + 2);
        const float2 v1a = f2add(f2add(ya0, ya1), ya2), v1b = f2add(f2add(yb0, yb1), yb2);
        const float2 v2a = f2fma(ya2, ya2, f2fma(ya1, ya1, f2mul(ya0, ya0)));
        const float2 v2b = f2fma(yb2, yb2, f2fma(yb1, yb1, f2mul(yb0, yb0)));
        const float2 v3a = f2fma(xa2, ya2, f2fma(xa1, ya1, f2mul(xa0, ya0)));
        const float2 v3b = f2fma(xb2, yb2, f2fma(xb1, yb1, f2mul(xb0, yb0)));
        const float2 s1 = hsum3(v1a, v1b), s2 = hsum3(v2a, v2b), s3 = hsum3(v3a, v3b);
        // the target's window statistics from the rows just loaded: mu_x, mu_x^2 + c1, sigma_x + c2
        const float2 w1a = f2add(f2add(xa0, xa1), xa2), w1b = f2add(f2add(xb0, xb1), xb2);
        const float2 w2a = f2fma(xa2, xa2, f2fma(xa1, xa1, f2mul(xa0, xa0)));
        const float2 w2b = f2fma(xb2, xb2, f2fma(xb1, xb1, f2mul(xb0, xb0)));
        const float2 mux = f2mul(hsum3(w1a, w1b), inv_cnt);
        const float2 mux2 = f2mul(mux, mux);
        const float2 mux2c = f2add(mux2, f2s(kC1));
        const float2 sgxc = f2add(f2fma(hsum3(w2a, w2b), inv_cnt, f2neg(mux2)), f2s(kC2));
        const float2 muy = f2mul(s1, inv_cnt);
        const float2 muy2 = f2mul(muy, muy);
        const float2 mxy = f2mul(mux, muy);
        const float2 sgy = f2fma(s2, inv_cnt, f2neg(muy2));
        const float2 sgxy = f2fma(s3, inv_cnt, f2neg(mxy));
        const float2 a1 = f2fma(f2s(2.f), mxy, f2s(kC1));
        const float2 a2 = f2fma(f2s(2.f), sgxy, f2s(kC2));
        const float2 b1 = f2add(mux2c, muy2);
        const float2 b2 = f2add(sgxc, sgy);
        const float2 den = f2mul(b1, b2);
        const float2 r12 = f2(rcp_nr(den.x), rcp_nr(den.y));
        const float2 ssv = f2mul(f2mul(a1, a2), r12);
        const float2 lv = f2fma(f2s(-0.5f), ssv, f2s(0.5f));
        float v0 = fminf(fmaxf(lv.x, 0.f), 1.f), v1 = fminf(fmaxf(lv.y, 0.f), 1.f);
        const bool pass0 = v0 == lv.x, pass1 = v1 == lv.y;       // clip_by_value passes the gradient inside [0,1]
        if (bk0) v0 = 0.f;
        if (bk1) v1 = 0.f;
        if constexpr (MODE == 0) update(iS, c, v0, v1);
        else if constexpr (MODE == 2) { vmin[iS][c][0] = v0; vmin[iS][c][1] = v1; }
        else {
          float g0, g1;
          share(iS, c, v0, v1, pass0, pass1, coefS, g0, g1);
          if constexpr (COEF) {
            // h = dL/d ssim = -g/2; Hh = 2 h / (#taps b1 b2); A, 2B, C as in k_fused
            const float2 Hh = f2mul(f2(-g0 * inv_cnt.x, -g1 * inv_cnt.y), r12);
            const float2 Hs = f2mul(Hh, ssv);
            const float2 t1 = f2mul(mux, f2add(a2, f2neg(a1)));
            const float2 t2 = f2mul(muy, f2add(b2, f2neg(b1)));
            *reinterpret_cast<float2*>(sA + so) = f2fma(Hh, t1, f2neg(f2mul(Hs, t2)));
            *reinterpret_cast<float2*>(sB + so) = f2mul(f2neg(Hs), b1);
            *reinterpret_cast<float2*>(sC + so) = f2mul(Hh, a1);
          }
        }
      }
    }
  };

  // ---- adjoint of one source: dL/dS on the centre strip from the coefficient planes, then (levels below full
  // resolution) the up-sampling adjoint reduced inside the tile ----------------------------------------------------
  int cw = 0, ch = 0, X_lo = 0, FW = 1, Y_lo = 0, FHt = 1;
  float inv_FW = 0.f, inv_chFW = 0.f, inv_FHFW = 0.f;
  if constexpr (GRAD) {
    // low-resolution footprint of the tile's centre and, per footprint column / row, the tile columns / rows feeding it
    cw = min(kFCW, W - tx0); ch = min(kFCH, H - ty0);        // in-image centre extent
    X_lo = tX0[2]; FW = tX1[2 + cw - 1] - X_lo + 1;
    Y_lo = tY0[2]; FHt = min(tY1[2 + ch - 1] - Y_lo + 1, 16);
    if (!identity) {
      auto ranges = [](const int* lo, const int* hi, int n, int V, int& ra, int& rb) {
        int a0 = 1, a1 = 0, b0 = 1, b1 = 0;
        bool fa = false, fb = false;
        for (int o = 0; o < n; ++o) {
          if (lo[o + 2] == V) { if (!fa) { a0 = o; fa = true; } a1 = o; }
          if (hi[o + 2] == V) { if (!fb) { b0 = o; fb = true; } b1 = o; }
        }
        ra = a0 | (a1 << 8); rb = b0 | (b1 << 8);
      };
      if (tid < FW) ranges(tX0, tX1, cw, X_lo + tid, rXa[tid], rXb[tid]);
      else if (tid >= 64 && tid < 64 + FHt) ranges(tY0, tY1, ch, Y_lo + (tid - 64), rYa[tid - 64], rYb[tid - 64]);
      // (visible to the reduction passes behind the barriers of the first source)
    }
    XPT_MS_ASSERT(FW >= 1 && (identity || FW <= kMSTPitch) && FHt >= 1 && FHt <= 16 && cw >= 1 && ch >= 1 && ch <= kFCH);
    inv_FW = 1.f / (float)FW; inv_chFW = 1.f / (float)(ch * FW); inv_FHFW = 1.f / (float)(FHt * FW);
  }
  auto adjoint = [&](int m) {
    float* glow = m < a.N ? a.gsynth[l] + ((size_t)b * a.N + m) * h * w * 3
                          : a.gstereo[l] + ((size_t)b * a.NS + (m - a.N)) * h * w * 3;
    if (g_active) {           // warps 0..12 own one centre row each
      const int rrow = (cyy + 2) * kFP + c0 + 2;
      float2 g[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float2 gc = f2s(0.f);
        if (ssim) {
          const float2 yv = lds2(sy + c * kFRegion + rrow), xv = lds2(sx + c * kFRegion + rrow);
          const int so = c * kFStats + cyy * kFP + c0;
          float2 va, vb;
          va = f2add(f2add(lds2(sA + so), lds2(sA + so + kFP)), lds2(sA + so + 2 * kFP));
          vb = f2add(f2add(lds2(sA + so + 2), lds2(sA + so + kFP + 2)), lds2(sA + so + 2 * kFP + 2));
          const float2 sa = hsum3(va, vb);
          va = f2add(f2add(lds2(sB + so), lds2(sB + so + kFP)), lds2(sB + so + 2 * kFP));
          vb = f2add(f2add(lds2(sB + so + 2), lds2(sB + so + kFP + 2)), lds2(sB + so + 2 * kFP + 2));
          const float2 sb = hsum3(va, vb);
          va = f2add(f2add(lds2(sC + so), lds2(sC + so + kFP)), lds2(sC + so + 2 * kFP));
          vb = f2add(f2add(lds2(sC + so + 2), lds2(sC + so + kFP + 2)), lds2(sC + so + 2 * kFP + 2));
          const float2 sc = hsum3(va, vb);
          gc = f2fma(yv, sb, f2fma(xv, sc, sa));               // sB holds 2 dL/dP(y^2)
        }
        if (do_l1) {
          const int so = c * kFStats + (cyy + 1) * kFP + c0 + 1;
          gc = f2add(gc, f2(sLp[so], sLp[so + 1]));
        }
        g[c] = gc;
      }
      const int gy = ty0 + cyy;
      const bool in0 = tx0 + c0 < W, in1 = tx0 + c0 + 1 < W;     // (gy < H: g_active)
      if (identity) {
        float* o = glow + ((size_t)gy * w + tx0 + c0) * 3;
        if (in0) { o[0] = g[0].x; o[1] = g[1].x; o[2] = g[2].x; }
        if (in1) { o[3] = g[0].y; o[4] = g[1].y; o[5] = g[2].y; }
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          *reinterpret_cast<float2*>(sG + c * kFCentre + cyy * kFCP + c0) = f2(in0 ? g[c].x : 0.f, in1 ? g[c].y : 0.f);
      }
    }
    if (!identity) {
      __syncthreads();
      // adjoint of the up-sampling, horizontal pass: T[c][r][X] = sum over the tile's columns o of wx(o, X) g[c][r][o]
      for (int j = tid; j < 3 * ch * FW; j += kFThreads) {
        const int c = (int)(((float)j + 0.5f) * inv_chFW), k = j - c * ch * FW;      // exact for these small integers
        const int r = (int)(((float)k + 0.5f) * inv_FW), Xi = k - r * FW;
        const int pa = rXa[Xi], pb = rXb[Xi];
        XPT_MS_ASSERT(c >= 0 && c < 3 && r >= 0 && r < ch && Xi >= 0 && Xi < FW && (pa >> 8) < cw && (pb >> 8) < cw);
        const float* gr = sG + c * kFCentre + r * kFCP;
        float v = 0.f;
        for (int o = pa & 0xff; o <= (pa >> 8); ++o) v = fmaf(1.f - tFX[o + 2], gr[o], v);
        for (int o = pb & 0xff; o <= (pb >> 8); ++o) v = fmaf(tFX[o + 2], gr[o], v);
        sT[(c * kFCH + r) * kMSTPitch + Xi] = v;
      }
      __syncthreads();
      // vertical pass and ONE atomic per low-resolution value of the footprint
      for (int j = tid; j < 3 * FHt * FW; j += kFThreads) {
        const int c = (int)(((float)j + 0.5f) * inv_FHFW), k = j - c * FHt * FW;
        const int Yi = (int)(((float)k + 0.5f) * inv_FW), Xi = k - Yi * FW;
        const int pa = rYa[Yi], pb = rYb[Yi];
        XPT_MS_ASSERT(c >= 0 && c < 3 && Yi >= 0 && Yi < FHt && Xi >= 0 && Xi < FW && (pa >> 8) < ch && (pb >> 8) < ch);
        XPT_MS_ASSERT(Y_lo + Yi < h && X_lo + Xi < w);
        const float* tc = sT + c * kFCH * kMSTPitch + Xi;
        float v = 0.f;
        for (int r = pa & 0xff; r <= (pa >> 8); ++r) v = fmaf(1.f - tFY[r + 2], tc[r * kMSTPitch], v);
        for (int r = pb & 0xff; r <= (pb >> 8); ++r) v = fmaf(tFY[r + 2], tc[r * kMSTPitch], v);
        if (v != 0.f) atomicAdd(glow + ((size_t)(Y_lo + Yi) * w + (X_lo + Xi)) * 3 + c, v);
      }
    } else {
      __syncthreads();      // the next source's up-sampling rewrites sy
    }
  };

  auto reduce_loss = [&](float (&lsum)[NM]) {
#pragma unroll
    for (int k = 0; k < NM; ++k) {
      lsum[k] = warp_sum(lsum[k]);
      if (lane == 0) red[k * 16 + wid] = lsum[k];
    }
    __syncthreads();
    if (tid < NM) {
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < kFThreads / 32; ++k) v += red[tid * 16 + k];
      (tid == 0 ? a.loss_part : a.loss_part2)[(size_t)b * a.S * a.tiles + blockIdx.y] = v * a.norm[l];
    }
  };

  if constexpr (CMB) {
    for (int m = 0; m < nsrc; ++m) {
      upsample_flow(m);
      __syncthreads();
      strip_terms(m, std::integral_constant<int, 2>{});
      __syncthreads();
      upsample(m);
      __syncthreads();
      strip_terms(m, std::integral_constant<int, 3>{});
      __syncthreads();
      if constexpr (GRAD) adjoint(m);
    }
    reduce_loss(lacc);
  } else {
    // ---- sweep 1: minimum, tie count and first winner per (pixel, channel) --------------------------------------
    for (int m = 0; m < nsrc; ++m) {
      upsample(m);
      __syncthreads();
      strip_terms(m, std::integral_constant<int, 0>{});
      __syncthreads();
    }
    // loss of this tile: sum of the minima over its in-image centre pixels
    {
      float lsum[NM];
#pragma unroll
      for (int k = 0; k < NM; ++k) {
        lsum[k] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (s_centre.x != 0.f) lsum[k] += vmin[k][c][0];
          if (s_centre.y != 0.f) lsum[k] += vmin[k][c][1];
        }
      }
      reduce_loss(lsum);
    }

    // ---- sweep 2: gradient of the winners ----------------------------------------------------------------------
    // (sources in reverse order: the up-sampled tile of the last source of sweep 1 is still in sy)
    if constexpr (GRAD) {
      __syncthreads();          // reduce_loss's scratch reads are done; sy of source nsrc-1 is complete since sweep 1's barrier
      for (int m = nsrc - 1; m >= 0; --m) {
        if (m != nsrc - 1) {
          upsample(m);
          __syncthreads();
        }
        strip_terms(m, std::integral_constant<int, 1>{});
        __syncthreads();
        adjoint(m);
      }
    }
  }
}

}  // namespace xpt
