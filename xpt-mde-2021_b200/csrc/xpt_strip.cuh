// xpt_strip.cuh -- the streaming strip kernel of the total-loss path (north-star kernels 1+2+3+4), round 2.
//
// Same arithmetic per sample as k_fused (xpt_fused.cuh), different machine mapping:
//
//   * work item ("piece") = rows [ya, yb) of one 64-column strip (60 centre columns + halo 2) of one
//     (snippet, level).  A CTA owns a static, equal-cost list of pieces and MARCHES DOWN their rows; nothing is
//     warped twice vertically (k_fused re-warped 39 % halo samples per 64x13 tile) and the list is cut so that every
//     CTA of the grid carries the same number of row chunks (no partial last wave);
//   * the CTA is a software pipeline of warp ROLES, one 32-lane warp per role instance, lane = 2 adjacent columns:
//       L  loader      target rows (3 planes) + depth row -> shared-memory ring
//       Y  warp        one per source: projection in packed FP32 (FFMA2, both columns per instruction), 2 x 4
//                      16-byte gathers, bilinear value + Jacobian -> rings
//       S  statistics  one per (source, channel): 3x3 window sums through warp shuffles (horizontal) and register
//                      windows (vertical) -- no overlapping shared-memory re-reads --, SSIM + L1, the adjoint
//                      coefficients, THEIR 3x3 box sums the same way, and dL/dS of one channel
//       G  adjoint     one per source: contracts dL/dS with the cached Jacobian, projection adjoint, pose sums kept
//                      in registers for the whole piece (one warp reduction per piece instead of one per tile row)
//       O  output      sums dL/ddepth over the sources in a fixed order, edge-aware smoothness forward + backward
//     role k works on row chunk (tick - lag_k); one CTA barrier per tick of 2 rows separates producers from
//     consumers, every ring is indexed by the global row number, so the schedule is static and deadlock-free;
//   * camera geometry comes from the ctx's own global buffers (geoK / geoT written by the pyramid launch), staged
//     per piece into a few shared-memory words: no process-wide __constant__ block, no upload, no batch chunking.
//
// Citations: SURVEY.md Appendix A (reference synthesize_base.py:66-178, bilinear_interp.py:34-147,
// loss_util.py:6-96, losses.py:147-154,386-440).
#pragma once
#include "xpt_fused.cuh"

namespace xpt {

constexpr int kSW = 64;                 // region columns of a strip (32 lanes x 2)
constexpr int kSCWMax = 60;             // centre columns (halo 2 on both sides)
constexpr int kSTRows = 16;             // ring rows: target/depth/disparity (L -> Y, S, O)
constexpr int kSTPlanes = 5;            // x0 x1 x2 depth disparity
constexpr int kSYRows = 8;              //            warped value (Y -> S; S re-reads the two rows before)
constexpr int kSJRows = 8;              //            Jacobian     (Y -> G)
constexpr int kSGRows = 4;              //            dL/dS        (S -> G)
constexpr int kSDRows = 4;              //            dL/ddepth    (G -> O)
constexpr int kSJPlanes = 10;           // gu[3], gv[3], u, v, 1/den, D
constexpr int kSLagY = 1, kSLagB = 2, kSLagX = 2, kSLagS = 3, kSLagG = 4, kSLagO = 5;   // ticks behind the loader
constexpr int kSXRows = 4, kSXPlanes = 11;   // target-statistics ring (X -> S): mux[3], mux^2+c1 [3], sigma_x+c2 [3], 1/#taps, centre flags
constexpr int kSLagEnd = kSLagO + 1;    // + one tick in which the output stage writes the last piece's loss record

struct StripPiece { int b, l, x0, cw, ya, yb, slot, nch; };     // nch = ceil((yb - ya + 4) / 2) row chunks
struct StripCta { int first, count, chunks, pad; };

struct StripArgs {
  LevelTable lt;
  int B, N, S;
  const StripPiece* pieces;
  const StripCta* ctas;
  const float* geoK;                   // [B][S][18]
  const float* geoT;                   // [B][N][12]
  const float* depth[kMaxScales];
  const float* disp[kMaxScales];
  const float4* src4[kMaxScales];      // RGBx source levels [B,N,h,w]
  int do_l1, do_ssim, do_smooth;
  int logit;                           // depth[] holds logits (XPT_FLAG_DEPTH_LOGIT)
  float norm_photo[kMaxScales];
  float norm_sm_x[kMaxScales];
  float norm_sm_y[kMaxScales];
  float grad_factor;
  float gcoef_l1, gcoef_ssim, gcoef_smooth;
  float* loss_part; int slots_per_b;   // [B][slots][3]
  float* pose_part;                    // [B][slots][N][12]
  float* d_depth[kMaxScales];
  float* d_disp[kMaxScales];
};

template <int NS>
struct StripSmem {
  static constexpr int T = 0;                                    // [5][16][64]: x0 x1 x2 depth disparity
  static constexpr int kY = 0;                                   // per source: [4][8][64] y0 y1 y2 notblack
  static constexpr int kJ = kY + 4 * kSYRows * kSW;              //             [10][8][64]
  static constexpr int kG = kJ + kSJPlanes * kSJRows * kSW;      //             [3][4][64]
  static constexpr int kD = kG + 3 * kSGRows * kSW;              //             [4][64]
  static constexpr int kSrc = kD + kSDRows * kSW;                // floats per source
  static constexpr int src0 = T + kSTPlanes * kSTRows * kSW;
  static constexpr int X = src0 + NS * kSrc;                     // [11][4][64] target statistics of a row (X -> S)
  static constexpr int geo = X + kSXPlanes * kSXRows * kSW;      // [NS][32]: K rows 0-1, inv K rows 0-1, [R|t]
  static constexpr int geoG = geo + NS * 32;                     // [NS][48]: [R|t] of a G warp's source, K, inv K
  static constexpr int loss = geoG + NS * 48;                    // [2][3 NS][2]
  static constexpr int lossO = loss + 2 * 3 * NS * 2;            // [2][2]: smoothness sums of the two output warps
  static constexpr int stage = (lossO + 4 + 3) & ~3;             // [NS][16 taps + 8 aux][32 lanes] float4: Y's gathers (cp.async)
  static constexpr int kStage = 24 * 32 * 4;                     //   and the per-sample weights / coordinates kept across the tick
  static constexpr int kFloats = stage + NS * kStage;
  static constexpr size_t kBytes = sizeof(float) * kFloats;
  static constexpr int kWarps = 4 + NS + NS + 3 * NS;   // L X O0 O1 | G | Y | S
  static constexpr int kThreads = 32 * kWarps;
};

// XPT_STRIP_PROF (profiles/strip_roles_time.py only): every warp adds up the cycles between leaving a tick barrier and
// arriving at the next one = its busy time per tick; the slowest role is the critical path of the pipeline
#ifdef XPT_STRIP_PROF
__device__ long long g_strip_busy[148 * 8 * 32];
__device__ __forceinline__ long long strip_clock() { long long t; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)); return t; }
struct StripTimer {
  long long busy, t0;
  __device__ __forceinline__ StripTimer() : busy(0), t0(strip_clock()) {}
  __device__ __forceinline__ ~StripTimer() { if ((threadIdx.x & 31) == 0 && blockIdx.x < 148 * 8) g_strip_busy[blockIdx.x * 32 + (threadIdx.x >> 5)] = busy; }
};
#define STRIP_TIMER() StripTimer strip_timer_
// BAR.SYNC does not block at issue (the warp stalls at the next shared-memory access): touch shared memory after the
// barrier and make the clock read depend on that load, so t0 is the release time
__device__ __forceinline__ long long strip_clock_after_smem() {
  unsigned v; long long t;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(0u) : "memory");
  asm volatile("{ .reg .u64 c; mov.u64 c, %%clock64; and.b32 %1, %1, 0; cvt.u64.u32 %0, %1; add.u64 %0, %0, c; }" : "=l"(t), "+r"(v));
  return t;
}
#define strip_bar() do { strip_timer_.busy += strip_clock() - strip_timer_.t0; asm volatile("bar.sync 0;" ::: "memory"); strip_timer_.t0 = strip_clock_after_smem(); } while (0)
#else
#define STRIP_TIMER() ((void)0)
__device__ __forceinline__ void strip_bar() { asm volatile("bar.sync 0;" ::: "memory"); }
#endif
__device__ __forceinline__ void sts2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

__device__ __forceinline__ StripPiece load_piece(const StripPiece* __restrict__ tab, int i) {
  const int4* q = reinterpret_cast<const int4*>(tab + i);
  const int4 a = __ldg(q), b = __ldg(q + 1);
  StripPiece p;
  p.b = a.x; p.l = a.y; p.x0 = a.z; p.cw = a.w; p.ya = b.x; p.yb = b.y; p.slot = b.z; p.nch = b.w;
  return p;
}

// pixel rays of a column pair, the operation order of k_fused: fma(k0, x, k1*y) + k2
__device__ __forceinline__ float2 ray_pair(float k0, float k1, float k2, float2 fx, float fy) {
  return f2add(f2fma(f2s(k0), fx, f2s(k1 * fy)), f2s(k2));
}

// 1 / #in-image taps of the 3x3 window of a pixel whose row has cy and whose column has cx in-image neighbours
__device__ __forceinline__ float strip_box_inv(int cy, int cx) { return (cy > 0 && cx > 0) ? box_inv(cy * cx) : 0.f; }
__device__ __forceinline__ int strip_cnt(int g, int n) {      // in-image members of {g-1, g, g+1}; 0 when g is outside
  return ((unsigned)g < (unsigned)n) ? (min(g + 1, n - 1) - max(g - 1, 0) + 1) : 0;
}

// ---------------------------------------------------------------------------------------------------------
// L: loader of the target / depth rows.  The rows of chunk t+1 are requested while chunk t is being stored, so the
// HBM latency of a row sits behind one whole tick.
// ---------------------------------------------------------------------------------------------------------
struct StripRow { float v[6]; float d[2]; float e[2]; };     // two adjacent pixels: (A.c0 A.c1 A.c2 B.c0 B.c1 B.c2), depth, disparity

__device__ __forceinline__ StripRow strip_load_row(const float* __restrict__ tg, const float* __restrict__ dp,
                                                   const float* __restrict__ ep, int H, int W, int gy, int gx) {
  StripRow r;
#pragma unroll
  for (int k = 0; k < 6; ++k) r.v[k] = 0.f;
  r.d[0] = r.d[1] = r.e[0] = r.e[1] = 0.f;
  if ((unsigned)gy < (unsigned)H) {
    const long long o = (long long)gy * W + gx;
    if ((unsigned)gx < (unsigned)W) {
      r.v[0] = __ldg(tg + o * 3); r.v[1] = __ldg(tg + o * 3 + 1); r.v[2] = __ldg(tg + o * 3 + 2);
      r.d[0] = __ldg(dp + o);
      if (ep) r.e[0] = __ldg(ep + o);
    }
    if ((unsigned)(gx + 1) < (unsigned)W) {
      r.v[3] = __ldg(tg + o * 3 + 3); r.v[4] = __ldg(tg + o * 3 + 4); r.v[5] = __ldg(tg + o * 3 + 5);
      r.d[1] = __ldg(dp + o + 1);
      if (ep) r.e[1] = __ldg(ep + o + 1);
    }
  }
  return r;
}

template <int NS>
__device__ __forceinline__ void strip_role_l(const StripArgs& a, float* smem, const int lane, const StripCta cta) {
  using SM = StripSmem<NS>;
  float* const T = smem + SM::T;
  const int total = cta.chunks, pend = cta.first + cta.count;
  int pi = cta.first, ci = 0;
  StripPiece p = load_piece(a.pieces, pi);
  StripRow nx[2];
  auto fetch = [&]() {            // rows of chunk (pi, ci)
    const Level& L = a.lt.lv[p.l];
    const float* const tg = L.tgt + (long long)p.b * L.tgt_bs;
    const float* const dp = a.depth[p.l] + (long long)p.b * L.H * L.W;
    const float* const ep = (a.do_smooth && a.disp[p.l]) ? a.disp[p.l] + (long long)p.b * L.H * L.W : nullptr;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      nx[r] = strip_load_row(tg, dp, ep, L.H, L.W, p.ya - 2 + 2 * ci + r, p.x0 - 2 + 2 * lane);
      if (a.logit) {           // InverseSigmoidActivation at load; pixels outside the image keep depth 0 = "no sample"
        const int gy = p.ya - 2 + 2 * ci + r, gx = p.x0 - 2 + 2 * lane;
        const bool in0 = (unsigned)gy < (unsigned)L.H && (unsigned)gx < (unsigned)L.W;
        const bool in1 = (unsigned)gy < (unsigned)L.H && (unsigned)(gx + 1) < (unsigned)L.W;
        nx[r].d[0] = in0 ? depth_of_logit(nx[r].d[0]) : 0.f;
        nx[r].d[1] = in1 ? depth_of_logit(nx[r].d[1]) : 0.f;
      }
    }
    if (++ci == p.nch) { ci = 0; if (++pi < pend) p = load_piece(a.pieces, pi); }
  };
  fetch();
  STRIP_TIMER();
  for (int t = 0; t < total + kSLagEnd; ++t) {
    if (t < total) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float* const q = T + ((2 * t + r) & (kSTRows - 1)) * kSW + 2 * lane;
        sts2(q, f2(nx[r].v[0], nx[r].v[3])); sts2(q + kSTRows * kSW, f2(nx[r].v[1], nx[r].v[4]));
        sts2(q + 2 * kSTRows * kSW, f2(nx[r].v[2], nx[r].v[5])); sts2(q + 3 * kSTRows * kSW, f2(nx[r].d[0], nx[r].d[1]));
        sts2(q + 4 * kSTRows * kSW, f2(nx[r].e[0], nx[r].e[1]));
      }
      if (t + 1 < total) fetch();
    }
    strip_bar();
  }
}

// ---------------------------------------------------------------------------------------------------------
// O: output stage -- dL/ddepth summed over the sources in a fixed order, edge-aware smoothness (losses.py:409-440)
// forward + backward on rows walked top to bottom, loss record of the piece.  Rows are requested one tick ahead.
// ---------------------------------------------------------------------------------------------------------
template <int NS, bool DERIVE>
__device__ __forceinline__ void strip_role_o(const StripArgs& a, float* smem, const int lane, const int k,
                                             const StripCta cta) {
  using SM = StripSmem<NS>;
  const float* const T = smem + SM::T + 2 * lane;
  const int total = cta.chunks, pend = cta.first + cta.count;
  int opi = cta.first, oci = 0;
  StripPiece op = load_piece(a.pieces, opi);
  float lsum_sm = 0.f;
  // the loss record of a finished piece is written one tick later by warp 0 (both warps' sums are in shared memory then)
  int pend_b = -1, pend_slot = 0, pend_l = 0, pend_par = 0;
  STRIP_TIMER();
  for (int t = 0; t < total + kSLagEnd; ++t) {
    const int co = t - kSLagO;
    if (k == 0 && pend_b >= 0) {
      if (lane == 0) {
        const float* lb = smem + SM::loss + pend_par * (3 * NS * 2);
        float l1 = 0.f, ss = 0.f;
        for (int q = 0; q < 3 * NS; ++q)
          if (q < 3 * a.N) { l1 += lb[2 * q]; ss += lb[2 * q + 1]; }
        const float* lo = smem + SM::lossO + pend_par * 2;
        float* out = a.loss_part + ((size_t)pend_b * a.slots_per_b + pend_slot) * 3;
        out[0] = l1 * a.norm_photo[pend_l]; out[1] = ss * a.norm_photo[pend_l]; out[2] = lo[0] + lo[1];
      }
      pend_b = -1;
    }
    if (co >= 0 && co < total) {
      const Level& L = a.lt.lv[op.l];
      const int H = L.H, W = L.W, Lr = op.yb - op.ya;
      const int gx = op.x0 - 2 + 2 * lane;
      const bool cen0 = 2 * lane >= 2 && 2 * lane < 2 + op.cw && gx < W;
      const bool cen1 = 2 * lane + 1 >= 2 && 2 * lane + 1 < 2 + op.cw && gx + 1 < W;
      const int s = 2 * co + k, sg = 2 * oci + k;
      const int gy = op.ya + sg - 4;                         // row whose outputs are due at this step
      if (sg >= 4 && sg < Lr + 4) {
        float gd0 = 0.f, gd1 = 0.f, z0 = 0.f, z1 = 0.f;      // dL/ddisp of the smoothness term; disparity of the row
        if (a.do_smooth) {
          const float k3 = a.grad_factor;
          const float nx_ = a.norm_sm_x[op.l], ny_ = a.norm_sm_y[op.l];
          const float gcx = a.gcoef_smooth * nx_, gcy = a.gcoef_smooth * ny_;
          // rows gy-1, gy, gy+1 (region rows sg-3 .. sg-1) were staged by the loader; zero outside the image
          float tr[3][6], dr[3][2];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float* q = T + ((s - 3 + j) & (kSTRows - 1)) * kSW;
            const float2 c0 = lds2(q), c1 = lds2(q + kSTRows * kSW), c2 = lds2(q + 2 * kSTRows * kSW);
            const float2 dd = lds2(q + (DERIVE ? 3 : 4) * kSTRows * kSW);
            tr[j][0] = c0.x; tr[j][1] = c1.x; tr[j][2] = c2.x; tr[j][3] = c0.y; tr[j][4] = c1.y; tr[j][5] = c2.y;
            dr[j][0] = dd.x; dr[j][1] = dd.y;
            if (DERIVE) {
#pragma unroll
              for (int i = 0; i < 2; ++i) dr[j][i] = dr[j][i] > 0.00001f ? __frcp_rn(dr[j][i]) : 0.f;   // safe_reciprocal_number
            }
          }
          const bool up_in = gy >= 1, dn_in = gy + 1 < H;
          // vertical pairs (gy-1, gy) and (gy, gy+1) of both columns
          float ty_up[2] = {0.f, 0.f}, ty[2] = {0.f, 0.f}, sdy[2] = {0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (up_in) {
              float e = 0.f;
#pragma unroll
              for (int c = 0; c < 3; ++c) e += fabsf((tr[0][3 * i + c] - tr[1][3 * i + c]) * k3);
              const float w = expf(-(e * (1.f / 3.f)));
              const float sd = (dr[0][i] - dr[1][i]) * w;
              ty_up[i] = gcy * sgnf(sd) * w;
            }
            if (dn_in) {
              float e = 0.f;
#pragma unroll
              for (int c = 0; c < 3; ++c) e += fabsf((tr[1][3 * i + c] - tr[2][3 * i + c]) * k3);
              const float w = expf(-(e * (1.f / 3.f)));
              const float sd = (dr[1][i] - dr[2][i]) * w;
              sdy[i] = sd;
              ty[i] = gcy * sgnf(sd) * w;
            }
          }
          // horizontal pairs of row gy: (A, B) inside the lane, (B, right lane's A) across lanes
          float rt[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) rt[c] = __shfl_down_sync(0xffffffffu, tr[1][c], 1);
          const float rd = __shfl_down_sync(0xffffffffu, dr[1][0], 1);
          float txA = 0.f, txB = 0.f, sdxA = 0.f, sdxB = 0.f;
          if ((unsigned)gx < (unsigned)W && gx + 1 < W) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) e += fabsf((tr[1][c] - tr[1][3 + c]) * k3);
            const float w = expf(-(e * (1.f / 3.f)));
            sdxA = (dr[1][0] - dr[1][1]) * w;
            txA = gcx * sgnf(sdxA) * w;
          }
          if (gx + 1 >= 0 && gx + 2 < W && lane < 31) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) e += fabsf((tr[1][3 + c] - rt[c]) * k3);
            const float w = expf(-(e * (1.f / 3.f)));
            sdxB = (dr[1][1] - rd) * w;
            txB = gcx * sgnf(sdxB) * w;
          }
          const float txL = __shfl_up_sync(0xffffffffu, txB, 1);       // pair (left lane's B, A)
          // forward terms of the pairs a pixel starts, gradient of all four pairs it is part of (losses.py:409-440)
          if (cen0) lsum_sm += fabsf(sdxA) * nx_ + fabsf(sdy[0]) * ny_;
          if (cen1) lsum_sm += fabsf(sdxB) * nx_ + fabsf(sdy[1]) * ny_;
          gd0 = txA; gd0 += ty[0]; gd0 -= (lane > 0 ? txL : 0.f); gd0 -= ty_up[0];
          gd1 = txB; gd1 += ty[1]; gd1 -= txA; gd1 -= ty_up[1];
          z0 = dr[1][0]; z1 = dr[1][1];
        }
        float g0 = 0.f, g1 = 0.f;
        if (DERIVE && a.do_smooth) { g0 = -(gd0 * z0) * z0; g1 = -(gd1 * z1) * z1; }   // d disp / d depth = -disp^2
#pragma unroll
        for (int n = 0; n < NS; ++n)
          if (n < a.N) {
            const float2 v = lds2(smem + SM::src0 + n * SM::kSrc + SM::kD + ((s - 2) & (kSDRows - 1)) * kSW + 2 * lane);
            g0 += v.x; g1 += v.y;
          }
        const long long o = (long long)op.b * H * W + (long long)gy * W + gx;
        if (a.logit) {           // dL/dlogit = dL/ddepth * d depth / d logit
          const float2 D = lds2(T + 3 * kSTRows * kSW + ((s - 2) & (kSTRows - 1)) * kSW);
          g0 *= ddepth_dlogit(D.x); g1 *= ddepth_dlogit(D.y);
        }
        if (a.d_depth[op.l]) {
          if (cen0) a.d_depth[op.l][o] = g0;
          if (cen1) a.d_depth[op.l][o + 1] = g1;
        }
        if (!DERIVE && a.do_smooth && a.d_disp[op.l]) {
          if (cen0) a.d_disp[op.l][o] = gd0;
          if (cen1) a.d_disp[op.l][o + 1] = gd1;
        }
      }
      if (++oci == op.nch) {
        // piece done: this warp's smoothness sum -> shared memory; the record is written next tick
        const float sm = warp_sum(lsum_sm);
        if (lane == 0) smem[SM::lossO + (opi & 1) * 2 + k] = sm;
        pend_b = op.b; pend_slot = op.slot; pend_l = op.l; pend_par = opi & 1;
        lsum_sm = 0.f; oci = 0;
        if (++opi < pend) op = load_piece(a.pieces, opi);
      }
    }
    strip_bar();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Y: inverse warp of one source, two region rows per tick, two adjacent columns per lane.  The sixteen 16-byte
// gathers of a tick go global -> shared with cp.async (no registers in flight), both rows at once.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(float* dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int NS>
__device__ __forceinline__ void strip_role_y(const StripArgs& a, float* smem, const int lane, const int n,
                                             const StripCta cta) {
  using SM = StripSmem<NS>;
  const float* const T = smem + SM::T;
  float* const Yr = smem + SM::src0 + n * SM::kSrc + SM::kY;
  float* const Jr = smem + SM::src0 + n * SM::kSrc + SM::kJ;
  float* const geo = smem + SM::geo + n * 32;
  float* const stg = smem + SM::stage + n * SM::kStage + lane * 4;     // [16 taps + 8 aux][lane] float4
  const int total = cta.chunks, pend = cta.first + cta.count;
  const bool live = n < a.N;
  int pi = cta.first, ci = 0;
  StripPiece p = load_piece(a.pieces, pi);
  bool fresh = true;

  STRIP_TIMER();
  for (int t = 0; t < total + kSLagEnd; ++t) {
    const int cb = t - kSLagB, c = t - kSLagY;
    const bool doB = live && cb >= 0 && cb < total, doA = live && c >= 0 && c < total;
    const Level& L = a.lt.lv[p.l];
    const int H = L.H, W = L.W;
    if (doA && fresh) {
      // camera geometry of this piece -> 24 shared-memory words of this warp (K rows 0-1, inv K rows 0-1, [R|t])
      __syncwarp();
      if (lane < 6) geo[lane] = __ldg(a.geoK + (size_t)(p.b * a.S + p.l) * kGeoK + lane);
      else if (lane < 12) geo[lane] = __ldg(a.geoK + (size_t)(p.b * a.S + p.l) * kGeoK + 9 + (lane - 6));
      else if (lane < 24) geo[lane] = __ldg(a.geoT + (size_t)(p.b * a.N + n) * kGeoT + (lane - 12));
      __syncwarp();
      fresh = false;
    }
    const float4* const img4 = a.src4[p.l] + (size_t)(p.b * a.N + n) * H * W;
    const float fx0 = (float)(p.x0 - 2 + 2 * lane);
    const float2 fx = f2(fx0, fx0 + 1.f);
    const float wlim = (float)(W - 2), hlim = (float)(H - 2);
    // Row by row: consume the gathers of row r of chunk t-2 (phase B), then request the gathers of row r of chunk
    // t-1 into the half of the staging buffer that was just read (phase A): every gather has about half a tick to land.
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      // ---- phase B: bilinear value + Jacobian of two samples -> rings -------------------------------------------
      if (doB) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // all but the most recent request have landed
        const int s = 2 * cb + r;
        float2 yxy[2], guxy[2], gvxy[2];
        float yz[2], guz[2], gvz[2], su[2], sv[2], si[2], Dd[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int q = 2 * r + k;
          const float4 t0 = lds4(stg + (4 * q + 0) * 128), t1 = lds4(stg + (4 * q + 1) * 128);
          const float4 t2 = lds4(stg + (4 * q + 2) * 128), t3 = lds4(stg + (4 * q + 3) * 128);
          const float4 wq = lds4(stg + (16 + 2 * q) * 128), cq = lds4(stg + (17 + 2 * q) * 128);
          const float wuf = wq.x, wuc = wq.y, wvf = wq.z, wvc = wq.w;
          su[k] = cq.x; sv[k] = cq.y; si[k] = cq.z; Dd[k] = cq.w;
          // I0 = (vf,uf), I1 = (vc,uf), I2 = (vf,uc), I3 = (vc,uc)   (bilinear_interp.py:125-128)
          const float w0 = wuf * wvf, w1 = wuf * wvc, w2 = wuc * wvf, w3 = wuc * wvc;
          const float2 a0 = f2(t0.x, t0.y), a1 = f2(t1.x, t1.y), a2 = f2(t2.x, t2.y), a3 = f2(t3.x, t3.y);
          yxy[k] = f2fma(a3, f2s(w3), f2fma(a2, f2s(w2), f2fma(a0, f2s(w0), f2mul(a1, f2s(w1)))));
          yz[k] = fmaf(t3.z, w3, fmaf(t2.z, w2, fmaf(t0.z, w0, t1.z * w1)));
          const float2 d20 = f2sub(a2, a0), d31 = f2sub(a3, a1), d10 = f2sub(a1, a0), d32 = f2sub(a3, a2);
          guxy[k] = f2fma(f2s(wvf), d20, f2mul(f2s(wvc), d31));
          gvxy[k] = f2fma(f2s(wuf), d10, f2mul(f2s(wuc), d32));
          guz[k] = fmaf(wvf, t2.z - t0.z, wvc * (t3.z - t1.z));
          gvz[k] = fmaf(wuf, t1.z - t0.z, wuc * (t3.z - t2.z));
        }
        float* const q = Yr + (s & (kSYRows - 1)) * kSW + 2 * lane;
        sts2(q, f2(yxy[0].x, yxy[1].x));
        sts2(q + kSYRows * kSW, f2(yxy[0].y, yxy[1].y));
        sts2(q + 2 * kSYRows * kSW, f2(yz[0], yz[1]));
        // mean_c(synth) == 0 marks an invalid pixel for the losses (loss_util.py:15-16)
        const float nb0 = (((yxy[0].x + yxy[0].y) + yz[0]) == 0.f) ? 0.f : 1.f;
        const float nb1 = (((yxy[1].x + yxy[1].y) + yz[1]) == 0.f) ? 0.f : 1.f;
        sts2(q + 3 * kSYRows * kSW, f2(nb0, nb1));
        float* const j = Jr + (s & (kSJRows - 1)) * kSW + 2 * lane;
        constexpr int JP = kSJRows * kSW;
        sts2(j, f2(guxy[0].x, guxy[1].x));
        sts2(j + JP, f2(guxy[0].y, guxy[1].y));
        sts2(j + 2 * JP, f2(guz[0], guz[1]));
        sts2(j + 3 * JP, f2(gvxy[0].x, gvxy[1].x));
        sts2(j + 4 * JP, f2(gvxy[0].y, gvxy[1].y));
        sts2(j + 5 * JP, f2(gvz[0], gvz[1]));
        sts2(j + 6 * JP, f2(su[0], su[1]));
        sts2(j + 7 * JP, f2(sv[0], sv[1]));
        sts2(j + 8 * JP, f2(si[0], si[1]));
        sts2(j + 9 * JP, f2(Dd[0], Dd[1]));
      }
      // ---- phase A: projection + taps of one row, its eight gathers requested (global -> shared) ------------------
      if (doA) {
        const int s = 2 * c + r, sg = 2 * ci + r;
        const float fy = (float)(p.ya - 2 + sg);
        const float2 D = lds2(T + (3 * kSTRows + (s & (kSTRows - 1))) * kSW + 2 * lane);
        float2 pu, pv, inv;
        {
          const float4 gA = lds4(geo), gB = lds4(geo + 4), gC = lds4(geo + 8);      // K0..K5 | Ki0..Ki5
          const float4 tA = lds4(geo + 12), tB = lds4(geo + 16), tC = lds4(geo + 20); // R (9), t (3)
          const float2 r0 = ray_pair(gB.z, gB.w, gC.x, fx, fy);
          const float2 r1 = ray_pair(gC.y, gC.z, gC.w, fx, fy);
          const float2 X0 = f2mul(r0, D), X1 = f2mul(r1, D);
          // reference order: X = ray D, Y = R X + t, p = K_s Y, (u, v) = p.xy / (p.z + 1e-10)
          const float2 Y0 = f2add(f2fma(f2s(tA.z), D, f2fma(f2s(tA.x), X0, f2mul(f2s(tA.y), X1))), f2s(tC.y));
          const float2 Y1 = f2add(f2fma(f2s(tB.y), D, f2fma(f2s(tA.w), X0, f2mul(f2s(tB.x), X1))), f2s(tC.z));
          const float2 Y2 = f2add(f2fma(f2s(tC.x), D, f2fma(f2s(tB.z), X0, f2mul(f2s(tB.w), X1))), f2s(tC.w));
          const float2 p0 = f2fma(f2s(gA.z), Y2, f2fma(f2s(gA.x), Y0, f2mul(f2s(gA.y), Y1)));
          const float2 p1 = f2fma(f2s(gB.y), Y2, f2fma(f2s(gA.w), Y0, f2mul(f2s(gB.x), Y1)));
          const float2 den = f2add(Y2, f2s(1e-10f));
          // perspective divide: one reciprocal per column, refined once, Markstein correction of both quotients
          float ra, rb;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(den.x));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(den.y));
          const float2 nden = f2neg(den);
          float2 rr = f2(ra, rb);
          rr = f2fma(f2fma(nden, rr, f2s(1.f)), rr, rr);
          const float2 q0 = f2mul(p0, rr), q1 = f2mul(p1, rr);
          pu = f2fma(f2fma(nden, q0, p0), rr, q0);
          pv = f2fma(f2fma(nden, q1, p1), rr, q1);
          inv = rr;
        }
        const float us[2] = {pu.x, pu.y}, vs[2] = {pv.x, pv.y}, Ds[2] = {D.x, D.y}, is[2] = {inv.x, inv.y};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int q = 2 * r + k;
          const float uf = floorf(us[k]), vf = floorf(vs[k]);
          // valid <=> 0 <= floor(u) <= W-2 and 0 <= floor(v) <= H-2 and D != 0 (bilinear_interp.py:53-76); NaN -> false
          const bool val = (uf >= 0.f) && (uf <= wlim) && (vf >= 0.f) && (vf <= hlim) && (Ds[k] != 0.f);
          const int iu = val ? (int)uf : 0, iv = val ? (int)vf : 0;
          // weights * valid_mask (bilinear_interp.py:100): an invalid sample gathers texel (0,0) with zero weights
          float4 wq, cq;
          wq.x = val ? (uf + 1.f) - us[k] : 0.f;
          wq.y = val ? us[k] - uf : 0.f;
          wq.z = val ? (vf + 1.f) - vs[k] : 0.f;
          wq.w = val ? vs[k] - vf : 0.f;
          cq.x = val ? us[k] : 0.f; cq.y = val ? vs[k] : 0.f; cq.z = val ? is[k] : 0.f; cq.w = Ds[k];
          const float4* tp = img4 + (iv * W + iu);
          cp_async16(stg + (4 * q + 0) * 128, tp);
          cp_async16(stg + (4 * q + 1) * 128, tp + W);
          cp_async16(stg + (4 * q + 2) * 128, tp + 1);
          cp_async16(stg + (4 * q + 3) * 128, tp + W + 1);
          *reinterpret_cast<float4*>(stg + (16 + 2 * q) * 128) = wq;
          *reinterpret_cast<float4*>(stg + (17 + 2 * q) * 128) = cq;
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");        // one group per row, empty when nothing was requested
    }
    if (doA && ++ci == p.nch) { ci = 0; fresh = true; if (++pi < pend) p = load_piece(a.pieces, pi); }
    strip_bar();
  }
}

// ---------------------------------------------------------------------------------------------------------
// window machinery shared by the X and S roles
// ---------------------------------------------------------------------------------------------------------
struct Win2 {        // sliding 3-row window of a row quantity h: sum(h[r-2], h[r-1], h[r]) in the order ((a + b) + c)
  float2 p1, p2;     // p1 = h[r-1], p2 = h[r-2] + h[r-1]
  __device__ __forceinline__ float2 push(float2 h) {
    const float2 v = f2add(p2, h);
    p2 = f2add(p1, h); p1 = h;
    return v;
  }
  __device__ __forceinline__ void clear() { p1 = f2s(0.f); p2 = f2s(0.f); }
};

// horizontal 3-sums of a row quantity held as one column pair per lane: (left + q.x + q.y, q.x + q.y + right)
__device__ __forceinline__ float2 hsum_nb(float2 q, float2 lr) { return f2add(lr, f2s(q.x + q.y)); }
__device__ __forceinline__ float2 shfl_lr(float2 q) {   // (right column of the left lane, left column of the right lane)
  return f2(__shfl_up_sync(0xffffffffu, q.y, 1), __shfl_down_sync(0xffffffffu, q.x, 1));
}

// per-lane column constants of a piece: 1/#taps for window rows with 3 / 2 in-image rows, centre-column flags
struct StripCols { float2 ic3, ic2, cen; };
__device__ __forceinline__ StripCols strip_cols(const StripPiece& p, int W, int lane) {
  const int gx = p.x0 - 2 + 2 * lane;
  const int cx0 = strip_cnt(gx, W), cx1 = strip_cnt(gx + 1, W);
  StripCols c;
  c.ic3 = f2(strip_box_inv(3, cx0), strip_box_inv(3, cx1));
  c.ic2 = f2(strip_box_inv(2, cx0), strip_box_inv(2, cx1));
  c.cen = f2((2 * lane >= 2 && 2 * lane < 2 + p.cw && gx < W) ? 1.f : 0.f,
             (2 * lane + 1 >= 2 && 2 * lane + 1 < 2 + p.cw && gx + 1 < W) ? 1.f : 0.f);
  return c;
}

// ---------------------------------------------------------------------------------------------------------
// X: window statistics of the TARGET rows (the same for every source): mu_x, mu_x^2 + c1, sigma_x + c2 of the three
// channels (loss_util.py:70-84), plus the row's 1/#taps and centre flags, one tick ahead of the statistics warps
// ---------------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void strip_role_x(const StripArgs& a, float* smem, const int lane, const StripCta cta) {
  using SM = StripSmem<NS>;
  const float* const T = smem + SM::T + 2 * lane;
  float* const X = smem + SM::X + 2 * lane;
  const int total = cta.chunks, pend = cta.first + cta.count;
  int pi = cta.first, ci = 0;
  StripPiece p = load_piece(a.pieces, pi);
  StripCols col = strip_cols(p, a.lt.lv[p.l].W, lane);
  int H = a.lt.lv[p.l].H, Lr = p.yb - p.ya;
  Win2 wx[3], wxx[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { wx[c].clear(); wxx[c].clear(); }
  STRIP_TIMER();
  for (int t = 0; t < total + kSLagEnd; ++t) {
    const int c = t - kSLagX;
    if (c >= 0 && c < total) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int s = 2 * c + r, sg = 2 * ci + r;
        // statistics row sg-1 (image row gy1): 1/#taps, centre flags of that row
        const int gy1 = p.ya - 3 + sg;
        const int cy = strip_cnt(gy1, H);
        const float2 ic = cy == 3 ? col.ic3 : (cy == 2 ? col.ic2 : f2s(0.f));
        const bool row_c = sg >= 3 && sg < Lr + 3;                  // a centre row of this piece
        float* const q = X + ((s - 1) & (kSXRows - 1)) * kSW;
        constexpr int XP = kSXRows * kSW;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float2 x = lds2(T + ch * kSTRows * kSW + (s & (kSTRows - 1)) * kSW);
          const float2 xn = shfl_lr(x);
          const float2 Vx = wx[ch].push(hsum_nb(x, xn));
          const float2 Vxx = wxx[ch].push(hsum_nb(f2mul(x, x), f2mul(xn, xn)));
          const float2 mux = f2mul(Vx, ic), mux2 = f2mul(mux, mux);
          sts2(q + ch * XP, mux);
          sts2(q + (3 + ch) * XP, f2add(mux2, f2s(kC1)));
          sts2(q + (6 + ch) * XP, f2add(f2fma(Vxx, ic, f2neg(mux2)), f2s(kC2)));
        }
        sts2(q + 9 * XP, ic);
        sts2(q + 10 * XP, row_c ? col.cen : f2s(0.f));
      }
      if (++ci == p.nch) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) { wx[ch].clear(); wxx[ch].clear(); }
        ci = 0;
        if (++pi < pend) {
          p = load_piece(a.pieces, pi);
          col = strip_cols(p, a.lt.lv[p.l].W, lane);
          H = a.lt.lv[p.l].H; Lr = p.yb - p.ya;
        }
      }
    }
    strip_bar();
  }
}

// ---------------------------------------------------------------------------------------------------------
// S: window statistics of the warped rows, SSIM + L1, adjoint coefficients and their box sums for one (source, channel)
// ---------------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void strip_role_s(const StripArgs& a, float* smem, const int lane, const int n, const int ch,
                                             const StripCta cta) {
  using SM = StripSmem<NS>;
  const float* const Tx = smem + SM::T + ch * kSTRows * kSW + 2 * lane;
  const float* const Xs = smem + SM::X + 2 * lane;
  const float* const Yc = smem + SM::src0 + n * SM::kSrc + SM::kY + ch * kSYRows * kSW + 2 * lane;
  const float* const Ynb = smem + SM::src0 + n * SM::kSrc + SM::kY + 3 * kSYRows * kSW + 2 * lane;
  float* const Gc = smem + SM::src0 + n * SM::kSrc + SM::kG + ch * kSGRows * kSW + 2 * lane;
  const int total = cta.chunks, pend = cta.first + cta.count;
  const bool live = n < a.N;
  int pi = cta.first, ci = 0;
  StripPiece p = load_piece(a.pieces, pi);
  float cl1 = a.gcoef_l1 * a.norm_photo[p.l], hss2 = -a.gcoef_ssim * a.norm_photo[p.l];

  Win2 wy, wyy, wxy, wA, wB, wC;
  wy.clear(); wyy.clear(); wxy.clear(); wA.clear(); wB.clear(); wC.clear();
  float2 l1t2 = f2s(0.f);                                    // L1 gradient term of row sg-2
  float2 ls_l1 = f2s(0.f), ls_ss = f2s(0.f);

  STRIP_TIMER();
  for (int t = 0; t < total + kSLagEnd; ++t) {
    const int c = t - kSLagS;
    if (live && c >= 0 && c < total) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int s = 2 * c + r;
        // rows sg (window sums), sg-1 (L1, loss masks) and sg-2 (dL/dS) straight from the rings: no row history in registers
        const float2 y = lds2(Yc + (s & (kSYRows - 1)) * kSW), x = lds2(Tx + (s & (kSTRows - 1)) * kSW);
        const float2 y1 = lds2(Yc + ((s - 1) & (kSYRows - 1)) * kSW), x1 = lds2(Tx + ((s - 1) & (kSTRows - 1)) * kSW);
        const float2 nb1 = lds2(Ynb + ((s - 1) & (kSYRows - 1)) * kSW);
        const float2 y2 = lds2(Yc + ((s - 2) & (kSYRows - 1)) * kSW), x2 = lds2(Tx + ((s - 2) & (kSTRows - 1)) * kSW);
        // target statistics, 1/#taps and centre flags of statistics row sg-1 (X role)
        const float* const xq = Xs + ((s - 1) & (kSXRows - 1)) * kSW;
        constexpr int XP = kSXRows * kSW;
        const float2 mux = lds2(xq + ch * XP), MUX2C = lds2(xq + (3 + ch) * XP), SGXC = lds2(xq + (6 + ch) * XP);
        const float2 ic = lds2(xq + 9 * XP);
        const float2 cnt_w = f2mul(lds2(xq + 10 * XP), nb1);
        float2 g;
        {
          // ---- row sums of y, y^2, xy over columns (c-1, c, c+1) -----------------------------------------------
          const float2 yn = shfl_lr(y), xn = shfl_lr(x);
          const float2 hy = hsum_nb(y, yn);
          const float2 hyy = hsum_nb(f2mul(y, y), f2mul(yn, yn));
          const float2 hxy = hsum_nb(f2mul(x, y), f2mul(xn, yn));
          const float2 Vy = wy.push(hy), Vyy = wyy.push(hyy), Vxy = wxy.push(hxy);
          // ---- SSIM, loss, adjoint coefficients of row sg-1 (loss_util.py:52-96) -----------------------------
          const float2 muy = f2mul(Vy, ic), muy2 = f2mul(muy, muy), mxy = f2mul(mux, muy);
          const float2 sgy = f2fma(Vyy, ic, f2neg(muy2));
          const float2 sgxy = f2fma(Vxy, ic, f2neg(mxy));
          const float2 a1 = f2fma(f2s(2.f), mxy, f2s(kC1));
          const float2 a2 = f2fma(f2s(2.f), sgxy, f2s(kC2));
          const float2 b1 = f2add(MUX2C, muy2);
          const float2 b2 = f2add(SGXC, sgy);
          const float2 den = f2mul(b1, b2);
          const float2 r12 = f2(rcp_nr(den.x), rcp_nr(den.y));
          const float2 ssim = f2mul(f2mul(a1, a2), r12);
          const float2 lv = f2fma(f2s(-0.5f), ssim, f2s(0.5f));
          const float lc0 = fminf(fmaxf(lv.x, 0.f), 1.f), lc1 = fminf(fmaxf(lv.y, 0.f), 1.f);
          ls_ss = f2fma(cnt_w, f2(lc0, lc1), ls_ss);
          // clip_by_value passes the gradient inside [0,1]; Hh = 2 h / (#taps b1 b2), h = dTotal/d ssim
          const float2 hlive = f2mul(f2mul(ic, f2s(hss2)), nb1);
          const float2 Hh = f2mul(f2(lc0 == lv.x ? hlive.x : 0.f, lc1 == lv.y ? hlive.y : 0.f), r12);
          const float2 Hs = f2mul(Hh, ssim);
          const float2 t1 = f2mul(mux, f2add(a2, f2neg(a1)));
          const float2 t2 = f2mul(muy, f2add(b2, f2neg(b1)));
          const float2 Av = f2fma(Hh, t1, f2neg(f2mul(Hs, t2)));
          const float2 Bv = f2mul(f2neg(Hs), b1);              // 2 dL/dP(y^2)
          const float2 Cv = f2mul(Hh, a1);
          // ---- 3x3 box sums of the coefficients (adjoint of the SAME-padded mean) -> dL/dS of row sg-2 -----------
          const float2 sa = wA.push(hsum_nb(Av, shfl_lr(Av)));
          const float2 sb = wB.push(hsum_nb(Bv, shfl_lr(Bv)));
          const float2 sc = wC.push(hsum_nb(Cv, shfl_lr(Cv)));
          g = f2fma(y2, sb, f2fma(x2, sc, sa));
        }
        // ---- L1 (loss_util.py:6-25): loss of row sg-1, gradient term of row sg-2 ------------------------------
        // (a term that is switched off has a zero coefficient -- hss2 or cl1 -- and its loss sum is dropped below:
        //  no branch around the window state, so the row-to-row renaming costs no register moves)
        {
          const float2 d1 = f2sub(y1, x1);
          ls_l1 = f2fma(cnt_w, f2(fabsf(d1.x), fabsf(d1.y)), ls_l1);
          g = f2add(g, l1t2);
          const float2 nbc = f2mul(nb1, f2s(cl1));
          l1t2 = f2(d1.x == 0.f ? 0.f : copysignf(nbc.x, d1.x), d1.y == 0.f ? 0.f : copysignf(nbc.y, d1.y));
        }
        sts2(Gc + ((s - 2) & (kSGRows - 1)) * kSW, g);
      }
      if (++ci == p.nch) {
        // piece done: loss sums of this (source, channel) -> scratch of the output stage; windows restart
        const float v0 = warp_sum(ls_l1.x + ls_l1.y), v1 = warp_sum(ls_ss.x + ls_ss.y);
        if (lane == 0) {
          float* lb = smem + SM::loss + (pi & 1) * (3 * NS * 2) + (n * 3 + ch) * 2;
          lb[0] = a.do_l1 ? v0 : 0.f; lb[1] = a.do_ssim ? v1 : 0.f;
        }
        ls_l1 = f2s(0.f); ls_ss = f2s(0.f);
        wy.clear(); wyy.clear(); wxy.clear(); wA.clear(); wB.clear(); wC.clear();
        l1t2 = f2s(0.f);
        ci = 0;
        if (++pi < pend) {
          p = load_piece(a.pieces, pi);
          cl1 = a.gcoef_l1 * a.norm_photo[p.l]; hss2 = -a.gcoef_ssim * a.norm_photo[p.l];
        }
      }
    }
    strip_bar();
  }
}

// ---------------------------------------------------------------------------------------------------------
// G: dL/dS x Jacobian -> projection adjoint: dL/ddepth per pixel, dL/dR, dL/dt sums per piece.
// One warp per source.
// ---------------------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void strip_role_g(const StripArgs& a, float* smem, const int lane, const int gw,
                                             const StripCta cta) {
  using SM = StripSmem<NS>;
  constexpr int NG = 1;                         // sources of this warp
  float* const geo = smem + SM::geoG + gw * 48;  // [R|t] x 2 sources, K rows 0-1, inv K rows 0-1
  const int total = cta.chunks, pend = cta.first + cta.count;
  const int n0 = NG * gw;
  const bool live = n0 < a.N;
  int pi = cta.first, ci = 0;
  StripPiece p = load_piece(a.pieces, pi);
  bool fresh = true;
  float acc[NG][12];
#pragma unroll
  for (int q = 0; q < NG; ++q)
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[q][k] = 0.f;
  float2 cen = f2s(0.f), fx = f2s(0.f);
  int Lr = 0;

  STRIP_TIMER();
  for (int t = 0; t < total + kSLagEnd; ++t) {
    const int c = t - kSLagG;
    if (live && c >= 0 && c < total) {
      if (fresh) {
        // [R|t] of the two sources (24 words), then K rows 0-1 and inv K rows 0-1 (12 words)
        __syncwarp();
        if (lane < 12 * NG) {
          if (n0 + lane / 12 < a.N) geo[lane] = __ldg(a.geoT + (size_t)(p.b * a.N + n0 + lane / 12) * kGeoT + lane % 12);
        } else if (lane >= 24) {
          const int k = lane - 24;
          if (k < 6) geo[24 + k] = __ldg(a.geoK + (size_t)(p.b * a.S + p.l) * kGeoK + k);
        }
        if (lane < 6) geo[30 + lane] = __ldg(a.geoK + (size_t)(p.b * a.S + p.l) * kGeoK + 9 + lane);
        __syncwarp();
        const int W = a.lt.lv[p.l].W, gx = p.x0 - 2 + 2 * lane;
        cen = f2((2 * lane >= 2 && 2 * lane < 2 + p.cw && gx < W) ? 1.f : 0.f,
                 (2 * lane + 1 >= 2 && 2 * lane + 1 < 2 + p.cw && gx + 1 < W) ? 1.f : 0.f);
        fx = f2((float)gx, (float)gx + 1.f);
        Lr = p.yb - p.ya;
        fresh = false;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int s = 2 * c + r, sg = 2 * ci + r;
        if (sg >= 4 && sg < Lr + 4) {                 // centre row sg-4 of the piece = region row sg-2
          const int kg = ((s - 2) & (kSGRows - 1)) * kSW + 2 * lane;
          const int kj = ((s - 2) & (kSJRows - 1)) * kSW + 2 * lane;
          constexpr int GP = kSGRows * kSW, JP = kSJRows * kSW;
          const float fy = (float)(p.ya + sg - 4);
          const float4 kA = lds4(geo + 24), kB = lds4(geo + 28), kC = lds4(geo + 32);     // K0..K5 | Ki0..Ki5
          const float gk[6] = {kA.x, kA.y, kA.z, kA.w, kB.x, kB.y};
          const float2 r0 = ray_pair(kB.z, kB.w, kC.x, fx, fy);
          const float2 r1 = ray_pair(kC.y, kC.z, kC.w, fx, fy);
#pragma unroll
          for (int q = 0; q < NG; ++q) {
            if (n0 + q < a.N) {
              const float* const Jr = smem + SM::src0 + (n0 + q) * SM::kSrc + SM::kJ;
              const float* const Gr = smem + SM::src0 + (n0 + q) * SM::kSrc + SM::kG;
              float* const Dr = smem + SM::src0 + (n0 + q) * SM::kSrc + SM::kD;
              const float2 g0 = f2mul(lds2(Gr + kg), cen), g1 = f2mul(lds2(Gr + GP + kg), cen), g2 = f2mul(lds2(Gr + 2 * GP + kg), cen);
              const float2 gu2 = f2fma(g2, lds2(Jr + 2 * JP + kj), f2fma(g1, lds2(Jr + JP + kj), f2mul(g0, lds2(Jr + kj))));
              const float2 gv2 = f2fma(g2, lds2(Jr + 5 * JP + kj), f2fma(g1, lds2(Jr + 4 * JP + kj), f2mul(g0, lds2(Jr + 3 * JP + kj))));
              const float2 U = lds2(Jr + 6 * JP + kj), V = lds2(Jr + 7 * JP + kj), I = lds2(Jr + 8 * JP + kj), D = lds2(Jr + 9 * JP + kj);
              // samples without a valid warp cached zero Jacobians and 1/den = 0: exact zeros below
              const float2 gp0 = f2mul(gu2, I), gp1 = f2mul(gv2, I);
              const float2 gp2 = f2mul(f2neg(f2fma(gv2, V, f2mul(gu2, U))), I);
              const float2 X0 = f2mul(r0, D), X1 = f2mul(r1, D);
              // dL/dY = K_s^T dL/dp (last row of K_s is (0,0,1))
              const float2 gY0 = f2fma(f2s(gk[0]), gp0, f2mul(f2s(gk[3]), gp1));
              const float2 gY1 = f2fma(f2s(gk[1]), gp0, f2mul(f2s(gk[4]), gp1));
              const float2 gY2 = f2add(f2fma(f2s(gk[2]), gp0, f2mul(f2s(gk[5]), gp1)), gp2);
              const float2 gY[3] = {gY0, gY1, gY2};
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                acc[q][3 * i + 0] = fmaf(gY[i].y, X0.y, fmaf(gY[i].x, X0.x, acc[q][3 * i + 0]));
                acc[q][3 * i + 1] = fmaf(gY[i].y, X1.y, fmaf(gY[i].x, X1.x, acc[q][3 * i + 1]));
                acc[q][3 * i + 2] = fmaf(gY[i].y, D.y, fmaf(gY[i].x, D.x, acc[q][3 * i + 2]));
                acc[q][9 + i] += gY[i].x + gY[i].y;
              }
              // dL/dX = R^T dL/dY, dL/dD = ray . dL/dX
              const float4 ta = lds4(geo + q * 12), tb = lds4(geo + q * 12 + 4);
              const float t8 = geo[q * 12 + 8];
              const float2 gX0 = f2fma(f2s(tb.z), gY2, f2fma(f2s(ta.x), gY0, f2mul(f2s(ta.w), gY1)));
              const float2 gX1 = f2fma(f2s(tb.w), gY2, f2fma(f2s(ta.y), gY0, f2mul(f2s(tb.x), gY1)));
              const float2 gX2 = f2fma(f2s(t8), gY2, f2fma(f2s(ta.z), gY0, f2mul(f2s(tb.y), gY1)));
              const float2 gD = f2add(f2fma(gX0, r0, f2mul(gX1, r1)), gX2);
              sts2(Dr + ((s - 2) & (kSDRows - 1)) * kSW + 2 * lane, gD);
            }
          }
        }
      }
      if (++ci == p.nch) {
        // piece done: 12 pose sums of (b, n) -> one partial record per source
#pragma unroll
        for (int q = 0; q < NG; ++q) {
          float v[16];
#pragma unroll
          for (int k = 0; k < 12; ++k) { v[k] = acc[q][k]; acc[q][k] = 0.f; }
          v[12] = v[13] = v[14] = v[15] = 0.f;
          const float tot = warp_reduce16(v, lane);
          const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
          if ((lane & 1) == 0 && idx < 12 && n0 + q < a.N)
            a.pose_part[(((size_t)p.b * a.slots_per_b + p.slot) * a.N + n0 + q) * 12 + idx] = tot;
        }
        ci = 0; fresh = true;
        if (++pi < pend) p = load_piece(a.pieces, pi);
      }
    }
    strip_bar();
  }
}

// Warp order: L (X) O0 O1 | G... | Y... | S...  For four sources the 24 warps are six aligned warpgroups -- {L, X, O0, O1},
// {G0..G3}, {Y0..Y3} and three of statistics warps -- and the register file is re-divided with setmaxnreg: the
// statistics role carries 12 sliding-window pairs (88 registers), the warp role 72, the adjoint role 64, the rest 80
// (24 warps x 80 at launch = 12 x 88 + 4 x 72 + 4 x 64 + 4 x 80: the pool is used up exactly).
// registers per thread of the four-source layout: statistics / warp / adjoint / {L, X, O0, O1} warps; launch = 80
#ifndef XPT_REG_S
#define XPT_REG_S 88
#define XPT_REG_Y 72
#define XPT_REG_G 64
#define XPT_REG_W 80
#endif
static_assert(12 * XPT_REG_S + 4 * XPT_REG_Y + 4 * XPT_REG_G + 4 * XPT_REG_W == 24 * 80, "setmaxnreg: the register pool of the CTA must be used up exactly");
template <int NS, bool DERIVE>
__global__ void __launch_bounds__(StripSmem<NS>::kThreads, NS == 1 ? 2 : 1) k_strip(const __grid_constant__ StripArgs a) {
  extern __shared__ __align__(16) float smem[];
  const StripCta cta = a.ctas[blockIdx.x];
  if (cta.count == 0) return;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The first steps of a CTA read ring rows "before the beginning" (row s-1, s-2 of step 0) with zero weights: the
  // rings must hold finite numbers then, not whatever an earlier kernel left in shared memory (0 x NaN = NaN).
  for (int i = threadIdx.x; i < StripSmem<NS>::kFloats; i += StripSmem<NS>::kThreads) smem[i] = 0.f;
  __syncthreads();
  if (NS == 4) {
    // one setmaxnreg per warpgroup, executed by its four warps together, then the warps part into their roles
    if (wid < 4) {
      if (XPT_REG_W < 80) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(XPT_REG_W));
      if (wid == 0) strip_role_l<NS>(a, smem, lane, cta);
      else if (wid == 1) strip_role_x<NS>(a, smem, lane, cta);
      else strip_role_o<NS, DERIVE>(a, smem, lane, wid - 2, cta);
    } else if (wid < 8) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(XPT_REG_G));
      strip_role_g<NS>(a, smem, lane, wid - 4, cta);
    } else if (wid < 12) {
      if (XPT_REG_Y > 80) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(XPT_REG_Y));
      else if (XPT_REG_Y < 80) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(XPT_REG_Y));
      strip_role_y<NS>(a, smem, lane, wid - 8, cta);
    } else {
      if (XPT_REG_S > 80) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(XPT_REG_S));
      else if (XPT_REG_S < 80) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(XPT_REG_S));
      strip_role_s<NS>(a, smem, lane, (wid - 12) / 3, (wid - 12) % 3, cta);
    }
    return;
  }
  if (wid == 0) strip_role_l<NS>(a, smem, lane, cta);
  else if (wid == 1) strip_role_x<NS>(a, smem, lane, cta);
  else if (wid < 4) strip_role_o<NS, DERIVE>(a, smem, lane, wid - 2, cta);
  else if (wid < 4 + NS) strip_role_g<NS>(a, smem, lane, wid - 4, cta);
  else if (wid < 4 + 2 * NS) strip_role_y<NS>(a, smem, lane, wid - 4 - NS, cta);
  else strip_role_s<NS>(a, smem, lane, (wid - 4 - 2 * NS) / 3, (wid - 4 - 2 * NS) % 3, cta);
}

}  // namespace xpt
