// xpt_fused.cuh -- the fused tile kernel of the total-loss path (north-star kernels 1+2+3+4):
// inverse warp -> shared-memory tile with halo -> L1 + SSIM (+ smoothness) -> SSIM/L1 adjoint
// -> bilinear + projection adjoint, one launch for all scales, all sources, forward AND backward.
//
// Design (DESIGN.md "k_fused"):
//   * one CTA (512 threads, 2 CTAs/SM = 32 warps/SM) = one 64x13 centre tile of (level, snippet);
//     all N sources are looped inside so the target tile and its statistics are loaded once;
//   * camera geometry (K_s, inv K_s, [R|t]) sits in the constant bank: the compiler keeps it in
//     UNIFORM registers (LDCU -> FFMA R, R, UR, R), so the projection costs no vector registers;
//   * the 3x3 SSIM windows and their adjoint are evaluated on 2-pixel strips per thread from 8-byte
//     aligned shared-memory rows (LDS.64), vertical sums first, with Blackwell's packed FP32 pipe
//     (FFMA2 / FADD2 / FMUL2 on float2) doing both pixels of the strip per instruction;
//   * the bilinear Jacobian (dS/du, dS/dv per channel) and (u, v, 1/den) of the centre samples are
//     cached in shared memory by the forward phase, so the backward phase gathers nothing again;
//   * region pitch is 68 floats and the centre width is 64 so that 384/832/1280-wide images (and
//     every 2^k-scaled level down to 64) tile without waste.
#pragma once
#include "xpt_kernels.cuh"

// XPT_EXP: timing ablations for profiles/ablate_variants.sh ONLY (results are wrong when non-zero):
//   1 = no CTA barriers inside the source loop, 2 = skip phase Y, 4 = skip phase S, 8 = skip phase G
#ifndef XPT_EXP
#define XPT_EXP 0
#endif
// XPT_YPIPE bit 0: phase Y keeps both samples of a thread in flight (eight texel loads issued together);
//           bit 1: the tile prologue issues the loads of its three iterations together
#ifndef XPT_YPIPE
#define XPT_YPIPE 2
#endif
// XPT_GRID_TILE_MAJOR 1: grid = (tiles, snippets) -- a snippet's tiles are consecutive CTAs, so the halo rows a tile shares
// with its vertical neighbour are still in L2 whatever the batch size (A/B against grid = (snippets, tiles))
#ifndef XPT_GRID_TILE_MAJOR
#define XPT_GRID_TILE_MAJOR 2          // 2 = by grid size (host), 0 / 1 = forced
#endif
#if XPT_EXP & 1
#define XPT_SYNC() ((void)0)
#else
#define XPT_SYNC() __syncthreads()
#endif

namespace xpt {

constexpr int kFCW = 64, kFCH = 13;            // centre tile
constexpr int kFSW = 66, kFSH = 15;            // statistics region (halo 1)
constexpr int kFRW = 68, kFRH = 17;            // warped / target region (halo 2)
constexpr int kFP = 68;                        // row pitch (floats) of region and statistics arrays
constexpr int kFSStrips = 33;                  // 2-pixel strips per statistics row
constexpr int kFGStrips = 32;                  // 2-pixel strips per centre row
constexpr int kFThreads = 512;
constexpr int kFRegion = kFRH * kFP;           // 1156 floats per channel
constexpr int kFStats = kFSH * kFP;            // 1020
constexpr int kFCP = 64;                       // pitch of the centre arrays
constexpr int kFCentre = kFCH * kFCP;          // 832
constexpr int kFYIters = (kFRegion + kFThreads - 1) / kFThreads;   // 3 (tile load)
// warp phase: every thread takes two region samples (1024); the remaining 132 -- the last two halo rows, never
// centre samples -- are computed ONE SOURCE AHEAD by the three warps that idle in the adjoint phase (13..15),
// so the phase is two balanced iterations.
constexpr int kFYMain = 2 * kFThreads;         // 1024
constexpr int kFYTail = kFRegion - kFYMain;    // 132
constexpr int kFTailThreads = kFThreads - kFCH * 32;   // 96
static_assert(kFYTail <= 2 * kFTailThreads && kFYMain >= (kFCH + 2) * kFP, "tail samples must be halo samples of warps 13..15");
constexpr int kFRedSrc = 4;                    // sources whose pose partials are buffered before the cross-warp sum
constexpr int kFRedFloats = kFRedSrc * kFCH * 16;                  // 832 (>= 48 floats of loss scratch)


// Camera geometry in the constant bank (uniform-register operands of the projection).  The bank has a K region
// (records of S x 18 floats: K_s, inv K_s per level) and a [R|t] region (records of N x 12 floats); a ctx owns a run of
// records in each (xptwarp.cu: geo_slot_*).  A launch addresses its records through blockIdx.x itself -- the grid's x
// extent starts at the ctx's first K record and the CTAs below it exit at once -- so the K block's address stays
// blockIdx.x * S * 18 + level * 18 with a compile-time base (a run-time base offset cost 4 % of the kernel in spills).
constexpr int kGeoConstFloats = 15360;         // 60 KB
constexpr int kGeoKRegion = 9216;              // floats of the K region (60 %: S x 18 against N x 12 at S = N = 4)
__constant__ float c_geo[kGeoConstFloats];

template <bool GRAD>
struct FusedSmem {
  // offsets in floats
  static constexpr int sy = 0;                                  // [3][17][68] warped tile
  static constexpr int sx = sy + 3 * kFRegion + 4;              // [3][17][68] target tile
  static constexpr int sD = sx + 3 * kFRegion + 4;              // [17][68] depth (0 outside the image)
  static constexpr int sR0 = sD + kFRegion + 4;                 // [17][68] x2: ray = inv(K_s) (u,v,1), xy components
  static constexpr int sR1 = sR0 + kFRegion + 4;
  static constexpr int red = sR1 + kFRegion + 4;                // pose partials [kFRedSrc][13 warps][16] / loss scratch
  static constexpr int sA = red + kFRedFloats;                  // GRAD: [3][15][68] x3
  static constexpr int sB = sA + (GRAD ? 3 * kFStats + 4 : 0);
  static constexpr int sC = sB + (GRAD ? 3 * kFStats + 4 : 0);
  static constexpr int sGU = sC + (GRAD ? 3 * kFStats + 4 : 0); // GRAD: [3][13][64] dS_c/du
  static constexpr int sGV = sGU + (GRAD ? 3 * kFCentre : 0);
  static constexpr int sU = sGV + (GRAD ? 3 * kFCentre : 0);    // GRAD: [13][64] u, v, 1/den
  static constexpr int sV = sU + (GRAD ? kFCentre : 0);
  static constexpr int sI = sV + (GRAD ? kFCentre : 0);
  static constexpr int kFloats = sI + (GRAD ? kFCentre : 0);
  static constexpr size_t kBytes = sizeof(float) * kFloats;
};

// ---- packed FP32 (Blackwell FFMA2 / FADD2 / FMUL2) -----------------------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }

__device__ __forceinline__ float box_inv(int cnt) {   // 1 / #in-image taps of a 3x3 window
  return cnt == 9 ? (1.f / 9.f) : (cnt == 6 ? (1.f / 6.f) : (cnt == 4 ? 0.25f : 1.f / (float)cnt));
}

// sum of 16 per-thread accumulators over a warp in 16 shuffles: at each butterfly step a lane keeps
// half of the values and hands the other half to its partner.  Afterwards lane L holds the total of
// value index ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1).
__device__ __forceinline__ float warp_reduce16(float v[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool hi = lane & 16;
    float send = hi ? v[i] : v[i + 8];
    float keep = hi ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool hi = lane & 8;
    float send = hi ? v[i] : v[i + 4];
    float keep = hi ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool hi = lane & 4;
    float send = hi ? v[i] : v[i + 2];
    float keep = hi ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool hi = lane & 2;
    float send = hi ? v[0] : v[1];
    float keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

struct FusedArgs {
  LevelTable lt;                       // Level.tiles_x/y/slot_base describe the 64x13 tiling
  int B, N;
  int b_off;                           // snippet of blockIdx.x = 0 (first snippet of this launch minus the K record offset)
  int k_rec0;                          // first K record of the ctx's slot: CTAs with blockIdx.x < k_rec0 have no work
  int geo_t_off;                       // float offset such that [R|t] of (blockIdx.x, n) sits at geo_t_off + (blockIdx.x * N + n) * 12
  int tiles_per_b;
  int tile_major;                      // grid = (tiles, snippets) instead of (snippets, tiles)
  int first_tile[kMaxScales + 1];
  const float* depth[kMaxScales];
  const float* disp[kMaxScales];
  const float4* src4[kMaxScales];      // RGBx texels of the source levels [B,N,h,w] (written by the pyramid kernels)
  int do_l1, do_ssim, do_smooth;
  float norm_photo[kMaxScales];        // sw_s / (N*h*w*3)
  float norm_sm_x[kMaxScales];
  float norm_sm_y[kMaxScales];
  float grad_factor;
  float gcoef_l1, gcoef_ssim, gcoef_smooth;   // dTotal / d(per-snippet loss), incl. 1/global_batch and grad_scale
  float* loss_part; int slots_per_b;   // [B][slots][3]
  float* pose_part;                    // [B][slots][N][12]
  float* synth_out[kMaxScales];
  float* mask_out[kMaxScales];
  float* d_depth[kMaxScales];
  float* d_disp[kMaxScales];
  float4* d_src4[kMaxScales];          // DSRC: RGBx gradient texels of every source level [B,N,h,w], zeroed by the host
};

// Perspective divide of both coordinates with ONE reciprocal: r = 1/den refined once, then each
// quotient gets the remainder correction q' = q + (p - den*q)*r (Markstein), which reproduces the
// correctly rounded IEEE quotient the reference computes for normal-range operands.
__device__ __forceinline__ void div_pair(float p0, float p1, float den, float& u, float& v, float& r) {
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  r = fmaf(fmaf(-den, r, 1.f), r, r);
  const float q0 = p0 * r, q1 = p1 * r;
  u = fmaf(fmaf(-den, q0, p0), r, q0);
  v = fmaf(fmaf(-den, q1, p1), r, q1);
}

// 16-byte vector reduction into an RGBx gradient texel (sm_90+: red.global.add.v4.f32)
__device__ __forceinline__ void red_add_rgbx(float4* p, float x, float y, float z) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(0.f) : "memory");
}

// horizontal 3-sums of a 4-column vertical sum (a.x a.y b.x b.y) for the strip's two pixels
__device__ __forceinline__ float2 hsum3(float2 a, float2 b) {
  const float m = a.y + b.x;
  return make_float2(a.x + m, m + b.y);
}

// value of one warped region sample (no Jacobian): the same operation order as the main warp phase
__device__ __forceinline__ void halo_sample(const float* __restrict__ gk, const float* __restrict__ gt,
                                            const float4* __restrict__ img4, float D, float r0, float r1, int W, int H,
                                            float yv[3]) {
  const float X0 = r0 * D, X1 = r1 * D, X2 = D;
  const float Y0 = gt[0] * X0 + gt[1] * X1 + gt[2] * X2 + gt[9];
  const float Y1 = gt[3] * X0 + gt[4] * X1 + gt[5] * X2 + gt[10];
  const float Y2 = gt[6] * X0 + gt[7] * X1 + gt[8] * X2 + gt[11];
  const float p0 = gk[0] * Y0 + gk[1] * Y1 + gk[2] * Y2;
  const float p1 = gk[3] * Y0 + gk[4] * Y1 + gk[5] * Y2;
  const float den = Y2 + 1e-10f;
  float pu, pv, inv_den;
  div_pair(p0, p1, den, pu, pv, inv_den);
  const Taps tp = make_taps(pu, pv, D, W, H);
  yv[0] = yv[1] = yv[2] = 0.f;
  if (tp.valid) {
    const float4* tp0 = img4 + (tp.iv * W + tp.iu);
    const float4 t0 = __ldg(tp0), t2 = __ldg(tp0 + 1), t1 = __ldg(tp0 + W), t3 = __ldg(tp0 + W + 1);
    const float w0 = tp.w_uf * tp.w_vf, w1 = tp.w_uf * tp.w_vc, w2 = tp.w_uc * tp.w_vf, w3 = tp.w_uc * tp.w_vc;
    yv[0] = ((t0.x * w0 + t1.x * w1) + t2.x * w2) + t3.x * w3;
    yv[1] = ((t0.y * w0 + t1.y * w1) + t2.y * w2) + t3.y * w3;
    yv[2] = ((t0.z * w0 + t1.z * w1) + t2.z * w2) + t3.z * w3;
  }
}

// GRAD: backward in the same launch.  OUT: synth_ms / mask_ms are written.  DSRC: dL/dsource scatter.
// DERIVE 1: the disparity of the smoothness term is safe_reciprocal_number(depth) formed in the kernel.
// DERIVE 2: in addition depth[] holds the depth net's LOGITS (XPT_FLAG_DEPTH_LOGIT): InverseSigmoidActivation is
//           applied when the depth tile is loaded and d_depth[] receives dL/dlogit.
template <bool GRAD, bool OUT, bool DSRC, int DERIVE>
__global__ void __launch_bounds__(kFThreads, 2) k_fused(const __grid_constant__ FusedArgs a) {
  using SM = FusedSmem<GRAD>;
  extern __shared__ __align__(16) float smem[];
  float* const sy = smem + SM::sy;
  float* const sx = smem + SM::sx;
  float* const sD = smem + SM::sD;
  float* const sR0 = smem + SM::sR0;
  float* const sR1 = smem + SM::sR1;
  float* const red = smem + SM::red;
  float* const sA = smem + SM::sA;
  float* const sB = smem + SM::sB;
  float* const sC = smem + SM::sC;
  float* const sGU = smem + SM::sGU;
  float* const sGV = smem + SM::sGV;
  float* const sU = smem + SM::sU;
  float* const sV = smem + SM::sV;
  float* const sI = smem + SM::sI;

  // ---- which tile -------------------------------------------------------------
  // grid = (tiles, snippets): a snippet's tiles are consecutive CTAs (large levels first), so the halo rows a tile shares
  // with its vertical neighbour are still in L2 whatever the batch size -- at config 5 (B = 128) the snippet-major order
  // of round 1 re-read them from HBM (12.9 GB DRAM traffic for 6.4 GB algorithmic; -4.5 % kernel time with this order,
  // +-0 at config 2 / 3: profiles/r02_ab_grid_order.txt).  (A persistent grid drawing tiles from a ticket was measured in
  // round 2: +2 % time at config 2 and config 3 -- the ticket's two barriers per tile cost more than the tail.)
  // a.tile_major (host: grids of more than a few waves): see above; small grids keep the snippet-major order of round 1,
  // whose LAST CTAs are the cheap small-level tiles of all snippets (config 2: 97.3 vs 98-99 us)
  int t = a.tile_major ? blockIdx.x : blockIdx.y;
  const int bl = a.tile_major ? blockIdx.y : blockIdx.x;      // K record of this snippet inside the constant bank
  if (bl < a.k_rec0) return;
  const int b = a.b_off + bl;
  int l = 0;
  while (l + 1 < a.lt.S && t >= a.first_tile[l + 1]) ++l;
  t -= a.first_tile[l];
  const Level& L = a.lt.lv[l];
  const int H = L.H, W = L.W, P = H * W;
  const int ty0 = (t / L.tiles_x) * kFCH, tx0 = (t % L.tiles_x) * kFCW;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int slot = L.slot_base + t;
  const float* const gk = c_geo + (bl * a.lt.S + l) * kGeoK;     // K_s (9), inv K_s (9): uniform registers

  // ---- target tile, depth tile and pixel rays (halo 2); zero outside the image ---------------------
  // depth 0 marks "no sample": the reference's D != 0 validity test (bilinear_interp.py:53-76) then also
  // rejects the out-of-image halo, so the warp phase needs no bounds test of its own.
  {
    const float* tgt = L.tgt + b * L.tgt_bs;
    const float* dep = a.depth[l] + (long long)b * P;
#if XPT_YPIPE & 2
    // all loads of the three iterations are issued before the first store (clamped addresses, zero by select): one
    // exposed memory latency at the start of a tile instead of three
    float tv[kFYIters][3], td[kFYIters];
    bool tin[kFYIters];
#pragma unroll
    for (int it = 0; it < kFYIters; ++it) {
      const int i = min(tid + it * kFThreads, kFRegion - 1);
      const int ry = i / kFP, rx = i - ry * kFP;
      const int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
      tin[it] = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const long long o = (long long)min(max(gy, 0), H - 1) * W + min(max(gx, 0), W - 1);
      const float* p = tgt + o * 3;
      tv[it][0] = __ldg(p); tv[it][1] = __ldg(p + 1); tv[it][2] = __ldg(p + 2);
      td[it] = __ldg(dep + o);
    }
#pragma unroll
    for (int it = 0; it < kFYIters; ++it) {
      const int i = tid + it * kFThreads;
      if (i < kFRegion) {
        const int ry = i / kFP, rx = i - ry * kFP;
        const int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
        float d = td[it];
        if (DERIVE == 2) d = depth_of_logit(d);
        const float fx = (float)gx, fy = (float)gy;
        const float r0 = gk[9] * fx + gk[10] * fy + gk[11];
        const float r1 = gk[12] * fx + gk[13] * fy + gk[14];
        const bool in = tin[it];
        sx[i] = in ? tv[it][0] : 0.f; sx[kFRegion + i] = in ? tv[it][1] : 0.f; sx[2 * kFRegion + i] = in ? tv[it][2] : 0.f;
        sD[i] = in ? d : 0.f; sR0[i] = in ? r0 : 0.f; sR1[i] = in ? r1 : 0.f;
      }
    }
#else
#pragma unroll
    for (int it = 0; it < kFYIters; ++it) {
      const int i = tid + it * kFThreads;
      if (i < kFRegion) {
        const int ry = i / kFP, rx = i - ry * kFP;
        const int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, d = 0.f, r0 = 0.f, r1 = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
          const float* p = tgt + ((long long)gy * W + gx) * 3;
          v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2);
          d = __ldg(dep + (long long)gy * W + gx);
          if (DERIVE == 2) d = depth_of_logit(d);
          // reference order.  The last rows of K_s and inv(K_s) are exactly (0,0,1)
          // (synthesize_base.py:66-71), so ray.z = 1 and p.z = Y.z hold bit-exactly and are not recomputed.
          const float fx = (float)gx, fy = (float)gy;
          r0 = gk[9] * fx + gk[10] * fy + gk[11];
          r1 = gk[12] * fx + gk[13] * fy + gk[14];
        }
        sx[i] = v0; sx[kFRegion + i] = v1; sx[2 * kFRegion + i] = v2;
        sD[i] = d; sR0[i] = r0; sR1[i] = r1;
      }
    }
#endif
  }
  __syncthreads();

  // ---- strip coordinates ----------------------------------------------------------------
  // S phase: statistics row qy (0..14), columns q0, q0+1
  // warps 0..14 own one statistics row each (lane = strip: conflict-free 8-byte rows); the 33rd strip of
  // every row goes to lanes 0..14 of warp 15
  const int qy = wid < kFSH ? wid : lane, q0 = wid < kFSH ? lane * 2 : 2 * (kFSStrips - 1);
  // rows below the image (ragged last tile row, the small pyramid levels) are skipped altogether: the
  // statistics slots of a ragged tile are zeroed once, the skipped centre rows contribute nothing
  const bool s_active = (wid < kFSH || lane < kFSH) && (unsigned)(ty0 - 1 + qy) < (unsigned)H;
  // G phase: centre row cyy (0..12), columns c0, c0+1
  const int cyy = tid >> 5, c0 = (tid & 31) * 2;
  const bool g_active = cyy < kFCH && ty0 + cyy < H;
  if (GRAD && (ty0 == 0 || ty0 + kFCH + 1 > H)) {          // block-uniform: some statistics row lies outside the image
    for (int i = tid; i < SM::sGU - SM::sA; i += kFThreads) sA[i] = 0.f;
    __syncthreads();
  }

  float lsum_l1 = 0.f, lsum_ssim = 0.f, lsum_sm = 0.f;

  // the 132 tail samples of the region (see kFYTail) for source n, written straight into the warped tile: region
  // rows 15..16 are read by the statistics phase only, so warps 13..15 may fill them for source n+1 while the
  // other warps run the adjoint phase of source n
  const bool tail_warp = wid >= kFCH;
  const int tj = tid - kFCH * 32;
  auto tail_compute = [&](int n) {
    const float* const gtn = c_geo + a.geo_t_off + (bl * a.N + n) * kGeoT;
    const float4* const imgn = a.src4[l] + (size_t)(b * a.N + n) * P;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = kFYMain + tj + k * kFTailThreads;
      if (k == 0 || tj < kFYTail - kFTailThreads) {
        float yv[3];
        halo_sample(gk, gtn, imgn, sD[i], sR0[i], sR1[i], W, H, yv);
        sy[i] = yv[0]; sy[kFRegion + i] = yv[1]; sy[2 * kFRegion + i] = yv[2];
      }
    }
  };
  if (GRAD && tail_warp && !(XPT_EXP & 2)) tail_compute(0);      // overlaps with the smoothness rows of warps 0..12


  // ---- smoothness on the centre strip (losses.py:409-440) ------------------------------------
  // a.disp[l] == NULL: the disparity is safe_reciprocal_number(depth) (utils/util_funcs.py:157-160, what
  // model_wrappers.py:47-48 feeds in), taken from the depth tile already in shared memory; its gradient is
  // folded into d_depth (d disp / d depth = -disp^2), so neither disp_ms nor d_disp_ms touches HBM.
  float gD[2] = {0.f, 0.f};
  if (a.do_smooth && g_active) {
    const float* dsp = DERIVE ? nullptr : a.disp[l] + (long long)b * P;
    const float nx = a.norm_sm_x[l], ny = a.norm_sm_y[l];
    const float gcx = a.gcoef_smooth * nx, gcy = a.gcoef_smooth * ny;
    const float k3 = a.grad_factor;
    const int gy = ty0 + cyy;
    auto disp_at = [&](int ri, int yy, int xx) -> float {       // region index / image coordinates of the same pixel
      if (!DERIVE) return __ldg(dsp + (long long)yy * W + xx);
      const float D = sD[ri];
      return D > 0.00001f ? __frcp_rn(D) : 0.f;
    };
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const int gx = tx0 + c0 + o;
      if (gy < H && gx < W) {
        const int ri = (cyy + 2) * kFP + (c0 + o + 2);
        const float d = disp_at(ri, gy, gx);
        float gd = 0.f;
        if (gx + 1 < W) {
          float e = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) e += fabsf((sx[c * kFRegion + ri] - sx[c * kFRegion + ri + 1]) * k3);
          const float w = expf(-(e * (1.f / 3.f)));
          const float sd = (d - disp_at(ri + 1, gy, gx + 1)) * w;
          lsum_sm += fabsf(sd) * nx;
          gd += gcx * sgnf(sd) * w;
        }
        if (gy + 1 < H) {
          float e = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) e += fabsf((sx[c * kFRegion + ri] - sx[c * kFRegion + ri + kFP]) * k3);
          const float w = expf(-(e * (1.f / 3.f)));
          const float sd = (d - disp_at(ri + kFP, gy + 1, gx)) * w;
          lsum_sm += fabsf(sd) * ny;
          gd += gcy * sgnf(sd) * w;
        }
        if (GRAD) {
          if (gx >= 1) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) e += fabsf((sx[c * kFRegion + ri - 1] - sx[c * kFRegion + ri]) * k3);
            const float w = expf(-(e * (1.f / 3.f)));
            const float sd = (disp_at(ri - 1, gy, gx - 1) - d) * w;
            gd -= gcx * sgnf(sd) * w;
          }
          if (gy >= 1) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) e += fabsf((sx[c * kFRegion + ri - kFP] - sx[c * kFRegion + ri]) * k3);
            const float w = expf(-(e * (1.f / 3.f)));
            const float sd = (disp_at(ri - kFP, gy - 1, gx) - d) * w;
            gd -= gcy * sgnf(sd) * w;
          }
          if (DERIVE) gD[o] = -(gd * d) * d;
          else if (a.d_disp[l]) a.d_disp[l][(long long)b * P + gy * W + gx] = gd;
        }
      }
    }
  }

  // ---- per-strip invariants of the S phase ------------------------------------------------------
  const float cl1 = a.gcoef_l1 * a.norm_photo[l];
  const float hss2 = -a.gcoef_ssim * a.norm_photo[l];           // 2 * dTotal/d ssim at a contributing pixel
  float2 inv_cnt, s_hc, s_centre;          // per pixel of the strip: 1/#taps, in-image * hss2/#taps, counted in the loss (0/1)
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
    s_hc = f2(iv[0] * hss2, iv[1] * hss2);                       // 0 outside the image (iv = 0)
    // keep the six values in registers: the compiler would otherwise re-derive them for every source
    asm volatile("" : "+f"(inv_cnt.x), "+f"(inv_cnt.y), "+f"(s_hc.x), "+f"(s_hc.y), "+f"(s_centre.x), "+f"(s_centre.y));
  }
  // window statistics of the target (x) for this strip: evaluated once, kept across the N sources
  // (premixed with the SSIM constants: mux, mux^2 + c1, sigma_x + c2)
  float2 MUX[3], MUX2C[3], SGXC[3];
  if (a.do_ssim && s_active) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* px = sx + c * kFRegion + qy * kFP + q0;
      const float2 a0 = lds2(px), b0 = lds2(px + 2), a1 = lds2(px + kFP), b1 = lds2(px + kFP + 2);
      const float2 a2 = lds2(px + 2 * kFP), b2 = lds2(px + 2 * kFP + 2);
      const float2 v1a = f2add(f2add(a0, a1), a2), v1b = f2add(f2add(b0, b1), b2);
      const float2 v2a = f2fma(a2, a2, f2fma(a1, a1, f2mul(a0, a0))), v2b = f2fma(b2, b2, f2fma(b1, b1, f2mul(b0, b0)));
      const float2 s1 = hsum3(v1a, v1b);
      const float2 s2 = hsum3(v2a, v2b);
      const float2 mu = f2mul(s1, inv_cnt);
      const float2 mu2 = f2mul(mu, mu);
      MUX[c] = mu;
      MUX2C[c] = f2add(mu2, f2s(kC1));
      SGXC[c] = f2add(f2fma(s2, inv_cnt, f2neg(mu2)), f2s(kC2));
    }
  }

  for (int n = 0; n < a.N; ++n) {
    const float* const gt = c_geo + a.geo_t_off + (bl * a.N + n) * kGeoT;   // [R|t]: uniform registers
    const float4* const img4 = a.src4[l] + (size_t)(b * a.N + n) * P;

    // ---- phase Y: inverse warp of the region into shared memory ---------------------------------
    // GRAD: two balanced iterations, the tail comes from warps 13..15 (above / adjoint phase); forward only: no idle
    // phase to hide the tail in, so a third, partly filled iteration takes it
#if XPT_YPIPE & 1
    if (GRAD && !(XPT_EXP & 2)) {
      // both samples of a thread in flight together: the two projection chains run first, the eight texel loads are
      // issued back to back (unconditional: a sample without a valid warp reads texel (0,0) and is zeroed by select),
      // then both are consumed -- one exposed gather latency per source instead of two
      constexpr int kS = kFYMain / kFThreads;
      float pu[kS], pv[kS], inv_den[kS];
      Taps tp[kS];
      float4 tx[kS][4];
#pragma unroll
      for (int it = 0; it < kS; ++it) {
        const int i = tid + it * kFThreads;
        const float D = sD[i];
        const float X0 = sR0[i] * D, X1 = sR1[i] * D, X2 = D;
        const float Y0 = gt[0] * X0 + gt[1] * X1 + gt[2] * X2 + gt[9];
        const float Y1 = gt[3] * X0 + gt[4] * X1 + gt[5] * X2 + gt[10];
        const float Y2 = gt[6] * X0 + gt[7] * X1 + gt[8] * X2 + gt[11];
        const float p0 = gk[0] * Y0 + gk[1] * Y1 + gk[2] * Y2;
        const float p1 = gk[3] * Y0 + gk[4] * Y1 + gk[5] * Y2;
        const float den = Y2 + 1e-10f;
        div_pair(p0, p1, den, pu[it], pv[it], inv_den[it]);
        tp[it] = make_taps(pu[it], pv[it], D, W, H);
        const float4* tp0 = img4 + (tp[it].iv * W + tp[it].iu);      // (0,0) when not valid
        tx[it][0] = __ldg(tp0); tx[it][2] = __ldg(tp0 + 1); tx[it][1] = __ldg(tp0 + W); tx[it][3] = __ldg(tp0 + W + 1);
      }
#pragma unroll
      for (int it = 0; it < kS; ++it) {
        const int i = tid + it * kFThreads;
        const int ry = i / kFP, rx = i - ry * kFP;
        const bool centre = (unsigned)(ry - 2) < (unsigned)kFCH && (unsigned)(rx - 2) < (unsigned)kFCW;
        const Taps& q = tp[it];
        const float4 t0 = tx[it][0], t1 = tx[it][1], t2 = tx[it][2], t3 = tx[it][3];
        const float I0[3] = {t0.x, t0.y, t0.z}, I1[3] = {t1.x, t1.y, t1.z}, I2[3] = {t2.x, t2.y, t2.z}, I3[3] = {t3.x, t3.y, t3.z};
        const float w0 = q.w_uf * q.w_vf, w1 = q.w_uf * q.w_vc, w2 = q.w_uc * q.w_vf, w3 = q.w_uc * q.w_vc;
        float yv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float y = ((I0[c] * w0 + I1[c] * w1) + I2[c] * w2) + I3[c] * w3;
          yv[c] = q.valid ? y : 0.f;
        }
        sy[i] = yv[0]; sy[kFRegion + i] = yv[1]; sy[2 * kFRegion + i] = yv[2];
        if (centre) {
          const int ci = i - (2 * kFP + 2) - (ry - 2) * (kFP - kFCP);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float gu = q.w_vf * (I2[c] - I0[c]) + q.w_vc * (I3[c] - I1[c]);
            const float gv = q.w_uf * (I1[c] - I0[c]) + q.w_uc * (I3[c] - I2[c]);
            sGU[c * kFCentre + ci] = q.valid ? gu : 0.f;
            sGV[c * kFCentre + ci] = q.valid ? gv : 0.f;
          }
          sU[ci] = q.valid ? pu[it] : 0.f; sV[ci] = q.valid ? pv[it] : 0.f; sI[ci] = q.valid ? inv_den[it] : 0.f;
          if (OUT) {
            const int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
            if (gy < H && gx < W) {
              const long long o = (long long)(b * a.N + n) * P + gy * W + gx;
              if (a.synth_out[l]) { float* so = a.synth_out[l] + o * 3; so[0] = yv[0]; so[1] = yv[1]; so[2] = yv[2]; }
              if (a.mask_out[l]) a.mask_out[l][o] = q.valid ? 1.f : 0.f;
            }
          }
        }
      }
    } else
#endif
    {
#pragma unroll
    for (int it = 0; it < (GRAD ? kFYMain / kFThreads : kFYIters); ++it) {
      const int i = tid + it * kFThreads;
      if ((GRAD || i < kFRegion) && !(XPT_EXP & 2)) {
        const int ry = i / kFP, rx = i - ry * kFP;
        const bool centre = (unsigned)(ry - 2) < (unsigned)kFCH && (unsigned)(rx - 2) < (unsigned)kFCW;
        float yv[3] = {0.f, 0.f, 0.f};
        float gu[3] = {0.f, 0.f, 0.f}, gv[3] = {0.f, 0.f, 0.f};
        float su = 0.f, sv = 0.f, si = 0.f;
        const float D = sD[i];
        const float X0 = sR0[i] * D, X1 = sR1[i] * D, X2 = D;
        const float Y0 = gt[0] * X0 + gt[1] * X1 + gt[2] * X2 + gt[9];
        const float Y1 = gt[3] * X0 + gt[4] * X1 + gt[5] * X2 + gt[10];
        const float Y2 = gt[6] * X0 + gt[7] * X1 + gt[8] * X2 + gt[11];
        const float p0 = gk[0] * Y0 + gk[1] * Y1 + gk[2] * Y2;
        const float p1 = gk[3] * Y0 + gk[4] * Y1 + gk[5] * Y2;
        const float den = Y2 + 1e-10f;
        float pu, pv, inv_den;
        div_pair(p0, p1, den, pu, pv, inv_den);
        const Taps tp = make_taps(pu, pv, D, W, H);
        if (tp.valid) {
          // four 16-byte texel loads: I0 = (vf,uf), I1 = (vc,uf), I2 = (vf,uc), I3 = (vc,uc)   (bilinear_interp.py:125-128)
          const float4* tp0 = img4 + (tp.iv * W + tp.iu);
          const float4 t0 = __ldg(tp0), t2 = __ldg(tp0 + 1), t1 = __ldg(tp0 + W), t3 = __ldg(tp0 + W + 1);
          const float I0[3] = {t0.x, t0.y, t0.z}, I1[3] = {t1.x, t1.y, t1.z}, I2[3] = {t2.x, t2.y, t2.z}, I3[3] = {t3.x, t3.y, t3.z};
          const float w0 = tp.w_uf * tp.w_vf, w1 = tp.w_uf * tp.w_vc, w2 = tp.w_uc * tp.w_vf, w3 = tp.w_uc * tp.w_vc;
#pragma unroll
          for (int c = 0; c < 3; ++c) yv[c] = ((I0[c] * w0 + I1[c] * w1) + I2[c] * w2) + I3[c] * w3;
          if (GRAD && centre) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              gu[c] = tp.w_vf * (I2[c] - I0[c]) + tp.w_vc * (I3[c] - I1[c]);
              gv[c] = tp.w_uf * (I1[c] - I0[c]) + tp.w_uc * (I3[c] - I2[c]);
            }
            su = pu; sv = pv; si = inv_den;
          }
        }
        sy[i] = yv[0]; sy[kFRegion + i] = yv[1]; sy[2 * kFRegion + i] = yv[2];
        if (centre) {
          if (GRAD) {
            const int ci = i - (2 * kFP + 2) - (ry - 2) * (kFP - kFCP);
            sGU[ci] = gu[0]; sGU[kFCentre + ci] = gu[1]; sGU[2 * kFCentre + ci] = gu[2];
            sGV[ci] = gv[0]; sGV[kFCentre + ci] = gv[1]; sGV[2 * kFCentre + ci] = gv[2];
            sU[ci] = su; sV[ci] = sv; sI[ci] = si;
          }
          if (OUT) {
            const int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
            if (gy < H && gx < W) {
              const long long o = (long long)(b * a.N + n) * P + gy * W + gx;
              if (a.synth_out[l]) { float* so = a.synth_out[l] + o * 3; so[0] = yv[0]; so[1] = yv[1]; so[2] = yv[2]; }
              if (a.mask_out[l]) a.mask_out[l][o] = tp.valid ? 1.f : 0.f;
            }
          }
        }
      }
    }
    }
    XPT_SYNC();

    // ---- phase S: L1 + SSIM (and the SSIM adjoint coefficients) on a 2-pixel strip ----------------
    if (s_active && !(XPT_EXP & 4)) {
      // the strip's pixels sit at region (qy+1, q0+1) and (qy+1, q0+2)
      const int mid = (qy + 1) * kFP + q0;
      float2 hlive, cnt_w;      // hlive: hss2/#taps where in-image and not black; cnt_w: counted in the loss and not black
      {
        // mean_c(synth) == 0 (loss_util.py:15-16); the compiler merges these loads with the window rows below
        const float2 m0a = lds2(sy + mid), m0b = lds2(sy + mid + 2);
        const float2 m1a = lds2(sy + kFRegion + mid), m1b = lds2(sy + kFRegion + mid + 2);
        const float2 m2a = lds2(sy + 2 * kFRegion + mid), m2b = lds2(sy + 2 * kFRegion + mid + 2);
        const bool bk0 = ((m0a.y + m1a.y) + m2a.y) == 0.f;
        const bool bk1 = ((m0b.x + m1b.x) + m2b.x) == 0.f;
        hlive = f2(bk0 ? 0.f : s_hc.x, bk1 ? 0.f : s_hc.y);
        cnt_w = f2(bk0 ? 0.f : s_centre.x, bk1 ? 0.f : s_centre.y);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* py = sy + c * kFRegion + qy * kFP + q0;
        const float* px = sx + c * kFRegion + qy * kFP + q0;
        const float2 ya0 = lds2(py), yb0 = lds2(py + 2), ya1 = lds2(py + kFP), yb1 = lds2(py + kFP + 2);
        const float2 ya2 = lds2(py + 2 * kFP), yb2 = lds2(py + 2 * kFP + 2);
        const float2 xa0 = lds2(px), xb0 = lds2(px + 2), xa1 = lds2(px + kFP), xb1 = lds2(px + kFP + 2);
        const float2 xa2 = lds2(px + 2 * kFP), xb2 = lds2(px + 2 * kFP + 2);
        if (a.do_l1) {
          lsum_l1 = fmaf(cnt_w.x, fabsf(ya1.y - xa1.y), lsum_l1);
          lsum_l1 = fmaf(cnt_w.y, fabsf(yb1.x - xb1.x), lsum_l1);
        }
        if (a.do_ssim) {
          const float2 v1a = f2add(f2add(ya0, ya1), ya2), v1b = f2add(f2add(yb0, yb1), yb2);
          const float2 v2a = f2fma(ya2, ya2, f2fma(ya1, ya1, f2mul(ya0, ya0)));
          const float2 v2b = f2fma(yb2, yb2, f2fma(yb1, yb1, f2mul(yb0, yb0)));
          const float2 v3a = f2fma(xa2, ya2, f2fma(xa1, ya1, f2mul(xa0, ya0)));
          const float2 v3b = f2fma(xb2, yb2, f2fma(xb1, yb1, f2mul(xb0, yb0)));
          const float2 s1 = hsum3(v1a, v1b), s2 = hsum3(v2a, v2b), s3 = hsum3(v3a, v3b);
          const float2 mux = MUX[c];
          const float2 muy = f2mul(s1, inv_cnt);
          const float2 muy2 = f2mul(muy, muy);
          const float2 mxy = f2mul(mux, muy);
          const float2 sgy = f2fma(s2, inv_cnt, f2neg(muy2));
          const float2 sgxy = f2fma(s3, inv_cnt, f2neg(mxy));
          const float2 a1 = f2fma(f2s(2.f), mxy, f2s(kC1));
          const float2 a2 = f2fma(f2s(2.f), sgxy, f2s(kC2));
          const float2 b1 = f2add(MUX2C[c], muy2);
          const float2 b2 = f2add(SGXC[c], sgy);
          const float2 den = f2mul(b1, b2);
          const float2 r12 = f2(rcp_nr(den.x), rcp_nr(den.y));
          const float2 ssim = f2mul(f2mul(a1, a2), r12);
          const float2 lv = f2fma(f2s(-0.5f), ssim, f2s(0.5f));
          const float lc0 = fminf(fmaxf(lv.x, 0.f), 1.f), lc1 = fminf(fmaxf(lv.y, 0.f), 1.f);
          lsum_ssim = fmaf(cnt_w.x, lc0, lsum_ssim);
          lsum_ssim = fmaf(cnt_w.y, lc1, lsum_ssim);
          if (GRAD) {
            // clip_by_value passes the gradient inside [0,1] (there the clamp is the identity).
            // With Hh = 2 h / (#taps b1 b2), h = dTotal/d ssim:
            //   dL/dP(xy) = Hh a1,  2 dL/dP(y^2) = -Hh ssim b1,
            //   dL/dmu_y  = Hh [ mux (a2 - a1) - ssim muy (b2 - b1) ]
            const float2 Hh = f2mul(f2(lc0 == lv.x ? hlive.x : 0.f, lc1 == lv.y ? hlive.y : 0.f), r12);
            const float2 Hs = f2mul(Hh, ssim);
            const float2 t1 = f2mul(mux, f2add(a2, f2neg(a1)));
            const float2 t2 = f2mul(muy, f2add(b2, f2neg(b1)));
            const float2 Av = f2fma(Hh, t1, f2neg(f2mul(Hs, t2)));
            const float2 Bv = f2mul(f2neg(Hs), b1);
            const float2 Cv = f2mul(Hh, a1);
            const int so = c * kFStats + qy * kFP + q0;
            *reinterpret_cast<float2*>(sA + so) = Av;
            *reinterpret_cast<float2*>(sB + so) = Bv;
            *reinterpret_cast<float2*>(sC + so) = Cv;
          }
        }
      }
    }

    // ---- phase G: dL/dS on the centre strip, pushed through the bilinear + projection adjoint ------
    if (GRAD) {
      XPT_SYNC();
      if (tail_warp && n + 1 < a.N && !(XPT_EXP & 2)) tail_compute(n + 1);   // warps 13..15 have no adjoint rows
      if (g_active && !(XPT_EXP & 8)) {     // warp-uniform: warps 0..12 own one centre row each
        float acc[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = 0.f;
        const int rrow = (cyy + 2) * kFP + c0 + 2;      // centre pixel in region coordinates
        float2 g[3];
        {
          float2 yv[3], xv[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) { yv[c] = lds2(sy + c * kFRegion + rrow); xv[c] = lds2(sx + c * kFRegion + rrow); }
          // L1 term: cl1 * sign(y - x), 0 where the synthesised pixel is black
          const float2 nbc = f2((((yv[0].x + yv[1].x) + yv[2].x) == 0.f) ? 0.f : cl1,
                                (((yv[0].y + yv[1].y) + yv[2].y) == 0.f) ? 0.f : cl1);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float2 gc = f2s(0.f);
            if (a.do_ssim) {
              // centre (cyy, c0+o) = statistics (cyy+1, c0+o+1): window = statistics rows cyy..cyy+2, cols c0+o..c0+o+2
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
              gc = f2fma(yv[c], sb, f2fma(xv[c], sc, sa));       // sB holds 2 dL/dP(y^2)
            }
            if (a.do_l1) {
              const float2 d = f2add(yv[c], f2neg(xv[c]));
              gc = f2add(gc, f2(d.x == 0.f ? 0.f : copysignf(nbc.x, d.x), d.y == 0.f ? 0.f : copysignf(nbc.y, d.y)));
            }
            g[c] = gc;
          }
        }
        const int ci = cyy * kFCP + c0;
        float2 gu2 = f2s(0.f), gv2 = f2s(0.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          gu2 = f2fma(g[c], lds2(sGU + c * kFCentre + ci), gu2);
          gv2 = f2fma(g[c], lds2(sGV + c * kFCentre + ci), gv2);
        }
        const float2 U2 = lds2(sU + ci), V2 = lds2(sV + ci), I2v = lds2(sI + ci);
        const float2 D2 = lds2(sD + rrow), R02 = lds2(sR0 + rrow), R12 = lds2(sR1 + rrow);
        // samples without a valid warp (incl. everything outside the image) cached zero Jacobians and 1/den = 0:
        // their contributions below are exact zeros, so no bounds test is needed
        const float2 gp0 = f2mul(gu2, I2v), gp1 = f2mul(gv2, I2v);
        const float2 gp2 = f2mul(f2neg(f2fma(gv2, V2, f2mul(gu2, U2))), I2v);
        const float2 X02 = f2mul(R02, D2), X12 = f2mul(R12, D2);
        const float gp0s[2] = {gp0.x, gp0.y}, gp1s[2] = {gp1.x, gp1.y}, gp2s[2] = {gp2.x, gp2.y};
        const float X0s[2] = {X02.x, X02.y}, X1s[2] = {X12.x, X12.y}, Ds[2] = {D2.x, D2.y};
        const float r0s[2] = {R02.x, R02.y}, r1s[2] = {R12.x, R12.y};
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const float X0 = X0s[o], X1 = X1s[o], X2 = Ds[o];
          const float gY0 = gk[0] * gp0s[o] + gk[3] * gp1s[o];          // K_s^T with last row (0,0,1)
          const float gY1 = gk[1] * gp0s[o] + gk[4] * gp1s[o];
          const float gY2 = gk[2] * gp0s[o] + gk[5] * gp1s[o] + gp2s[o];
          acc[0] += gY0 * X0; acc[1] += gY0 * X1; acc[2] += gY0 * X2;
          acc[3] += gY1 * X0; acc[4] += gY1 * X1; acc[5] += gY1 * X2;
          acc[6] += gY2 * X0; acc[7] += gY2 * X1; acc[8] += gY2 * X2;
          acc[9] += gY0; acc[10] += gY1; acc[11] += gY2;
          const float gX0 = gt[0] * gY0 + gt[3] * gY1 + gt[6] * gY2;
          const float gX1 = gt[1] * gY0 + gt[4] * gY1 + gt[7] * gY2;
          const float gX2 = gt[2] * gY0 + gt[5] * gY1 + gt[8] * gY2;
          gD[o] += gX0 * r0s[o] + gX1 * r1s[o] + gX2;
          if (DSRC && a.d_src4[l]) {
            // dL/dsource: re-derive the taps from the cached coordinates (bit-identical to the forward) and add the
            // four weighted texels with ONE 16-byte vector reduction each (REDG.E.ADD.F32x4) into the RGBx gradient level
            const float Us = o ? U2.y : U2.x, Vs = o ? V2.y : V2.x, inv = o ? I2v.y : I2v.x;
            const Taps tp = make_taps(Us, Vs, Ds[o], W, H);
            if (tp.valid && inv != 0.f) {
              float4* p = a.d_src4[l] + (size_t)(b * a.N + n) * P + (tp.iv * W + tp.iu);
              const float w0 = tp.w_uf * tp.w_vf, w1 = tp.w_uf * tp.w_vc, w2 = tp.w_uc * tp.w_vf, w3 = tp.w_uc * tp.w_vc;
              const float gs[3] = {o ? g[0].y : g[0].x, o ? g[1].y : g[1].x, o ? g[2].y : g[2].x};
              red_add_rgbx(p, w0 * gs[0], w0 * gs[1], w0 * gs[2]);
              red_add_rgbx(p + 1, w2 * gs[0], w2 * gs[1], w2 * gs[2]);
              red_add_rgbx(p + W, w1 * gs[0], w1 * gs[1], w1 * gs[2]);
              red_add_rgbx(p + W + 1, w3 * gs[0], w3 * gs[1], w3 * gs[2]);
            }
          }
        }
        // 12 pose accumulators: warp reduction in 16 shuffles; the cross-warp sum is deferred (below)
        const float tot = warp_reduce16(acc, lane);
        if ((lane & 1) == 0)
          red[((n & (kFRedSrc - 1)) * kFCH + wid) * 16 +
              (((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1))] = tot;
      }
    }
    XPT_SYNC();           // sy / sA.. / sGU.. are rewritten by the next source; pose partials are complete
    if (GRAD && ((n & (kFRedSrc - 1)) == kFRedSrc - 1 || n == a.N - 1)) {
      // deterministic cross-warp sum of the buffered sources.  The buffer is next written in a later G phase,
      // i.e. behind two more barriers that these threads also have to pass.
      const int n0 = n & ~(kFRedSrc - 1);
      const int k = tid >> 4, j = tid & 15;
      if (k <= n - n0 && j < 12) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kFCH; ++w)
          if (ty0 + w < H) v += red[(k * kFCH + w) * 16 + j];       // rows below the image were skipped
        a.pose_part[(((size_t)b * a.slots_per_b + slot) * a.N + (n0 + k)) * 12 + j] = v;
      }
    }
  }

  if (GRAD && g_active && a.d_depth[l]) {
    const int gy = ty0 + cyy;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const int gx = tx0 + c0 + o;
      if (gy < H && gx < W)
        a.d_depth[l][(long long)b * P + gy * W + gx] = DERIVE == 2 ? gD[o] * ddepth_dlogit(sD[(cyy + 2) * kFP + c0 + o + 2]) : gD[o];
    }
  }

  // ---- loss sums -> one partial record per tile -----------------------------------------------------
  {
    const float v0 = warp_sum(lsum_l1), v1 = warp_sum(lsum_ssim), v2 = warp_sum(lsum_sm);
    __syncthreads();      // the last pose sum has read the scratch
    if (lane == 0) { red[wid * 3] = v0; red[wid * 3 + 1] = v1; red[wid * 3 + 2] = v2; }
    __syncthreads();
    if (tid < 3) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kFThreads / 32; ++w) v += red[w * 3 + tid];
      const float nrm = (tid == 2) ? 1.f : a.norm_photo[l];
      a.loss_part[((size_t)b * a.slots_per_b + slot) * 3 + tid] = v * nrm;
    }
  }
}

}  // namespace xpt
