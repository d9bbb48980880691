// xpt_fused.cuh -- the fused tile kernel of the total-loss path (north-star kernels 1+2+3+4):
// inverse warp -> shared-memory tile with halo -> L1 + SSIM (+ smoothness) -> SSIM/L1 adjoint
// -> bilinear + projection adjoint, one launch for all scales, all sources, forward AND backward.
//
// Design (DESIGN.md "k_fused"):
//   * one CTA = one 64x13 centre tile of (level, snippet); all N sources are looped inside so the
//     target tile and its window statistics are loaded / computed once;
//   * the 3x3 SSIM windows and their adjoint are evaluated on 4-pixel strips per thread from
//     16-byte aligned shared-memory rows (LDS.128 + LDS.64), vertical sums first, so each input is
//     loaded once per strip instead of nine times per pixel;
//   * the bilinear Jacobian (dS/du, dS/dv per channel) and (u, v, 1/den) of the centre samples are
//     cached in shared memory by the forward phase, so the backward phase gathers nothing again;
//   * region widths are multiples of 4 floats and the centre width is 64 so that 384/832/1280-wide
//     images (and every 2^k-scaled level down to 64) tile without waste.
#pragma once
#include "xpt_kernels.cuh"

namespace xpt {

constexpr int kFCW = 64, kFCH = 13;            // centre tile
constexpr int kFSW = 66, kFSH = 15;            // statistics region (halo 1)
constexpr int kFRW = 68, kFRH = 17;            // warped / target region (halo 2)
constexpr int kFP = 68;                        // row pitch (floats) of region and statistics arrays
constexpr int kFStrips = 17;                   // 4-pixel strips per statistics row
constexpr int kFThreads = 256;
constexpr int kFRegion = kFRH * kFP;           // 1156 floats per channel
constexpr int kFStats = kFSH * kFP;            // 1020
constexpr int kFCP = 64;                       // pitch of the centre arrays
constexpr int kFCentre = kFCH * kFCP;          // 832
constexpr int kFYIters = (kFRegion + kFThreads - 1) / kFThreads;   // 5

template <bool GRAD>
struct FusedSmem {
  // offsets in floats
  static constexpr int sy = 0;                                  // [3][17][68] warped tile
  static constexpr int sx = sy + 3 * kFRegion + 8;              // [3][17][68] target tile (+8: strip over-read)
  static constexpr int sD = sx + 3 * kFRegion + 8;              // [17][68] depth
  static constexpr int red = sD + kFRegion + 4;                 // 128 floats of reduction scratch
  static constexpr int sA = red + 128;                          // GRAD: [3][15][68] x3
  static constexpr int sB = sA + (GRAD ? 3 * kFStats + 8 : 0);
  static constexpr int sC = sB + (GRAD ? 3 * kFStats + 8 : 0);
  static constexpr int sGU = sC + (GRAD ? 3 * kFStats + 8 : 0); // GRAD: [3][13][64] dS_c/du
  static constexpr int sGV = sGU + (GRAD ? 3 * kFCentre : 0);
  static constexpr int sU = sGV + (GRAD ? 3 * kFCentre : 0);    // GRAD: [13][64] u, v, 1/den
  static constexpr int sV = sU + (GRAD ? kFCentre : 0);
  static constexpr int sI = sV + (GRAD ? kFCentre : 0);
  static constexpr int kFloats = sI + (GRAD ? kFCentre : 0);
  static constexpr size_t kBytes = sizeof(float) * kFloats;
};

__device__ __forceinline__ void lds6(const float* p, float v[6]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float2 b = *reinterpret_cast<const float2*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y;
}

__device__ __forceinline__ void lds4u(const float* p, float v[4]) {   // 8-byte aligned
  const float2 a = *reinterpret_cast<const float2*>(p);
  const float2 b = *reinterpret_cast<const float2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ float box_inv(int cnt) {   // 1 / #in-image taps of a 3x3 window
  return cnt == 9 ? (1.f / 9.f) : (cnt == 6 ? (1.f / 6.f) : (cnt == 4 ? 0.25f : 1.f / (float)cnt));
}

// sum of 12 per-thread accumulators over a warp in 16 shuffles: at each butterfly step a lane keeps
// half of the values and hands the other half to its partner.  Result k lands in lane (k*2) % 32 ... we
// only need the totals somewhere deterministic: afterwards lane L holds total[index_of(L)] in v[0].
__device__ __forceinline__ float warp_reduce16(float v[16], int lane) {
  // step 1: exchange 8 values with lane^16
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
  // lane L now holds the total of value index ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1)
  return v[0];
}

struct FusedArgs {
  LevelTable lt;                       // Level.tiles_x/y/slot_base describe the 64x13 tiling
  int B, N;
  int tiles_per_b;
  int first_tile[kMaxScales + 1];
  const float* geoK; const float* geoT;
  const float* depth[kMaxScales];
  const float* disp[kMaxScales];
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
  float* d_src[kMaxScales];
  long long d_src_bs[kMaxScales], d_src_fs[kMaxScales];
};

template <bool GRAD>
__global__ void __launch_bounds__(kFThreads, 2) k_fused(FusedArgs a) {
  using SM = FusedSmem<GRAD>;
  extern __shared__ __align__(16) float smem[];
  float* const sy = smem + SM::sy;
  float* const sx = smem + SM::sx;
  float* const sD = smem + SM::sD;
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
  int t = blockIdx.x;
  const int b = blockIdx.y;
  int l = 0;
  while (l + 1 < a.lt.S && t >= a.first_tile[l + 1]) ++l;
  t -= a.first_tile[l];
  const Level& L = a.lt.lv[l];
  const int H = L.H, W = L.W, P = H * W;
  const int ty0 = (t / L.tiles_x) * kFCH, tx0 = (t % L.tiles_x) * kFCW;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int slot = L.slot_base + t;

  float K[9], Ki[9];
  {
    const float* gk = a.geoK + ((size_t)b * a.lt.S + l) * kGeoK;
#pragma unroll
    for (int k = 0; k < 9; ++k) { K[k] = __ldg(gk + k); Ki[k] = __ldg(gk + 9 + k); }
  }

  // ---- target tile and depth tile (halo 2), zero outside the image --------------------
  {
    const float* tgt = L.tgt + b * L.tgt_bs;
    const float* dep = a.depth[l] + (long long)b * P;
#pragma unroll
    for (int it = 0; it < kFYIters; ++it) {
      int i = tid + it * kFThreads;
      if (i < kFRegion) {
        int ry = i / kFP, rx = i - ry * kFP;
        int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
        float v0 = 0.f, v1 = 0.f, v2 = 0.f, d = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
          const float* p = tgt + ((long long)gy * W + gx) * 3;
          v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2);
          d = __ldg(dep + (long long)gy * W + gx);
        }
        sx[i] = v0; sx[kFRegion + i] = v1; sx[2 * kFRegion + i] = v2;
        sD[i] = d;
      }
    }
  }
  __syncthreads();

  // ---- strip coordinates ----------------------------------------------------------------
  // S phase: statistics row qy (0..14), columns q0..q0+3
  const int qy = tid / kFStrips, q0 = (tid - qy * kFStrips) * 4;
  const bool s_active = tid < kFSH * kFStrips;
  // G phase: centre row cyy (0..12), columns c0..c0+3
  const int cyy = tid >> 4, c0 = (tid & 15) * 4;
  const bool g_active = cyy < kFCH;

  float lsum_l1 = 0.f, lsum_ssim = 0.f, lsum_sm = 0.f;

  // ---- smoothness on the centre strip (losses.py:409-440) ------------------------------------
  if (a.do_smooth && g_active) {
    const float* dsp = a.disp[l] + (long long)b * P;
    const float nx = a.norm_sm_x[l], ny = a.norm_sm_y[l];
    const float gcx = a.gcoef_smooth * nx, gcy = a.gcoef_smooth * ny;
    const float k3 = a.grad_factor;
    const int gy = ty0 + cyy;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int gx = tx0 + c0 + o;
      if (gy < H && gx < W) {
        const int ri = (cyy + 2) * kFP + (c0 + o + 2);
        const float d = __ldg(dsp + (long long)gy * W + gx);
        float gd = 0.f;
        if (gx + 1 < W) {
          float e = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) e += fabsf((sx[c * kFRegion + ri] - sx[c * kFRegion + ri + 1]) * k3);
          float w = expf(-(e / 3.f));
          float sd = (d - __ldg(dsp + (long long)gy * W + gx + 1)) * w;
          lsum_sm += fabsf(sd) * nx;
          gd += gcx * sgnf(sd) * w;
        }
        if (gy + 1 < H) {
          float e = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) e += fabsf((sx[c * kFRegion + ri] - sx[c * kFRegion + ri + kFP]) * k3);
          float w = expf(-(e / 3.f));
          float sd = (d - __ldg(dsp + (long long)(gy + 1) * W + gx)) * w;
          lsum_sm += fabsf(sd) * ny;
          gd += gcy * sgnf(sd) * w;
        }
        if (GRAD) {
          if (gx >= 1) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) e += fabsf((sx[c * kFRegion + ri - 1] - sx[c * kFRegion + ri]) * k3);
            float w = expf(-(e / 3.f));
            float sd = (__ldg(dsp + (long long)gy * W + gx - 1) - d) * w;
            gd -= gcx * sgnf(sd) * w;
          }
          if (gy >= 1) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) e += fabsf((sx[c * kFRegion + ri - kFP] - sx[c * kFRegion + ri]) * k3);
            float w = expf(-(e / 3.f));
            float sd = (__ldg(dsp + (long long)(gy - 1) * W + gx) - d) * w;
            gd -= gcy * sgnf(sd) * w;
          }
          if (a.d_disp[l]) a.d_disp[l][(long long)b * P + gy * W + gx] = gd;
        }
      }
    }
  }

  // ---- per-strip invariants of the S phase ------------------------------------------------------
  float inv_cnt[4];
  bool s_in[4], s_centre[4];
  {
    const int gy = ty0 - 1 + qy;
    const bool row_in = s_active && gy >= 0 && gy < H;
    const int cy = min(gy + 1, H - 1) - max(gy - 1, 0) + 1;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int q = q0 + o, gx = tx0 - 1 + q;
      s_in[o] = row_in && q < kFSW && gx >= 0 && gx < W;
      const int cx = min(gx + 1, W - 1) - max(gx - 1, 0) + 1;
      inv_cnt[o] = s_in[o] ? box_inv(cy * cx) : 0.f;
      s_centre[o] = s_in[o] && qy >= 1 && qy <= kFCH && q >= 1 && q <= kFCW;
    }
  }
  // window statistics of the target (x) for this strip: evaluated once, kept across the N sources
  float MUX[3][4], SGX[3][4];
  if (a.do_ssim && s_active) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* px = sx + c * kFRegion + qy * kFP + q0;
      float X0[6], X1[6], X2[6], v1[6], v2[6];
      lds6(px, X0); lds6(px + kFP, X1); lds6(px + 2 * kFP, X2);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        v1[j] = X0[j] + X1[j] + X2[j];
        v2[j] = fmaf(X2[j], X2[j], fmaf(X1[j], X1[j], X0[j] * X0[j]));
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        float mu = (v1[o] + v1[o + 1] + v1[o + 2]) * inv_cnt[o];
        MUX[c][o] = mu;
        SGX[c][o] = (v2[o] + v2[o + 1] + v2[o + 2]) * inv_cnt[o] - mu * mu;
      }
    }
  }

  // ---- per-strip invariants of the G phase ------------------------------------------------------
  float gD[4] = {0.f, 0.f, 0.f, 0.f};
  const float cl1 = a.gcoef_l1 * a.norm_photo[l];
  const float hss = -0.5f * a.gcoef_ssim * a.norm_photo[l];     // dTotal/d ssim at a contributing pixel

  for (int n = 0; n < a.N; ++n) {
    float T[12];
    {
      const float* gt = a.geoT + ((size_t)b * a.N + n) * kGeoT;
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] = __ldg(gt + k);
    }
    const float* img = L.src + b * L.src_bs + n * L.src_fs;

    // ---- phase Y: inverse warp of the region into shared memory ---------------------------------
#pragma unroll
    for (int it = 0; it < kFYIters; ++it) {
      const int i = tid + it * kFThreads;
      if (i < kFRegion) {
        const int ry = i / kFP, rx = i - ry * kFP;
        const int gy = ty0 - 2 + ry, gx = tx0 - 2 + rx;
        const bool centre = ry >= 2 && ry < 2 + kFCH && rx >= 2 && rx < 2 + kFCW;
        float yv[3] = {0.f, 0.f, 0.f};
        float gu[3] = {0.f, 0.f, 0.f}, gv[3] = {0.f, 0.f, 0.f};
        float su = 0.f, sv = 0.f, si = 0.f;
        bool valid = false;
        const bool inimg = gy >= 0 && gy < H && gx >= 0 && gx < W;
        if (inimg) {
          const float D = sD[i];
          float r0, r1, r2;
          ray_of_pixel(Ki, (float)gx, (float)gy, r0, r1, r2);
          const Proj pr = project(K, T, r0, r1, r2, D);
          const Taps tp = make_taps(pr.u, pr.v, D, W, H);
          valid = tp.valid;
          if (valid) {
            float I0[3], I1[3], I2[3], I3[3];
            gather_taps(img, W, tp, I0, I1, I2, I3);
            const float w0 = tp.w_uf * tp.w_vf, w1 = tp.w_uf * tp.w_vc, w2 = tp.w_uc * tp.w_vf, w3 = tp.w_uc * tp.w_vc;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              yv[c] = ((I0[c] * w0 + I1[c] * w1) + I2[c] * w2) + I3[c] * w3;
              if (GRAD) {
                gu[c] = tp.w_vf * (I2[c] - I0[c]) + tp.w_vc * (I3[c] - I1[c]);
                gv[c] = tp.w_uf * (I1[c] - I0[c]) + tp.w_uc * (I3[c] - I2[c]);
              }
            }
            su = pr.u; sv = pr.v; si = 1.f / pr.den;
          }
        }
        sy[i] = yv[0]; sy[kFRegion + i] = yv[1]; sy[2 * kFRegion + i] = yv[2];
        if (centre) {
          if (GRAD) {
            const int ci = (ry - 2) * kFCP + (rx - 2);
            sGU[ci] = gu[0]; sGU[kFCentre + ci] = gu[1]; sGU[2 * kFCentre + ci] = gu[2];
            sGV[ci] = gv[0]; sGV[kFCentre + ci] = gv[1]; sGV[2 * kFCentre + ci] = gv[2];
            sU[ci] = su; sV[ci] = sv; sI[ci] = si;
          }
          if (inimg) {
            const long long o = (long long)(b * a.N + n) * P + gy * W + gx;
            if (a.synth_out[l]) { float* so = a.synth_out[l] + o * 3; so[0] = yv[0]; so[1] = yv[1]; so[2] = yv[2]; }
            if (a.mask_out[l]) a.mask_out[l][o] = valid ? 1.f : 0.f;
          }
        }
      }
    }
    __syncthreads();

    // ---- phase S: L1 + SSIM (and the SSIM adjoint coefficients) on a 4-pixel strip ----------------
    if (s_active) {
      bool masked[4];
      {
        // mean_c(synth) == 0 (loss_util.py:15-16): the pixel itself sits at region (qy+1, q0+o+1)
        float m0[6], m1[6], m2[6];
        lds6(sy + (qy + 1) * kFP + q0, m0);
        lds6(sy + kFRegion + (qy + 1) * kFP + q0, m1);
        lds6(sy + 2 * kFRegion + (qy + 1) * kFP + q0, m2);
#pragma unroll
        for (int o = 0; o < 4; ++o) masked[o] = ((m0[o + 1] + m1[o + 1]) + m2[o + 1]) == 0.f;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* py = sy + c * kFRegion + qy * kFP + q0;
        const float* px = sx + c * kFRegion + qy * kFP + q0;
        float Y0[6], Y1[6], Y2[6], X0[6], X1[6], X2[6];
        lds6(py, Y0); lds6(py + kFP, Y1); lds6(py + 2 * kFP, Y2);
        lds6(px, X0); lds6(px + kFP, X1); lds6(px + 2 * kFP, X2);
        if (a.do_l1) {
#pragma unroll
          for (int o = 0; o < 4; ++o)
            if (s_centre[o] && !masked[o]) lsum_l1 += fabsf(Y1[o + 1] - X1[o + 1]);
        }
        if (a.do_ssim) {
          float v1[6], v2[6], v3[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            v1[j] = Y0[j] + Y1[j] + Y2[j];
            v2[j] = fmaf(Y2[j], Y2[j], fmaf(Y1[j], Y1[j], Y0[j] * Y0[j]));
            v3[j] = fmaf(X2[j], Y2[j], fmaf(X1[j], Y1[j], X0[j] * Y0[j]));
          }
          float Ao[4], Bo[4], Co[4];
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const float inv = inv_cnt[o];
            const float mux = MUX[c][o], sgx = SGX[c][o];
            const float muy = (v1[o] + v1[o + 1] + v1[o + 2]) * inv;
            const float sgy = (v2[o] + v2[o + 1] + v2[o + 2]) * inv - muy * muy;
            const float sgxy = (v3[o] + v3[o + 1] + v3[o + 2]) * inv - mux * muy;
            const float a1 = 2.f * mux * muy + kC1, a2 = 2.f * sgxy + kC2;
            const float b1 = mux * mux + muy * muy + kC1, b2 = sgx + sgy + kC2;
            const float r12 = 1.f / (b1 * b2);
            const float ssim = (a1 * a2) * r12;
            const float lv = (1.f - ssim) * 0.5f;
            const bool live = s_in[o] && !masked[o];
            if (s_centre[o] && !masked[o]) lsum_ssim += fminf(fmaxf(lv, 0.f), 1.f);
            float A = 0.f, Bq = 0.f, Cq = 0.f;
            if (GRAD && live && lv >= 0.f && lv <= 1.f) {        // clip_by_value passes the gradient inside [0,1]
              const float rb1 = r12 * b2, rb2 = r12 * b1;        // 1/b1, 1/b2
              const float hi = hss * inv;
              A = hi * ((2.f * mux * (a2 - a1)) * r12 - ssim * (2.f * muy) * (rb1 - rb2));
              Bq = hi * (-ssim * rb2);
              Cq = hi * (2.f * a1 * r12);
            }
            Ao[o] = A; Bo[o] = Bq; Co[o] = Cq;
          }
          if (GRAD) {
            const int so = c * kFStats + qy * kFP + q0;
            *reinterpret_cast<float4*>(sA + so) = make_float4(Ao[0], Ao[1], Ao[2], Ao[3]);
            *reinterpret_cast<float4*>(sB + so) = make_float4(Bo[0], Bo[1], Bo[2], Bo[3]);
            *reinterpret_cast<float4*>(sC + so) = make_float4(Co[0], Co[1], Co[2], Co[3]);
          }
        }
      }
    }

    // ---- phase G: dL/dS on the centre strip, pushed through the bilinear + projection adjoint ------
    if (GRAD) {
      __syncthreads();
      float acc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = 0.f;
      if (g_active) {
        const int gy = ty0 + cyy;
        const int rrow = (cyy + 2) * kFP + c0 + 2;      // centre pixel in region coordinates
        float yv[3][4], xv[3][4], g[3][4];
        bool masked[4];
#pragma unroll
        for (int c = 0; c < 3; ++c) { lds4u(sy + c * kFRegion + rrow, yv[c]); lds4u(sx + c * kFRegion + rrow, xv[c]); }
#pragma unroll
        for (int o = 0; o < 4; ++o) masked[o] = ((yv[0][o] + yv[1][o]) + yv[2][o]) == 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {0.f, 0.f, 0.f, 0.f};
          if (a.do_ssim) {
            // centre (cyy, c0+o) = statistics (cyy+1, c0+o+1): its window is statistics rows cyy..cyy+2,
            // columns c0+o..c0+o+2
            const int so = c * kFStats + cyy * kFP + c0;
            float R0[6], R1[6], R2[6], v[6];
            lds6(sA + so, R0); lds6(sA + so + kFP, R1); lds6(sA + so + 2 * kFP, R2);
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] = R0[j] + R1[j] + R2[j];
#pragma unroll
            for (int o = 0; o < 4; ++o) sa[o] = v[o] + v[o + 1] + v[o + 2];
            lds6(sB + so, R0); lds6(sB + so + kFP, R1); lds6(sB + so + 2 * kFP, R2);
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] = R0[j] + R1[j] + R2[j];
#pragma unroll
            for (int o = 0; o < 4; ++o) sb[o] = v[o] + v[o + 1] + v[o + 2];
            lds6(sC + so, R0); lds6(sC + so + kFP, R1); lds6(sC + so + 2 * kFP, R2);
#pragma unroll
            for (int j = 0; j < 6; ++j) v[j] = R0[j] + R1[j] + R2[j];
#pragma unroll
            for (int o = 0; o < 4; ++o) sc[o] = v[o] + v[o + 1] + v[o + 2];
          }
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            float gc = sa[o] + 2.f * yv[c][o] * sb[o] + xv[c][o] * sc[o];
            if (a.do_l1 && !masked[o]) gc += cl1 * sgnf(yv[c][o] - xv[c][o]);
            g[c][o] = gc;
          }
        }
        const int ci = cyy * kFCP + c0;
        float GUv[3][4], GVv[3][4], U[4], V[4], IV[4];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float4 p = *reinterpret_cast<const float4*>(sGU + c * kFCentre + ci);
          const float4 q = *reinterpret_cast<const float4*>(sGV + c * kFCentre + ci);
          GUv[c][0] = p.x; GUv[c][1] = p.y; GUv[c][2] = p.z; GUv[c][3] = p.w;
          GVv[c][0] = q.x; GVv[c][1] = q.y; GVv[c][2] = q.z; GVv[c][3] = q.w;
        }
        {
          const float4 p = *reinterpret_cast<const float4*>(sU + ci);
          const float4 q = *reinterpret_cast<const float4*>(sV + ci);
          const float4 r = *reinterpret_cast<const float4*>(sI + ci);
          U[0] = p.x; U[1] = p.y; U[2] = p.z; U[3] = p.w;
          V[0] = q.x; V[1] = q.y; V[2] = q.z; V[3] = q.w;
          IV[0] = r.x; IV[1] = r.y; IV[2] = r.z; IV[3] = r.w;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int gx = tx0 + c0 + o;
          if (gy < H && gx < W) {
            const float gu = g[0][o] * GUv[0][o] + g[1][o] * GUv[1][o] + g[2][o] * GUv[2][o];
            const float gv = g[0][o] * GVv[0][o] + g[1][o] * GVv[1][o] + g[2][o] * GVv[2][o];
            const float D = sD[rrow + o];
            float r0, r1, r2;
            ray_of_pixel(Ki, (float)gx, (float)gy, r0, r1, r2);
            const float X0 = r0 * D, X1 = r1 * D, X2 = r2 * D;
            const float inv = IV[o];
            const float gp0 = gu * inv, gp1 = gv * inv, gp2 = -(gu * U[o] + gv * V[o]) * inv;
            const float gY0 = K[0] * gp0 + K[3] * gp1 + K[6] * gp2;
            const float gY1 = K[1] * gp0 + K[4] * gp1 + K[7] * gp2;
            const float gY2 = K[2] * gp0 + K[5] * gp1 + K[8] * gp2;
            acc[0] += gY0 * X0; acc[1] += gY0 * X1; acc[2] += gY0 * X2;
            acc[3] += gY1 * X0; acc[4] += gY1 * X1; acc[5] += gY1 * X2;
            acc[6] += gY2 * X0; acc[7] += gY2 * X1; acc[8] += gY2 * X2;
            acc[9] += gY0; acc[10] += gY1; acc[11] += gY2;
            const float gX0 = T[0] * gY0 + T[3] * gY1 + T[6] * gY2;
            const float gX1 = T[1] * gY0 + T[4] * gY1 + T[7] * gY2;
            const float gX2 = T[2] * gY0 + T[5] * gY1 + T[8] * gY2;
            gD[o] += gX0 * r0 + gX1 * r1 + gX2 * r2;
            if (a.d_src[l]) {
              // dL/dsource: re-derive the taps from the cached coordinates (bit-identical to the forward)
              const Taps tp = make_taps(U[o], V[o], D, W, H);
              if (tp.valid && inv != 0.f) {
                float* dimg = a.d_src[l] + b * a.d_src_bs[l] + n * a.d_src_fs[l];
                const float w0 = tp.w_uf * tp.w_vf, w1 = tp.w_uf * tp.w_vc, w2 = tp.w_uc * tp.w_vf, w3 = tp.w_uc * tp.w_vc;
                float* p = dimg + ((long long)tp.iv * W + tp.iu) * 3;
                float* q = p + (long long)W * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  atomicAdd(p + c, w0 * g[c][o]);
                  atomicAdd(p + 3 + c, w2 * g[c][o]);
                  atomicAdd(q + c, w1 * g[c][o]);
                  atomicAdd(q + 3 + c, w3 * g[c][o]);
                }
              }
            }
          }
        }
      }
      // 12 pose accumulators: warp reduction in 16 shuffles, then one deterministic cross-warp sum
      const float tot = warp_reduce16(acc, lane);
      if ((lane & 1) == 0) red[wid * 16 + (((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1))] = tot;
      __syncthreads();
      if (tid < 12) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kFThreads / 32; ++w) v += red[w * 16 + tid];
        a.pose_part[(((size_t)b * a.slots_per_b + slot) * a.N + n) * 12 + tid] = v;
      }
    }
    __syncthreads();      // sy / sA.. / sGU.. / red are rewritten by the next source
  }

  if (GRAD && g_active && a.d_depth[l]) {
    const int gy = ty0 + cyy;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int gx = tx0 + c0 + o;
      if (gy < H && gx < W) a.d_depth[l][(long long)b * P + gy * W + gx] = gD[o];
    }
  }

  // ---- loss sums -> one partial record per tile -----------------------------------------------------
  {
    const float v0 = warp_sum(lsum_l1), v1 = warp_sum(lsum_ssim), v2 = warp_sum(lsum_sm);
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
