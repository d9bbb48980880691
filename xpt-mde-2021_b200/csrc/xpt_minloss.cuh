// xpt_minloss.cuh -- per-pixel minimum over sources at full resolution, the ROUND-1 kernel (k_photo_min).  Since round 2 the
// default is k_min_strip (xpt_minstrip.cuh); this one serves XPT_FLAG_MIN_TILES (A/B parity tests) and scale sets with a
// level between full and half resolution.  The argument struct and up_taps() below are shared by both kernels.
//
// Per-pixel minimum over sources at full resolution:
// MonoDepth2LossMultiScale (reference losses.py:198-232) and MoALossMultiScale (:282-321).
//
// For every scale the synthesised views [B,N,h,w,3] (plus, for MoA, the stereo synthesis [B,1,h,w,3]) are
// bilinearly up-sampled to the full resolution (resize_bilinear, losses.py:377-383: tf.image.resize with
// half-pixel centres), the photometric term is evaluated PER PIXEL AND CHANNEL against the full-resolution
// target (loss_util.py with reduce=False), the minimum over the sources is taken (tf.reduce_min, axis=1) and
// averaged over H, W, 3.  Quirk kept: a black (invalid) synthesised pixel has loss 0 and wins.
//
// One CTA = one 32x16 full-resolution tile of one (scale, snippet).  The up-sampled tile is produced on the fly
// from the low-resolution synthesis (never materialised in HBM).  Sweep 1 over the sources finds the minimum
// and the number of ties per (pixel, channel); with GRAD a second sweep re-evaluates each source, routes the
// upstream gradient to the winners (tf.reduce_min splits it equally among ties), applies the SSIM / L1 adjoint
// and scatters through the adjoint of the up-sampling into d_synth (fp32 atomics: the one non-deterministic
// summation order of this row).
#pragma once
#include "xpt_kernels.cuh"

namespace xpt {

struct MinLossArgs {
  int B, N, NS;                        // temporal sources, stereo sources (0 or 1)
  int H, W;                            // full resolution
  int S;
  int h[kMaxScales], w[kMaxScales];
  const float* synth[kMaxScales];      // [B,N,h,w,3]
  const float* stereo[kMaxScales];     // [B,NS,h,w,3] (NS > 0)
  const float* target; long long tgt_bs;
  int method;                          // 0 L1, 1 L2, 2 SSIM
  float norm[kMaxScales];              // sw_s / (H*W*3)
  const float* gbatch;                 // upstream dL/d loss_batch[b] (NULL = 1)
  int tiles_x, tiles;                  // full-resolution tiling
  float* loss_part;                    // [B][S*tiles]
  float* gsynth[kMaxScales];           // GRAD: [B,N,h,w,3], zeroed by the host, accumulated atomically
  float* gstereo[kMaxScales];          // GRAD: [B,NS,h,w,3]
  // CombinedLossMultiScale (losses.py:235-279): instead of the minimum over sources, every source's term counts
  // where it is smaller than the term of the flow-warped view `cmb_flow` [B,N,cmb_h,cmb_w,3] (level 0 of
  // warped_target_ms), both up-sampled to H x W; norm[] then carries 1/(N*H*W*3).  NULL = min-over-sources mode.
  const float* cmb_flow; int cmb_h, cmb_w;
  // k_min_strip<.., PAIR>: L1 and SSIM of one loss set in one launch -- loss_part holds L1, loss_part2 SSIM, the
  // gradient is pair_c_l1 dL1 + pair_c_ssim dSSIM (the upstream weights of the two losses, uniform over the batch)
  float* loss_part2; float pair_c_l1, pair_c_ssim;
};

template <bool GRAD>
struct MinLossSmem {
  using P = PhotoSmem<GRAD>;
  // PhotoSmem layout (sx, sy, mu_x, sigma_x, A, B, C, red) + minimum and tie count per (channel, stats position)
  static constexpr int kMinOff = P::kFloats;
  static constexpr int kFloats = P::kFloats + 3 * P::kStats * 2;
  static constexpr size_t kBytes = sizeof(float) * kFloats;
};

// tf.image.resize(bilinear, half-pixel centres) source taps of one output coordinate
__device__ __forceinline__ void up_taps(int o, int in_size, float scale, int& lo, int& hi, float& lerp) {
  const float pos = ((float)o + 0.5f) * scale - 0.5f;
  const float fl = floorf(pos);
  lo = max((int)fl, 0);
  hi = min((int)ceilf(pos), in_size - 1);
  lerp = pos - fl;
}

template <bool GRAD>
__global__ void __launch_bounds__(kPhotoThreads, 3) k_photo_min(MinLossArgs a) {
  using SM = PhotoSmem<GRAD>;
  constexpr int HL = SM::HL, RW = SM::RW, SW = SM::SW;
  constexpr int HS = HL - 1;
  extern __shared__ float smem[];
  float* sx = smem;
  float* sy = sx + 3 * SM::kRegion;
  float* smx = sy + 3 * SM::kRegion;
  float* ssx = smx + 3 * SM::kStats;
  float* sA = ssx + 3 * SM::kStats;
  float* sB = sA + (GRAD ? 3 * SM::kStats : 0);
  float* sC = sB + (GRAD ? 3 * SM::kStats : 0);
  float* red = sC + (GRAD ? 3 * SM::kStats : 0);
  float* smin = smem + MinLossSmem<GRAD>::kMinOff;      // [3][kStats]
  float* scnt = smin + 3 * SM::kStats;                  // [3][kStats] number of sources attaining the minimum

  const int l = blockIdx.x / a.tiles, t = blockIdx.x % a.tiles;
  const int b = blockIdx.y;
  const int H = a.H, W = a.W, h = a.h[l], w = a.w[l];
  const int ty0 = (t / a.tiles_x) * kTH, tx0 = (t % a.tiles_x) * kTW;
  const int tid = threadIdx.x;
  const float sc_y = (float)h / (float)H, sc_x = (float)w / (float)W;
  const bool ssim = a.method == 2;

  // ---- target tile + window statistics -----------------------------------------------------------------
  const float* tgt = a.target + b * a.tgt_bs;
  for (int i = tid; i < SM::kRegion; i += kPhotoThreads) {
    const int ry = i / RW, rx = i % RW;
    const int gy = ty0 + ry - HL, gx = tx0 + rx - HL;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const float* p = tgt + ((long long)gy * W + gx) * 3;
      v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2);
    }
    sx[i] = v0; sx[SM::kRegion + i] = v1; sx[2 * SM::kRegion + i] = v2;
  }
  for (int i = tid; i < 3 * SM::kStats; i += kPhotoThreads) { smin[i] = 3.0e38f; scnt[i] = 0.f; }
  __syncthreads();
  if (ssim) {
    for (int i = tid; i < SM::kStats; i += kPhotoThreads) {
      const int qy = i / SW, qx = i % SW;
      const int gy = ty0 + qy - HS, gx = tx0 + qx - HS;
      const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const int cy = min(gy + 1, H - 1) - max(gy - 1, 0) + 1, cx = min(gx + 1, W - 1) - max(gx - 1, 0) + 1;
      const float inv = in ? 1.f / (float)(cy * cx) : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* px = sx + c * SM::kRegion + qy * RW + qx;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) { const float v = px[dy * RW + dx]; s1 += v; s2 += v * v; }
        const float mu = s1 * inv;
        smx[c * SM::kStats + i] = mu;
        ssx[c * SM::kStats + i] = s2 * inv - mu * mu;
      }
    }
  }
  const float gb = (GRAD && a.gbatch) ? __ldg(a.gbatch + b) : 1.f;
  const float coef = gb * a.norm[l];
  const int nsrc = a.N + a.NS;

  // up-sampled region of a low-resolution view [h,w,3] into sy (zero outside the image)
  auto load_low = [&](const float* low, int h, int w) {
    const float sc_y = (float)h / (float)H, sc_x = (float)w / (float)W;
    for (int i = tid; i < SM::kRegion; i += kPhotoThreads) {
      const int ry = i / RW, rx = i % RW;
      const int gy = ty0 + ry - HL, gx = tx0 + rx - HL;
      float yv[3] = {0.f, 0.f, 0.f};
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        int y0, y1, x0, x1; float fy, fx;
        up_taps(gy, h, sc_y, y0, y1, fy);
        up_taps(gx, w, sc_x, x0, x1, fx);
        const float* r0 = low + (size_t)y0 * w * 3;
        const float* r1 = low + (size_t)y1 * w * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float tl = __ldg(r0 + x0 * 3 + c), tr = __ldg(r0 + x1 * 3 + c);
          const float bl = __ldg(r1 + x0 * 3 + c), br = __ldg(r1 + x1 * 3 + c);
          const float top = tl + (tr - tl) * fx;
          const float bot = bl + (br - bl) * fx;
          yv[c] = top + (bot - top) * fy;
        }
      }
      sy[i] = yv[0]; sy[SM::kRegion + i] = yv[1]; sy[2 * SM::kRegion + i] = yv[2];
    }
  };
  auto load_region = [&](int m) {
    load_low(m < a.N ? a.synth[l] + ((size_t)b * a.N + m) * h * w * 3
                     : a.stereo[l] + ((size_t)b * a.NS + (m - a.N)) * h * w * 3, h, w);
  };

  // per-pixel, per-channel loss of the source currently in sy at stats position i (in-image positions only);
  // for SSIM also returns what the adjoint needs
  struct Term { float val; float dm, dq, dr; bool pass; };
  auto term = [&](int i, int c, bool masked, float inv) -> Term {
    Term r; r.val = 0.f; r.dm = r.dq = r.dr = 0.f; r.pass = false;
    const int qy = i / SW, qx = i % SW;
    const int ri = (qy + 1) * RW + (qx + 1);
    if (masked) return r;                                   // tf.where(mask, 0, loss): value 0, no gradient
    if (!ssim) {
      const float df = sy[c * SM::kRegion + ri] - sx[c * SM::kRegion + ri];
      r.val = a.method == 0 ? fabsf(df) : df * df;
      r.pass = true;
      r.dm = df;
      return r;
    }
    const float* px = sx + c * SM::kRegion + qy * RW + qx;
    const float* py = sy + c * SM::kRegion + qy * RW + qx;
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float yv = py[dy * RW + dx], xv = px[dy * RW + dx];
        s1 += yv; s2 += yv * yv; s3 += xv * yv;
      }
    const float mux = smx[c * SM::kStats + i], sgx = ssx[c * SM::kStats + i];
    const float muy = s1 * inv;
    const float sgy = s2 * inv - muy * muy;
    const float sgxy = s3 * inv - mux * muy;
    const float a1 = 2.f * mux * muy + kC1, a2 = 2.f * sgxy + kC2;
    const float b1 = mux * mux + muy * muy + kC1, b2 = sgx + sgy + kC2;
    const float sv = (a1 * a2) / (b1 * b2);
    const float lv = (1.f - sv) * 0.5f;
    r.pass = (lv >= 0.f) && (lv <= 1.f);
    r.val = fminf(fmaxf(lv, 0.f), 1.f);
    if (GRAD) {
      const float ib = 1.f / (b1 * b2);
      r.dm = (2.f * mux * (a2 - a1)) * ib - sv * (2.f * muy) * (1.f / b1 - 1.f / b2);
      r.dq = -sv / b2;
      r.dr = 2.f * a1 * ib;
    }
    return r;
  };

  // dL/dS of the source in sy from the coefficient planes sA/sB/sC (SSIM: box-summed over the 3x3 window), pushed
  // through the adjoint of the up-sampling into the low-resolution gradient (fp32 atomics in L2; zero-weight taps are
  // skipped).  A level at full resolution (identity up-sampling) owns its pixels: plain stores.
  // (Staging the tile's contributions in shared memory first was measured and is slower: fp32 shared-memory atomics
  // serialise on the few low-resolution values a coarse level's tile maps to -- 394 vs 307 us at config 2.)
  const bool identity = (h == H) && (w == W);
  auto scatter = [&](int m) {
    float* glow = m < a.N ? a.gsynth[l] + ((size_t)b * a.N + m) * h * w * 3
                          : a.gstereo[l] + ((size_t)b * a.NS + (m - a.N)) * h * w * 3;
    for (int i = tid; i < kTW * kTH; i += kPhotoThreads) {
      const int cy_ = i / kTW, cx_ = i % kTW;
      const int gy = ty0 + cy_, gx = tx0 + cx_;
      if (gy >= H || gx >= W) continue;
      const int ri = (cy_ + HL) * RW + (cx_ + HL);
      const int si = (cy_ + HS) * SW + (cx_ + HS);
      float g[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (ssim) {
          const float yv = sy[c * SM::kRegion + ri], xv = sx[c * SM::kRegion + ri];
          float gc = 0.f;
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
              const int k = c * SM::kStats + si + dy * SW + dx;
              gc += sA[k] + 2.f * yv * sB[k] + xv * sC[k];
            }
          g[c] = gc;
        } else {
          g[c] = sA[c * SM::kStats + si];
        }
      }
      if (g[0] == 0.f && g[1] == 0.f && g[2] == 0.f) continue;
      if (identity) {
        float* o = glow + ((size_t)gy * w + gx) * 3;
        o[0] = g[0]; o[1] = g[1]; o[2] = g[2];
        continue;
      }
      // adjoint of the up-sampling: top = tl + (tr - tl) fx, out = top + (bot - top) fy
      int y0, y1, x0, x1; float fy, fx;
      up_taps(gy, h, sc_y, y0, y1, fy);
      up_taps(gx, w, sc_x, x0, x1, fx);
      const float wt[4] = {(1.f - fy) * (1.f - fx), (1.f - fy) * fx, fy * (1.f - fx), fy * fx};
      const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (wt[k] == 0.f) continue;
        float* o = glow + ((size_t)ys[k] * w + xs[k]) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) atomicAdd(o + c, wt[k] * g[c]);
      }
    }
  };

  // ---- CombinedLossMultiScale: per source, the static term where it beats the flow term ----------------------
  if (a.cmb_flow) {
    float lsum = 0.f;
    for (int m = 0; m < a.N; ++m) {
      __syncthreads();
      load_low(a.cmb_flow + ((size_t)b * a.N + m) * a.cmb_h * a.cmb_w * 3, a.cmb_h, a.cmb_w);
      __syncthreads();
      for (int i = tid; i < SM::kStats; i += kPhotoThreads) {       // flow term of this source -> smin
        const int qy = i / SW, qx = i % SW;
        const int gy = ty0 + qy - HS, gx = tx0 + qx - HS;
        if (!(gy >= 0 && gy < H && gx >= 0 && gx < W)) continue;
        const int ri = (qy + 1) * RW + (qx + 1);
        const bool masked = ((sy[ri] + sy[SM::kRegion + ri]) + sy[2 * SM::kRegion + ri]) == 0.f;
        const int cy = min(gy + 1, H - 1) - max(gy - 1, 0) + 1, cx = min(gx + 1, W - 1) - max(gx - 1, 0) + 1;
        const float inv = 1.f / (float)(cy * cx);
#pragma unroll
        for (int c = 0; c < 3; ++c) smin[c * SM::kStats + i] = term(i, c, masked, inv).val;
      }
      __syncthreads();
      load_region(m);
      __syncthreads();
      for (int i = tid; i < SM::kStats; i += kPhotoThreads) {
        const int qy = i / SW, qx = i % SW;
        const int gy = ty0 + qy - HS, gx = tx0 + qx - HS;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
        const bool centre = in && qy >= HS && qy < HS + kTH && qx >= HS && qx < HS + kTW;
        const int ri = (qy + 1) * RW + (qx + 1);
        const bool masked = ((sy[ri] + sy[SM::kRegion + ri]) + sy[2 * SM::kRegion + ri]) == 0.f;
        const int cy = min(gy + 1, H - 1) - max(gy - 1, 0) + 1, cx = min(gx + 1, W - 1) - max(gx - 1, 0) + 1;
        const float inv = in ? 1.f / (float)(cy * cx) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float A = 0.f, Bq = 0.f, Cq = 0.f;
          if (in) {
            const Term r = term(i, c, masked, inv);
            const bool keep = r.val < smin[c * SM::kStats + i];     // tf.cast(static_loss < flow_loss): a constant
            if (keep && centre) lsum += r.val;
            if (GRAD && keep && r.pass && !masked) {
              if (ssim) { const float hh = -0.5f * coef; A = hh * r.dm * inv; Bq = hh * r.dq * inv; Cq = hh * r.dr * inv; }
              else A = a.method == 0 ? coef * sgnf(r.dm) : coef * 2.f * r.dm;
            }
          }
          if (GRAD) { sA[c * SM::kStats + i] = A; sB[c * SM::kStats + i] = Bq; sC[c * SM::kStats + i] = Cq; }
        }
      }
      if constexpr (GRAD) {
        __syncthreads();
        scatter(m);
      }
    }
    lsum = warp_sum(lsum);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = lsum;
    __syncthreads();
    if (tid == 0) {
      float v = 0.f;
      for (int k = 0; k < kPhotoThreads / 32; ++k) v += red[k];
      a.loss_part[(size_t)b * a.S * a.tiles + blockIdx.x] = v * a.norm[l];
    }
    return;
  }

  // ---- sweep 1: minimum and tie count per (pixel, channel) over the stats region -------------------------
  for (int m = 0; m < nsrc; ++m) {
    load_region(m);
    __syncthreads();
    for (int i = tid; i < SM::kStats; i += kPhotoThreads) {
      const int qy = i / SW, qx = i % SW;
      const int gy = ty0 + qy - HS, gx = tx0 + qx - HS;
      if (!(gy >= 0 && gy < H && gx >= 0 && gx < W)) continue;
      const int ri = (qy + 1) * RW + (qx + 1);
      const bool masked = ((sy[ri] + sy[SM::kRegion + ri]) + sy[2 * SM::kRegion + ri]) == 0.f;
      const int cy = min(gy + 1, H - 1) - max(gy - 1, 0) + 1, cx = min(gx + 1, W - 1) - max(gx - 1, 0) + 1;
      const float inv = 1.f / (float)(cy * cx);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = term(i, c, masked, inv).val;
        const float cur = smin[c * SM::kStats + i];
        if (v < cur) { smin[c * SM::kStats + i] = v; scnt[c * SM::kStats + i] = 1.f; }
        else if (v == cur) scnt[c * SM::kStats + i] += 1.f;
      }
    }
    __syncthreads();
  }
  // loss of this tile: sum of the minima over its in-image centre pixels
  {
    float lsum = 0.f;
    for (int i = tid; i < kTW * kTH; i += kPhotoThreads) {
      const int cy_ = i / kTW, cx_ = i % kTW;
      if (ty0 + cy_ < H && tx0 + cx_ < W) {
        const int si = (cy_ + HS) * SW + (cx_ + HS);
        lsum += (smin[si] + smin[SM::kStats + si]) + smin[2 * SM::kStats + si];
      }
    }
    lsum = warp_sum(lsum);
    if ((tid & 31) == 0) red[tid >> 5] = lsum;
    __syncthreads();
    if (tid == 0) {
      float v = 0.f;
      for (int k = 0; k < kPhotoThreads / 32; ++k) v += red[k];
      a.loss_part[(size_t)b * a.S * a.tiles + blockIdx.x] = v * a.norm[l];
    }
  }
  // ---- sweep 2: gradient of the winners -------------------------------------------------------------------
  if constexpr (GRAD)
  for (int m = 0; m < nsrc; ++m) {
    __syncthreads();
    load_region(m);
    __syncthreads();
    for (int i = tid; i < SM::kStats; i += kPhotoThreads) {
      const int qy = i / SW, qx = i % SW;
      const int gy = ty0 + qy - HS, gx = tx0 + qx - HS;
      const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
      const int ri = (qy + 1) * RW + (qx + 1);
      const bool masked = ((sy[ri] + sy[SM::kRegion + ri]) + sy[2 * SM::kRegion + ri]) == 0.f;
      const int cy = min(gy + 1, H - 1) - max(gy - 1, 0) + 1, cx = min(gx + 1, W - 1) - max(gx - 1, 0) + 1;
      const float inv = in ? 1.f / (float)(cy * cx) : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float A = 0.f, Bq = 0.f, Cq = 0.f;
        if (in) {
          const Term r = term(i, c, masked, inv);
          const bool win = r.val == smin[c * SM::kStats + i];
          if (win && r.pass && !masked) {
            const float g = coef / scnt[c * SM::kStats + i];         // tf.reduce_min: equal split among ties
            if (ssim) { const float hh = -0.5f * g; A = hh * r.dm * inv; Bq = hh * r.dq * inv; Cq = hh * r.dr * inv; }
            else A = a.method == 0 ? g * sgnf(r.dm) : g * 2.f * r.dm;   // L1 / L2: a purely local term
          }
        }
        sA[c * SM::kStats + i] = A; sB[c * SM::kStats + i] = Bq; sC[c * SM::kStats + i] = Cq;
      }
    }
    __syncthreads();
    scatter(m);
  }
}

// loss_batch[b] = sum of the per-tile partials (fp64)
__global__ void k_sum_slots(const float* __restrict__ part, int slots, float* __restrict__ out) {
  const int b = blockIdx.x;
  double v = 0.0;
  for (int i = threadIdx.x; i < slots; i += blockDim.x) v += (double)part[(size_t)b * slots + i];
  __shared__ double sh[128];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[b] = (float)sh[0];
}

}  // namespace xpt
