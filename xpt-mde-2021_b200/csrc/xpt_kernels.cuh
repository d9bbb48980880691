// xpt_kernels.cuh -- sm_100a device code for the xpt-mde view-synthesis + loss path.
//
// Reference semantics restated in SURVEY.md Appendix A; citations are to the
// reference checkout (model/synthesize/*.py, model/loss_and_metric/*.py,
// utils/convert_pose.py).  Everything is fp32 and channel-last; there is no dense
// contraction here, so no tensor cores: the kernels are gather / stencil /
// reduction work bounded by HBM and L1/shared-memory bandwidth.
#pragma once
#include <cuda.h>            // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

namespace xpt {

constexpr int kMaxScales = 8;
constexpr int kMaxSrc = 8;
constexpr float kC1 = 0.01f * 0.01f;   // loss_util.py:68
constexpr float kC2 = 0.03f * 0.03f;   // loss_util.py:69

// One pyramid level as the kernels see it.
struct Level {
  int s;              // integer down-scale
  int H, W;           // level size
  int tiles_x, tiles_y;
  int slot_base;      // first partial-sum slot of this level inside one snippet
  const float* src;   // source level [B,N,H,W,3]
  long long src_bs, src_fs;   // batch / frame strides in elements
  const float* tgt;   // target level [B,H,W,3]
  long long tgt_bs;
};

struct LevelTable {
  int S;
  Level lv[kMaxScales];
};

template <typename T>
struct PtrTable {
  T* p[kMaxScales];
};

// geometry scratch: per (b, level): K_s (9) + inv(K_s) (9); per (b, n): R (9) + t (3)
constexpr int kGeoK = 18;
constexpr int kGeoT = 12;

// model/build_model/model_factory.py:133-137 InverseSigmoidActivation, the depth nets' last op:
// depth = safe_reciprocal_number(sigmoid(x) + 0.01), and its derivative expressed through the depth itself
// (y = 1/depth = sigmoid + 0.01, d depth / d x = -depth^2 s (1 - s) with s = y - 0.01)
__device__ __forceinline__ float depth_of_logit(float x) {
  const float s = __fdiv_rn(1.f, 1.f + expf(-x));
  const float y = s + 0.01f;
  return y > 0.00001f ? __fdiv_rn(1.f, y) : 0.f;
}
__device__ __forceinline__ float ddepth_dlogit(float D) {
  if (!(D > 0.f)) return 0.f;
  const float s = __fdiv_rn(1.f, D) - 0.01f;
  return -(D * D) * (s * (1.f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// geometry: pose twist -> [R|t] (utils/convert_pose.py:32-71, negated skew
// matrix :53-56, identity below 1e-8 :65) and per-level intrinsics
// (synthesize_base.py:66-71) with their inverses (synthesize_base.py:138).
// Evaluated in fp64 and rounded once: B*N + B*S tiny problems.
// ---------------------------------------------------------------------------
struct GeoArgs {
  const float* pose; const float* intrinsic;
  float* geoK; float* geoT; float* matr_out;
  int B, N, S;
  int s[kMaxScales];
};

__device__ __forceinline__ void geometry_item(const GeoArgs& ga, int i) {
  const float* __restrict__ pose = ga.pose;
  const float* __restrict__ intrinsic = ga.intrinsic;
  float* __restrict__ geoK = ga.geoK;
  float* __restrict__ geoT = ga.geoT;
  float* __restrict__ matr_out = ga.matr_out;
  const int B = ga.B, N = ga.N;
  if (pose && i < B * N) {
    const float* p = pose + (size_t)i * 6;
    double w1 = p[3], w2 = p[4], w3 = p[5];
    float thf = sqrtf(p[3] * p[3] + p[4] * p[4] + p[5] * p[5]);
    double th = sqrt(w1 * w1 + w2 * w2 + w3 * w3);
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (!(fabsf(thf) < 1e-8f)) {
      w1 /= th; w2 /= th; w3 /= th;
      double Wm[9] = {0, w3, -w2, -w3, 0, w1, w2, -w1, 0};
      double sn = sin(th), cs = 1.0 - cos(th);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double ww = 0;
          for (int k = 0; k < 3; ++k) ww += Wm[r * 3 + k] * Wm[k * 3 + c];
          R[r * 3 + c] += Wm[r * 3 + c] * sn + ww * cs;
        }
    }
    if (geoT) {
      float* g = geoT + (size_t)i * kGeoT;
      for (int k = 0; k < 9; ++k) g[k] = (float)R[k];
      g[9] = p[0]; g[10] = p[1]; g[11] = p[2];
    }
    if (matr_out) {
      float* m = matr_out + (size_t)i * 16;
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) m[r * 4 + c] = (float)R[r * 3 + c];
        m[r * 4 + 3] = p[r];
      }
      m[12] = 0.f; m[13] = 0.f; m[14] = 0.f; m[15] = 1.f;
    }
  }
  if (geoK && intrinsic && i < B * ga.S) {
    int b = i / ga.S, l = i % ga.S;
    const float* K = intrinsic + (size_t)b * 9;
    float sc = (float)ga.s[l];
    float Ks[9];
    for (int k = 0; k < 6; ++k) Ks[k] = K[k] / sc;      // rows 0-1 divided, fp32 like the reference
    Ks[6] = 0.f; Ks[7] = 0.f; Ks[8] = 1.f;
    double a = Ks[0], bb = Ks[1], c = Ks[2], d = Ks[3], e = Ks[4], f = Ks[5], g = Ks[6], h = Ks[7], k9 = Ks[8];
    double det = a * (e * k9 - f * h) - bb * (d * k9 - f * g) + c * (d * h - e * g);
    double inv[9] = {(e * k9 - f * h) / det, (c * h - bb * k9) / det, (bb * f - c * e) / det,
                     (f * g - d * k9) / det, (a * k9 - c * g) / det, (c * d - a * f) / det,
                     (d * h - e * g) / det, (bb * g - a * h) / det, (a * e - bb * d) / det};
    float* o = geoK + (size_t)i * kGeoK;
    for (int k = 0; k < 9; ++k) { o[k] = Ks[k]; o[9 + k] = (float)inv[k]; }
  }
}

__global__ void k_geometry(GeoArgs ga) { geometry_item(ga, blockIdx.x * blockDim.x + threadIdx.x); }

// ---------------------------------------------------------------------------
// pyramids: tf.image.resize(bilinear), TF2 half-pixel centres, no antialias
// (synthesize_base.py:74-85, util_funcs.py:163-175).  For an integer factor s the
// sample position is s*d + (s-1)/2: even s -> the centre 2x2 with weights 1/2,
// evaluated as nested lerps like TF's kernel; odd s -> the centre pixel.
// One thread per output pixel of (level, frame); frame N of a snippet is the target.
// ---------------------------------------------------------------------------
struct PyramidArgs {
  const float* source; long long src_bs, src_fs;
  const float* target; long long tgt_bs;
  int B, N, H, W;
  int S;                       // number of levels
  int s[kMaxScales];
  float* src_out[kMaxScales];  // [B,N,h,w,3] (NULL for s == 1)
  float4* src4_out[kMaxScales];// [B,N,h,w] RGBx texels for the fused kernel's 16-byte gathers (all levels; NULL = not wanted)
  float* tgt_out[kMaxScales];  // [B,h,w,3]   (NULL = not wanted)
  int with_geometry;           // grid row S carries the camera geometry (overlaps with the pyramid rows)
  GeoArgs geo;
};

__global__ void k_pyramid(PyramidArgs a) {
  if (a.with_geometry && (int)blockIdx.y == a.S) {
    geometry_item(a.geo, blockIdx.x * blockDim.x + threadIdx.x);
    return;
  }
  const int l = blockIdx.y;
  const int s = a.s[l];
  const int h = a.H / s, w = a.W / s;
  const int nfr = a.N + 1;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)a.B * nfr * h * w;
  if (idx >= total) return;
  int x = (int)(idx % w);
  long long r = idx / w;
  int y = (int)(r % h); r /= h;
  int f = (int)(r % nfr);
  int b = (int)(r / nfr);
  const float* in;
  float* out = nullptr;
  float4* out4 = nullptr;
  if (f < a.N) {
    if (a.src_out[l] == nullptr && a.src4_out[l] == nullptr) return;
    in = a.source + b * a.src_bs + f * a.src_fs;
    const long long o = ((long long)(b * a.N + f) * h + y) * w + x;
    if (a.src_out[l]) out = a.src_out[l] + o * 3;
    if (a.src4_out[l]) out4 = a.src4_out[l] + o;
  } else {
    if (a.tgt_out[l] == nullptr || a.target == nullptr) return;
    in = a.target + b * a.tgt_bs;
    out = a.tgt_out[l] + (((long long)b * h + y) * w + x) * 3;
  }
  const long long rowst = (long long)a.W * 3;
  float v[3];
  if (s == 1) {
    const float* p = in + y * rowst + x * 3;
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
  } else if (s & 1) {
    const float* p = in + (long long)(y * s + s / 2) * rowst + (x * s + s / 2) * 3;
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
  } else {
    const float* p0 = in + (long long)(y * s + s / 2 - 1) * rowst + (x * s + s / 2 - 1) * 3;
    const float* p1 = p0 + rowst;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float tl = __ldg(p0 + c), tr = __ldg(p0 + 3 + c), bl = __ldg(p1 + c), br = __ldg(p1 + 3 + c);
      float top = tl + (tr - tl) * 0.5f;
      float bot = bl + (br - bl) * 0.5f;
      v[c] = top + (bot - top) * 0.5f;
    }
  }
  if (out) { out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; }
  if (out4) *out4 = make_float4(v[0], v[1], v[2], 0.f);
}

// Fast path for scales within {1,2,4,8}: ONE pass over the full-resolution frames.  A CTA stages an
// 8-row x 128-pixel tile in shared memory with 16-byte loads and emits the s=2, 4, 8 outputs of
// that tile (centre 2x2 of each s x s block, TF's nested lerps), so every input byte is read from
// HBM once instead of once per level.  Extra z-slice: the camera geometry (overlaps with the copy).
constexpr int kPyrTW = 128, kPyrTH = 8, kPyrThreads = 256;

struct PyramidTiledArgs {
  const float* source; long long src_bs, src_fs;
  const float* target; long long tgt_bs;       // target may be NULL
  int B, N, H, W;
  float* src_out[4];   // index = log2(s): [1] s=2, [2] s=4, [3] s=8 (NULL = level absent)
  float4* src4_out[4]; // RGBx texels of the source levels, index = log2(s) incl. [0] = full resolution (NULL = not wanted)
  float* tgt_out[4];
  int with_geometry;
  GeoArgs geo;
};

// tile -> outputs.  FULL: the tile is TW wide, so every division below has a compile-time divisor.
template <bool FULL, int TW = kPyrTW>
__device__ __forceinline__ void pyramid_emit(const PyramidTiledArgs& a, const float (*tile)[TW * 3], int tw_rt,
                                             bool is_tgt, long long frame, int x0, int y0) {
  const int tw = FULL ? TW : tw_rt;
  if (!is_tgt) {
    // RGBx texels (16 bytes) of every source level for the fused kernel's 128-bit gathers
#pragma unroll
    for (int lg = 0; lg <= 3; ++lg) {
      float4* o4 = a.src4_out[lg];
      if (o4 == nullptr) continue;
      const int s = 1 << lg;
      const int oh = kPyrTH >> lg, ow = tw >> lg;
      const int Hs = a.H >> lg, Ws = a.W >> lg;
      o4 += (frame * Hs + (y0 >> lg)) * Ws + (x0 >> lg);
      for (int e = threadIdx.x; e < oh * ow; e += kPyrThreads) {
        const int oy = e / ow, ox = e - oy * ow;
        float v[3];
        if (lg == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) v[c] = tile[oy][ox * 3 + c];
        } else {
          const int ry = oy * s + s / 2 - 1, rx = (ox * s + s / 2 - 1) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float tl = tile[ry][rx + c], tr = tile[ry][rx + 3 + c], bl = tile[ry + 1][rx + c], br = tile[ry + 1][rx + 3 + c];
            const float top = tl + (tr - tl) * 0.5f;
            const float bot = bl + (br - bl) * 0.5f;
            v[c] = top + (bot - top) * 0.5f;
          }
        }
        o4[(long long)oy * Ws + ox] = make_float4(v[0], v[1], v[2], 0.f);
      }
    }
  }
#pragma unroll
  for (int lg = 1; lg <= 3; ++lg) {
    float* outp = is_tgt ? a.tgt_out[lg] : a.src_out[lg];
    if (outp == nullptr) continue;
    const int s = 1 << lg;
    const int oh = kPyrTH >> lg;                 // output rows of this tile
    const int Hs = a.H >> lg, Ws = a.W >> lg;
    float* o = outp + (frame * Hs + (y0 >> lg)) * Ws * 3 + (x0 >> lg) * 3;
    const int roww = (tw >> lg) * 3;             // FULL: 192 / 96 / 48
    for (int e = threadIdx.x; e < oh * roww; e += kPyrThreads) {
      const int oy = e / roww, rem = e - oy * roww;      // rem = 3*ox + c inside the tile's output row
      const int ox = rem / 3, c = rem - ox * 3;
      const int ry = oy * s + s / 2 - 1, rx = (ox * s + s / 2 - 1) * 3 + c;
      const float tl = tile[ry][rx], tr = tile[ry][rx + 3], bl = tile[ry + 1][rx], br = tile[ry + 1][rx + 3];
      const float top = tl + (tr - tl) * 0.5f;
      const float bot = bl + (br - bl) * 0.5f;
      o[(long long)oy * Ws * 3 + rem] = top + (bot - top) * 0.5f;
    }
  }
}

#ifndef XPT_PYR_MINB
#define XPT_PYR_MINB 6
#endif
__global__ void __launch_bounds__(kPyrThreads, XPT_PYR_MINB) k_pyramid_tiled(const __grid_constant__ PyramidTiledArgs a) {
  __shared__ __align__(16) float tile[kPyrTH][kPyrTW * 3];
  const int nfr = a.N + 1;
  if ((int)blockIdx.z == a.B * nfr) {           // geometry slice
    if (a.with_geometry && blockIdx.y == 0) geometry_item(a.geo, blockIdx.x * blockDim.x + threadIdx.x);
    return;
  }
  const int b = blockIdx.z / nfr, f = blockIdx.z % nfr;
  const bool is_tgt = f == a.N;
  if (is_tgt && a.target == nullptr) return;
  const float* in = is_tgt ? a.target + b * a.tgt_bs : a.source + b * a.src_bs + f * a.src_fs;
  const int x0 = blockIdx.x * kPyrTW, y0 = blockIdx.y * kPyrTH;
  const int tw = min(kPyrTW, a.W - x0);         // multiple of 8
  const float* in0 = in + ((long long)y0 * a.W + x0) * 3;
  const long long rowst = (long long)a.W * 3;
  const long long frame = is_tgt ? (long long)b : (long long)(b * a.N + f);
  if (tw == kPyrTW) {
    constexpr int kRowF4 = kPyrTW * 3 / 4;      // 96 float4 per tile row: all index math is compile-time
#pragma unroll
    for (int k = 0; k < kPyrTH * kRowF4 / kPyrThreads; ++k) {
      const int i = threadIdx.x + k * kPyrThreads;
      const int r = i / kRowF4, c4 = i - r * kRowF4;
      *reinterpret_cast<float4*>(&tile[r][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(in0 + r * rowst) + c4);
    }
    __syncthreads();
    pyramid_emit<true>(a, tile, tw, is_tgt, frame, x0, y0);
  } else {
    const int row_f4 = tw * 3 / 4;
    for (int i = threadIdx.x; i < kPyrTH * row_f4; i += kPyrThreads) {
      const int r = i / row_f4, c4 = i - r * row_f4;
      *reinterpret_cast<float4*>(&tile[r][c4 * 4]) = __ldg(reinterpret_cast<const float4*>(in0 + r * rowst) + c4);
    }
    __syncthreads();
    pyramid_emit<false>(a, tile, tw, is_tgt, frame, x0, y0);
  }
}

// ---- the same pass with the tile loaded by the TMA unit -------------------------------------------------------
// One elected thread arms an mbarrier with the tile's byte count and issues ONE bulk tensor copy
// (cp.async.bulk.tensor -> UTMALDG): an 8-row x 64-pixel x 3-channel box of the frame lands in shared memory without
// a single LDG / STS of the CTA; columns and rows beyond the frame are zero-filled by the unit.  The tensor maps
// describe the caller's frames as they lie in HBM: source [B][N][H][W*3] (any batch / frame stride, e.g. the
// image5d[:, :-1] view), target [B][H][W*3].
constexpr int kPyrTmaTW = 64;                  // box = 192 floats x 8 rows (a box dimension may not exceed 256 elements)
constexpr int kPyrTmaBoxes = 2;                // boxes per CTA: a 128-pixel tile like k_pyramid_tiled, two copies in flight

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kPyrThreads, XPT_PYR_MINB) k_pyramid_tma(const __grid_constant__ PyramidTiledArgs a,
                                                                            const __grid_constant__ CUtensorMap tm_src,
                                                                            const __grid_constant__ CUtensorMap tm_tgt) {
  __shared__ __align__(128) float tile[kPyrTmaBoxes][kPyrTH][kPyrTmaTW * 3];
  __shared__ __align__(8) unsigned long long mbar[kPyrTmaBoxes];
  const int nfr = a.N + 1;
  if ((int)blockIdx.z == a.B * nfr) {           // geometry slice
    if (a.with_geometry && blockIdx.y == 0) geometry_item(a.geo, blockIdx.x * blockDim.x + threadIdx.x);
    return;
  }
  const int b = blockIdx.z / nfr, f = blockIdx.z % nfr;
  const bool is_tgt = f == a.N;
  if (is_tgt && a.target == nullptr) return;
  const int x0 = blockIdx.x * (kPyrTmaTW * kPyrTmaBoxes), y0 = blockIdx.y * kPyrTH;
  if (x0 >= a.W) return;
  // one mbarrier per box: the levels of box 0 are emitted while box 1 is still in flight
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kPyrTmaBoxes; ++k)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[k])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    constexpr unsigned kBoxBytes = kPyrTH * kPyrTmaTW * 3 * sizeof(float);
#pragma unroll
    for (int k = 0; k < kPyrTmaBoxes; ++k) {
      const int xk = x0 + k * kPyrTmaTW;
      if (xk >= a.W) continue;
      const unsigned bar = smem_u32(&mbar[k]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBoxBytes) : "memory");
      const unsigned dst = smem_u32(&tile[k][0][0]);
      if (is_tgt)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(&tm_tgt), "r"(bar), "r"(xk * 3), "r"(y0), "r"(b) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"(&tm_src), "r"(bar), "r"(xk * 3), "r"(y0), "r"(f), "r"(b) : "memory");
    }
  }
  const long long frame = is_tgt ? (long long)b : (long long)(b * a.N + f);
#pragma unroll
  for (int k = 0; k < kPyrTmaBoxes; ++k) {
    const int xk = x0 + k * kPyrTmaTW;
    if (xk >= a.W) continue;
    // every thread waits for the box's transaction bytes (phase 0 of its barrier)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(&mbar[k])), "r"(0) : "memory");
    const int tw = min(kPyrTmaTW, a.W - xk);    // multiple of 8
    if (tw == kPyrTmaTW) pyramid_emit<true, kPyrTmaTW>(a, tile[k], tw, is_tgt, frame, xk, y0);
    else pyramid_emit<false, kPyrTmaTW>(a, tile[k], tw, is_tgt, frame, xk, y0);
  }
}

// ---- persistent, double-buffered form of k_pyramid_tma --------------------------------------------------------------
// grid = a few CTAs per SM; every CTA walks the list of (frame, row band, 128-pixel column pair) items with a stride of
// gridDim.x and keeps TWO stages of boxes in shared memory: while the levels of item i are emitted, the bulk copies of
// item i+1 are already in flight (one mbarrier per stage, phase = use count parity).  The grid-shaped kernel above runs
// its CTAs in lock step -- all load, then all store, then a partial second wave.  MEASURED SLOWER than the grid form
// (18.9 vs 17.9 us at config 2, 94.6 vs 86.0 us at config 3): the per-item CTA barrier and the 40-register budget of six
// resident CTAs cost more than the overlap gains; opt-in through XPT_PYRAMID=tma_persistent, parity-tested.
constexpr int kPyrStages = 2;

struct PyrItem { int b, f, x0, y0; bool is_tgt, skip; };

__device__ __forceinline__ PyrItem pyr_item(const PyramidTiledArgs& a, int item, int ntx, int nty) {
  PyrItem it;
  const int nfr = a.N + 1;
  const int tx = item % ntx, r = item / ntx;
  const int ty = r % nty, z = r / nty;
  it.b = z / nfr; it.f = z % nfr;
  it.is_tgt = it.f == a.N;
  it.skip = it.is_tgt && a.target == nullptr;
  it.x0 = tx * (kPyrTmaTW * kPyrTmaBoxes); it.y0 = ty * kPyrTH;
  return it;
}

__global__ void __launch_bounds__(kPyrThreads, XPT_PYR_MINB) k_pyramid_tma_p(const __grid_constant__ PyramidTiledArgs a,
                                                                              const __grid_constant__ CUtensorMap tm_src,
                                                                              const __grid_constant__ CUtensorMap tm_tgt,
                                                                              int ntx, int nty, int total) {
  __shared__ __align__(128) float tile[kPyrStages][kPyrTmaBoxes][kPyrTH][kPyrTmaTW * 3];
  __shared__ __align__(8) unsigned long long mbar[kPyrStages];
  if (a.with_geometry) geometry_item(a.geo, blockIdx.x * blockDim.x + threadIdx.x);       // (the host sizes the grid for it)
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kPyrStages; ++k)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar[k])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr unsigned kBoxBytes = kPyrTH * kPyrTmaTW * 3 * sizeof(float);
  auto issue = [&](int item, int stage) {            // thread 0: arm the stage's barrier and start its copies
    const PyrItem it = pyr_item(a, item, ntx, nty);
    if (it.skip) return;
    const unsigned bar = smem_u32(&mbar[stage]);
    int nbox = 0;
#pragma unroll
    for (int k = 0; k < kPyrTmaBoxes; ++k) nbox += (it.x0 + k * kPyrTmaTW < a.W) ? 1 : 0;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kBoxBytes * nbox) : "memory");
#pragma unroll
    for (int k = 0; k < kPyrTmaBoxes; ++k) {
      const int xk = it.x0 + k * kPyrTmaTW;
      if (xk >= a.W) continue;
      const unsigned dst = smem_u32(&tile[stage][k][0][0]);
      if (it.is_tgt)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(&tm_tgt), "r"(bar), "r"(xk * 3), "r"(it.y0), "r"(it.b) : "memory");
      else
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"(&tm_src), "r"(bar), "r"(xk * 3), "r"(it.y0), "r"(it.f), "r"(it.b) : "memory");
    }
  };
  int item = blockIdx.x;
  if (item < total && threadIdx.x == 0) issue(item, 0);
  unsigned phase[kPyrStages] = {0u, 0u};
  for (int n = 0; item < total; item += gridDim.x, ++n) {
    const int stage = n & 1;
    const int next = item + gridDim.x;
    // stage^1 was last read in the previous iteration, which ended with a CTA barrier: it may be refilled now
    if (next < total && threadIdx.x == 0) issue(next, stage ^ 1);
    const PyrItem it = pyr_item(a, item, ntx, nty);
    if (!it.skip) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "WAIT_%=:\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
          "@p bra DONE_%=;\n\t"
          "bra WAIT_%=;\n\t"
          "DONE_%=:\n\t}" ::"r"(smem_u32(&mbar[stage])), "r"(phase[stage]) : "memory");
      phase[stage] ^= 1u;
      const long long frame = it.is_tgt ? (long long)it.b : (long long)(it.b * a.N + it.f);
#pragma unroll
      for (int k = 0; k < kPyrTmaBoxes; ++k) {
        const int xk = it.x0 + k * kPyrTmaTW;
        if (xk >= a.W) continue;
        const int tw = min(kPyrTmaTW, a.W - xk);    // multiple of 8
        if (tw == kPyrTmaTW) pyramid_emit<true, kPyrTmaTW>(a, tile[stage][k], tw, it.is_tgt, frame, xk, it.y0);
        else pyramid_emit<false, kPyrTmaTW>(a, tile[stage][k], tw, it.is_tgt, frame, xk, it.y0);
      }
    }
    __syncthreads();
  }
}

// adjoint of the source pyramid: d_source[full] += resize^T(d_source_level) for s > 1
// (each level pixel spreads 1/4 to its centre 2x2 / 1 to the centre pixel).
struct PyramidAdjArgs {
  float* d_source;             // [B,N,H,W,3] dense, already holding the level-0 gradient
  int BN, H, W, S;
  int s[kMaxScales];
  const float* d_level[kMaxScales];   // [B*N,h,w,3] for s > 1
};

__global__ void k_pyramid_adjoint(PyramidAdjArgs a) {
  // one thread per full-res pixel of one frame: gathers from every level that touches it
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)a.BN * a.H * a.W;
  if (idx >= total) return;
  int x = (int)(idx % a.W);
  long long r = idx / a.W;
  int y = (int)(r % a.H);
  int f = (int)(r / a.H);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int l = 0; l < a.S; ++l) {
    int s = a.s[l];
    if (s == 1 || a.d_level[l] == nullptr) continue;
    int h = a.H / s, w = a.W / s;
    int ys = y / s, xs = x / s, ry = y % s, rx = x % s;
    float wt;
    if (s & 1) {
      if (ry != s / 2 || rx != s / 2) continue;
      wt = 1.f;
    } else {
      if ((ry != s / 2 - 1 && ry != s / 2) || (rx != s / 2 - 1 && rx != s / 2)) continue;
      wt = 0.25f;
    }
    const float* g = a.d_level[l] + (((long long)f * h + ys) * w + xs) * 3;
    acc[0] += wt * g[0]; acc[1] += wt * g[1]; acc[2] += wt * g[2];
  }
  float* o = a.d_source + idx * 3;
  o[0] += acc[0]; o[1] += acc[1]; o[2] += acc[2];
}

// dL/dsource of the fused path: the kernel scatters into RGBx gradient levels (one 16-byte reduction per tap);
// this pass folds them -- level 0 plus the adjoint of the source pyramid for s > 1 (each level pixel spreads 1/4 to
// its centre 2x2, 1 to the centre pixel for odd s) -- into the dense 3-channel d_source, which it WRITES (no memset
// and no read-modify-write of d_source).  One thread per full-resolution pixel of one frame.
struct DsourceFinishArgs {
  float* d_source;             // [B*N,H,W,3] dense
  int BN, H, W, S;
  int s[kMaxScales];
  const float4* d_level4[kMaxScales];   // [B*N,h,w] RGBx
};

__global__ void __launch_bounds__(256) k_dsource_finish(DsourceFinishArgs a) {
  // grid = (ceil(W / 256), H, B*N): no integer division by runtime image sizes
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= a.W) return;
  const int y = blockIdx.y, f = blockIdx.z;
  const long long idx = ((long long)f * a.H + y) * a.W + x;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int l = 0; l < a.S; ++l) {
    const int s = a.s[l];
    if (a.d_level4[l] == nullptr) continue;
    int h, w, ys, xs, ry, rx;
    if ((s & (s - 1)) == 0) {                  // power of two (the usual 1, 2, 4, 8): shifts and masks
      const int lg = 31 - __clz(s);
      h = a.H >> lg; w = a.W >> lg; ys = y >> lg; xs = x >> lg; ry = y & (s - 1); rx = x & (s - 1);
    } else {
      h = a.H / s; w = a.W / s; ys = y / s; xs = x / s; ry = y % s; rx = x % s;
    }
    float wt = 1.f;
    if (s > 1) {
      if (s & 1) {
        if (ry != s / 2 || rx != s / 2) continue;
      } else {
        if ((ry != s / 2 - 1 && ry != s / 2) || (rx != s / 2 - 1 && rx != s / 2)) continue;
        wt = 0.25f;
      }
    }
    const float4 g = __ldg(a.d_level4[l] + ((long long)f * h + ys) * w + xs);
    acc[0] += wt * g.x; acc[1] += wt * g.y; acc[2] += wt * g.z;
  }
  float* o = a.d_source + idx * 3;
  o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2];
}

// ---------------------------------------------------------------------------
// inverse warp of one target pixel into one source frame.
// Order of operations follows the reference:
//   r = inv(K_s) (u,v,1)   synthesize_base.py:138
//   X = r * D              :141
//   Y = R X + t            :158
//   p = K_s Y              :173
//   (u',v') = p.xy / (p.z + 1e-10)   :176-177  -- no positive-depth test
// ---------------------------------------------------------------------------
struct Proj {
  float u, v, den;     // source pixel coordinates and the divisor
  float X0, X1, X2;    // target-frame 3-D point
};

__device__ __forceinline__ void ray_of_pixel(const float* __restrict__ Ki, float uu, float vv,
                                             float& r0, float& r1, float& r2) {
  r0 = Ki[0] * uu + Ki[1] * vv + Ki[2];
  r1 = Ki[3] * uu + Ki[4] * vv + Ki[5];
  r2 = Ki[6] * uu + Ki[7] * vv + Ki[8];
}

// 1/x: MUFU.RCP + one Newton step (<= 1 ulp of the correctly rounded reciprocal)
__device__ __forceinline__ float rcp_nr(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(fmaf(-x, r, 1.f), r, r);
}

// FAST = false: IEEE division like the reference's `pixel_coords / (z + 1e-10)`;
// FAST = true : one reciprocal + two multiplies (<= 2 ulp on u, v), `inv` returned for the backward.
template <bool FAST = false>
__device__ __forceinline__ Proj project(const float* __restrict__ K, const float* __restrict__ T,
                                        float r0, float r1, float r2, float D, float* inv_out = nullptr) {
  Proj o;
  o.X0 = r0 * D; o.X1 = r1 * D; o.X2 = r2 * D;
  float Y0 = T[0] * o.X0 + T[1] * o.X1 + T[2] * o.X2 + T[9];
  float Y1 = T[3] * o.X0 + T[4] * o.X1 + T[5] * o.X2 + T[10];
  float Y2 = T[6] * o.X0 + T[7] * o.X1 + T[8] * o.X2 + T[11];
  float p0 = K[0] * Y0 + K[1] * Y1 + K[2] * Y2;
  float p1 = K[3] * Y0 + K[4] * Y1 + K[5] * Y2;
  float p2 = K[6] * Y0 + K[7] * Y1 + K[8] * Y2;
  o.den = p2 + 1e-10f;
  if (FAST) {
    const float inv = rcp_nr(o.den);
    o.u = p0 * inv;
    o.v = p1 * inv;
    if (inv_out) *inv_out = inv;
  } else {
    o.u = p0 / o.den;
    o.v = p1 / o.den;
  }
  return o;
}

// bilinear tap set of one sample (bilinear_interp.py:34-100)
struct Taps {
  bool valid;
  int iu, iv;                  // floor(u), floor(v)
  float w_uf, w_uc, w_vf, w_vc;
};

__device__ __forceinline__ Taps make_taps(float u, float v, float D, int W, int H) {
  Taps t;
  float uf = floorf(u), vf = floorf(v);
  // valid <=> 0 <= floor(u) <= W-2 and 0 <= floor(v) <= H-2 and D != 0 (:53-76); NaN compares false
  t.valid = (uf >= 0.f) && (uf <= (float)(W - 2)) && (vf >= 0.f) && (vf <= (float)(H - 2)) && (D != 0.f);
  t.iu = t.valid ? (int)uf : 0;
  t.iv = t.valid ? (int)vf : 0;
  t.w_uf = (uf + 1.f) - u;     // u_ceil - u
  t.w_uc = u - uf;
  t.w_vf = (vf + 1.f) - v;
  t.w_vc = v - vf;
  return t;
}

__device__ __forceinline__ void gather_taps(const float* __restrict__ img, int W, const Taps& t,
                                            float I0[3], float I1[3], float I2[3], float I3[3]) {
  // I0 = (vf,uf), I1 = (vc,uf), I2 = (vf,uc), I3 = (vc,uc)   (bilinear_interp.py:125-128)
  const float* p = img + ((long long)t.iv * W + t.iu) * 3;
  const float* q = p + (long long)W * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    I0[c] = __ldg(p + c);
    I2[c] = __ldg(p + 3 + c);
    I1[c] = __ldg(q + c);
    I3[c] = __ldg(q + 3 + c);
  }
}

__device__ __forceinline__ void warp_sample(const float* __restrict__ img, int W, int H, float u, float v,
                                            float D, float y[3], bool& valid) {
  Taps t = make_taps(u, v, D, W, H);
  valid = t.valid;
  if (!t.valid) { y[0] = y[1] = y[2] = 0.f; return; }
  float I0[3], I1[3], I2[3], I3[3];
  gather_taps(img, W, t, I0, I1, I2, I3);
  float w0 = t.w_uf * t.w_vf, w1 = t.w_uf * t.w_vc, w2 = t.w_uc * t.w_vf, w3 = t.w_uc * t.w_vc;
#pragma unroll
  for (int c = 0; c < 3; ++c) y[c] = ((I0[c] * w0 + I1[c] * w1) + I2[c] * w2) + I3[c] * w3;
}

// ---------------------------------------------------------------------------
// standalone synthesis forward (SynthesizeMultiScale): one thread per target pixel
// of (level, b), looping over the N sources so depth / ray are computed once.
// ---------------------------------------------------------------------------
struct WarpFwdArgs {
  LevelTable lt;
  int B, N;
  const float* geoK; const float* geoT;
  const float* depth[kMaxScales];
  float* synth[kMaxScales];
  float* mask[kMaxScales];     // may be NULL
};

__global__ void __launch_bounds__(256) k_warp_fwd(WarpFwdArgs a) {
  const int l = blockIdx.y;
  const Level& L = a.lt.lv[l];
  const int P = L.H * L.W;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.B * P) return;
  int b = (int)(idx / P);
  int pix = (int)(idx % P);
  int y = pix / L.W, x = pix % L.W;
  const float* gk = a.geoK + ((size_t)b * a.lt.S + l) * kGeoK;
  float K[9], Ki[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { K[k] = __ldg(gk + k); Ki[k] = __ldg(gk + 9 + k); }
  float D = __ldg(a.depth[l] + idx);
  float r0, r1, r2;
  ray_of_pixel(Ki, (float)x, (float)y, r0, r1, r2);
  for (int n = 0; n < a.N; ++n) {
    const float* gt = a.geoT + ((size_t)b * a.N + n) * kGeoT;
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = __ldg(gt + k);
    Proj pr = project(K, T, r0, r1, r2, D);
    const float* img = L.src + b * L.src_bs + n * L.src_fs;
    float yv[3]; bool valid;
    warp_sample(img, L.W, L.H, pr.u, pr.v, D, yv, valid);
    long long o = ((long long)(b * a.N + n) * P + pix);
    float* so = a.synth[l] + o * 3;
    so[0] = yv[0]; so[1] = yv[1]; so[2] = yv[2];
    if (a.mask[l]) a.mask[l][o] = valid ? 1.f : 0.f;
  }
}

// ---------------------------------------------------------------------------
// bilinear + projection adjoint for one sample with upstream g = dL/dS.
// Returns dL/dD contribution; accumulates dL/dR (9) and dL/dt (3) into acc[12];
// optionally scatters dL/dsource with red.global.add.f32.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float sample_adjoint(const float* __restrict__ img, float* __restrict__ dimg,
                                                int W, int H, const float* __restrict__ K,
                                                const float* __restrict__ T, float r0, float r1, float r2,
                                                float D, const Proj& pr, const float g[3], float acc[12]) {
  Taps t = make_taps(pr.u, pr.v, D, W, H);
  if (!t.valid) return 0.f;
  float I0[3], I1[3], I2[3], I3[3];
  gather_taps(img, W, t, I0, I1, I2, I3);
  float gu = 0.f, gv = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    gu += g[c] * (t.w_vf * (I2[c] - I0[c]) + t.w_vc * (I3[c] - I1[c]));
    gv += g[c] * (t.w_uf * (I1[c] - I0[c]) + t.w_uc * (I3[c] - I2[c]));
  }
  if (dimg) {
    float w0 = t.w_uf * t.w_vf, w1 = t.w_uf * t.w_vc, w2 = t.w_uc * t.w_vf, w3 = t.w_uc * t.w_vc;
    float* p = dimg + ((long long)t.iv * W + t.iu) * 3;
    float* q = p + (long long)W * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicAdd(p + c, w0 * g[c]);
      atomicAdd(p + 3 + c, w2 * g[c]);
      atomicAdd(q + c, w1 * g[c]);
      atomicAdd(q + 3 + c, w3 * g[c]);
    }
  }
  float inv = 1.f / pr.den;
  float gp0 = gu * inv, gp1 = gv * inv, gp2 = -(gu * pr.u + gv * pr.v) * inv;
  float gY0 = K[0] * gp0 + K[3] * gp1 + K[6] * gp2;
  float gY1 = K[1] * gp0 + K[4] * gp1 + K[7] * gp2;
  float gY2 = K[2] * gp0 + K[5] * gp1 + K[8] * gp2;
  acc[0] += gY0 * pr.X0; acc[1] += gY0 * pr.X1; acc[2] += gY0 * pr.X2;
  acc[3] += gY1 * pr.X0; acc[4] += gY1 * pr.X1; acc[5] += gY1 * pr.X2;
  acc[6] += gY2 * pr.X0; acc[7] += gY2 * pr.X1; acc[8] += gY2 * pr.X2;
  acc[9] += gY0; acc[10] += gY1; acc[11] += gY2;
  float gX0 = T[0] * gY0 + T[3] * gY1 + T[6] * gY2;
  float gX1 = T[1] * gY0 + T[4] * gY1 + T[7] * gY2;
  float gX2 = T[2] * gY0 + T[5] * gY1 + T[8] * gY2;
  return gX0 * r0 + gX1 * r1 + gX2 * r2;
}

// block-level reduction of 12 pose accumulators; thread 0 stores them.
template <int NT>
__device__ __forceinline__ void block_reduce_store12(float acc[12], float* __restrict__ out, float* smem /* >= 12*NT/32 */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    float v = warp_sum(acc[k]);
    if (lane == 0) smem[wid * 12 + k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float v = 0.f;
    for (int w = 0; w < NT / 32; ++w) v += smem[w * 12 + threadIdx.x];
    out[threadIdx.x] = v;
  }
  __syncthreads();
}

// standalone synthesis backward: chunk of kWarpBwdChunk pixels of one (level, b) per block.
constexpr int kWarpBwdThreads = 256;
constexpr int kWarpBwdPix = 4;
constexpr int kWarpBwdChunk = kWarpBwdThreads * kWarpBwdPix;

struct WarpBwdArgs {
  LevelTable lt;
  int B, N;
  const float* geoK; const float* geoT;
  const float* depth[kMaxScales];
  const float* gsynth[kMaxScales];     // dL/d synth
  float* d_depth[kMaxScales];
  float* d_src[kMaxScales];            // per-level dL/d source level (NULL = off), same layout as Level.src
  long long d_src_bs[kMaxScales], d_src_fs[kMaxScales];
  float* pose_part;                    // [B][slots][N][12]
  int slots_per_b;
  int chunk_base[kMaxScales];          // first chunk slot of each level
};

__global__ void __launch_bounds__(kWarpBwdThreads) k_warp_bwd(WarpBwdArgs a) {
  __shared__ float red[12 * kWarpBwdThreads / 32];
  const int l = blockIdx.z;
  const Level& L = a.lt.lv[l];
  const int P = L.H * L.W;
  const int chunk = blockIdx.x;
  if (chunk * kWarpBwdChunk >= P) return;
  const int b = blockIdx.y;
  const float* gk = a.geoK + ((size_t)b * a.lt.S + l) * kGeoK;
  float K[9], Ki[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { K[k] = __ldg(gk + k); Ki[k] = __ldg(gk + 9 + k); }
  float gD[kWarpBwdPix];
  float D[kWarpBwdPix];
#pragma unroll
  for (int j = 0; j < kWarpBwdPix; ++j) {
    int pix = chunk * kWarpBwdChunk + j * kWarpBwdThreads + threadIdx.x;
    gD[j] = 0.f;
    D[j] = pix < P ? __ldg(a.depth[l] + (long long)b * P + pix) : 0.f;
  }
  for (int n = 0; n < a.N; ++n) {
    const float* gt = a.geoT + ((size_t)b * a.N + n) * kGeoT;
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = __ldg(gt + k);
    const float* img = L.src + b * L.src_bs + n * L.src_fs;
    float* dimg = a.d_src[l] ? a.d_src[l] + b * a.d_src_bs[l] + n * a.d_src_fs[l] : nullptr;
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
#pragma unroll
    for (int j = 0; j < kWarpBwdPix; ++j) {
      int pix = chunk * kWarpBwdChunk + j * kWarpBwdThreads + threadIdx.x;
      if (pix < P) {
        int y = pix / L.W, x = pix % L.W;
        float r0, r1, r2;
        ray_of_pixel(Ki, (float)x, (float)y, r0, r1, r2);
        Proj pr = project(K, T, r0, r1, r2, D[j]);
        const float* gp = a.gsynth[l] + ((long long)(b * a.N + n) * P + pix) * 3;
        float g[3] = {__ldg(gp), __ldg(gp + 1), __ldg(gp + 2)};
        gD[j] += sample_adjoint(img, dimg, L.W, L.H, K, T, r0, r1, r2, D[j], pr, g, acc);
      }
    }
    float* out = a.pose_part + (((size_t)b * a.slots_per_b + a.chunk_base[l] + chunk) * a.N + n) * 12;
    block_reduce_store12<kWarpBwdThreads>(acc, out, red);
  }
#pragma unroll
  for (int j = 0; j < kWarpBwdPix; ++j) {
    int pix = chunk * kWarpBwdChunk + j * kWarpBwdThreads + threadIdx.x;
    if (pix < P && a.d_depth[l]) a.d_depth[l][(long long)b * P + pix] = gD[j];
  }
}

// ---------------------------------------------------------------------------
// pose epilogue: deterministic fp64 reduction of the per-block partials, then the
// adjoint of Rodrigues' formula: with S = -[w]x, a = sin(th)/th, b = (1-cos th)/th^2,
// R = I + a S + b S^2  =>  dR/dw_k = a E_k + a'(w_k/th) S + b (E_k S + S E_k) + b'(w_k/th) S^2.
// ---------------------------------------------------------------------------
__device__ void pose_block(const float* __restrict__ pose_part, int slots_per_b, int slots_used,
                           const float* __restrict__ pose, float* __restrict__ d_pose, int N, int bn) {
  __shared__ double red[4][12];
  const float scale = 1.0f;
  const int b = bn / N, n = bn % N;
  double acc[12];
  for (int k = 0; k < 12; ++k) acc[k] = 0.0;
  for (int sl = threadIdx.x; sl < slots_used; sl += blockDim.x) {
    const float* p = pose_part + (((size_t)b * slots_per_b + sl) * N + n) * 12;
    for (int k = 0; k < 12; ++k) acc[k] += (double)p[k];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int k = 0; k < 12; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    // threads 0..2 each take one rotation component k (the serial fp64 chain is the epilogue's latency);
    // thread 0 also writes the translation part
    const int k = threadIdx.x;
    double G[12];
    for (int j = 0; j < 12; ++j) G[j] = (red[0][j] + red[1][j] + red[2][j] + red[3][j]) * (double)scale;
    const float* p = pose + (size_t)bn * 6;
    double w[3] = {p[3], p[4], p[5]};
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double th = sqrt(th2);
    float* o = d_pose + (size_t)bn * 6;
    if (k == 0) { o[0] = (float)G[9]; o[1] = (float)G[10]; o[2] = (float)G[11]; }
    if (th < 1e-8) {
      // the reference's gradient is NaN here: propagate that, do not invent a value
      o[3 + k] = __int_as_float(0x7fc00000);
    } else {
      double Sm[9] = {0, w[2], -w[1], -w[2], 0, w[0], w[1], -w[0], 0};
      double S2[9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double v = 0;
          for (int j = 0; j < 3; ++j) v += Sm[r * 3 + j] * Sm[j * 3 + c];
          S2[r * 3 + c] = v;
        }
      double sn = sin(th), cs = cos(th);
      double ca = sn / th, cb = (1.0 - cs) / th2;
      double da = (th * cs - sn) / th2;                    // d(sin th / th)/d th
      double db = (th * sn - 2.0 * (1.0 - cs)) / (th2 * th);   // d((1-cos th)/th^2)/d th
      double E[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      // E_k = d S / d w_k
      if (k == 0) { E[5] = 1; E[7] = -1; }
      if (k == 1) { E[2] = -1; E[6] = 1; }
      if (k == 2) { E[1] = 1; E[3] = -1; }
      double dth = w[k] / th;
      double sacc = 0;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          double es = 0;
          for (int j = 0; j < 3; ++j) es += E[r * 3 + j] * Sm[j * 3 + c] + Sm[r * 3 + j] * E[j * 3 + c];
          double dR = ca * E[r * 3 + c] + da * dth * Sm[r * 3 + c] + cb * es + db * dth * S2[r * 3 + c];
          sacc += G[r * 3 + c] * dR;
        }
      o[3 + k] = (float)sacc;
    }
  }
}

// ---------------------------------------------------------------------------
// photometric tile kernel.  One CTA = one (level, snippet, 32x16 tile), looping
// over the N sources.  The warped tile y (with halo) lives in shared memory:
//   FUSED : y is produced by the inverse warp in-kernel (north-star kernels 1+2+3)
//   !FUSED: y is read from a synth tensor (standalone PhotometricLossMultiScale)
// L1 (loss_util.py:6-25) and SSIM (loss_util.py:52-96) are evaluated from that
// tile; with GRAD the SSIM/L1 adjoint is evaluated from the same tile (halo 2) and
// either written out as dL/dsynth (!FUSED) or pushed straight through the bilinear
// + projection adjoint (FUSED).
// ---------------------------------------------------------------------------
constexpr int kTW = 32, kTH = 16, kPhotoThreads = 256;
constexpr int kPixPerThread = kTW * kTH / kPhotoThreads;   // 2

struct PhotoArgs {
  LevelTable lt;
  int B, N;
  int tiles_per_b;                     // sum over levels of tiles
  int first_tile[kMaxScales + 1];      // prefix of tiles per level (per snippet)
  // inputs
  const float* synth[kMaxScales];      // !FUSED: [B,N,h,w,3]
  const float* geoK; const float* geoT;
  const float* depth[kMaxScales];      // FUSED
  const float* disp[kMaxScales];       // FUSED + smoothness
  // loss configuration
  int l1_kind;                         // 0 none, 1 = L1, 2 = L2 (reported in column 0)
  int do_ssim;
  int do_smooth;                       // FUSED only
  float norm_photo[kMaxScales];        // sw_s / (N*h*w*3)
  float norm_sm_x[kMaxScales];         // sw_s * 0.5 / (h*(w-1)) / s
  float norm_sm_y[kMaxScales];         // sw_s * 0.5 / ((h-1)*w) / s
  float grad_factor;
  // gradient coefficients: dTotal/d(per-snippet loss k) ; multiplied by gbatch[b] when non-NULL
  float gcoef_l1, gcoef_ssim, gcoef_smooth;
  const float* gbatch;
  // outputs
  float* loss_part;                    // [B][slots][3]
  int slots_per_b;
  float* synth_out[kMaxScales];        // FUSED optional
  float* mask_out[kMaxScales];         // FUSED optional
  float* gsynth[kMaxScales];           // !FUSED GRAD: dL/d synth
  float* d_depth[kMaxScales];          // FUSED GRAD
  float* d_disp[kMaxScales];           // FUSED GRAD + smoothness
  float* d_src[kMaxScales];            // FUSED GRAD optional
  long long d_src_bs[kMaxScales], d_src_fs[kMaxScales];
  float* pose_part;                    // FUSED GRAD: [B][slots][N][12]
};

template <bool GRAD>
struct PhotoSmem {
  static constexpr int HL = GRAD ? 2 : 1;
  static constexpr int RW = kTW + 2 * HL, RH = kTH + 2 * HL;        // y / x region
  static constexpr int SW = kTW + 2 * (HL - 1), SH = kTH + 2 * (HL - 1);   // stats region
  static constexpr int kRegion = RW * RH, kStats = SW * SH;
  static constexpr int kFloats = 3 * kRegion * 2 + 3 * kStats * 2 + (GRAD ? 3 * kStats * 3 : 0) + 128;
  static constexpr size_t kBytes = sizeof(float) * kFloats;
};

__device__ __forceinline__ float sgnf(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }

template <bool FUSED, bool GRAD>
__global__ void __launch_bounds__(kPhotoThreads) k_photo(PhotoArgs a) {
  using SM = PhotoSmem<GRAD>;
  constexpr int HL = SM::HL, RW = SM::RW, SW = SM::SW;
  constexpr int HS = HL - 1;     // halo of the stats region
  extern __shared__ float smem[];
  float* sx = smem;                       // [3][RH][RW] target
  float* sy = sx + 3 * SM::kRegion;       // [3][RH][RW] warped
  float* smx = sy + 3 * SM::kRegion;      // [3][SH][SW] mu_x
  float* ssx = smx + 3 * SM::kStats;      // [3][SH][SW] sigma_x
  float* sA = ssx + 3 * SM::kStats;       // GRAD: [3][SH][SW] each
  float* sB = sA + (GRAD ? 3 * SM::kStats : 0);
  float* sC = sB + (GRAD ? 3 * SM::kStats : 0);
  float* red = sC + (GRAD ? 3 * SM::kStats : 0);   // 128 floats

  // ---- which tile -------------------------------------------------------
  int t = blockIdx.x;
  const int b = blockIdx.y;
  int l = 0;
  while (l + 1 < a.lt.S && t >= a.first_tile[l + 1]) ++l;
  t -= a.first_tile[l];
  const Level& L = a.lt.lv[l];
  const int H = L.H, W = L.W, P = H * W;
  const int ty0 = (t / L.tiles_x) * kTH, tx0 = (t % L.tiles_x) * kTW;
  const int tid = threadIdx.x;
  const int slot = L.slot_base + t;

  float K[9], Ki[9];
  if (FUSED) {
    const float* gk = a.geoK + ((size_t)b * a.lt.S + l) * kGeoK;
#pragma unroll
    for (int k = 0; k < 9; ++k) { K[k] = __ldg(gk + k); Ki[k] = __ldg(gk + 9 + k); }
  }

  // ---- target tile + its window statistics (shared by all N sources) ------
  const float* tgt = L.tgt + b * L.tgt_bs;
  for (int i = tid; i < SM::kRegion; i += kPhotoThreads) {
    int ry = i / RW, rx = i % RW;
    int gy = ty0 + ry - HL, gx = tx0 + rx - HL;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      const float* p = tgt + ((long long)gy * W + gx) * 3;
      v0 = __ldg(p); v1 = __ldg(p + 1); v2 = __ldg(p + 2);
    }
    sx[i] = v0; sx[SM::kRegion + i] = v1; sx[2 * SM::kRegion + i] = v2;
  }
  __syncthreads();
  if (a.do_ssim) {
    for (int i = tid; i < SM::kStats; i += kPhotoThreads) {
      int qy = i / SW, qx = i % SW;
      int gy = ty0 + qy - HS, gx = tx0 + qx - HS;
      bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
      int cy = min(gy + 1, H - 1) - max(gy - 1, 0) + 1, cx = min(gx + 1, W - 1) - max(gx - 1, 0) + 1;
      float inv = in ? 1.f / (float)(cy * cx) : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* px = sx + c * SM::kRegion + qy * RW + qx;     // window top-left in region coords
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) { float v = px[dy * RW + dx]; s1 += v; s2 += v * v; }
        float mu = s1 * inv;
        smx[c * SM::kStats + i] = mu;
        ssx[c * SM::kStats + i] = s2 * inv - mu * mu;
      }
    }
  }

  float lsum_l1 = 0.f, lsum_ssim = 0.f, lsum_sm = 0.f;
  const float gb = (GRAD && a.gbatch) ? __ldg(a.gbatch + b) : 1.f;

  // ---- smoothness on this tile (losses.py:409-440), FUSED only -------------
  if (FUSED && a.do_smooth) {
    const float* dsp = a.disp[l] + (long long)b * P;
    const float nx = a.norm_sm_x[l], ny = a.norm_sm_y[l];
    const float gcx = a.gcoef_smooth * gb * nx, gcy = a.gcoef_smooth * gb * ny;
    const float k3 = a.grad_factor;
#pragma unroll
    for (int j = 0; j < kPixPerThread; ++j) {
      int ci = tid + j * kPhotoThreads;
      int cy_ = ci / kTW, cx_ = ci % kTW;
      int gy = ty0 + cy_, gx = tx0 + cx_;
      if (gy < H && gx < W) {
        int ri = (cy_ + HL) * RW + (cx_ + HL);
        float d = __ldg(dsp + (long long)gy * W + gx);
        float gd = 0.f;
        // forward differences owned by this pixel
        if (gx + 1 < W) {
          float e = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) e += fabsf((sx[c * SM::kRegion + ri] - sx[c * SM::kRegion + ri + 1]) * k3);
          float w = expf(-(e / 3.f));
          float sd = (d - __ldg(dsp + (long long)gy * W + gx + 1)) * w;
          lsum_sm += fabsf(sd) * nx;
          gd += gcx * sgnf(sd) * w;
        }
        if (gy + 1 < H) {
          float e = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) e += fabsf((sx[c * SM::kRegion + ri] - sx[c * SM::kRegion + ri + RW]) * k3);
          float w = expf(-(e / 3.f));
          float sd = (d - __ldg(dsp + (long long)(gy + 1) * W + gx)) * w;
          lsum_sm += fabsf(sd) * ny;
          gd += gcy * sgnf(sd) * w;
        }
        if (GRAD) {
          // differences owned by the left / upper neighbour, where this pixel is the subtrahend
          if (gx >= 1) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) e += fabsf((sx[c * SM::kRegion + ri - 1] - sx[c * SM::kRegion + ri]) * k3);
            float w = expf(-(e / 3.f));
            float sd = (__ldg(dsp + (long long)gy * W + gx - 1) - d) * w;
            gd -= gcx * sgnf(sd) * w;
          }
          if (gy >= 1) {
            float e = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) e += fabsf((sx[c * SM::kRegion + ri - RW] - sx[c * SM::kRegion + ri]) * k3);
            float w = expf(-(e / 3.f));
            float sd = (__ldg(dsp + (long long)(gy - 1) * W + gx) - d) * w;
            gd -= gcy * sgnf(sd) * w;
          }
          if (a.d_disp[l]) a.d_disp[l][(long long)b * P + gy * W + gx] = gd;
        }
      }
    }
  }

  // ---- per-thread centre pixels -------------------------------------------
  float gD[kPixPerThread];
  float Dc[kPixPerThread];
#pragma unroll
  for (int j = 0; j < kPixPerThread; ++j) {
    gD[j] = 0.f; Dc[j] = 0.f;
    if (FUSED && GRAD) {
      int ci = tid + j * kPhotoThreads;
      int gy = ty0 + ci / kTW, gx = tx0 + ci % kTW;
      if (gy < H && gx < W) Dc[j] = __ldg(a.depth[l] + (long long)b * P + gy * W + gx);
    }
  }
  const float cl1 = a.gcoef_l1 * gb * a.norm_photo[l];
  const float cssim = a.gcoef_ssim * gb * a.norm_photo[l];

  for (int n = 0; n < a.N; ++n) {
    float T[12];
    const float* img = nullptr;
    if (FUSED) {
      const float* gt = a.geoT + ((size_t)b * a.N + n) * kGeoT;
#pragma unroll
      for (int k = 0; k < 12; ++k) T[k] = __ldg(gt + k);
      img = L.src + b * L.src_bs + n * L.src_fs;
    }
    // ---- phase Y: warped tile with halo into shared memory -----------------
    for (int i = tid; i < SM::kRegion; i += kPhotoThreads) {
      int ry = i / RW, rx = i % RW;
      int gy = ty0 + ry - HL, gx = tx0 + rx - HL;
      float yv[3] = {0.f, 0.f, 0.f};
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        if (FUSED) {
          float D = __ldg(a.depth[l] + (long long)b * P + gy * W + gx);
          float r0, r1, r2;
          ray_of_pixel(Ki, (float)gx, (float)gy, r0, r1, r2);
          Proj pr = project(K, T, r0, r1, r2, D);
          bool valid;
          warp_sample(img, W, H, pr.u, pr.v, D, yv, valid);
          bool centre = ry >= HL && ry < HL + kTH && rx >= HL && rx < HL + kTW;
          if (centre) {
            long long o = (long long)(b * a.N + n) * P + gy * W + gx;
            if (a.synth_out[l]) { float* so = a.synth_out[l] + o * 3; so[0] = yv[0]; so[1] = yv[1]; so[2] = yv[2]; }
            if (a.mask_out[l]) a.mask_out[l][o] = valid ? 1.f : 0.f;
          }
        } else {
          const float* p = a.synth[l] + ((long long)(b * a.N + n) * P + gy * W + gx) * 3;
          yv[0] = __ldg(p); yv[1] = __ldg(p + 1); yv[2] = __ldg(p + 2);
        }
      }
      sy[i] = yv[0]; sy[SM::kRegion + i] = yv[1]; sy[2 * SM::kRegion + i] = yv[2];
    }
    __syncthreads();

    // ---- phase S: per-pixel L1 / SSIM over the stats region -----------------
    for (int i = tid; i < SM::kStats; i += kPhotoThreads) {
      int qy = i / SW, qx = i % SW;
      int gy = ty0 + qy - HS, gx = tx0 + qx - HS;
      bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
      bool centre = in && qy >= HS && qy < HS + kTH && qx >= HS && qx < HS + kTW;
      int ri = (qy + 1) * RW + (qx + 1);          // this pixel in region coords
      float y0 = sy[ri], y1 = sy[SM::kRegion + ri], y2 = sy[2 * SM::kRegion + ri];
      bool masked = ((y0 + y1) + y2) == 0.f;       // mean_c(synth) == 0 (loss_util.py:15-16)
      if (centre && !masked && a.l1_kind) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float df = sy[c * SM::kRegion + ri] - sx[c * SM::kRegion + ri];
          lsum_l1 += (a.l1_kind == 1) ? fabsf(df) : df * df;
        }
      }
      if (a.do_ssim) {
        int cy = min(gy + 1, H - 1) - max(gy - 1, 0) + 1, cx = min(gx + 1, W - 1) - max(gx - 1, 0) + 1;
        float inv = in ? 1.f / (float)(cy * cx) : 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float A = 0.f, Bq = 0.f, Cq = 0.f;
          if (in && !masked) {
            const float* px = sx + c * SM::kRegion + qy * RW + qx;
            const float* py = sy + c * SM::kRegion + qy * RW + qx;
            float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                float yv = py[dy * RW + dx], xv = px[dy * RW + dx];
                s1 += yv; s2 += yv * yv; s3 += xv * yv;
              }
            float mux = smx[c * SM::kStats + i], sgx = ssx[c * SM::kStats + i];
            float muy = s1 * inv;
            float sgy = s2 * inv - muy * muy;
            float sgxy = s3 * inv - mux * muy;
            float a1 = 2.f * mux * muy + kC1, a2 = 2.f * sgxy + kC2;
            float b1 = mux * mux + muy * muy + kC1, b2 = sgx + sgy + kC2;
            float ssim = (a1 * a2) / (b1 * b2);
            float lv = (1.f - ssim) * 0.5f;
            bool pass = (lv >= 0.f) && (lv <= 1.f);          // clip_by_value gradient
            float lc = fminf(fmaxf(lv, 0.f), 1.f);
            if (centre) lsum_ssim += lc;
            if (GRAD && pass) {
              // h = dL/d ssim at this pixel
              float h = -0.5f * cssim;
              float ib1b2 = 1.f / (b1 * b2);
              float dm = (2.f * mux * (a2 - a1)) * ib1b2 - ssim * (2.f * muy) * (1.f / b1 - 1.f / b2);
              float dq = -ssim / b2;
              float dr = 2.f * a1 * ib1b2;
              A = h * dm * inv; Bq = h * dq * inv; Cq = h * dr * inv;
            }
          }
          if (GRAD) { sA[c * SM::kStats + i] = A; sB[c * SM::kStats + i] = Bq; sC[c * SM::kStats + i] = Cq; }
        }
      }
    }

    // ---- phase G: dL/dy at the centre pixels, then out or through the warp --
    if (GRAD) {
      __syncthreads();
      float acc[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) acc[k] = 0.f;
#pragma unroll
      for (int j = 0; j < kPixPerThread; ++j) {
        int ci = tid + j * kPhotoThreads;
        int cy_ = ci / kTW, cx_ = ci % kTW;
        int gy = ty0 + cy_, gx = tx0 + cx_;
        if (gy < H && gx < W) {
          int ri = (cy_ + HL) * RW + (cx_ + HL);
          float yv[3] = {sy[ri], sy[SM::kRegion + ri], sy[2 * SM::kRegion + ri]};
          bool masked = ((yv[0] + yv[1]) + yv[2]) == 0.f;
          float g[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float gc = 0.f;
            float xv = sx[c * SM::kRegion + ri];
            if (a.do_ssim) {
              // centre pixel (cy_,cx_) has stats coords (cy_+HS, cx_+HS); its window starts one up-left
              const float* pA = sA + c * SM::kStats + cy_ * SW + cx_;
              const float* pB = sB + c * SM::kStats + cy_ * SW + cx_;
              const float* pC = sC + c * SM::kStats + cy_ * SW + cx_;
              float sa = 0.f, sb = 0.f, sc = 0.f;
#pragma unroll
              for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) { sa += pA[dy * SW + dx]; sb += pB[dy * SW + dx]; sc += pC[dy * SW + dx]; }
              gc = sa + 2.f * yv[c] * sb + xv * sc;
            }
            if (a.l1_kind && !masked) {
              float df = yv[c] - xv;
              gc += cl1 * ((a.l1_kind == 1) ? sgnf(df) : 2.f * df);
            }
            g[c] = gc;
          }
          if (FUSED) {
            float r0, r1, r2;
            ray_of_pixel(Ki, (float)gx, (float)gy, r0, r1, r2);
            Proj pr = project(K, T, r0, r1, r2, Dc[j]);
            float* dimg = a.d_src[l] ? a.d_src[l] + b * a.d_src_bs[l] + n * a.d_src_fs[l] : nullptr;
            gD[j] += sample_adjoint(img, dimg, W, H, K, T, r0, r1, r2, Dc[j], pr, g, acc);
          } else {
            float* go = a.gsynth[l] + ((long long)(b * a.N + n) * P + gy * W + gx) * 3;
            go[0] = g[0]; go[1] = g[1]; go[2] = g[2];
          }
        }
      }
      if (FUSED) {
        float* out = a.pose_part + (((size_t)b * a.slots_per_b + slot) * a.N + n) * 12;
        block_reduce_store12<kPhotoThreads>(acc, out, red);
      }
    }
    __syncthreads();     // sy / sA.. are rewritten by the next source
  }

  if (FUSED && GRAD) {
#pragma unroll
    for (int j = 0; j < kPixPerThread; ++j) {
      int ci = tid + j * kPhotoThreads;
      int gy = ty0 + ci / kTW, gx = tx0 + ci % kTW;
      if (gy < H && gx < W && a.d_depth[l]) a.d_depth[l][(long long)b * P + gy * W + gx] = gD[j];
    }
  }

  // ---- block reduction of the loss sums -> one partial record per tile ------
  {
    float v0 = warp_sum(lsum_l1), v1 = warp_sum(lsum_ssim), v2 = warp_sum(lsum_sm);
    const int lane = tid & 31, wid = tid >> 5;
    if (lane == 0) { red[wid * 3] = v0; red[wid * 3 + 1] = v1; red[wid * 3 + 2] = v2; }
    __syncthreads();
    if (tid < 3) {
      float v = 0.f;
      for (int w = 0; w < kPhotoThreads / 32; ++w) v += red[w * 3 + tid];
      float nrm = (tid == 2) ? 1.f : a.norm_photo[l];
      a.loss_part[((size_t)b * a.slots_per_b + slot) * 3 + tid] = v * nrm;
    }
  }
}

// ---------------------------------------------------------------------------
// standalone smoothness (SmoothenessLossMultiScale on given tensors):
// one thread per pixel of (level, b); forward sums + gather-form backward.
// ---------------------------------------------------------------------------
struct SmoothArgs {
  int S, B;
  int H[kMaxScales], W[kMaxScales];
  const float* disp[kMaxScales];
  const float* tgt[kMaxScales]; long long tgt_bs[kMaxScales];
  float norm_x[kMaxScales], norm_y[kMaxScales];
  float grad_factor;
  const float* gbatch;       // NULL = ones
  float* d_disp[kMaxScales]; // NULL = forward only
  float* loss_part;          // [B][slots][3] (column 2)
  int slots_per_b;
  int chunk_base[kMaxScales];
};

__device__ __forceinline__ float smooth_weight(const float* __restrict__ p, const float* __restrict__ q, float k3) {
  float e = fabsf((__ldg(p) - __ldg(q)) * k3) + fabsf((__ldg(p + 1) - __ldg(q + 1)) * k3) +
            fabsf((__ldg(p + 2) - __ldg(q + 2)) * k3);
  return expf(-(e / 3.f));
}

__global__ void __launch_bounds__(256) k_smooth(SmoothArgs a) {
  __shared__ float red[8];
  const int l = blockIdx.z, b = blockIdx.y;
  const int H = a.H[l], W = a.W[l], P = H * W;
  if ((int)blockIdx.x * 256 >= P) return;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  float ls = 0.f;
  if (pix < P) {
    int y = pix / W, x = pix % W;
    const float* dsp = a.disp[l] + (long long)b * P;
    const float* img = a.tgt[l] + b * a.tgt_bs[l];
    const float* ip = img + (long long)pix * 3;
    float d = __ldg(dsp + pix);
    float gb = a.gbatch ? __ldg(a.gbatch + b) : 1.f;
    float gd = 0.f;
    float nx = a.norm_x[l], ny = a.norm_y[l], k3 = a.grad_factor;
    if (x + 1 < W) {
      float w = smooth_weight(ip, ip + 3, k3);
      float sd = (d - __ldg(dsp + pix + 1)) * w;
      ls += fabsf(sd) * nx; gd += gb * nx * sgnf(sd) * w;
    }
    if (y + 1 < H) {
      float w = smooth_weight(ip, ip + (long long)W * 3, k3);
      float sd = (d - __ldg(dsp + pix + W)) * w;
      ls += fabsf(sd) * ny; gd += gb * ny * sgnf(sd) * w;
    }
    if (a.d_disp[l]) {
      if (x >= 1) {
        float w = smooth_weight(ip - 3, ip, k3);
        float sd = (__ldg(dsp + pix - 1) - d) * w;
        gd -= gb * nx * sgnf(sd) * w;
      }
      if (y >= 1) {
        float w = smooth_weight(ip - (long long)W * 3, ip, k3);
        float sd = (__ldg(dsp + pix - W) - d) * w;
        gd -= gb * ny * sgnf(sd) * w;
      }
      a.d_disp[l][(long long)b * P + pix] = gd;
    }
  }
  float v = warp_sum(ls);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    float* o = a.loss_part + ((size_t)b * a.slots_per_b + a.chunk_base[l] + blockIdx.x) * 3;
    o[0] = 0.f; o[1] = 0.f; o[2] = s;
  }
}

// ---------------------------------------------------------------------------
// loss epilogue: deterministic fp64 sum of the per-tile partials.
//   loss_batch[k][b] = sum_slots part[b][slot][k]          (merge_multi_scale_losses, losses.py:147-154)
//   mean_k = sum_b loss_batch[k][b] / global_batch        (compute_average_loss, losses.py:49)
//   total  = sum_k w_k * mean_k                            (losses.py:50-54)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pose_epilogue(const float* __restrict__ pose_part, int slots_per_b,
                                                       int slots_used, const float* __restrict__ pose,
                                                       float* __restrict__ d_pose, int N, float /*scale*/) {
  pose_block(pose_part, slots_per_b, slots_used, pose, d_pose, N, blockIdx.x);
}

__global__ void __launch_bounds__(256) k_loss_epilogue(const float* __restrict__ loss_part, int slots_per_b,
                                                       int slots_used, int B, float inv_global_batch, float w0,
                                                       float w1, float w2, float* __restrict__ losses,
                                                       float* __restrict__ loss_batch) {
  __shared__ double red[8][3];
  __shared__ double tot[3];
  if (threadIdx.x < 3) tot[threadIdx.x] = 0.0;
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int sl = threadIdx.x; sl < slots_used; sl += blockDim.x) {
      const float* p = loss_part + ((size_t)b * slots_per_b + sl) * 3;
      acc[0] += p[0]; acc[1] += p[1]; acc[2] += p[2];
    }
    for (int k = 0; k < 3; ++k) {
      double v = acc[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double v = 0.0;
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      if (loss_batch) loss_batch[threadIdx.x * B + b] = (float)v;
      tot[threadIdx.x] += v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && losses) {
    double m0 = tot[0] * inv_global_batch, m1 = tot[1] * inv_global_batch, m2 = tot[2] * inv_global_batch;
    losses[0] = (float)(w0 * m0 + w1 * m1 + w2 * m2);
    losses[1] = (float)m0; losses[2] = (float)m1; losses[3] = (float)m2;
  }
}

// ---------------------------------------------------------------------------
// merged epilogue of the fused path: blocks [0, B*N) reduce the pose partials and apply the Rodrigues
// adjoint (same code as k_pose_epilogue); blocks [B*N, B*N+B) reduce one snippet's loss partials; the
// last of those to finish (atomic ticket) sums the snippets in fixed order -> losses[4].
// ---------------------------------------------------------------------------
struct EpilogueArgs {
  const float* pose_part; const float* loss_part;
  int slots_per_b, slots_used, B, N;
  const float* pose; float* d_pose;            // d_pose may be NULL (forward only)
  float inv_global_batch, w0, w1, w2;
  float* losses; float* loss_batch;            // loss_batch may be NULL
  double* loss_sum_b;                          // [B][3] scratch
  unsigned int* ticket;                        // zero before the first launch; self-resetting
  // multi-rank (reference distributer.py:93-110): nranks > 1 sums the 4 loss scalars over the ranks INSIDE this kernel,
  // through peer memory over NVLink -- no collective launch behind the step (see loss_exchange below)
  int nranks, rank;
  float* const* peer_inbox;                    // device table [nranks]: every rank's inbox [2][nranks][kInboxSlot] (own included)
  unsigned int* seq;                           // this rank's step counter
  unsigned int* exchange_error;                // set when a peer did not arrive within the spin budget
  long long spin_budget;                       // clock64 ticks
};
constexpr int kInboxSlot = 8;                  // floats per (parity, source rank): 4 x (value, step number) pairs (32 bytes)
constexpr int kMaxRanks = 64;

// The path's only exchange, fused into the epilogue: every rank PUSHES its 4 loss scalars into slot [step parity][rank]
// of every rank's inbox with peer stores (P2P over NVLink / NVSwitch, mapped through CUDA IPC), then waits until its
// own inbox holds this step's record of every rank and adds them up in rank order -- so all ranks obtain the
// bit-identical sum.  One thread per peer.  Each scalar travels as ONE aligned 8-byte store {value, step number}
// (single-copy atomic), so a record validates itself and neither side needs a fence: the latency is one NVLink write.
// Two parity slots are enough: a rank cannot start step s+2 before every rank has published step s+1, i.e. after
// every rank has consumed step s.  A peer that does not arrive within the budget (ranks issuing different numbers of
// steps) raises *exchange_error instead of hanging the GPU.
__device__ __forceinline__ void loss_exchange(const EpilogueArgs& a, const float loc[4], unsigned s, float (*got)[4], int r) {
  const int par = (int)(s & 1u);
  {
    float* slot = a.peer_inbox[r] + (size_t)(par * a.nranks + a.rank) * kInboxSlot;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      asm volatile("st.relaxed.sys.global.v2.b32 [%0], {%1, %2};" ::"l"(slot + 2 * k), "r"(__float_as_uint(loc[k])), "r"(s) : "memory");
  }
  {
    const float* slot = a.peer_inbox[a.rank] + (size_t)(par * a.nranks + r) * kInboxSlot;
    const long long t0 = clock64();
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned v = 0u, seen = 0u;
      while (ok) {
        asm volatile("ld.relaxed.sys.global.v2.b32 {%0, %1}, [%2];" : "=r"(v), "=r"(seen) : "l"(slot + 2 * k) : "memory");
        if (seen == s) break;
        if (clock64() - t0 > a.spin_budget) ok = false;
      }
      got[r][k] = ok ? __uint_as_float(v) : 0.f;
    }
    if (!ok) *a.exchange_error = 1u;
  }
}

__global__ void __launch_bounds__(128) k_epilogue(EpilogueArgs a) {
  __shared__ double red[4][3];
  __shared__ bool last;
  const int nb_pose = a.d_pose ? a.B * a.N : 0;
  if ((int)blockIdx.x < nb_pose) {
    pose_block(a.pose_part, a.slots_per_b, a.slots_used, a.pose, a.d_pose, a.N, blockIdx.x);
    return;
  }
  const int b = blockIdx.x - nb_pose;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int sl = threadIdx.x; sl < a.slots_used; sl += blockDim.x) {
    const float* p = a.loss_part + ((size_t)b * a.slots_per_b + sl) * 3;
    acc[0] += p[0]; acc[1] += p[1]; acc[2] += p[2];
  }
  for (int k = 0; k < 3; ++k) {
    double v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    a.loss_sum_b[b * 3 + threadIdx.x] = v;
    if (a.loss_batch) a.loss_batch[threadIdx.x * a.B + b] = (float)v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(a.ticket, 1u) == (unsigned)(a.B - 1);
  __syncthreads();
  if (!last) return;                            // block-uniform
  __shared__ float loc[4];
  __shared__ unsigned step;
  __shared__ float got[kMaxRanks][4];
  if (threadIdx.x == 0) {
    __threadfence();
    double tot[3] = {0.0, 0.0, 0.0};
    for (int bb = 0; bb < a.B; ++bb)
      for (int k = 0; k < 3; ++k) tot[k] += ((volatile double*)a.loss_sum_b)[bb * 3 + k];
    const double m0 = tot[0] * a.inv_global_batch, m1 = tot[1] * a.inv_global_batch, m2 = tot[2] * a.inv_global_batch;
    loc[0] = (float)(a.w0 * m0 + a.w1 * m1 + a.w2 * m2);
    loc[1] = (float)m0; loc[2] = (float)m1; loc[3] = (float)m2;
    *a.ticket = 0u;
    if (a.nranks > 1) { step = *a.seq + 1u; *a.seq = step; }
    else { a.losses[0] = loc[0]; a.losses[1] = loc[1]; a.losses[2] = loc[2]; a.losses[3] = loc[3]; }
  }
  if (a.nranks <= 1) return;
  __syncthreads();
  if ((int)threadIdx.x < a.nranks) loss_exchange(a, loc, step, got, threadIdx.x);
  __syncthreads();
  if (threadIdx.x < 4) {
    float v = 0.f;
    for (int r = 0; r < a.nranks; ++r) v += got[r][threadIdx.x];      // rank order: the same sum on every rank
    a.losses[threadIdx.x] = v;
  }
}

// device staging -> pinned host memory THROUGH its device mapping, 16 bytes per thread and fully coalesced
// (each warp emits 512 contiguous bytes): the small outputs of the host-buffer entry point leave in one launch
struct DrainArgs {
  int n;
  const float* src[16];
  float* dst[16];
  long long count[16];
};

__global__ void __launch_bounds__(256) k_drain(DrainArgs a) {
  const int seg = blockIdx.y;
  const float* __restrict__ src = a.src[seg];
  float* __restrict__ dst = a.dst[seg];
  const long long n = a.count[seg];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0) {
    const long long n4 = n >> 2;
    for (long long i = t; i < n4; i += stride) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
    for (long long i = (n4 << 2) + t; i < n; i += stride) dst[i] = src[i];
  } else {
    for (long long i = t; i < n; i += stride) dst[i] = src[i];
  }
}

// ---------------------------------------------------------------------------
// stereo rig poses (losses.py:105-140, :481-495; utils/convert_pose.py:151-168)
// ---------------------------------------------------------------------------
// general 4x4 inverse (tf.linalg.inv): Gauss-Jordan with partial pivoting in fp64, rounded once
__device__ inline void invert4x4(const float* __restrict__ m, float* __restrict__ out) {
  double a[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) { a[r][c] = m[r * 4 + c]; a[r][4 + c] = r == c ? 1.0 : 0.0; }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    for (int r = col + 1; r < 4; ++r)
      if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
    if (piv != col)
      for (int c = 0; c < 8; ++c) { double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
    const double inv = 1.0 / a[col][col];
    for (int c = 0; c < 8; ++c) a[col][c] *= inv;
    for (int r = 0; r < 4; ++r) {
      if (r == col) continue;
      const double f = a[r][col];
      for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
    }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) out[r * 4 + c] = (float)a[r][4 + c];
}

// pose_matr2rvec_batch: fp32, the reference's operation order (theta = acos((tr R - 1)/2) is ill-conditioned
// for the small rig rotations, so the SAME roundings as the reference matter more than extra precision)
__device__ inline void matr2rvec(const float* __restrict__ m, float* __restrict__ out) {
  const float tr = (m[0] + m[5]) + m[10];
  const float theta = acosf((tr - 1.f) / 2.f);
  const float ax[3] = {m[6] - m[9], m[8] - m[2], m[1] - m[4]};
  out[0] = m[3]; out[1] = m[7]; out[2] = m[11];
  if (fabsf(theta) < 0.00001f) {
    for (int k = 0; k < 3; ++k) out[3 + k] = ax[k] / 2.f;
  } else {
    const float den = 2.f * sinf(theta);
    for (int k = 0; k < 3; ++k) out[3 + k] = ax[k] / den * theta;
  }
}

__global__ void k_pose_matr2rvec(const float* __restrict__ matr, int count, int invert, float* __restrict__ rvec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float m[16];
  if (invert) invert4x4(matr + (size_t)i * 16, m);
  else for (int k = 0; k < 16; ++k) m[k] = matr[(size_t)i * 16 + k];
  matr2rvec(m, rvec + (size_t)i * 6);
}

// StereoPoseLoss: loss[b] = mean_n [ MSE6(rvec(T_LR), pose_lr[b,n]) + MSE6(rvec(inv T_LR), pose_rl[b,n]) ]
// and its gradient w.r.t. both predictions scaled by gscale[b] (NULL = 1).  One thread per snippet.
__global__ void k_stereo_pose_loss(const float* __restrict__ T_LR, const float* __restrict__ pose_lr,
                                   const float* __restrict__ pose_rl, int B, int n, float* __restrict__ loss_batch,
                                   const float* __restrict__ gscale, float* __restrict__ d_lr, float* __restrict__ d_rl) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float inv[16], lr[6], rl[6];
  matr2rvec(T_LR + (size_t)b * 16, lr);
  invert4x4(T_LR + (size_t)b * 16, inv);
  matr2rvec(inv, rl);
  const float g = (gscale ? gscale[b] : 1.f) * (2.f / (6.f * (float)n));
  float acc = 0.f;
  for (int j = 0; j < n; ++j) {
    float s0 = 0.f, s1 = 0.f;
    for (int k = 0; k < 6; ++k) {
      const size_t o = ((size_t)b * n + j) * 6 + k;
      const float e0 = lr[k] - pose_lr[o], e1 = rl[k] - pose_rl[o];
      s0 += e0 * e0; s1 += e1 * e1;
      if (d_lr) d_lr[o] = -g * e0;
      if (d_rl) d_rl[o] = -g * e1;
    }
    acc += s0 / 6.f + s1 / 6.f;
  }
  if (loss_batch) loss_batch[b] = acc / (float)n;
}

__global__ void k_fill(float* __restrict__ p, long long n, float v) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace xpt
