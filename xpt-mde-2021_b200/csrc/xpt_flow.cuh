// xpt_flow.cuh -- optical-flow warping (reference model/synthesize/flow_warping.py:11-71):
// warped[b,n,y,x,:] = bilinear(source level [b,n], (x, y) - flow[b,n,y,x,:]) with the validity rule of
// bilinear_interp.py:53-76 (all four taps inside the image; there is no depth here), and its backward:
// dL/dflow = -(dL/du, dL/dv) through the bilinear weights, dL/dsource scattered with fp32 atomics.
//
// One thread per (b, n, pixel) of a level; blockIdx.y = level.  The flow levels are small (PWC-Net emits
// H/4 .. H/32, flow_net.py:44-48), so this is a latency-bound gather: 8 bytes of flow in, four 12-byte taps from
// the (L2-resident) source level, 12 bytes out per sample.
#pragma once
#include "xpt_kernels.cuh"

namespace xpt {

struct FlowWarpArgs {
  LevelTable lt;                       // Level.src / src_bs / src_fs: the source level of each flow level
  int B, N;
  const float* flow[kMaxScales];       // [B,N,h,w,2]  (u, v) flow from source to target
  float* warped[kMaxScales];           // fwd: [B,N,h,w,3]
  float* mask[kMaxScales];             // fwd: [B,N,h,w,1] optional
  const float* gwarped[kMaxScales];    // bwd: dL/d warped
  float* d_flow[kMaxScales];           // bwd: [B,N,h,w,2]
  float* d_src[kMaxScales];            // bwd: per-level dL/d source level (NULL = off)
  long long d_src_bs[kMaxScales], d_src_fs[kMaxScales];
};

template <bool BWD>
__global__ void __launch_bounds__(256) k_flow_warp(FlowWarpArgs a) {
  const int l = blockIdx.y;
  const Level& L = a.lt.lv[l];
  const int P = L.H * L.W;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.B * a.N * P) return;
  const int pix = (int)(idx % P);
  const int bn = (int)(idx / P);
  const int b = bn / a.N, n = bn - b * a.N;
  const int y = pix / L.W, x = pix - y * L.W;
  const float2 fl = __ldg(reinterpret_cast<const float2*>(a.flow[l]) + idx);
  // pixel_coords = uvgrid - uvflow (flow_warping.py:69)
  const float u = (float)x - fl.x, v = (float)y - fl.y;
  const Taps t = make_taps(u, v, 1.f, L.W, L.H);
  const float* img = L.src + b * L.src_bs + n * L.src_fs;
  if (!BWD) {
    float yv[3] = {0.f, 0.f, 0.f};
    if (t.valid) {
      float I0[3], I1[3], I2[3], I3[3];
      gather_taps(img, L.W, t, I0, I1, I2, I3);
      const float w0 = t.w_uf * t.w_vf, w1 = t.w_uf * t.w_vc, w2 = t.w_uc * t.w_vf, w3 = t.w_uc * t.w_vc;
#pragma unroll
      for (int c = 0; c < 3; ++c) yv[c] = ((I0[c] * w0 + I1[c] * w1) + I2[c] * w2) + I3[c] * w3;
    }
    float* o = a.warped[l] + idx * 3;
    o[0] = yv[0]; o[1] = yv[1]; o[2] = yv[2];
    if (a.mask[l]) a.mask[l][idx] = t.valid ? 1.f : 0.f;
  } else {
    float gu = 0.f, gv = 0.f;
    if (t.valid) {
      const float* gp = a.gwarped[l] + idx * 3;
      const float g[3] = {__ldg(gp), __ldg(gp + 1), __ldg(gp + 2)};
      float I0[3], I1[3], I2[3], I3[3];
      gather_taps(img, L.W, t, I0, I1, I2, I3);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        gu += g[c] * (t.w_vf * (I2[c] - I0[c]) + t.w_vc * (I3[c] - I1[c]));
        gv += g[c] * (t.w_uf * (I1[c] - I0[c]) + t.w_uc * (I3[c] - I2[c]));
      }
      if (a.d_src[l]) {
        float* dimg = a.d_src[l] + b * a.d_src_bs[l] + n * a.d_src_fs[l];
        const float w0 = t.w_uf * t.w_vf, w1 = t.w_uf * t.w_vc, w2 = t.w_uc * t.w_vf, w3 = t.w_uc * t.w_vc;
        float* p = dimg + ((long long)t.iv * L.W + t.iu) * 3;
        float* q = p + (long long)L.W * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          atomicAdd(p + c, w0 * g[c]);
          atomicAdd(p + 3 + c, w2 * g[c]);
          atomicAdd(q + c, w1 * g[c]);
          atomicAdd(q + 3 + c, w3 * g[c]);
        }
      }
    }
    if (a.d_flow[l]) reinterpret_cast<float2*>(a.d_flow[l])[idx] = make_float2(-gu, -gv);
  }
}

// L2Regularizer (losses.py:522-534): sum(w^2)/2 of one weight tensor, accumulated in fp64 into *acc by the last
// block standing (deterministic two-stage sum: per-block partials, then one thread adds them in order).
__global__ void __launch_bounds__(256) k_l2_partial(const float* __restrict__ w, long long n, double* __restrict__ part) {
  __shared__ double sh[256];
  double v = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double x = (double)__ldg(w + i);
    v += x * x;
  }
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

// out[0] (+)= 0.5 * sum(part[0..n))
__global__ void k_l2_finish(const double* __restrict__ part, int n, float* __restrict__ out, int accumulate) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double v = 0.0;
    for (int i = 0; i < n; ++i) v += part[i];
    const float r = (float)(0.5 * v);
    out[0] = accumulate ? out[0] + r : r;
  }
}

// d_w[i] = scale[0] * w[i]   (adjoint of sum(w^2)/2 with upstream gradient scale[0])
__global__ void __launch_bounds__(256) k_scale_by(const float* __restrict__ w, long long n, const float* __restrict__ scale,
                                                  float* __restrict__ out) {
  const float s = __ldg(scale);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = s * __ldg(w + i);
}

}  // namespace xpt
