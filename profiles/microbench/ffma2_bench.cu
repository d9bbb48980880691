// Microbenchmark: issue throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// 8 independent chains per thread; reports achieved lane-FMA per clock per SM.
#include <cuda_runtime.h>
#include <cstdio>

template <bool PACKED>
__global__ void k(float* out, int iters, float a, float b) {
  float2 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = make_float2(threadIdx.x * 0.001f + i, blockIdx.x * 0.002f - i);
  const float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) {
        v[i] = __ffma2_rn(v[i], aa, bb);
      } else {
        v[i].x = fmaf(v[i].x, aa.x, bb.x);
        v[i].y = fmaf(v[i].y, aa.y, bb.y);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool PACKED>
double run(int blocks, int threads, int iters) {
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PACKED><<<blocks, threads>>>(out, iters, 1.0001f, 0.0001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<PACKED><<<blocks, threads>>>(out, iters, 1.0001f, 0.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(out);
  return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = 256, blocks = sms * (warps * 32 / threads > 0 ? warps * 32 / threads : 1);
    if (warps * 32 < threads) { threads = warps * 32; blocks = sms; }
    int iters = 20000;
    double ms_s = run<false>(blocks, threads, iters), ms_p = run<true>(blocks, threads, iters);
    double fmas = (double)blocks * threads * iters * 16.0;
    printf("warps/SM=%2d  scalar FFMA: %.3f ms  %.1f GFMA/s   packed FFMA2: %.3f ms  %.1f GFMA/s   ratio %.2f\n", warps,
           ms_s, fmas / ms_s / 1e6, ms_p, fmas / ms_p / 1e6, ms_s / ms_p);
  }
  return 0;
}
