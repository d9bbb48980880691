// What costs H2D throughput in a chunked copy/compute pipeline?  8 x 2.95 MB pinned copies:
//  (a) back to back, (b) + event record after each, (c) + a compute stream waiting on each event and running a
//  50 us kernel, (d) + a D2H copy per chunk on a third stream, (e) two alternating H2D streams.
// build: nvcc -O2 -arch=sm_100a -o profiles/microbench/h2d_pipe profiles/microbench/h2d_pipe.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__global__ void spin(long long cycles) { long long t0 = clock64(); while (clock64() - t0 < cycles) {} }
int main() {
  const size_t chunk = 2949120, n = 8;
  char *h, *d, *h2, *d2; cudaHostAlloc(&h, chunk * n, 0); cudaMalloc(&d, chunk * n); memset(h, 1, chunk * n);
  cudaHostAlloc(&h2, 400000 * n, 0); cudaMalloc(&d2, 400000 * n);
  cudaStream_t sa, sb, sc, so; cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&so, cudaStreamNonBlocking);
  cudaEvent_t ev[8], dn[8], t0, t1; for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  for (auto& e : dn) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  cudaEventCreate(&t0); cudaEventCreate(&t1);
  const long long cyc = 50 * 1965;   // ~50 us
  for (int mode = 0; mode < 5; ++mode) {
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaDeviceSynchronize();
      cudaEventRecord(t0, sc);
      cudaStreamWaitEvent(sa, t0, 0); cudaStreamWaitEvent(sb, t0, 0); cudaStreamWaitEvent(so, t0, 0);
      for (size_t k = 0; k < n; ++k) {
        cudaStream_t s = (mode == 4 && (k & 1)) ? sb : sa;
        cudaMemcpyAsync(d + k * chunk, h + k * chunk, chunk, cudaMemcpyHostToDevice, s);
        if (mode >= 1) cudaEventRecord(ev[k], s);
        if (mode >= 2) { cudaStreamWaitEvent(sc, ev[k], 0); spin<<<148, 128, 0, sc>>>(cyc); }
        if (mode >= 3) { cudaEventRecord(dn[k], sc); cudaStreamWaitEvent(so, dn[k], 0);
                         cudaMemcpyAsync(h2 + k * 400000, d2 + k * 400000, 400000, cudaMemcpyDeviceToHost, so); }
      }
      if (mode < 2) { cudaEventRecord(ev[0], sa); cudaStreamWaitEvent(sc, ev[0], 0); }
      if (mode == 4) { cudaEventRecord(ev[1], sb); cudaStreamWaitEvent(sc, ev[1], 0); }
      if (mode >= 3) { cudaEventRecord(dn[0], so); cudaStreamWaitEvent(sc, dn[0], 0); }
      cudaEventRecord(t1, sc); cudaEventSynchronize(t1);
      float ms; cudaEventElapsedTime(&ms, t0, t1); if (ms < best) best = ms;
    }
    const char* names[] = {"back-to-back", "+event records", "+compute stream waits (50us kernels)", "+D2H per chunk", "two alternating H2D streams (+compute +D2H)"};
    printf("%-46s total %7.1f us  (copies alone would be %.1f us at 55 GB/s)\n", names[mode], best * 1e3f, chunk * n / 55e3);
  }
  return 0;
}
