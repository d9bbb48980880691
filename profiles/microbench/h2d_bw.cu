// Pinned host -> device copy bandwidth by allocation flavour and copy size (bounds bench.py's e2e).
// build: nvcc -O2 -o profiles/microbench/h2d_bw profiles/microbench/h2d_bw.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
static float run(void* d, void* h, size_t bytes, int reps, bool h2d, cudaStream_t st) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaMemcpyAsync(h2d ? d : h, h2d ? h : d, bytes, h2d ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; ++i) cudaMemcpyAsync(h2d ? d : h, h2d ? h : d, bytes, h2d ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, st);
  cudaEventRecord(e1, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return (float)(bytes * (double)reps / (ms * 1e-3) / 1e9);
}
int main() {
  cudaStream_t st; cudaStreamCreate(&st);
  const size_t sizes[] = {1u << 20, 6u << 20, 28u << 20, 128u << 20};
  void* d; cudaMalloc(&d, 128u << 20);
  const struct { const char* name; unsigned flags; } kinds[] = {
      {"default", cudaHostAllocDefault}, {"portable", cudaHostAllocPortable}, {"write-combined", cudaHostAllocWriteCombined}};
  for (auto& k : kinds) {
    void* h; if (cudaHostAlloc(&h, 128u << 20, k.flags) != cudaSuccess) { printf("%s: alloc failed\n", k.name); continue; }
    memset(h, 1, 128u << 20);
    for (size_t s : sizes)
      printf("%-15s %4zu MB  h2d %6.1f GB/s   d2h %6.1f GB/s\n", k.name, s >> 20, run(d, h, s, 20, true, st), run(d, h, s, 20, false, st));
    cudaFreeHost(h);
  }
  // registered malloc memory (what a torch/TF host tensor that is pinned after the fact looks like)
  void* m = aligned_alloc(4096, 128u << 20); memset(m, 1, 128u << 20);
  if (cudaHostRegister(m, 128u << 20, cudaHostRegisterDefault) == cudaSuccess) {
    for (size_t s : sizes) printf("%-15s %4zu MB  h2d %6.1f GB/s   d2h %6.1f GB/s\n", "registered", s >> 20, run(d, m, s, 20, true, st), run(d, m, s, 20, false, st));
    cudaHostUnregister(m);
  }
  return 0;
}
