// Microbenchmark: how many bytes does one prefetch.global.L2 bring into L2?
// A buffer is prefetched at a given spacing (after L2 was flushed), then read
// in full; the read time tells whether the whole buffer was L2-resident.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o prefetch_gran prefetch_gran.cu
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void flush(float4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void prefetch(const char* p, size_t bytes, size_t spacing) {
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * spacing;
       i < bytes; i += (size_t)gridDim.x * blockDim.x * spacing)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + i));
}
__global__ void rd(const float4* p, size_t n, float* out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldcs(p + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 12345.678f) *out = acc;
}

int main() {
  const size_t bytes = 8u << 20, fbytes = 512u << 20;
  char *buf, *fl;
  float* out;
  cudaMalloc(&buf, bytes);
  cudaMalloc(&fl, fbytes);
  cudaMalloc(&out, 4);
  cudaMemset(buf, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t spacings[] = {0, 32, 64, 128, 256, 512};
  for (size_t sp : spacings) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      flush<<<148 * 8, 256>>>((float4*)fl, fbytes / 16);
      if (sp) prefetch<<<148 * 4, 256>>>(buf, bytes, sp);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      // One small CTA: latency-bound, so L2 hits and misses differ clearly.
      rd<<<1, 128>>>((const float4*)buf, bytes / 16, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("prefetch spacing %4zu B: read of %zu MiB takes %.4f ms = %.0f GB/s (%s)\n",
           sp, bytes >> 20, best, bytes / best / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
