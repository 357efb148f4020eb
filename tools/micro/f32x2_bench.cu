// Microbenchmark: issue rate of packed FP32 (FFMA2/FADD2) against scalar
// FFMA/FADD on sm_100a, alone and mixed with integer ALU work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_bench f32x2_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float2 unpack(u64 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

constexpr int ITER = 4096;

// MODE 0: 16 scalar FFMA chains.  1: 8 FFMA2 chains (same flops).
// 2: 16 scalar FADD chains.       3: 8 FADD2 chains.
// 4: 16 FFMA + 8 integer ops per iteration.  5: 8 FFMA2 + 8 integer ops.
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s, int n) {
  float a[16];
  u64 p[8];
  unsigned z[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pack(a[2 * i], a[2 * i + 1]), z[i] = threadIdx.x + i;
  const u64 sp = pack(s, s * 1.0001f);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (MODE == 0 || MODE == 4) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.5f + a[(i + 1) & 15] * 0.f);
      } else if (MODE == 1 || MODE == 5) {
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], sp, sp);
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = a[i] + s;
      } else if (MODE == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = add2(p[i], sp);
      }
      if (MODE >= 4) {
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = (z[i] ^ (z[i] << 3)) + it;
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 q = unpack(p[i]);
    acc += q.x + q.y + (float)z[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, float* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = 148 * 8;
  k<MODE><<<grid, 256>>>(out, 1.0001f, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, 1.0001f, ITER);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // 16 fp32 values updated 4 times per iteration per thread
  const double ops = (double)grid * 256 * ITER * 4 * 16;
  printf("%-28s %8.3f ms  %7.2f T value-updates/s  (%s)\n", name, ms,
         ops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<0>("FFMA x16", out);
  run<1>("FFMA2 x8", out);
  run<2>("FADD x16", out);
  run<3>("FADD2 x8", out);
  run<4>("FFMA x16 + 16 int", out);
  run<5>("FFMA2 x8 + 16 int", out);
  return 0;
}
