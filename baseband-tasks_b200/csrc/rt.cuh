// Thin runtime layer: device allocation, copies and kernel launches.
//
// Built with nvcc this maps straight onto the CUDA runtime.  Built with
// -DBBT_EMULATE (tests/emu only) the same kernels run on host threads so that
// index arithmetic can be checked on machines without a GPU; that build is
// test infrastructure and is never loaded by the product package.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string>

#if defined(BBT_EMULATE)
// --------------------------------------------------------------- emulation
#include <math.h>
#include <stdlib.h>
#include <string.h>
struct bbt_emu_dim3 {
  unsigned x, y, z;
  bbt_emu_dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1)
      : x(x_), y(y_), z(z_) {}
};
typedef bbt_emu_dim3 dim3;
struct alignas(16) float4 {
  float x, y, z, w;
};
extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
void* bbt_emu_smem();
void bbt_emu_syncthreads();
#include <functional>
void bbt_emu_launch(dim3 grid, dim3 block, size_t smem,
                    const std::function<void()>& body);
#define BBT_GLOBAL static
#define BBT_DEV static inline
#define BBT_LAUNCH_BOUNDS(t, b)
#define BBT_SMEM(type) reinterpret_cast<type*>(bbt_emu_smem())
#define BBT_RESTRICT __restrict__
#define BBT_LAUNCH(kernel, grid, block, smem, stream, ...)            \
  do {                                                                \
    bbt::prof_count();                                                \
    bbt_emu_launch(grid, block, smem, [=]() { kernel(__VA_ARGS__); }); \
  } while (0)
#define BBT_SET_SMEM(kernel, bytes) 0
typedef void* bbt_stream_t;
namespace bbt {
void prof_count();
extern thread_local const char* prof_next_name;
inline float atomic_add(float* p, float v) {
  uint32_t* ip = reinterpret_cast<uint32_t*>(p);
  uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), nw;
  float f;
  do {
    memcpy(&f, &old, 4);
    f += v;
    memcpy(&nw, &f, 4);
  } while (!__atomic_compare_exchange_n(ip, &old, nw, false, __ATOMIC_RELAXED,
                                        __ATOMIC_RELAXED));
  return f;
}
inline unsigned long long atomic_add(unsigned long long* p,
                                     unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
inline unsigned atomic_add(unsigned* p, unsigned v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
}  // namespace bbt
float* bbt_emu_scratch();
namespace bbt {
// Exchange with the neighbouring thread (lane ^ 1), as __shfl_xor_sync does.
inline float shfl_xor(float v, int d) {
  float* s = bbt_emu_scratch();
  s[threadIdx.x] = v;
  bbt_emu_syncthreads();
  const float r = s[threadIdx.x ^ d];
  bbt_emu_syncthreads();
  return r;
}
inline float shfl_xor1(float v) { return shfl_xor(v, 1); }
inline void sincospi_d(double x, double* s, double* c) {
  *s = sin(M_PI * x);
  *c = cos(M_PI * x);
}
inline double dmul(double a, double b) { return a * b; }
inline double dadd(double a, double b) { return a + b; }
inline double ddiv(double a, double b) { return a / b; }
inline int dev_alloc(void** p, size_t bytes) {
  *p = malloc(bytes ? bytes : 1);
  return *p ? 0 : -1;
}
inline void dev_free(void* p) { free(p); }
inline int h2d(void* dst, const void* src, size_t bytes, bbt_stream_t) {
  memcpy(dst, src, bytes);
  return 0;
}
inline int dev_zero(void* dst, size_t bytes, bbt_stream_t) {
  memset(dst, 0, bytes);
  return 0;
}
inline const char* launch_error() { return nullptr; }
inline int sm_count() { return 4; }
}  // namespace bbt
#else
// --------------------------------------------------------------------- CUDA
#include <cuda_runtime.h>
#define BBT_GLOBAL static __global__
#define BBT_DEV __device__ __forceinline__
#define BBT_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define BBT_SMEM(type) reinterpret_cast<type*>(bbt_dyn_smem)
#define BBT_RESTRICT __restrict__
// Every launch is counted; with profiling on (bbt_profile_enable) it is also
// bracketed by CUDA events on its own stream.
#define BBT_LAUNCH(kernel, grid, block, smem, stream, ...)   \
  do {                                                       \
    bbt::ProfScope bbt_prof_scope(#kernel, stream);          \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);      \
  } while (0)
#define BBT_SET_SMEM(kernel, bytes) \
  bbt::set_max_smem(reinterpret_cast<const void*>(kernel), (bytes))
typedef cudaStream_t bbt_stream_t;
extern __shared__ float4 bbt_dyn_smem[];
namespace bbt {
void prof_count();
void prof_begin(const char* name, cudaStream_t stream);
void prof_end();
// Name used for the next launch instead of the stringified kernel symbol.
extern thread_local const char* prof_next_name;
struct ProfScope {
  ProfScope(const char* name, cudaStream_t stream) {
    prof_count();
    prof_begin(prof_next_name ? prof_next_name : name, stream);
    prof_next_name = nullptr;
  }
  ~ProfScope() { prof_end(); }
};
__device__ __forceinline__ float atomic_add(float* p, float v) {
  return atomicAdd(p, v);
}
__device__ __forceinline__ unsigned long long atomic_add(
    unsigned long long* p, unsigned long long v) {
  return atomicAdd(p, v);
}
__device__ __forceinline__ unsigned atomic_add(unsigned* p, unsigned v) {
  return atomicAdd(p, v);
}
__device__ __forceinline__ void sincospi_d(double x, double* s, double* c) {
  sincospi(x, s, c);
}
__device__ __forceinline__ float shfl_xor(float v, int d) {
  return __shfl_xor_sync(0xffffffffu, v, d);
}
__device__ __forceinline__ float shfl_xor1(float v) {
  return __shfl_xor_sync(0xffffffffu, v, 1);
}
// Explicitly rounded double arithmetic (no FMA contraction): the fold-bin
// assignment must match the oracle's numpy float64 operations bit for bit.
__device__ __forceinline__ double dmul(double a, double b) {
  return __dmul_rn(a, b);
}
__device__ __forceinline__ double dadd(double a, double b) {
  return __dadd_rn(a, b);
}
__device__ __forceinline__ double ddiv(double a, double b) {
  return __ddiv_rn(a, b);
}
inline int dev_alloc(void** p, size_t bytes) {
  return (int)cudaMalloc(p, bytes ? bytes : 1);
}
inline void dev_free(void* p) { cudaFree(p); }
inline int h2d(void* dst, const void* src, size_t bytes, bbt_stream_t s) {
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
}
inline int dev_zero(void* dst, size_t bytes, bbt_stream_t s) {
  return (int)cudaMemsetAsync(dst, 0, bytes, s);
}
inline const char* launch_error() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
int sm_count();  // cached per device (bbt_core.cu)
int set_max_smem(const void* kernel, size_t bytes);
}  // namespace bbt
#endif
