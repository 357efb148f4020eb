// Bulk asynchronous copies (TMA) and mbarriers for sm_100a: thin wrappers of
// the PTX the streaming kernels use to move whole tiles between HBM and
// shared memory without spending registers or LSU issue slots.
//   cp.async.bulk             contiguous runs (16-byte multiples)
//   cp.async.bulk.tensor.Nd   strided boxes described by a CUtensorMap
//   cp.async.bulk.prefetch*   the same ranges into L2 only
// Completion of loads is counted in bytes on an mbarrier (expect_tx); stores
// are tracked with bulk groups.
//
// With -DBBT_EMULATE (tests/emu) the copies are done synchronously by the
// issuing thread and an mbarrier is a counter of completed phases, so that the
// index logic of the pipelines can be checked on a machine without a GPU.
#pragma once
#include <stdint.h>

#if defined(BBT_EMULATE)
#include <string.h>

#include <atomic>
#include <thread>
namespace bbt {
// Emulated tensor map: up to 3 dimensions of 8-byte... (any element size).
struct TensorMap {
  char* base;
  uint64_t dim[3];       // elements, fastest first
  uint64_t stride[3];    // bytes; stride[0] = element size
  uint32_t box[3];
  int rank;
};
struct alignas(8) Mbar {
  std::atomic<uint32_t> phases_done;
  uint32_t pad;
};
inline void mbar_init(Mbar* b, int) { b->phases_done.store(0); }
inline void mbar_expect_tx(Mbar*, uint32_t) {}
inline void mbar_complete_emu(Mbar* b) { b->phases_done.fetch_add(1); }
// Complete a phase without any bytes (an empty tile).
inline void mbar_arrive(Mbar* b) { mbar_complete_emu(b); }
// Wait until phase number `k` (0-based count of completed phases) is done.
inline void mbar_wait(Mbar* b, uint32_t parity, uint32_t k_phase) {
  (void)parity;
  while (b->phases_done.load() <= k_phase) std::this_thread::yield();
}
inline void fence_proxy_async() {}
inline void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                      Mbar* bar, bool last = true) {
  memcpy(smem_dst, gsrc, bytes);
  if (last) mbar_complete_emu(bar);
}
inline void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  memcpy(gdst, smem_src, bytes);
}
inline void bulk_commit() {}
inline void bulk_wait_read0() {}
inline void bulk_wait_all0() {}
inline void bulk_prefetch_l2(const void*, uint32_t) {}
// Named barriers (bar.sync / bar.arrive with an id and a thread count): only
// orderings between groups of warps that help performance are expressed with
// them, so the emulation needs nothing for them.
inline void named_bar_sync(int, int) {}
inline void named_bar_arrive(int, int) {}
inline void tensor_prefetch_2d(const TensorMap*, int, int) {}
inline void tensor_prefetch_3d(const TensorMap*, int, int, int) {}
inline void tensor_copy_emu(const TensorMap* m, void* smem, const int* c,
                            bool load) {
  const uint64_t es = m->stride[0];
  char* s = static_cast<char*>(smem);
  const uint32_t b0 = m->box[0], b1 = m->rank > 1 ? m->box[1] : 1,
                 b2 = m->rank > 2 ? m->box[2] : 1;
  for (uint32_t k = 0; k < b2; ++k)
    for (uint32_t j = 0; j < b1; ++j)
      for (uint32_t i = 0; i < b0; ++i) {
        const uint64_t x = c[0] + i, y = m->rank > 1 ? c[1] + j : 0,
                       z = m->rank > 2 ? c[2] + k : 0;
        const bool in = x < m->dim[0] && (m->rank < 2 || y < m->dim[1]) &&
                        (m->rank < 3 || z < m->dim[2]);
        char* sp = s + ((uint64_t)(k * b1 + j) * b0 + i) * es;
        char* gp = m->base + x * es + (m->rank > 1 ? y * m->stride[1] : 0) +
                   (m->rank > 2 ? z * m->stride[2] : 0);
        if (load) {
          if (in) memcpy(sp, gp, es); else memset(sp, 0, es);
        } else if (in) {
          memcpy(gp, sp, es);
        }
      }
}
inline void tensor_load_2d(void* smem, const TensorMap* m, int c0, int c1,
                           Mbar* bar, bool last = true) {
  const int c[3] = {c0, c1, 0};
  tensor_copy_emu(m, smem, c, true);
  if (last) mbar_complete_emu(bar);
}
inline void tensor_load_3d(void* smem, const TensorMap* m, int c0, int c1,
                           int c2, Mbar* bar, bool last = true) {
  const int c[3] = {c0, c1, c2};
  tensor_copy_emu(m, smem, c, true);
  if (last) mbar_complete_emu(bar);
}
inline void tensor_store_2d(const TensorMap* m, int c0, int c1,
                            const void* smem) {
  const int c[3] = {c0, c1, 0};
  tensor_copy_emu(m, const_cast<void*>(smem), c, false);
}
inline void tensor_store_3d(const TensorMap* m, int c0, int c1, int c2,
                            const void* smem) {
  const int c[3] = {c0, c1, c2};
  tensor_copy_emu(m, const_cast<void*>(smem), c, false);
}
}  // namespace bbt
#define BBT_TMAP_PARAM const bbt::TensorMap
#else
// --------------------------------------------------------------------- CUDA
#include <cuda.h>
#include <cuda_runtime.h>
namespace bbt {
typedef CUtensorMap TensorMap;
typedef unsigned long long Mbar;
#if defined(__CUDACC__)
// Named barriers: `threads` (a multiple of 32) threads of the CTA take part,
// arriving (not waiting) or synchronising (waiting) on barrier `id` (1-15;
// 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int threads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(Mbar* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)),
               "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(Mbar* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(b)),
               "r"(bytes)
               : "memory");
}
// Complete a phase without any bytes (an empty tile).
__device__ __forceinline__ void mbar_arrive(Mbar* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b))
               : "memory");
}
// k_phase is only used by the emulation.
__device__ __forceinline__ void mbar_wait(Mbar* b, uint32_t parity, uint32_t) {
  const unsigned bar = smem_u32(b);
  unsigned ok = 0;
  for (unsigned spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spins > (1u << 26)) __trap();  // a lost copy must not hang the GPU
  }
}
// Order generic-proxy writes to shared memory before a bulk store reads them.
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc,
                                          uint32_t bytes, Mbar* bar,
                                          bool = true) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src,
                                           uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::
                   "l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// All committed bulk stores have finished READING shared memory.
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all0() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* g, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tensor_prefetch_2d(const TensorMap* m, int c0,
                                                   int c1) {
  asm volatile(
      "cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(m),
      "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tensor_prefetch_3d(const TensorMap* m, int c0,
                                                   int c1, int c2) {
  asm volatile(
      "cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::
          "l"(m),
      "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tensor_load_2d(void* smem, const TensorMap* m,
                                               int c0, int c1, Mbar* bar,
                                               bool = true) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::"
      "complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem)),
      "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tensor_load_3d(void* smem, const TensorMap* m,
                                               int c0, int c1, int c2,
                                               Mbar* bar, bool = true) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::"
      "complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(smem)),
      "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tensor_store_2d(const TensorMap* m, int c0,
                                                int c1, const void* smem) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group "
      "[%0, {%1, %2}], [%3];" ::"l"(m),
      "r"(c0), "r"(c1), "r"(smem_u32(smem))
      : "memory");
}
__device__ __forceinline__ void tensor_store_3d(const TensorMap* m, int c0,
                                                int c1, int c2,
                                                const void* smem) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group "
      "[%0, {%1, %2, %3}], [%4];" ::"l"(m),
      "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(smem))
      : "memory");
}
#endif  // __CUDACC__
}  // namespace bbt
#define BBT_TMAP_PARAM const __grid_constant__ bbt::TensorMap
#endif

namespace bbt {
// Describe a tensor of `rank` (<= 3) dimensions of elem_bytes-sized elements
// (dims fastest first, strides in bytes for dimensions 1.., multiples of 16)
// and the box one copy moves.  Returns 0 on success.
int make_tensor_map(TensorMap* map, void* base, int elem_bytes, int rank,
                    const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);
}  // namespace bbt
