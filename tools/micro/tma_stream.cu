// Microbenchmark: what HBM bandwidth do the access patterns of the
// dedispersion passes reach when the tiles are moved by TMA (bulk copies
// through shared memory) instead of per-thread LDG/STG?
//   row   : contiguous 128 KB tiles (the planar row pass, C4)
//   col   : tiles of 1024 rows x 128 B, rows 128 KB apart (the column passes)
// Each pattern is copied in -> out by (a) a per-thread LDG/STG kernel of the
// same shape as the product kernels (512 threads, 32 x 8 B per thread) and
// (b) a persistent kernel, one CTA per SM, with an NSTAGE ring of tiles filled
// by cp.async.bulk[.tensor] and drained by bulk stores; optionally every
// thread also reads its 32 values from the tile into registers and writes
// them back (TOUCH), as a kernel that computes on the tile would.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../baseband-tasks_b200/csrc/tma.cuh"

using namespace bbt;

namespace bbt {
int make_tensor_map(TensorMap* map, void* base, int elem_bytes, int rank,
                    const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                         const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*,
                         CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault,
                                &q) != cudaSuccess || !p)
      return -1;
    fn = (Fn)p;
  }
  cuuint64_t d[3];
  cuuint64_t s[2];
  cuuint32_t b[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) d[i] = dims[i], b[i] = box[i];
  for (int i = 1; i < rank; ++i) s[i - 1] = strides_bytes[i];
  CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                           : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                             : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = fn(map, dt, rank, base, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r;
}
}  // namespace bbt

constexpr int kThreads = 512;
constexpr int kTile = 128 * 1024;       // bytes
constexpr int kRowBytes = 128;          // column pattern: bytes per row of tile
constexpr int kRows = kTile / kRowBytes;  // 1024

extern __shared__ float4 dyn_smem[];

// ------------------------------------------------------------- LDG / STG
// COL = false: tile t is the contiguous range [t*kTile, (t+1)*kTile).
// COL = true : tile t is rows r < 1024 at in + r*pitch + t*128.
template <bool COL>
__global__ void __launch_bounds__(kThreads, 1)
    ldg_copy(const float2* in, float2* out, long long pitch_elems, int tiles,
             int ahead) {
  const int tile = blockIdx.x;
  const int tid = threadIdx.x;
  float2 v[32];
  if (COL) {
    const int g = tid & 15, t = tid >> 4;  // 16 lanes of 8 B, 32 row groups
    const long long t0 = (long long)(tile % 2048) * 16 +
                         (long long)(tile / 2048) * kRows * pitch_elems;
    const float2* p = in + t0 + g + (long long)t * pitch_elems;
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = __ldcs(p + (long long)e * 32 * pitch_elems);
    if (tile + ahead < tiles) {
      const int nt = tile + ahead;
      const float2* q = in + (long long)(nt % 2048) * 16 +
                        (long long)(nt / 2048) * kRows * pitch_elems;
      for (int r = tid; r < kRows; r += kThreads)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + (long long)r * pitch_elems));
    }
    float2* o = out + t0 + g + (long long)t * pitch_elems;
#pragma unroll
    for (int e = 0; e < 32; ++e) o[(long long)e * 32 * pitch_elems] = v[e];
  } else {
    const float2* p = in + (long long)tile * (kTile / 8) + tid;
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = __ldcs(p + e * kThreads);
    if (tile + ahead < tiles) {
      const float2* q = in + (long long)(tile + ahead) * (kTile / 8);
      for (int i = tid * 16; i < kTile / 8; i += kThreads * 16)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + i));
    }
    float2* o = out + (long long)tile * (kTile / 8) + tid;
#pragma unroll
    for (int e = 0; e < 32; ++e) o[e * kThreads] = v[e];
  }
}

// ------------------------------------------------------------------- TMA
// Persistent: CTA b handles tiles b, b + grid, ...  Thread 0 keeps NSTAGE
// loads in flight; all threads wait for a stage, optionally pass the tile
// through registers, then thread 0 stores the stage and, once the store has
// read it, refills it.  A column tile is TILE / 128 rows of 128 bytes.
template <bool COL, int TILE, int NSTAGE, bool TOUCH>
__global__ void __launch_bounds__(kThreads, 1)
    tma_copy(const char* in, char* out, BBT_TMAP_PARAM map_in,
             BBT_TMAP_PARAM map_out, int tiles) {
  char* smem = reinterpret_cast<char*>(dyn_smem);
  Mbar* bars = reinterpret_cast<Mbar*>(smem + (size_t)NSTAGE * TILE);
  const int tid = threadIdx.x;
  constexpr int NBOX = TILE / (256 * 128);   // boxes of 256 rows
  constexpr int kColsPerSlab = 2048;
  if (tid == 0)
    for (int s = 0; s < NSTAGE; ++s) mbar_init(bars + s, 1);
  __syncthreads();
  auto issue = [&](int tile, int s) {
    mbar_expect_tx(bars + s, TILE);
    char* dst = smem + (size_t)s * TILE;
    if (COL) {
      const int col = tile % kColsPerSlab, rb = tile / kColsPerSlab;
      for (int q = 0; q < NBOX; ++q)
        tensor_load_2d(dst + q * (TILE / NBOX), &map_in, col * 16,
                       (rb * NBOX + q) * 256, bars + s);
    } else {
      for (int q = 0; q < NBOX; ++q)
        bulk_load(dst + q * (TILE / NBOX),
                  in + (size_t)tile * TILE + q * (TILE / NBOX), TILE / NBOX,
                  bars + s);
    }
  };
  const int first = blockIdx.x, step = gridDim.x;
  if (tid == 0)
    for (int s = 0; s < NSTAGE; ++s)
      if (first + s * step < tiles) issue(first + s * step, s);
  int k = 0;
  for (int tile = first; tile < tiles; tile += step, ++k) {
    const int s = k % NSTAGE;
    mbar_wait(bars + s, (k / NSTAGE) & 1, k / NSTAGE);
    float2* buf = reinterpret_cast<float2*>(smem + (size_t)s * TILE);
    if (TOUCH) {
      constexpr int E = TILE / 8 / kThreads;
      float2 v[E];
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = buf[tid + e * kThreads];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        v[e].x += 1.f;
        buf[tid + e * kThreads] = v[e];
      }
      fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0) {
      if (COL) {
        const int col = tile % kColsPerSlab, rb = tile / kColsPerSlab;
        for (int q = 0; q < NBOX; ++q)
          tensor_store_2d(&map_out, col * 16, (rb * NBOX + q) * 256,
                          smem + (size_t)s * TILE + q * (TILE / NBOX));
      } else {
        for (int q = 0; q < NBOX; ++q)
          bulk_store(out + (size_t)tile * TILE + q * (TILE / NBOX),
                     smem + (size_t)s * TILE + q * (TILE / NBOX), TILE / NBOX);
      }
      bulk_commit();
      const int next = tile + NSTAGE * step;
      if (next < tiles) {
        bulk_wait_read0();
        issue(next, s);
      }
    }
  }
  if (tid == 0) bulk_wait_all0();
}

static float time_ms(void (*launch)(void*), void* ctx, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch(ctx);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch(ctx);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

struct Ctx {
  const char* in;
  char* out;
  CUtensorMap map_in, map_out;
  size_t bytes;
  long long pitch_elems;
};

template <bool COL, int TILE, int NSTAGE, bool TOUCH>
void launch_tma(void* c) {
  Ctx* x = (Ctx*)c;
  const size_t smem = (size_t)NSTAGE * TILE + 64;
  cudaFuncSetAttribute(tma_copy<COL, TILE, NSTAGE, TOUCH>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tma_copy<COL, TILE, NSTAGE, TOUCH><<<148, kThreads, smem>>>(
      x->in, x->out, x->map_in, x->map_out, (int)(x->bytes / TILE));
}
template <bool COL>
void launch_ldg(void* c) {
  Ctx* x = (Ctx*)c;
  const int tiles = (int)(x->bytes / kTile);
  ldg_copy<COL><<<tiles, kThreads>>>((const float2*)x->in, (float2*)x->out,
                                     x->pitch_elems, tiles, 148);
}

#define RUN(name, fn)                                                        \
  do {                                                                       \
    float ms = time_ms(fn, &c, 5);                                           \
    printf("%-40s %.3f ms  %.0f GB/s (%s)\n", name, ms, gb / ms * 1e3,       \
           cudaGetErrorString(cudaGetLastError()));                          \
  } while (0)

int main() {
  // One matrix of 8192 rows x 256 KB, in and out (2 GiB each).  The row
  // pattern walks it linearly; the column pattern takes tiles of R rows x
  // 128 B (rows 256 KB apart).
  const size_t bytes = 2ull << 30;
  const long long pitch = 256 * 1024;            // bytes per matrix row
  const long long rows = bytes / pitch;          // 8192 rows
  char *in, *out;
  cudaMalloc(&in, bytes);
  cudaMalloc(&out, bytes);
  cudaMemset(in, 1, bytes);
  cudaMemset(out, 0, bytes);
  Ctx c;
  c.in = in;
  c.out = out;
  c.bytes = bytes;
  c.pitch_elems = pitch / 8;
  uint64_t dims[2] = {(uint64_t)(pitch / 8), (uint64_t)rows};
  uint64_t strides[2] = {8, (uint64_t)pitch};
  uint32_t box[2] = {16, 256};
  int rc = make_tensor_map(&c.map_in, in, 8, 2, dims, strides, box);
  rc |= make_tensor_map(&c.map_out, out, 8, 2, dims, strides, box);
  if (rc) {
    printf("tensor map creation failed: %d\n", rc);
    return 1;
  }
  const double gb = 2.0 * bytes / 1e9;
  constexpr int K = 1024;
  RUN("row LDG/STG 128K tiles", launch_ldg<false>);
  RUN("row TMA 128K x1", (launch_tma<false, 128 * K, 1, false>));
  RUN("row TMA 128K x1 touch", (launch_tma<false, 128 * K, 1, true>));
  RUN("row TMA 64K x3", (launch_tma<false, 64 * K, 3, false>));
  RUN("row TMA 64K x3 touch", (launch_tma<false, 64 * K, 3, true>));
  RUN("row TMA 32K x6", (launch_tma<false, 32 * K, 6, false>));
  RUN("row TMA 32K x6 touch", (launch_tma<false, 32 * K, 6, true>));
  RUN("col LDG/STG 1024x128B tiles", launch_ldg<true>);
  RUN("col TMA 1024x128B x1", (launch_tma<true, 128 * K, 1, false>));
  RUN("col TMA 1024x128B x1 touch", (launch_tma<true, 128 * K, 1, true>));
  RUN("col TMA 512x128B x3", (launch_tma<true, 64 * K, 3, false>));
  RUN("col TMA 512x128B x3 touch", (launch_tma<true, 64 * K, 3, true>));
  RUN("col TMA 256x128B x6", (launch_tma<true, 32 * K, 6, false>));
  RUN("col TMA 256x128B x6 touch", (launch_tma<true, 32 * K, 6, true>));
  return 0;
}
