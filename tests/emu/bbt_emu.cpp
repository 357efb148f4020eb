// Host-thread emulation of the CUDA execution model, for index-logic tests on
// machines without a GPU.  TEST INFRASTRUCTURE ONLY: the product package never
// loads the library built from this file (tests/emu/build.sh).
// One OS thread plays one CUDA thread; blocks run one after another.
#include <algorithm>
#include <barrier>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define BBT_EMULATE 1
#include "../../baseband-tasks_b200/csrc/rt.cuh"

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
static thread_local std::barrier<>* t_barrier = nullptr;
static thread_local std::barrier<>* t_warp_barrier = nullptr;
static thread_local void* t_smem = nullptr;
static thread_local float* t_scratch = nullptr;

void* bbt_emu_smem() { return t_smem; }
float* bbt_emu_scratch() { return t_scratch; }
void bbt_emu_syncthreads() {
  if (t_barrier) t_barrier->arrive_and_wait();
}
void bbt_emu_syncwarp() {
  if (t_warp_barrier) t_warp_barrier->arrive_and_wait();
}

void bbt_emu_launch(dim3 grid, dim3 block, size_t smem,
                    const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  std::barrier<> bar(nthreads);
  // One barrier per warp (32 consecutive threads) for __syncwarp.
  std::vector<std::unique_ptr<std::barrier<>>> warp_bars;
  for (unsigned w = 0; w * 32 < nthreads; ++w)
    warp_bars.emplace_back(std::make_unique<std::barrier<>>(
        std::min(32u, nthreads - w * 32)));
  std::vector<char> shared(smem + 64);
  std::vector<float> scratch(nthreads + 1);
  std::vector<std::thread> pool;
  pool.reserve(nthreads);
  for (unsigned tid = 0; tid < nthreads; ++tid) {
    pool.emplace_back([&, tid]() {
      t_barrier = &bar;
      t_warp_barrier = warp_bars[tid / 32].get();
      t_smem = shared.data();
      t_scratch = scratch.data();
      blockDim = block;
      gridDim = grid;
      threadIdx = dim3(tid % block.x, (tid / block.x) % block.y,
                       tid / (block.x * block.y));
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx = dim3(bx, by, bz);
            body();
            bar.arrive_and_wait();
          }
      t_barrier = nullptr;
      t_warp_barrier = nullptr;
    });
  }
  for (auto& t : pool) t.join();
}
