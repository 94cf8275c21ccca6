// Thread-per-CUDA-thread shim: runs a __global__ kernel written against the CUDA execution model on the CPU, one OS
// thread per CUDA thread of a block (blocks run one after the other), __syncthreads / __syncwarp mapped to std::barrier,
// dynamic shared memory to a per-block buffer.  Test infrastructure for kernels that could not be run on a GPU when
// they were written (tests/test_glj_multi_host.py); it checks indexing, barrier placement and arithmetic — not timing,
// and not memory-model subtleties that only real warps expose.
#pragma once
#include <barrier>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

namespace vbshim {
struct Dim3 {
  unsigned x = 1, y = 1, z = 1;
};
struct BlockState {
  std::barrier<> all;
  std::vector<std::unique_ptr<std::barrier<>>> warps;
  std::vector<unsigned char> smem;
  explicit BlockState(int nthreads, size_t smem_bytes) : all(nthreads), smem(smem_bytes + 64) {
    for (int w = 0; w * 32 < nthreads; ++w) {
      const int n = nthreads - w * 32 < 32 ? nthreads - w * 32 : 32;
      warps.emplace_back(new std::barrier<>(n));
    }
  }
};
inline thread_local Dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local BlockState* t_block = nullptr;
inline unsigned char* dynamic_smem() {
  unsigned char* p = t_block->smem.data();
  return p + ((64 - (reinterpret_cast<uintptr_t>(p) & 63)) & 63);
}

template <class Kernel, class Args>
void launch(Kernel kern, Dim3 grid, int nthreads, size_t smem_bytes, const Args& args) {
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        BlockState block(nthreads, smem_bytes);
        std::vector<std::thread> threads;
        for (int t = 0; t < nthreads; ++t)
          threads.emplace_back([&, t] {
            t_threadIdx = Dim3{static_cast<unsigned>(t), 0, 0};
            t_blockIdx = Dim3{bx, by, bz};
            t_blockDim = Dim3{static_cast<unsigned>(nthreads), 1, 1};
            t_gridDim = grid;
            t_block = &block;
            kern(args);
            block.warps[t / 32]->arrive_and_drop();   // a thread that has left the kernel no longer takes part in barriers
            block.all.arrive_and_drop();
          });
        for (auto& th : threads) th.join();
      }
}
}  // namespace vbshim

#define threadIdx vbshim::t_threadIdx
#define blockIdx vbshim::t_blockIdx
#define blockDim vbshim::t_blockDim
#define gridDim vbshim::t_gridDim
inline void __syncthreads() { vbshim::t_block->all.arrive_and_wait(); }
inline void __syncwarp() { vbshim::t_block->warps[vbshim::t_threadIdx.x / 32]->arrive_and_wait(); }
