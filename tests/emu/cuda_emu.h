// cuda_emu.h -- a tiny SIMT emulator for spicedb-kubeapi-proxy_b200/csrc/kernels.cuh. TEST INFRASTRUCTURE ONLY:
// it is never linked into libzgpu.so, and the product has no CPU evaluation path.
//
// Why: the build container has no GPU. The check kernel is warp-synchronous code (ballots, shuffles, warp
// reductions around a shared-memory stack), so its LOGIC can be executed faithfully on the CPU: every CUDA thread
// of a block is an OS thread, every warp collective is a rendezvous of the warp's 32 threads. That lets the CPU
// test-suite run the very kernel source the GPU runs -- against the oracle -- before a GPU is available. It says
// nothing about performance or about hardware memory ordering; the -m gpu parity tests remain the gate.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct uint4 {
  uint32_t x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
struct dim3 {
  unsigned x = 1, y = 1, z = 1;
};

namespace zg_emu {

struct Barrier {  // sense-reversing barrier for `n` threads
  std::mutex m;
  std::condition_variable cv;
  unsigned n = 0, waiting = 0, gen = 0;
  void init(unsigned k) { n = k; waiting = 0; gen = 0; }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned g = gen;
    if (++waiting == n) {
      waiting = 0;
      ++gen;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return gen != g; });
    }
  }
};

struct Warp {
  Barrier bar;
  unsigned long long slot[32];
};

struct Block {
  std::vector<Warp> warps;
  Barrier bar;
  std::vector<uint8_t> smem;
  uint32_t static_u32[64];  // the one static __shared__ array a kernel may declare (ZG_BLOCK_SHARED_U32)
};

struct ThreadState {
  dim3 threadIdx, blockIdx, blockDim, gridDim;
  Block* block = nullptr;
  Warp* warp = nullptr;
  unsigned lane = 0;
};
inline ThreadState& ts() {
  static thread_local ThreadState t;
  return t;
}

// every lane publishes v, then reads what it needs: the building block of all collectives
template <class T>
inline void publish(T v) {
  static_assert(sizeof(T) <= 8, "collective operand too wide");
  ThreadState& t = ts();
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  t.warp->slot[t.lane] = raw;
  t.warp->bar.wait();
}
template <class T>
inline T read_lane(unsigned src) {
  unsigned long long raw = ts().warp->slot[src & 31];
  T v;
  std::memcpy(&v, &raw, sizeof(T));
  return v;
}
inline void done() { ts().warp->bar.wait(); }  // nobody overwrites a slot before everybody has read

// runs kernel(params) over a grid of `grid` blocks of `threads` threads with `smem_bytes` of dynamic shared memory;
// blocks run one after the other (a persistent kernel does not care), the threads of a block concurrently
template <class K, class P>
void launch(K kernel, const P& params, unsigned grid, unsigned threads, size_t smem_bytes) {
  for (unsigned b = 0; b < grid; ++b) {
    Block blk;
    blk.warps = std::vector<Warp>((threads + 31) / 32);
    for (size_t w = 0; w < blk.warps.size(); ++w) blk.warps[w].bar.init(std::min(32u, threads - unsigned(w) * 32));
    blk.bar.init(threads);
    blk.smem.assign(smem_bytes + 64, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < threads; ++t)
      th.emplace_back([&, t] {
        ThreadState& s = ts();
        s.threadIdx.x = t;
        s.blockIdx.x = b;
        s.blockDim.x = threads;
        s.gridDim.x = grid;
        s.block = &blk;
        s.warp = &blk.warps[t / 32];
        s.lane = t % 32;
        kernel(params);
      });
    for (auto& x : th) x.join();
  }
}

}  // namespace zg_emu

#define threadIdx (zg_emu::ts().threadIdx)
#define blockIdx (zg_emu::ts().blockIdx)
#define blockDim (zg_emu::ts().blockDim)
#define gridDim (zg_emu::ts().gridDim)
// `extern __shared__ __align__(16) uint8_t smem[];` in kernels.cuh becomes a pointer to the block's buffer
#define ZG_DYNAMIC_SMEM(name) uint8_t* name = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(zg_emu::ts().block->smem.data()) + 15) & ~uintptr_t(15))

#define ZG_BLOCK_SHARED_U32(name, n) uint32_t* name = zg_emu::ts().block->static_u32

template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __ldcg(const T* p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
// position of the n-th (1-based) set bit of mask at or above `base`; 0xFFFFFFFF if there is none (offset > 0 only)
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
  for (unsigned b = base; b < 32; ++b)
    if ((mask >> b) & 1u)
      if (--offset == 0) return b;
  return 0xFFFFFFFFu;
}
static inline int __ffs(unsigned v) { return __builtin_ffs(static_cast<int>(v)); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __syncthreads() { zg_emu::ts().block->bar.wait(); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { zg_emu::ts().warp->bar.wait(); }

template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
  zg_emu::publish(v);
  T r = zg_emu::read_lane<T>(static_cast<unsigned>(src));
  zg_emu::done();
  return r;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  zg_emu::publish(v);
  const unsigned lane = zg_emu::ts().lane;
  T r = lane >= delta ? zg_emu::read_lane<T>(lane - delta) : v;
  zg_emu::done();
  return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m) {
  zg_emu::publish(v);
  T r = zg_emu::read_lane<T>(zg_emu::ts().lane ^ static_cast<unsigned>(m));
  zg_emu::done();
  return r;
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
  zg_emu::publish<unsigned>(pred ? 1u : 0u);
  unsigned m = 0;
  for (unsigned i = 0; i < 32; ++i) m |= (zg_emu::read_lane<unsigned>(i) & 1u) << i;
  zg_emu::done();
  return m;
}
static inline unsigned __match_any_sync(unsigned, unsigned v) {
  zg_emu::publish(v);
  unsigned m = 0;
  for (unsigned i = 0; i < 32; ++i) m |= (zg_emu::read_lane<unsigned>(i) == v ? 1u : 0u) << i;
  zg_emu::done();
  return m;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __reduce_max_sync(unsigned, int v) {
  zg_emu::publish(v);
  int m = v;
  for (unsigned i = 0; i < 32; ++i) m = std::max(m, zg_emu::read_lane<int>(i));
  zg_emu::done();
  return m;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  zg_emu::publish(v);
  unsigned m = 0;
  for (unsigned i = 0; i < 32; ++i) m |= zg_emu::read_lane<unsigned>(i);
  zg_emu::done();
  return m;
}

static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicCAS(unsigned* p, unsigned expected, unsigned desired) {
  __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return expected;
}
static inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST);
}
using std::max;
using std::min;
