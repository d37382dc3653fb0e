// cuda_emu.h -- just enough of the CUDA device programming model to compile kubegpu_b200/csrc/*.cuh with
// g++ and run the kernels on the CPU: one OS thread per CUDA thread of a block, real barriers for
// __syncthreads(), per-warp rendezvous buffers for the warp collectives.  TEST INFRASTRUCTURE ONLY
// (tests/test_kernels_emulated.py): it lets kernel edits be checked against the oracle without a GPU; the
// bit-exactness claims of this repository rest on the B200 runs, not on this.
//
// Limits: blocks run one after the other; every thread of a block must reach the same __syncthreads()
// and every lane of a warp the same warp collective (true for the kernels here: collectives sit in
// warp-uniform control flow); a thread may return early only after its last barrier.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __constant__
#define __launch_bounds__(...)

struct int4 {
    int x, y, z, w;
};
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
struct alignas(8) uint2 {
    unsigned x, y;
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct alignas(16) uint4 {
    unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

namespace emu {

struct Warp {
    std::barrier<> bar{32};
    uint64_t buf[32];
};

struct Block {
    explicit Block(unsigned nthreads) : bar((std::ptrdiff_t)nthreads), warps((nthreads + 31) / 32) {
        for (auto &w : warps) w = std::make_unique<Warp>();
    }
    std::barrier<> bar;
    std::vector<std::unique_ptr<Warp>> warps;
    std::atomic<int> or_flag{0};
};

struct Ctx {
    dim3 tid, bid, bdim, gdim;
    Block *block = nullptr;
    int lane = 0, warp = 0;
};
inline thread_local Ctx ctx;

// Run `kernel` for every block of `grid` with `block.x` threads (block.x must be a multiple of 32).
inline void launch(dim3 grid, dim3 block, const std::function<void()> &kernel) {
    Block blk(block.x);
    std::vector<std::thread> team;
    for (unsigned t = 0; t < block.x; t++)
        team.emplace_back([&, t] {
            ctx.block = &blk;
            ctx.tid = dim3(t);
            ctx.bdim = block;
            ctx.gdim = grid;
            ctx.lane = (int)(t & 31);
            ctx.warp = (int)(t >> 5);
            for (unsigned by = 0; by < grid.y; by++)
                for (unsigned bx = 0; bx < grid.x; bx++) {
                    ctx.bid = dim3(bx, by);
                    blk.bar.arrive_and_wait();
                    kernel();
                    blk.bar.arrive_and_wait();
                }
        });
    for (auto &th : team) th.join();
}

inline Warp &my_warp() { return *ctx.block->warps[(size_t)ctx.warp]; }

template <class T, class Op>
inline T warp_all(T v, Op op) {        // every lane contributes v, every lane gets op over the 32 values
    Warp &w = my_warp();
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.buf[ctx.lane] = bits;
    w.bar.arrive_and_wait();
    T acc;
    std::memcpy(&acc, &w.buf[0], sizeof(T));
    for (int l = 1; l < 32; l++) {
        T x;
        std::memcpy(&x, &w.buf[l], sizeof(T));
        acc = op(acc, x);
    }
    w.bar.arrive_and_wait();
    return acc;
}

template <class T>
inline T warp_read(T v, int src_lane) {  // every lane contributes v and reads lane src_lane's value
    Warp &w = my_warp();
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    w.buf[ctx.lane] = bits;
    w.bar.arrive_and_wait();
    T out;
    std::memcpy(&out, &w.buf[src_lane & 31], sizeof(T));
    w.bar.arrive_and_wait();
    return out;
}

}  // namespace emu

#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::ctx.bdim)
#define gridDim (emu::ctx.gdim)

// ---- synchronisation and warp collectives ---------------------------------------------------------
inline void __syncthreads() { emu::ctx.block->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { emu::my_warp().bar.arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
    emu::Block &b = *emu::ctx.block;
    b.bar.arrive_and_wait();
    if (emu::ctx.tid.x == 0) b.or_flag.store(0);
    b.bar.arrive_and_wait();
    if (pred) b.or_flag.store(1);
    b.bar.arrive_and_wait();
    return b.or_flag.load();
}
inline int __syncthreads_and(int pred) { return !__syncthreads_or(!pred); }
inline unsigned __reduce_min_sync(unsigned, unsigned v) { return emu::warp_all(v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
inline unsigned __reduce_max_sync(unsigned, unsigned v) { return emu::warp_all(v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
inline unsigned __ballot_sync(unsigned, int pred) {
    return emu::warp_all((unsigned)(pred ? 1u << emu::ctx.lane : 0u), [](unsigned a, unsigned b) { return a | b; });
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::warp_read(v, src); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int = 32) { return emu::warp_read(v, emu::ctx.lane ^ lane_mask); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, int delta, int = 32) {      // lanes below delta keep their own value
    return emu::warp_read(v, emu::ctx.lane >= delta ? emu::ctx.lane - delta : emu::ctx.lane);
}

// ---- memory ------------------------------------------------------------------------------------------
template <class T>
inline T __ldg(const T *p) { return *p; }
template <typename T>
inline T __ldcv(const T *p) { return *reinterpret_cast<const volatile T *>(p); }
template <class T>
inline T __ldcg(const T *p) { return *p; }
template <class T>
inline unsigned long long __cvta_generic_to_shared(T *p) { return (unsigned long long)(uintptr_t)p; }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicMin(unsigned *p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}

inline unsigned long long atomicMin_system(unsigned long long *p, unsigned long long v) { return atomicMin(p, v); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- integer intrinsics ----------------------------------------------------------------------------------
inline long long clock64() { return 0; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return x == 0 ? 0 : __builtin_ctz(x) + 1; }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline unsigned __vimin3_u32(unsigned a, unsigned b, unsigned c) { unsigned m = a < b ? a : b; return m < c ? m : c; }
inline unsigned __vimax3_u32(unsigned a, unsigned b, unsigned c) { unsigned m = a > b ? a : b; return m > c ? m : c; }
inline unsigned __viaddmin_u32(unsigned a, unsigned b, unsigned c) { unsigned s = a + b; return s < c ? s : c; }
inline unsigned __viaddmax_u32(unsigned a, unsigned b, unsigned c) { unsigned s = a + b; return s > c ? s : c; }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t both = ((uint64_t)y << 32) | x;
    unsigned out = 0;
    for (int i = 0; i < 4; i++) out |= (unsigned)((both >> (8 * ((s >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return out;
}

// CUDA's min/max overload set as used by the kernels
template <class T>
inline T min(T a, T b) { return a < b ? a : b; }
template <class T>
inline T max(T a, T b) { return a > b ? a : b; }
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline long long min(long long a, int b) { return a < b ? a : (long long)b; }
inline long long min(int a, long long b) { return a < b ? (long long)a : b; }
inline long min(long a, int b) { return a < b ? a : (long)b; }
