// place_sequential.cuh -- K3: stateful sequential placement.
//
// Semantics (bit-exact against the CPU twin used by the tests):
//   for p = 0 .. P-1, in order:
//       key_p = min over nodes of (cost<<40 | node_id<<8 | S)   given the CURRENT free masks
//       if key_p != NO_FIT:  free_mask[node] &= ~S              (the pod takes those GPUs)
// This is what TakePodResources would make of a scheduling cycle if the reference's plugin
// tracked usage (it is a no-op there: gpuschedulerplugin/gpu_scheduler.go:57-63), and it has no
// snapshot-scoring collapse: every placement changes the state the next pod sees.
//
// Device data:  nodebest[9][Npad] uint32  (cost<<8 | S) of every node for k = 0..8, INF32 = no fit
//               tilebest[9][T]    uint64  min over each 128-node tile of (cost<<40|node_id<<8|S)
// place_init       : one block per tile: full enumeration for all 9 k (lane per node) -> both tables
// place_sequential : ONE persistent block of 1024 threads walks the pods: argmin over tilebest[k]
//                    (coalesced 8-byte loads + block reduction), commit, re-enumerate the winner
//                    node for all 9 k (warp per k, lane per subset), refresh its tile's 9 minima.
#pragma once
#include "score_pairs.cuh"

namespace kgpu {

constexpr int PLACE_TILE = 128;
#ifndef KGPU_PLACE_THREADS
#define KGPU_PLACE_THREADS 512      // 16 warps: cheaper barriers than 1024, still >= 9 warps (one per k)
#endif
constexpr int PLACE_THREADS = KGPU_PLACE_THREADS;

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(0xFFFFFFFFu, v, off));
    return v;
}

__device__ __forceinline__ unsigned long long wide_key(uint32_t nk, unsigned long long node_id) {
    return nk == INF32 ? ~0ull : (((unsigned long long)(nk >> 8) << 40) | (node_id << 8) | (nk & 0xFFu));
}

__global__ void __launch_bounds__(PLACE_TILE)
place_init(const int4 *__restrict__ topo4, const int32_t *__restrict__ free_mask, int64_t N, int64_t Npad,
           int64_t node_id_base, Weights W, PipeConsts pc, uint32_t *__restrict__ nodebest,
           unsigned long long *__restrict__ tilebest, int64_t T) {
    __shared__ int32_t sW[16];
    __shared__ unsigned long long sRed[9][PLACE_TILE / 32];
    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 16; i++) sW[i] = W.w[i];
    }
    __syncthreads();
    const int64_t tile = blockIdx.x;
    const int64_t node = tile * PLACE_TILE + tid;
    const bool valid = node < N;
    PairCosts C;
    uint32_t free;
    stage_node(topo4, free_mask, node, valid, sW, C, free);
#pragma unroll 1
    for (int k = 0; k <= 8; k++) {
        uint32_t key = node_key(k, C, pc, free, valid);
        if (key >= PEN) key = INF32;
        nodebest[(int64_t)k * Npad + node] = key;
        const unsigned long long w = warp_min_u64(wide_key(key, (unsigned long long)(node_id_base + node)));
        if ((tid & 31) == 0) sRed[k][tid >> 5] = w;
    }
    __syncthreads();
    if (tid < 9) {
        unsigned long long b = ~0ull;
#pragma unroll
        for (int w = 0; w < PLACE_TILE / 32; w++) b = min(b, sRed[tid][w]);
        tilebest[(int64_t)tid * T + tile] = b;
    }
}

// (cost<<8 | S) of one node for k GPUs, computed by a whole warp: lane per candidate subset.
// sCost = the node's weight-mapped 8x8 matrix in shared memory.
__device__ __forceinline__ uint32_t node_key_warp(int k, const int32_t *sCost, uint32_t fm, int lane) {
    if (k == 0) return 0u;
    const int nsub = c_nsub[k];
    uint32_t key = INF32;
    for (int s = lane; s < nsub; s += 32) {
        const uint32_t S = c_subsets[k][s];
        if (S & ~fm) continue;
        uint32_t cost = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = i + 1; j < 8; j++)
                if ((S & ((1u << i) | (1u << j))) == ((1u << i) | (1u << j))) cost += (uint32_t)sCost[i * 8 + j];
        key = min(key, (cost << 8) | S);
    }
    return __reduce_min_sync(0xFFFFFFFFu, key);
}

__global__ void __launch_bounds__(PLACE_THREADS, 1)
place_sequential(const int32_t *__restrict__ topo, int32_t *__restrict__ free_mask, int64_t N, int64_t Npad,
                 int64_t node_id_base, const int4 *__restrict__ pods4, int64_t P, Weights W,
                 uint32_t *__restrict__ nodebest, unsigned long long *__restrict__ tilebest, int64_t T,
                 unsigned long long *__restrict__ keys) {
    __shared__ int32_t sW[16];
    __shared__ int32_t sCost[64];
    __shared__ unsigned long long sRed[PLACE_THREADS / 32];
    __shared__ unsigned long long sWin;
    __shared__ uint32_t sFree;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 16; i++) sW[i] = W.w[i];
    }
    __syncthreads();

    for (int64_t p = 0; p < P; p++) {
        const int k = __ldg(pods4 + p).x;                 // block-uniform
        if (k < 0 || k > 8) {
            if (tid == 0) keys[p] = ~0ull;
            continue;
        }
        // 1. best tile for this k
        const unsigned long long *tb = tilebest + (int64_t)k * T;
        unsigned long long best = ~0ull;
        for (int64_t t = tid; t < T; t += PLACE_THREADS) best = min(best, tb[t]);
        best = warp_min_u64(best);
        if (lane == 0) sRed[warp] = best;
        __syncthreads();
        if (warp == 0) {
            unsigned long long b = warp_min_u64(lane < PLACE_THREADS / 32 ? sRed[lane] : ~0ull);
            if (lane == 0) sWin = b;
        }
        __syncthreads();
        const unsigned long long win = sWin;
        if (tid == 0) keys[p] = win;
        if (win == ~0ull || k == 0) {                      // nothing fits / nothing to take
            __syncthreads();
            continue;
        }
        // 2. commit: the pod takes GPUs S of that node
        const int64_t node = (int64_t)((win >> 8) & 0xFFFFFFFFull) - node_id_base;
        const uint32_t S = (uint32_t)(win & 0xFFull);
        if (tid < 64) sCost[tid] = sW[topo[node * 64 + tid] & 15];
        if (tid == 64) {
            const uint32_t fm = ((uint32_t)free_mask[node] & 0xFFu) & ~S;
            free_mask[node] = (int32_t)fm;
            sFree = fm;
        }
        __syncthreads();
        // 3+4. warp w handles k = w: re-enumerate the node (lane per subset), then refresh the minimum of the
        // node's tile for that k with the fresh value (the other 127 entries are unchanged in memory).
        if (warp <= 8) {
            const uint32_t nk = node_key_warp(warp, sCost, sFree, lane);
            if (lane == 0) nodebest[(int64_t)warp * Npad + node] = nk;
            const int64_t tile = node / PLACE_TILE;
            unsigned long long b = ~0ull;
#pragma unroll
            for (int j = 0; j < PLACE_TILE / 32; j++) {
                const int64_t n = tile * PLACE_TILE + lane + 32 * j;
                const uint32_t v = n == node ? nk : nodebest[(int64_t)warp * Npad + n];
                b = min(b, wide_key(v, (unsigned long long)(node_id_base + n)));
            }
            b = warp_min_u64(b);
            if (lane == 0) tilebest[(int64_t)warp * T + tile] = b;
        }
        __syncthreads();
    }
}

}  // namespace kgpu
