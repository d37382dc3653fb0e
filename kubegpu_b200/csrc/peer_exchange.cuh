// peer_exchange.cuh -- the multi-GPU key exchange over peer memory (NVLink / NVSwitch), for the
// one-process-per-GPU launch.  Replaces "NCCL all-gather of every rank's keys, then K2" by ONE kernel per rank:
//   * every rank owns slot arrays slots[2][G][max_pods] in memory its peers have mapped (CUDA IPC);
//   * store: thread p writes this rank's best key of pod p into slot [rank] of EVERY rank's array with a plain
//     8-byte store (coalesced, posted writes over NVLink: P * G * 8 bytes, 640 KB at P = 10k, G = 8) and resets the
//     rank's local key array behind itself (K1 of the next step accumulates into it with atomic min);
//   * sync: per block one system-scope fence (thread 0, after the block barrier: cumulative) and a ticket; the last
//     block publishes "rank r has stored epoch e" into every rank's flag array; EVERY block then waits until all G
//     flags of its own rank carry e;
//   * min: thread p takes the minimum over the G slots of pod p (local reads) -> final[p].
// Slot arrays are double buffered by epoch parity: a peer can run at most one epoch ahead (it passes this epoch's
// barrier only after this rank arrived at it, i.e. after this rank finished READING the previous epoch), and then it
// writes the other buffer.  No atomics on the data path, no cleaning: every slot is overwritten every epoch.
// Round-2 history: the first version pushed with 64-bit system-scope atomic min into one result array per rank and
// fenced per thread (correct, tests/test_gpu_multi.py; measured 18 us at 2 GPUs, 33 us at 8 -- no better than NCCL).
// P == 0 is a pure barrier (bench.py aligns the ranks with it before each timed step).
// A wait that exceeds ~10 s (a peer died, ranks out of step) raises *error instead of hanging the GPU.
#pragma once
#include <cstdint>

namespace kgpu {

constexpr int PEER_MAX_WORLD = 16;

struct PeerTable {
    unsigned long long *slots[PEER_MAX_WORLD];     // every rank's slot array of THIS epoch's parity: [G][max_pods]
    uint32_t *flags[PEER_MAX_WORLD];               // every rank's flag array [world]
};

__global__ void __launch_bounds__(256)
gather_and_min(unsigned long long *__restrict__ local_keys, int64_t P, int64_t max_pods, PeerTable peers, int rank, int world,
               uint32_t epoch, unsigned int *ticket, unsigned long long *__restrict__ final_keys, int *error) {
    __shared__ bool last;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) {
        const unsigned long long k = local_keys[p];
        local_keys[p] = ~0ull;                     // ready for the next step's K1
#pragma unroll
        for (int g = 0; g < PEER_MAX_WORLD; g++)
            if (g < world) peers.slots[g][(int64_t)rank * max_pods + p] = k;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                    // the block's stores (cumulative through the barrier) before the ticket
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {                                    // every store of this rank is ordered before what follows
        if (threadIdx.x == 0) *ticket = 0;         // for the next launch (stream order)
        if ((int)threadIdx.x < world) {
            __threadfence_system();
            volatile uint32_t *theirs = peers.flags[threadIdx.x];
            theirs[rank] = epoch;                  // "rank has stored epoch" into peer threadIdx.x
        }
    }
    if ((int)threadIdx.x < world) {                // all blocks: wait for every peer's flag in OUR flag array
        volatile uint32_t *mine = peers.flags[rank];
        const long long t0 = clock64();
        while ((int32_t)(mine[threadIdx.x] - epoch) < 0) {      // wrap-safe compare
            if (clock64() - t0 > 20000000000LL) { *error = 1; break; }   // ~10 s at 2 GHz
        }
    }
    __syncthreads();
    __threadfence_system();                        // acquire side: the flags were read before the slots are
    if (p < P) {
        // all G loads in flight together (they bypass L1: the slots were written by the peers into this device's memory;
        // a loop of dependent volatile loads would pay G L2 round trips one after the other)
        const unsigned long long *mine = peers.slots[rank];
        unsigned long long v[PEER_MAX_WORLD];
#pragma unroll
        for (int g = 0; g < PEER_MAX_WORLD; g++) v[g] = g < world ? __ldcv(mine + (int64_t)g * max_pods + p) : ~0ull;
        unsigned long long best = v[0];
#pragma unroll
        for (int g = 1; g < PEER_MAX_WORLD; g++) best = min(best, v[g]);
        final_keys[p] = best;
    }
}

}  // namespace kgpu
