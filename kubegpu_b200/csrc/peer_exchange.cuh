// peer_exchange.cuh -- the multi-GPU key exchange over peer memory (NVLink / NVSwitch), for the
// one-process-per-GPU launch.  Replaces "all-gather every rank's keys, then K2" (two launches through NCCL
// and a reduction kernel) by ONE kernel per rank:
//   * every rank owns a result array in memory its peers have mapped (CUDA IPC);
//   * push: thread p sends this rank's best key of pod p into EVERY rank's result array with a 64-bit
//     system-scope atomic min (P * G atomics of 8 bytes: 640 KB at P = 10k, G = 8; NO_FIT is not sent);
//   * sync: the last block to finish (atomic ticket) publishes "rank r has pushed epoch e" into every
//     rank's flag array and spins until all G ranks have published e.  When the kernel ends, this rank's
//     result array holds the global per-pod minimum, the same on every rank.
// Result arrays are double buffered by epoch parity; the OTHER buffer is cleaned by this kernel before the rank's
// arrival flag is set, so it is clean before any peer can reach the next epoch (they pass this epoch's barrier
// only after this rank arrived at it).
// Measured (round 2, 2 x B200, C2): keys bit-identical to one GPU and to the NCCL path (tests/test_gpu_multi.py).
#pragma once
#include <cstdint>

namespace kgpu {

constexpr int PEER_MAX_WORLD = 16;

struct PeerTable {
    unsigned long long *results[PEER_MAX_WORLD];   // every rank's result array of THIS epoch's parity (own entry included)
    uint32_t *flags[PEER_MAX_WORLD];               // every rank's flag array [world]
};

// The step is TWO launches per rank (K1 + this kernel): the kernel also does the housekeeping the host used to enqueue
// as memsets -- it resets the rank's local key array behind itself (K1 of the next step accumulates into it with atomic
// min) and cleans the OTHER result buffer (entries [0, clean_len): what the step before last left there) BEFORE the rank
// announces its arrival, so the buffer is clean before any peer can pass this epoch's barrier and push the next epoch.
__global__ void __launch_bounds__(256)
push_and_sync(unsigned long long *__restrict__ local_keys, int64_t P, PeerTable peers, int rank, int world,
              uint32_t epoch, unsigned int *ticket, unsigned long long *__restrict__ other_buffer, int64_t clean_len) {
    __shared__ bool last;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < clean_len; i += stride) other_buffer[i] = ~0ull;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) {
        const unsigned long long k = local_keys[p];
        local_keys[p] = ~0ull;                     // ready for the next step's K1
        if (k != ~0ull) {
#pragma unroll 1
            for (int g = 0; g < world; g++) atomicMin_system(peers.results[g] + p, k);
        }
    }
    __threadfence_system();                        // this thread's pushes (and its cleaning) before the ticket
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    // last block of this rank: every push of the rank is ordered before what follows
    if (threadIdx.x == 0) *ticket = 0;             // for the next launch (stream order)
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        volatile uint32_t *theirs = peers.flags[threadIdx.x];
        theirs[rank] = epoch;                      // "rank has pushed epoch" into peer threadIdx.x
        volatile uint32_t *mine = peers.flags[rank];
        while ((int32_t)(mine[threadIdx.x] - epoch) < 0) {}   // wait for peer threadIdx.x (wrap-safe compare)
    }
    __syncthreads();
    __threadfence_system();
}

}  // namespace kgpu
