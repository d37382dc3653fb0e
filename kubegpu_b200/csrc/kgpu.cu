// kgpu.cu -- C ABI of libkgpu (include/kgpu.h) over the sm_100a kernels in
// score_pairs.cuh.  No CPU fallback: every entry point needs a CUDA device.
#include "../../include/kgpu.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "multi_device.h"
#include "place_sequential.cuh"
#include "sparse_work.h"
#include "peer_exchange.cuh"
#include "score_pairs.cuh"
#include "score_pairs_sparse.cuh"

namespace {

thread_local std::string t_last_error;

// F2/F2N multipliers handed to the kernels as arguments so that ptxas cannot fold them
// (subset_dp_gen.cuh: pins those adds to the FMA pipe).
const kgpu::PipeConsts PC = {1u, 0xFFFFFFFFu};

const int32_t kDefaultWeights[16] = {64, 32, 16, 8, 4, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0};

}  // namespace

// One device's shard of the node array plus its scratch buffers.
struct kgpu_shard {
    int dev = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int32_t *d_topo = nullptr;       // [n][64]
    int32_t *d_free = nullptr;       // [n]
    int32_t *d_mem = nullptr;        // [n][8] MiB per GPU (0x7F7F7F7F = unconstrained until uploaded)
    int *d_flag = nullptr;           // "batch has memory-constrained pods"
    uint32_t *d_cpair = nullptr;     // K1s: [n][28] compacted scaled pair costs (see compact_nodes)
    uint32_t *d_perm = nullptr;      // K1s: [n] position -> GPU index
    bool compact_dirty = true;       // topology / free masks / weights changed since the cache was built
    int32_t *d_order = nullptr;      // K1s: slot -> node index, grouped by popcount(free), -1 = padding
    int64_t n_slots = 0, order_cap = 0;
    int64_t n = 0, cap = 0;
    int64_t node_id_base = 0;
    int32_t *d_pods = nullptr;       // [pcap][4]
    unsigned long long *d_keys = nullptr;    // [pcap]
    unsigned long long *d_gather = nullptr;  // [ndev][pcap] (multi-device only)
    unsigned long long *d_bestk = nullptr;   // [9] memo variant
    std::vector<uint8_t> tile_class;         // K1s: max free-GPU count per 128-slot tile of the order (host copy)
    std::vector<kgpu::SparseWorkItem> h_work;   // K1s work list for work_P pods (sparse_work.h)
    int4 *d_work = nullptr;
    int64_t work_cap = 0, work_P = -1;
    uint32_t *d_nodebest = nullptr;          // [views][9][Npad]  K3 tables
    unsigned long long *d_tilebest = nullptr;   // [views][9][T]
    int64_t place_cap = 0;
    int64_t pcap = 0;
};

struct kgpu_ctx {
    std::mutex mu;
    std::string err;
    std::vector<kgpu_shard> shards;
    int32_t W[16];
    int variant = KGPU_VARIANT_SPARSE;
    int64_t n_total = 0;
    int64_t launches = 0;
    double last_kernel_ms = 0.0;
    bool subsets_uploaded = false;
    kgpu::MultiDevice *multi = nullptr;   // NCCL communicator set, ndev > 1 only
    // peer-memory key exchange (kgpu_exchange_*): one allocation per rank, mapped by every peer:
    //   results[2][max_pods] uint64 | flags[PEER_MAX_WORLD] uint32 | ticket uint32 ;  local[] is private
    struct {
        int world = 0, rank = 0;
        int64_t max_pods = 0;
        void *base = nullptr;
        void *peer_base[kgpu::PEER_MAX_WORLD] = {};
        unsigned long long *local = nullptr;
        uint32_t epoch = 0;
        bool connected = false;
    } xch;
};

namespace {

int fail(kgpu_ctx *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_last_error = buf;
    if (h) h->err = buf;
    return code;
}

#define KGPU_CUDA(h, expr)                                                                          \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail((h), e__ == cudaErrorMemoryAllocation ? KGPU_ERR_NOMEM : KGPU_ERR_CUDA,      \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

int upload_subset_tables(kgpu_ctx *h) {
    uint8_t subs[9][70];
    uint8_t nsub[9];
    memset(subs, 0, sizeof subs);
    for (int k = 0; k <= 8; k++) {
        int c = 0;
        for (unsigned S = 0; S < 256; S++)
            if (__builtin_popcount(S) == k) subs[k][c++] = (uint8_t)S;
        nsub[k] = (uint8_t)c;
    }
    for (auto &s : h->shards) {
        KGPU_CUDA(h, cudaSetDevice(s.dev));
        KGPU_CUDA(h, cudaMemcpyToSymbol(kgpu::c_subsets, subs, sizeof subs));
        KGPU_CUDA(h, cudaMemcpyToSymbol(kgpu::c_nsub, nsub, sizeof nsub));
    }
    h->subsets_uploaded = true;
    return KGPU_OK;
}

int ensure_pod_capacity(kgpu_ctx *h, kgpu_shard &s, int64_t P) {
    if (P <= s.pcap) return KGPU_OK;
    int64_t cap = std::max<int64_t>(1024, P + P / 4);
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    if (s.d_pods) cudaFree(s.d_pods);
    if (s.d_keys) cudaFree(s.d_keys);
    if (s.d_gather) cudaFree(s.d_gather);
    s.d_pods = nullptr; s.d_keys = nullptr; s.d_gather = nullptr; s.pcap = 0;
    KGPU_CUDA(h, cudaMalloc(&s.d_pods, (size_t)cap * 16));
    KGPU_CUDA(h, cudaMalloc(&s.d_keys, (size_t)cap * 8));
    if (h->shards.size() > 1) KGPU_CUDA(h, cudaMalloc(&s.d_gather, (size_t)cap * 8 * h->shards.size()));
    s.pcap = cap;
    return KGPU_OK;
}

// Enqueue K1 for P pods on shard s: d_keys[p] = best placement over this shard's nodes.
// has_mem: 1 / 0 = the host knows whether some pod carries min_mem > 0; -1 = unknown (device buffers).
int launch_score(kgpu_ctx *h, kgpu_shard &s, const int32_t *d_pods, int64_t P, unsigned long long *d_keys,
                 cudaStream_t st, int has_mem) {
    if (P <= 0) return KGPU_OK;
    if ((reinterpret_cast<uintptr_t>(d_pods) & 15u) != 0)
        return fail(h, KGPU_ERR_INVALID, "pods pointer must be 16-byte aligned");
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    kgpu::Weights W;
    memcpy(W.w, h->W, sizeof W.w);
    const int4 *topo4 = reinterpret_cast<const int4 *>(s.d_topo);
    const int4 *pods4 = reinterpret_cast<const int4 *>(d_pods);

    const int4 *mem4 = reinterpret_cast<const int4 *>(s.d_mem);
    KGPU_CUDA(h, cudaMemsetAsync(d_keys, 0xFF, (size_t)P * 8, st));
    if (s.n == 0) return KGPU_OK;

    const bool wpp = h->variant == KGPU_VARIANT_WARP_PER_PAIR;
    const bool sparse = h->variant == KGPU_VARIANT_SPARSE;
    if (!wpp && has_mem != 0) {   // which pods go to K1m?  (flag read by its blocks; skipped when the host knows there are none)
        KGPU_CUDA(h, cudaMemsetAsync(s.d_flag, 0, sizeof(int), st));
        kgpu::any_mem_pod<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(pods4, P, s.d_flag);
        h->launches++;
    }
    if (h->variant == KGPU_VARIANT_MEMO_BY_K) {
        KGPU_CUDA(h, cudaMemsetAsync(s.d_bestk, 0xFF, 9 * 8, st));
        int blocks = (int)std::min<int64_t>((s.n + kgpu::LPN_THREADS - 1) / kgpu::LPN_THREADS, (int64_t)s.sm_count * 4);
        kgpu::memo_best_by_k<<<blocks, kgpu::LPN_THREADS, 0, st>>>(topo4, s.d_free, s.n, s.node_id_base, W, PC, s.d_bestk);
        kgpu::memo_gather<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(pods4, P, s.d_bestk, d_keys);
        h->launches += 2;
    }
    const int tile = wpp ? kgpu::WPP_TILE : kgpu::LPN_THREADS;
    const int64_t tiles = sparse ? s.n_slots / kgpu::SP_THREADS : (s.n + tile - 1) / tile;
    // Pod splits: enough blocks for ~8 waves of resident CTAs, but each block keeps
    // >= 128 pods so staging its node tile stays amortised.
    const int64_t resident = (int64_t)s.sm_count * (wpp ? 8 : sparse ? KGPU_SP_MINBLOCKS : KGPU_LPN_MINBLOCKS);
    static const int64_t waves = [] { const char *e = getenv("KGPU_WAVES"); int v = e ? atoi(e) : 0; return (int64_t)(v > 0 ? v : 8); }();
    int64_t splits = std::max<int64_t>(1, (waves * resident + tiles - 1) / tiles);
    splits = std::min<int64_t>(splits, std::max<int64_t>(1, P / 128));
    splits = std::min<int64_t>(splits, 65535);
    int64_t per = (P + splits - 1) / splits;
    per = (per + 31) / 32 * 32;
    splits = (P + per - 1) / per;
    if (tiles > 0x7FFFFFFFLL || per > 0x7FFFFFFFLL) return fail(h, KGPU_ERR_INVALID, "batch too large for one launch");
    dim3 grid((unsigned)tiles, (unsigned)splits);
    if (wpp) {
        kgpu::score_pairs_warp_per_pair<<<grid, kgpu::WPP_THREADS, 0, st>>>(topo4, s.d_free, s.d_mem, s.n, s.node_id_base, pods4, P,
                                                                            (int)per, W, PC, d_keys);
        h->launches++;
    } else {
        if (h->variant == KGPU_VARIANT_TILE_MEMO) {
            kgpu::score_pairs_lane_per_node<false, false><<<grid, kgpu::LPN_THREADS, 0, st>>>(
                topo4, s.d_free, mem4, s.d_flag, s.n, s.node_id_base, pods4, P, (int)per, W, PC, d_keys);
            h->launches++;
        } else if (h->variant == KGPU_VARIANT_LANE_PER_NODE) {
            kgpu::score_pairs_lane_per_node<true, false><<<grid, kgpu::LPN_THREADS, 0, st>>>(
                topo4, s.d_free, mem4, s.d_flag, s.n, s.node_id_base, pods4, P, (int)per, W, PC, d_keys);
            h->launches++;
        }
        if (sparse) {
            if (s.compact_dirty) {   // rebuild the compacted pair-cost cache (topology / masks / weights changed)
                kgpu::compact_nodes<<<(unsigned)((s.n + kgpu::SP_THREADS - 1) / kgpu::SP_THREADS), kgpu::SP_THREADS, 0, st>>>(
                    topo4, s.d_free, s.n, W, s.d_cpair, s.d_perm);
                h->launches++;
                s.compact_dirty = false;
            }
            const int4 *cpair4 = reinterpret_cast<const int4 *>(s.d_cpair);
            // Work list instead of the plain grid: auto = for small shards (few tiles per resident block), where
            // equal pod ranges are either too few or too short; KGPU_SP_WORKLIST=0/1 forces it off/on.
            static const int worklist_mode = [] { const char *e = getenv("KGPU_SP_WORKLIST"); return e ? atoi(e) : -1; }();
            const bool use_work = worklist_mode >= 0 ? worklist_mode != 0 : tiles * 4 < resident;
            const int4 *d_work = nullptr;
            if (use_work) {
                if (s.work_P != P) {
                    kgpu::build_sparse_work(s.tile_class, P, resident, s.h_work);
                    if ((int64_t)s.h_work.size() > s.work_cap) {
                        if (s.d_work) cudaFree(s.d_work);
                        s.d_work = nullptr; s.work_cap = 0;
                        KGPU_CUDA(h, cudaMalloc(&s.d_work, s.h_work.size() * sizeof(int4)));
                        s.work_cap = (int64_t)s.h_work.size();
                    }
                    static_assert(sizeof(kgpu::SparseWorkItem) == sizeof(int4), "work item layout");
                    KGPU_CUDA(h, cudaMemcpyAsync(s.d_work, s.h_work.data(), s.h_work.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
                    s.work_P = P;
                }
                d_work = s.d_work;
                grid = dim3((unsigned)s.h_work.size(), 1);
            }
            bool byte_keys = true;       // every cost < 2^16 <=> 28 * max weight < 65536
            for (int i = 0; i < 16; i++) byte_keys = byte_keys && h->W[i] <= 2340;
#define KGPU_LAUNCH_SPARSE(MEMF, BK)                                                                          \
    kgpu::score_pairs_sparse<true, MEMF, BK><<<grid, kgpu::SP_THREADS, 0, st>>>(                              \
        cpair4, s.d_perm, s.d_free, s.d_mem, s.d_order, s.d_flag, s.node_id_base, pods4, P, (int)per, d_work, PC, d_keys)
            if (byte_keys) KGPU_LAUNCH_SPARSE(false, true); else KGPU_LAUNCH_SPARSE(false, false);
            h->launches++;
            if (has_mem != 0) {
                if (byte_keys) KGPU_LAUNCH_SPARSE(true, true); else KGPU_LAUNCH_SPARSE(true, false);
                h->launches++;
            }
#undef KGPU_LAUNCH_SPARSE
        } else if (has_mem != 0) {   // K1m: the memory-constrained pods (its blocks exit at once if the flag is 0)
            kgpu::score_pairs_lane_per_node<true, true><<<grid, kgpu::LPN_THREADS, 0, st>>>(
                topo4, s.d_free, mem4, s.d_flag, s.n, s.node_id_base, pods4, P, (int)per, W, PC, d_keys);
            h->launches++;
        }
    }
    KGPU_CUDA(h, cudaGetLastError());
    return KGPU_OK;
}

int validate_topo(kgpu_ctx *h, const int32_t *topo, int64_t n) {
    for (int64_t i = 0; i < n * 64; i++)
        if ((uint32_t)topo[i] > 15u)
            return fail(h, KGPU_ERR_INVALID, "topo[%lld][%d] = %d outside the link-level domain 0..15",
                        (long long)(i / 64), (int)(i % 64), topo[i]);
    return KGPU_OK;
}

// Which shard holds local node index idx (contiguous split)?
kgpu_shard *shard_of(kgpu_ctx *h, int64_t idx, int64_t *local) {
    int64_t off = 0;
    for (auto &s : h->shards) {
        if (idx < off + s.n) { *local = idx - off; return &s; }
        off += s.n;
    }
    return nullptr;
}

void free_shard(kgpu_shard &s) {
    if (s.dev < 0) return;
    cudaSetDevice(s.dev);
    if (s.d_topo) cudaFree(s.d_topo);
    if (s.d_free) cudaFree(s.d_free);
    if (s.d_mem) cudaFree(s.d_mem);
    if (s.d_flag) cudaFree(s.d_flag);
    if (s.d_order) cudaFree(s.d_order);
    if (s.d_cpair) cudaFree(s.d_cpair);
    if (s.d_perm) cudaFree(s.d_perm);
    if (s.d_pods) cudaFree(s.d_pods);
    if (s.d_keys) cudaFree(s.d_keys);
    if (s.d_gather) cudaFree(s.d_gather);
    if (s.d_bestk) cudaFree(s.d_bestk);
    if (s.d_nodebest) cudaFree(s.d_nodebest);
    if (s.d_tilebest) cudaFree(s.d_tilebest);
    if (s.d_work) cudaFree(s.d_work);
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = kgpu_shard();
}

}  // namespace

extern "C" {

const char *kgpu_version(void) { return "0.1.0"; }

const char *kgpu_last_error(kgpu_t *h) {
    if (h) {
        std::lock_guard<std::mutex> g(h->mu);
        t_last_error = h->err;   // return storage that outlives the lock
    }
    return t_last_error.c_str();
}

int kgpu_create(const int *dev_ids, int ndev, kgpu_t **out) {
    if (!out) return fail(nullptr, KGPU_ERR_INVALID, "kgpu_create: out is NULL");
    *out = nullptr;
    if (!dev_ids || ndev < 1 || ndev > 64) return fail(nullptr, KGPU_ERR_INVALID, "kgpu_create: need 1..64 device ids");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, KGPU_ERR_CUDA, "kgpu_create: no usable CUDA device (%s); libkgpu has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    kgpu_ctx *h = new (std::nothrow) kgpu_ctx();
    if (!h) return fail(nullptr, KGPU_ERR_NOMEM, "kgpu_create: out of host memory");
    memcpy(h->W, kDefaultWeights, sizeof h->W);
    h->shards.resize((size_t)ndev);
    int rc = KGPU_OK;
    for (int i = 0; i < ndev && rc == KGPU_OK; i++) {
        kgpu_shard &s = h->shards[(size_t)i];
        if (dev_ids[i] < 0 || dev_ids[i] >= count) {
            rc = fail(nullptr, KGPU_ERR_INVALID, "kgpu_create: device id %d out of range (have %d)", dev_ids[i], count);
            break;
        }
        s.dev = dev_ids[i];
        auto step = [&](cudaError_t ce, const char *what) {
            if (ce != cudaSuccess && rc == KGPU_OK)
                rc = fail(nullptr, KGPU_ERR_CUDA, "kgpu_create: %s on device %d: %s", what, s.dev, cudaGetErrorString(ce));
        };
        step(cudaSetDevice(s.dev), "cudaSetDevice");
        step(cudaDeviceGetAttribute(&s.sm_count, cudaDevAttrMultiProcessorCount, s.dev), "query SM count");
        step(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate");
        step(cudaEventCreate(&s.ev0), "cudaEventCreate");
        step(cudaEventCreate(&s.ev1), "cudaEventCreate");
        step(cudaMalloc(&s.d_bestk, 9 * 8), "cudaMalloc");
        step(cudaMalloc(&s.d_flag, sizeof(int)), "cudaMalloc");
    }
    if (rc == KGPU_OK && ndev > 1) {
        std::vector<int> devs(dev_ids, dev_ids + ndev);
        std::string why;
        h->multi = kgpu::MultiDevice::create(devs, &why);
        if (!h->multi) rc = fail(nullptr, KGPU_ERR_COMM, "kgpu_create: %s", why.c_str());
    }
    if (rc == KGPU_OK) rc = upload_subset_tables(h);
    if (rc != KGPU_OK) {
        std::string keep = t_last_error;
        for (auto &s : h->shards) free_shard(s);
        delete h->multi;
        delete h;
        t_last_error = keep;
        return rc;
    }
    *out = h;
    return KGPU_OK;
}

int kgpu_destroy(kgpu_t *h) {
    if (!h) return KGPU_OK;
    for (auto &s : h->shards) {
        if (s.dev >= 0) { cudaSetDevice(s.dev); cudaStreamSynchronize(s.stream); }
    }
    delete h->multi;
    if (h->xch.base) {
        for (int r = 0; r < h->xch.world; r++)
            if (h->xch.connected && r != h->xch.rank && h->xch.peer_base[r]) cudaIpcCloseMemHandle(h->xch.peer_base[r]);
        cudaFree(h->xch.base);
        cudaFree(h->xch.local);
    }
    for (auto &s : h->shards) free_shard(s);
    delete h;
    return KGPU_OK;
}

int kgpu_set_weights(kgpu_t *h, const int32_t w[KGPU_NUM_LEVELS]) {
    if (!h || !w) return fail(h, KGPU_ERR_INVALID, "kgpu_set_weights: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    for (int i = 0; i < 16; i++)
        if (w[i] < 0 || w[i] > KGPU_MAX_WEIGHT)
            return fail(h, KGPU_ERR_INVALID, "kgpu_set_weights: w[%d] = %d outside 0..%d", i, w[i], KGPU_MAX_WEIGHT);
    memcpy(h->W, w, sizeof h->W);
    for (auto &s : h->shards) s.compact_dirty = true;
    return KGPU_OK;
}

int kgpu_get_weights(kgpu_t *h, int32_t w[KGPU_NUM_LEVELS]) {
    if (!h || !w) return fail(h, KGPU_ERR_INVALID, "kgpu_get_weights: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    memcpy(w, h->W, sizeof h->W);
    return KGPU_OK;
}

int kgpu_set_variant(kgpu_t *h, int variant) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_set_variant: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (variant == KGPU_VARIANT_AUTO) variant = KGPU_VARIANT_SPARSE;
    if (variant < KGPU_VARIANT_WARP_PER_PAIR || variant > KGPU_VARIANT_SPARSE)
        return fail(h, KGPU_ERR_INVALID, "kgpu_set_variant: unknown variant %d", variant);
    h->variant = variant;
    return KGPU_OK;
}

int kgpu_upload_nodes(kgpu_t *h, const int32_t *topo, const int32_t *free_mask, int64_t n, int64_t node_id_base) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_upload_nodes: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (n < 0 || (n > 0 && (!topo || !free_mask))) return fail(h, KGPU_ERR_INVALID, "kgpu_upload_nodes: bad arguments");
    if (node_id_base < 0 || node_id_base + n > 0xFFFFFFFFLL)
        return fail(h, KGPU_ERR_INVALID, "kgpu_upload_nodes: node ids must fit in 32 bits");
    int rc = validate_topo(h, topo, n);
    if (rc != KGPU_OK) return rc;
    const int64_t G = (int64_t)h->shards.size();
    const int64_t per = (n + G - 1) / G;     // contiguous ranges: shard g holds [g*per, ...)
    int64_t off = 0;
    for (auto &s : h->shards) {
        const int64_t cnt = std::max<int64_t>(0, std::min<int64_t>(per, n - off));
        KGPU_CUDA(h, cudaSetDevice(s.dev));
        if (cnt > s.cap) {
            if (s.d_topo) cudaFree(s.d_topo);
            if (s.d_free) cudaFree(s.d_free);
            if (s.d_mem) cudaFree(s.d_mem);
            if (s.d_cpair) cudaFree(s.d_cpair);
            if (s.d_perm) cudaFree(s.d_perm);
            s.d_topo = nullptr; s.d_free = nullptr; s.d_mem = nullptr; s.d_cpair = nullptr; s.d_perm = nullptr; s.cap = 0;
            KGPU_CUDA(h, cudaMalloc(&s.d_cpair, (size_t)cnt * 112));
            KGPU_CUDA(h, cudaMalloc(&s.d_perm, (size_t)cnt * 4));
            KGPU_CUDA(h, cudaMalloc(&s.d_topo, (size_t)cnt * 256));
            KGPU_CUDA(h, cudaMalloc(&s.d_free, (size_t)cnt * 4));
            KGPU_CUDA(h, cudaMalloc(&s.d_mem, (size_t)cnt * 32));
            s.cap = cnt;
        }
        // per-GPU memory is unconstrained (0x7F7F7F7F MiB) until kgpu_upload_gpu_memory says otherwise
        if (cnt > 0) KGPU_CUDA(h, cudaMemsetAsync(s.d_mem, 0x7F, (size_t)cnt * 32, s.stream));
        if (cnt > 0) {
            KGPU_CUDA(h, cudaMemcpyAsync(s.d_topo, topo + off * 64, (size_t)cnt * 256, cudaMemcpyHostToDevice, s.stream));
            KGPU_CUDA(h, cudaMemcpyAsync(s.d_free, free_mask + off, (size_t)cnt * 4, cudaMemcpyHostToDevice, s.stream));
        }
        s.n = cnt;
        s.compact_dirty = true;
        s.node_id_base = node_id_base + off;
        // K1s order: nodes grouped by number of free GPUs (8 first), each class in increasing node
        // index and padded to whole warps (32 slots), so the lanes of a warp share the bound F and are
        // in increasing node id (tie-break order inside a warp; across warps the flush compares ids).
        {
            std::vector<int32_t> order;
            order.reserve((size_t)cnt + 9 * 128);
            for (int f = 8; f >= 0; f--) {
                for (int64_t i = 0; i < cnt; i++)
                    if (__builtin_popcount((unsigned)free_mask[off + i] & 0xFFu) == f) order.push_back((int32_t)i);
                while (order.size() % 32) order.push_back(-1);
            }
            while (order.size() % kgpu::SP_THREADS) order.push_back(-1);
            if ((int64_t)order.size() > s.order_cap) {
                if (s.d_order) cudaFree(s.d_order);
                s.d_order = nullptr; s.order_cap = 0;
                KGPU_CUDA(h, cudaMalloc(&s.d_order, std::max<size_t>(1, order.size()) * 4));
                s.order_cap = (int64_t)order.size();
            }
            s.n_slots = (int64_t)order.size();
            s.tile_class.assign(order.size() / kgpu::SP_THREADS, 0);
            for (size_t sl = 0; sl < order.size(); sl++)
                if (order[sl] >= 0) {
                    const uint8_t f = (uint8_t)__builtin_popcount((unsigned)free_mask[off + order[sl]] & 0xFFu);
                    uint8_t &tc = s.tile_class[sl / kgpu::SP_THREADS];
                    tc = std::max(tc, f);
                }
            s.work_P = -1;
            if (!order.empty()) {
                KGPU_CUDA(h, cudaMemcpyAsync(s.d_order, order.data(), order.size() * 4, cudaMemcpyHostToDevice, s.stream));
                KGPU_CUDA(h, cudaStreamSynchronize(s.stream));   // `order` is a local
            }
        }
        off += cnt;
    }
    for (auto &s : h->shards) {
        KGPU_CUDA(h, cudaSetDevice(s.dev));
        KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
    }
    h->n_total = n;
    return KGPU_OK;
}

int kgpu_upload_gpu_memory(kgpu_t *h, const int32_t *mem_mib, int64_t n) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_upload_gpu_memory: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (n != h->n_total || (n > 0 && !mem_mib)) return fail(h, KGPU_ERR_INVALID, "kgpu_upload_gpu_memory: n must equal kgpu_num_nodes");
    for (int64_t i = 0; i < n * 8; i++)
        if (mem_mib[i] < 0) return fail(h, KGPU_ERR_INVALID, "kgpu_upload_gpu_memory: negative memory at node %lld", (long long)(i / 8));
    int64_t off = 0;
    for (auto &s : h->shards) {
        if (s.n > 0) {
            KGPU_CUDA(h, cudaSetDevice(s.dev));
            KGPU_CUDA(h, cudaMemcpyAsync(s.d_mem, mem_mib + off * 8, (size_t)s.n * 32, cudaMemcpyHostToDevice, s.stream));
            KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
        }
        off += s.n;
    }
    return KGPU_OK;
}

int kgpu_update_gpu_memory(kgpu_t *h, int64_t idx, const int32_t mem_mib[8]) {
    if (!h || !mem_mib) return fail(h, KGPU_ERR_INVALID, "kgpu_update_gpu_memory: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    if (idx < 0 || idx >= h->n_total) return fail(h, KGPU_ERR_INVALID, "kgpu_update_gpu_memory: index %lld out of range", (long long)idx);
    for (int i = 0; i < 8; i++)
        if (mem_mib[i] < 0) return fail(h, KGPU_ERR_INVALID, "kgpu_update_gpu_memory: negative memory");
    int64_t local = 0;
    kgpu_shard *s = shard_of(h, idx, &local);
    KGPU_CUDA(h, cudaSetDevice(s->dev));
    KGPU_CUDA(h, cudaMemcpyAsync(s->d_mem + local * 8, mem_mib, 32, cudaMemcpyHostToDevice, s->stream));
    KGPU_CUDA(h, cudaStreamSynchronize(s->stream));
    return KGPU_OK;
}

int kgpu_update_node(kgpu_t *h, int64_t idx, const int32_t topo[64], int32_t free_mask) {
    if (!h || !topo) return fail(h, KGPU_ERR_INVALID, "kgpu_update_node: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    if (idx < 0 || idx >= h->n_total) return fail(h, KGPU_ERR_INVALID, "kgpu_update_node: index %lld out of range", (long long)idx);
    int rc = validate_topo(h, topo, 1);
    if (rc != KGPU_OK) return rc;
    int64_t local = 0;
    kgpu_shard *s = shard_of(h, idx, &local);
    KGPU_CUDA(h, cudaSetDevice(s->dev));
    s->compact_dirty = true;
    KGPU_CUDA(h, cudaMemcpyAsync(s->d_topo + local * 64, topo, 256, cudaMemcpyHostToDevice, s->stream));
    KGPU_CUDA(h, cudaMemcpyAsync(s->d_free + local, &free_mask, 4, cudaMemcpyHostToDevice, s->stream));
    KGPU_CUDA(h, cudaStreamSynchronize(s->stream));
    return KGPU_OK;
}

int kgpu_set_free_mask(kgpu_t *h, int64_t idx, int32_t free_mask) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_set_free_mask: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (idx < 0 || idx >= h->n_total) return fail(h, KGPU_ERR_INVALID, "kgpu_set_free_mask: index %lld out of range", (long long)idx);
    int64_t local = 0;
    kgpu_shard *s = shard_of(h, idx, &local);
    s->compact_dirty = true;
    KGPU_CUDA(h, cudaSetDevice(s->dev));
    KGPU_CUDA(h, cudaMemcpyAsync(s->d_free + local, &free_mask, 4, cudaMemcpyHostToDevice, s->stream));
    KGPU_CUDA(h, cudaStreamSynchronize(s->stream));
    return KGPU_OK;
}

int kgpu_remove_node(kgpu_t *h, int64_t idx) { return kgpu_set_free_mask(h, idx, 0); }

int64_t kgpu_num_nodes(kgpu_t *h) {
    if (!h) return 0;
    std::lock_guard<std::mutex> g(h->mu);
    return h->n_total;
}

int kgpu_score_batch(kgpu_t *h, const int32_t *pods, int64_t P, uint64_t *out_keys) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (P < 0 || (P > 0 && (!pods || !out_keys))) return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch: bad arguments");
    if (P == 0) return KGPU_OK;
    const size_t G = h->shards.size();
    for (auto &s : h->shards) {
        int rc = ensure_pod_capacity(h, s, P);
        if (rc != KGPU_OK) return rc;
    }
    int has_mem = 0;
    for (int64_t p = 0; p < P && !has_mem; p++) has_mem = pods[4 * p + 3] > 0;
    // every device: pods H2D, K1 over its node shard (devices run concurrently)
    for (auto &s : h->shards) {
        KGPU_CUDA(h, cudaSetDevice(s.dev));
        KGPU_CUDA(h, cudaMemcpyAsync(s.d_pods, pods, (size_t)P * 16, cudaMemcpyHostToDevice, s.stream));
        KGPU_CUDA(h, cudaEventRecord(s.ev0, s.stream));
        int rc = launch_score(h, s, s.d_pods, P, s.d_keys, s.stream, has_mem);
        if (rc != KGPU_OK) return rc;
        KGPU_CUDA(h, cudaEventRecord(s.ev1, s.stream));
    }
    kgpu_shard &s0 = h->shards[0];
    if (G > 1) {
        // one all-gather of every shard's per-pod best over NVLink, then K2 on device 0
        std::vector<const void *> send(G);
        std::vector<void *> recv(G);
        std::vector<cudaStream_t> streams(G);
        for (size_t i = 0; i < G; i++) { send[i] = h->shards[i].d_keys; recv[i] = h->shards[i].d_gather; streams[i] = h->shards[i].stream; }
        std::string why;
        if (!h->multi->all_gather_u64(send, recv, (size_t)P, streams, &why))
            return fail(h, KGPU_ERR_COMM, "kgpu_score_batch: %s", why.c_str());
        KGPU_CUDA(h, cudaSetDevice(s0.dev));
        kgpu::reduce_shards<<<(unsigned)((P + 255) / 256), 256, 0, s0.stream>>>(s0.d_gather, (int)G, P, s0.d_keys);
        h->launches++;
        KGPU_CUDA(h, cudaGetLastError());
    }
    KGPU_CUDA(h, cudaSetDevice(s0.dev));
    KGPU_CUDA(h, cudaMemcpyAsync(out_keys, s0.d_keys, (size_t)P * 8, cudaMemcpyDeviceToHost, s0.stream));
    double worst = 0.0;
    for (auto &s : h->shards) {
        KGPU_CUDA(h, cudaSetDevice(s.dev));
        KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
        float ms = 0.f;
        KGPU_CUDA(h, cudaEventElapsedTime(&ms, s.ev0, s.ev1));
        worst = std::max(worst, (double)ms);
    }
    h->last_kernel_ms = worst;
    return KGPU_OK;
}

int kgpu_place_batch(kgpu_t *h, const int32_t *pods, int64_t P, uint64_t *out_keys) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_place_batch: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (h->shards.size() != 1) return fail(h, KGPU_ERR_STATE, "kgpu_place_batch: needs a single-device handle");
    if (P < 0 || (P > 0 && (!pods || !out_keys))) return fail(h, KGPU_ERR_INVALID, "kgpu_place_batch: bad arguments");
    if (P == 0) return KGPU_OK;
    // the batch's distinct memory requirements are the views the sequential kernels keep tables for
    kgpu::PlaceViews views;
    memset(&views, 0, sizeof views);
    views.n = 1;
    for (int64_t p = 0; p < P; p++) {
        const int32_t need = pods[4 * p + 3];
        if (need <= 0 || pods[4 * p] < 0 || pods[4 * p] > 8) continue;
        bool seen = false;
        for (int j = 1; j < views.n; j++) seen = seen || views.min_mem[j] == need;
        if (seen) continue;
        if (views.n == kgpu::PLACE_MAX_VIEWS)
            return fail(h, KGPU_ERR_INVALID, "kgpu_place_batch: more than %d distinct min_mem values in one batch", kgpu::PLACE_MAX_VIEWS - 1);
        views.min_mem[views.n++] = need;
    }
    kgpu_shard &s = h->shards[0];
    int rc = ensure_pod_capacity(h, s, P);
    if (rc != KGPU_OK) return rc;
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    KGPU_CUDA(h, cudaMemcpyAsync(s.d_pods, pods, (size_t)P * 16, cudaMemcpyHostToDevice, s.stream));
    if (s.n == 0) {
        KGPU_CUDA(h, cudaMemsetAsync(s.d_keys, 0xFF, (size_t)P * 8, s.stream));
    } else {
        const int64_t T = (s.n + kgpu::PLACE_TILE - 1) / kgpu::PLACE_TILE, Npad = T * kgpu::PLACE_TILE;
        if (Npad * views.n > s.place_cap) {
            if (s.d_nodebest) cudaFree(s.d_nodebest);
            if (s.d_tilebest) cudaFree(s.d_tilebest);
            s.d_nodebest = nullptr; s.d_tilebest = nullptr; s.place_cap = 0;
            KGPU_CUDA(h, cudaMalloc(&s.d_nodebest, (size_t)Npad * views.n * 9 * 4));
            KGPU_CUDA(h, cudaMalloc(&s.d_tilebest, (size_t)T * views.n * 9 * 8));
            s.place_cap = Npad * views.n;
        }
        kgpu::Weights W;
        memcpy(W.w, h->W, sizeof W.w);
        KGPU_CUDA(h, cudaEventRecord(s.ev0, s.stream));
        kgpu::place_init<<<dim3((unsigned)T, (unsigned)views.n), kgpu::PLACE_TILE, 0, s.stream>>>(
            reinterpret_cast<const int4 *>(s.d_topo), s.d_free, s.d_mem, s.n, Npad, s.node_id_base, W, PC, views, s.d_nodebest,
            s.d_tilebest, T);
        kgpu::place_sequential<<<1, kgpu::PLACE_THREADS, 0, s.stream>>>(s.d_topo, s.d_free, s.d_mem, s.n, Npad, s.node_id_base,
                                                                        reinterpret_cast<const int4 *>(s.d_pods), P, W, views,
                                                                        s.d_nodebest, s.d_tilebest, T, s.d_keys);
        h->launches += 2;
        s.compact_dirty = true;          // the free masks have changed on the device
        KGPU_CUDA(h, cudaGetLastError());
        KGPU_CUDA(h, cudaEventRecord(s.ev1, s.stream));
    }
    KGPU_CUDA(h, cudaMemcpyAsync(out_keys, s.d_keys, (size_t)P * 8, cudaMemcpyDeviceToHost, s.stream));
    KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
    if (s.n > 0) {
        float ms = 0.f;
        KGPU_CUDA(h, cudaEventElapsedTime(&ms, s.ev0, s.ev1));
        h->last_kernel_ms = ms;
    }
    return KGPU_OK;
}

int kgpu_get_free_masks(kgpu_t *h, int32_t *out_free_mask, int64_t n) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_get_free_masks: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (n != h->n_total || (n > 0 && !out_free_mask)) return fail(h, KGPU_ERR_INVALID, "kgpu_get_free_masks: n must equal kgpu_num_nodes");
    int64_t off = 0;
    for (auto &s : h->shards) {
        if (s.n > 0) {
            KGPU_CUDA(h, cudaSetDevice(s.dev));
            KGPU_CUDA(h, cudaMemcpyAsync(out_free_mask + off, s.d_free, (size_t)s.n * 4, cudaMemcpyDeviceToHost, s.stream));
            KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
        }
        off += s.n;
    }
    return KGPU_OK;
}

int kgpu_score_pairs(kgpu_t *h, const int64_t *node_idx, const int32_t *k, const int32_t *min_mem_mib, int64_t n,
                     uint32_t *out_node_keys) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_score_pairs: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (n < 0 || (n > 0 && (!node_idx || !k || !out_node_keys))) return fail(h, KGPU_ERR_INVALID, "kgpu_score_pairs: bad arguments");
    for (int64_t i = 0; i < n; i++)
        if (node_idx[i] < 0 || node_idx[i] >= h->n_total)
            return fail(h, KGPU_ERR_INVALID, "kgpu_score_pairs: node index %lld out of range", (long long)node_idx[i]);
    kgpu::Weights W;
    memcpy(W.w, h->W, sizeof W.w);
    // route each pair to the shard that holds its node
    int64_t off = 0;
    for (auto &s : h->shards) {
        std::vector<long long> idx;
        std::vector<int32_t> kk, mm;
        std::vector<int64_t> pos;
        for (int64_t i = 0; i < n; i++)
            if (node_idx[i] >= off && node_idx[i] < off + s.n) {
                idx.push_back(node_idx[i] - off); kk.push_back(k[i]); mm.push_back(min_mem_mib ? min_mem_mib[i] : 0); pos.push_back(i);
            }
        off += s.n;
        if (idx.empty()) continue;
        const size_t m = idx.size();
        KGPU_CUDA(h, cudaSetDevice(s.dev));
        long long *d_idx = nullptr; int32_t *d_k = nullptr, *d_mm = nullptr; uint32_t *d_out = nullptr;
        KGPU_CUDA(h, cudaMallocAsync(&d_idx, m * 8, s.stream));
        KGPU_CUDA(h, cudaMallocAsync(&d_k, m * 4, s.stream));
        KGPU_CUDA(h, cudaMallocAsync(&d_mm, m * 4, s.stream));
        KGPU_CUDA(h, cudaMemcpyAsync(d_mm, mm.data(), m * 4, cudaMemcpyHostToDevice, s.stream));
        KGPU_CUDA(h, cudaMallocAsync(&d_out, m * 4, s.stream));
        KGPU_CUDA(h, cudaMemcpyAsync(d_idx, idx.data(), m * 8, cudaMemcpyHostToDevice, s.stream));
        KGPU_CUDA(h, cudaMemcpyAsync(d_k, kk.data(), m * 4, cudaMemcpyHostToDevice, s.stream));
        kgpu::score_pair_list<<<(unsigned)((m + 127) / 128), 128, 0, s.stream>>>(
            reinterpret_cast<const int4 *>(s.d_topo), s.d_free, s.d_mem, s.n, d_idx, d_k, d_mm, (int64_t)m, W, PC, d_out);
        h->launches++;
        KGPU_CUDA(h, cudaGetLastError());
        std::vector<uint32_t> res(m);
        KGPU_CUDA(h, cudaMemcpyAsync(res.data(), d_out, m * 4, cudaMemcpyDeviceToHost, s.stream));
        KGPU_CUDA(h, cudaFreeAsync(d_idx, s.stream));
        KGPU_CUDA(h, cudaFreeAsync(d_k, s.stream));
        KGPU_CUDA(h, cudaFreeAsync(d_mm, s.stream));
        KGPU_CUDA(h, cudaFreeAsync(d_out, s.stream));
        KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
        for (size_t j = 0; j < m; j++) out_node_keys[pos[j]] = res[j];
    }
    return KGPU_OK;
}

int kgpu_score_batch_device(kgpu_t *h, const int32_t *d_pods, int64_t P, uint64_t *d_keys, void *stream) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch_device: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (h->shards.size() != 1) return fail(h, KGPU_ERR_STATE, "kgpu_score_batch_device: needs a single-device handle");
    if (P < 0 || (P > 0 && (!d_pods || !d_keys))) return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch_device: bad arguments");
    kgpu_shard &s = h->shards[0];
    return launch_score(h, s, d_pods, P, reinterpret_cast<unsigned long long *>(d_keys), (cudaStream_t)stream, -1);
}

int kgpu_score_batch_device_ex(kgpu_t *h, const int32_t *d_pods, int64_t P, uint64_t *d_keys, void *stream, int batch_flags) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch_device_ex: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (h->shards.size() != 1) return fail(h, KGPU_ERR_STATE, "kgpu_score_batch_device_ex: needs a single-device handle");
    if (P < 0 || (P > 0 && (!d_pods || !d_keys))) return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch_device_ex: bad arguments");
    kgpu_shard &s = h->shards[0];
    return launch_score(h, s, d_pods, P, reinterpret_cast<unsigned long long *>(d_keys), (cudaStream_t)stream,
                        (batch_flags & KGPU_BATCH_NO_MIN_MEM) ? 0 : -1);
}

namespace {
size_t xch_bytes(int64_t max_pods) { return (size_t)max_pods * 16 + kgpu::PEER_MAX_WORLD * 4 + 16; }
unsigned long long *xch_results(void *base, int64_t max_pods, int buf) { return reinterpret_cast<unsigned long long *>(base) + (int64_t)buf * max_pods; }
uint32_t *xch_flags(void *base, int64_t max_pods) { return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(base) + (size_t)max_pods * 16); }
}  // namespace

int kgpu_exchange_init(kgpu_t *h, int world, int rank, int64_t max_pods, unsigned char *out_handle) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_exchange_init: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    static_assert(sizeof(cudaIpcMemHandle_t) == KGPU_IPC_HANDLE_BYTES, "IPC handle size");
    if (h->shards.size() != 1) return fail(h, KGPU_ERR_STATE, "kgpu_exchange_init: needs a single-device handle");
    if (world < 1 || world > kgpu::PEER_MAX_WORLD || rank < 0 || rank >= world || max_pods < 1 || !out_handle)
        return fail(h, KGPU_ERR_INVALID, "kgpu_exchange_init: bad arguments (world <= %d)", kgpu::PEER_MAX_WORLD);
    if (h->xch.base) return fail(h, KGPU_ERR_STATE, "kgpu_exchange_init: already initialised");
    kgpu_shard &s = h->shards[0];
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    KGPU_CUDA(h, cudaMalloc(&h->xch.base, xch_bytes(max_pods)));
    KGPU_CUDA(h, cudaMalloc(&h->xch.local, (size_t)max_pods * 8));
    KGPU_CUDA(h, cudaMemset(h->xch.base, 0xFF, (size_t)max_pods * 16));                                   // both result buffers: NO_FIT
    KGPU_CUDA(h, cudaMemset(xch_flags(h->xch.base, max_pods), 0, kgpu::PEER_MAX_WORLD * 4 + 16));         // flags, ticket
    KGPU_CUDA(h, cudaDeviceSynchronize());
    cudaIpcMemHandle_t ipc;
    KGPU_CUDA(h, cudaIpcGetMemHandle(&ipc, h->xch.base));
    memcpy(out_handle, &ipc, sizeof ipc);
    h->xch.world = world; h->xch.rank = rank; h->xch.max_pods = max_pods; h->xch.epoch = 0; h->xch.connected = false;
    return KGPU_OK;
}

int kgpu_exchange_connect(kgpu_t *h, const unsigned char *handles) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_exchange_connect: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->xch.base) return fail(h, KGPU_ERR_STATE, "kgpu_exchange_connect: call kgpu_exchange_init first");
    if (!handles) return fail(h, KGPU_ERR_INVALID, "kgpu_exchange_connect: NULL handles");
    if (h->xch.connected) return fail(h, KGPU_ERR_STATE, "kgpu_exchange_connect: already connected");
    KGPU_CUDA(h, cudaSetDevice(h->shards[0].dev));
    for (int r = 0; r < h->xch.world; r++) {
        if (r == h->xch.rank) { h->xch.peer_base[r] = h->xch.base; continue; }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, handles + (size_t)r * KGPU_IPC_HANDLE_BYTES, sizeof ipc);
        KGPU_CUDA(h, cudaIpcOpenMemHandle(&h->xch.peer_base[r], ipc, cudaIpcMemLazyEnablePeerAccess));
    }
    h->xch.connected = true;
    return KGPU_OK;
}

int kgpu_score_batch_exchange(kgpu_t *h, const int32_t *d_pods, int64_t P, const uint64_t **d_final_keys, void *stream,
                              int batch_flags) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch_exchange: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->xch.connected) return fail(h, KGPU_ERR_STATE, "kgpu_score_batch_exchange: exchange not connected");
    if (P < 0 || P > h->xch.max_pods || (P > 0 && !d_pods) || !d_final_keys)
        return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch_exchange: bad arguments (P <= max_pods of kgpu_exchange_init)");
    kgpu_shard &s = h->shards[0];
    cudaStream_t st = (cudaStream_t)stream;
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    const uint32_t epoch = ++h->xch.epoch;
    const int buf = (int)(epoch & 1u);
    const int64_t mp = h->xch.max_pods;
    *d_final_keys = reinterpret_cast<const uint64_t *>(xch_results(h->xch.base, mp, buf));
    if (P == 0) return KGPU_OK;
    // the other buffer is what peers push into NEXT epoch: clean it before this rank can reach this epoch's barrier
    // (all max_pods entries: the next batch may be longer than this one)
    KGPU_CUDA(h, cudaMemsetAsync(xch_results(h->xch.base, mp, buf ^ 1), 0xFF, (size_t)mp * 8, st));
    const int rc = launch_score(h, s, d_pods, P, h->xch.local, st, (batch_flags & KGPU_BATCH_NO_MIN_MEM) ? 0 : -1);
    if (rc != KGPU_OK) return rc;
    kgpu::PeerTable tab;
    memset(&tab, 0, sizeof tab);
    for (int r = 0; r < h->xch.world; r++) {
        tab.results[r] = xch_results(h->xch.peer_base[r], mp, buf);
        tab.flags[r] = xch_flags(h->xch.peer_base[r], mp);
    }
    unsigned int *ticket = reinterpret_cast<unsigned int *>(xch_flags(h->xch.base, mp) + kgpu::PEER_MAX_WORLD);
    kgpu::push_and_sync<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(h->xch.local, P, tab, h->xch.rank, h->xch.world, epoch, ticket);
    h->launches++;
    KGPU_CUDA(h, cudaGetLastError());
    return KGPU_OK;
}

int kgpu_reduce_shards_device(kgpu_t *h, const uint64_t *d_gathered, int G, int64_t P, uint64_t *d_out, void *stream) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_reduce_shards_device: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (G < 1 || P < 0 || (P > 0 && (!d_gathered || !d_out))) return fail(h, KGPU_ERR_INVALID, "kgpu_reduce_shards_device: bad arguments");
    if (P == 0) return KGPU_OK;
    kgpu_shard &s = h->shards[0];
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    kgpu::reduce_shards<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long *>(d_gathered), G, P, reinterpret_cast<unsigned long long *>(d_out));
    h->launches++;
    KGPU_CUDA(h, cudaGetLastError());
    return KGPU_OK;
}

int64_t kgpu_kernel_launches(kgpu_t *h) {
    if (!h) return 0;
    std::lock_guard<std::mutex> g(h->mu);
    return h->launches;
}

double kgpu_last_kernel_ms(kgpu_t *h) {
    if (!h) return 0.0;
    std::lock_guard<std::mutex> g(h->mu);
    return h->last_kernel_ms;
}

}  // extern "C"
