// kgpu.cu -- C ABI of libkgpu (include/kgpu.h) over the sm_100a kernels in
// score_pairs.cuh.  No CPU fallback: every entry point needs a CUDA device.
#include "../../include/kgpu.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
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
#include "node_state.cuh"

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
    int4 *d_rec = nullptr;           // K1s: [tiles][7][SP_THREADS] compacted scaled pair costs, slot order (compact_nodes)
    uint32_t *d_meta = nullptr;      // K1s: [n_slots] position -> GPU index (3-bit fields) | free count << 24
    bool compact_dirty = true;       // topology / weights changed for every node since the cache was built
    bool order_dirty = true;         // the order (slots by free count) must be rebuilt before the next K1s launch
    int64_t stale_nodes = 0;         // nodes whose free count changed since the order was built (re-sort when many)
    int32_t *d_order = nullptr;      // K1s: slot -> node index, grouped by popcount(free), -1 = padding
    int32_t *d_slot_of = nullptr;    // K1s: node index -> slot
    int32_t *d_ord_cnt = nullptr, *d_ord_off = nullptr;   // counting sort scratch [9][blocks]
    long long *d_ord_meta = nullptr; // class counts [9], n_slots
    int64_t ord_nb_cap = 0;
    unsigned long long *d_bad = nullptr;   // validate_topo_dev: first out-of-domain element
    int32_t *d_upd = nullptr;        // state-change scratch: [cap] node indices | [cap] masks
    uint32_t *d_fit_patch = nullptr; // [9][cap] fit-table rows of the nodes a state change touched
    int64_t upd_cap = 0;
    uint32_t *d_fit = nullptr;       // (node, k) fit table on the device [9][n] and, patches, [9][upd]
    uint32_t *h_fit = nullptr;       // pinned host copy [9][n]: kgpu_fit_lookup / kgpu_score_pairs read it
    int64_t fit_cap = 0;
    bool fit_valid = false;
    int4 *d_query = nullptr;         // kgpu_score_pairs with min_mem: {node lo, node hi, k, min_mem}
    uint32_t *d_qout = nullptr;
    int64_t query_cap = 0;
    int32_t *d_free_scratch = nullptr;   // kgpu_place_batch_ex(KGPU_PLACE_DRY_RUN): the masks the dry run consumes
    int64_t free_scratch_cap = 0;
    int64_t class_count[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int64_t n_slots = 0, slot_cap = 0;
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
    bool tma_attr_done = false;                 // cudaFuncSetAttribute of the TMA instantiations done on this device
    uint32_t *d_nodebest = nullptr;          // [views][9][Npad]  K3 tables
    int32_t *d_half = nullptr;               // [Npad][96]  K3: half tables of every node (place_half_tables)
    unsigned long long *d_tilebest = nullptr;   // [views][9][T]
    int64_t place_cap = 0;
    int64_t pcap = 0;
};

struct kgpu_ctx {
    std::mutex mu;
    std::string err;
    std::vector<kgpu_shard> shards;
    int32_t W[16];
    int variant = KGPU_VARIANT_AUTO;
    int64_t n_total = 0;
    int64_t launches = 0;
    double last_kernel_ms = 0.0;
    bool subsets_uploaded = false;
    double last_upload_ms = 0.0;
    kgpu::MultiDevice *multi = nullptr;   // NCCL communicator set, ndev > 1 only
    // peer-memory key exchange (kgpu_exchange_*): one allocation per rank, mapped by every peer:
    //   slots[2][world][max_pods] uint64 | flags[PEER_MAX_WORLD] uint32 | ticket | error ;  local[] and final_keys[] are private
    struct {
        int world = 0, rank = 0;
        int64_t max_pods = 0;
        void *base = nullptr;
        void *peer_base[kgpu::PEER_MAX_WORLD] = {};
        unsigned long long *local = nullptr;
        unsigned long long *final_keys = nullptr;   // [2][max_pods]: the global keys of the last two epochs
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

// Classes of the K1s order are padded to whole warps (32).  Padding them to whole tiles (KGPU_ORDER_PAD=128: no block ever
// mixes classes) was measured on C2 shards of 12.5k .. 100k nodes and bought nothing (0.0778 vs 0.0758 ms at 12.5k, equal
// above), so the smaller order stays.
const int kOrderPad = [] { const char *e = getenv("KGPU_ORDER_PAD"); const int v = e ? atoi(e) : 0; return v == 128 ? (int)kgpu::SP_THREADS : 32; }();

// (Re)build the K1s order of shard s on the device: stable counting sort of the node indices by free-GPU
// count (node_state.cuh).  One 80-byte read-back tells the host the class sizes (grid size, work list).
int build_order(kgpu_ctx *h, kgpu_shard &s) {
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    const int64_t need_slots = s.n + 10 * kgpu::SP_THREADS;
    if (need_slots > s.slot_cap) {
        if (s.d_order) cudaFree(s.d_order);
        if (s.d_rec) cudaFree(s.d_rec);
        if (s.d_meta) cudaFree(s.d_meta);
        s.d_order = nullptr; s.d_rec = nullptr; s.d_meta = nullptr; s.slot_cap = 0;
        const int64_t cap = (need_slots + kgpu::SP_THREADS - 1) / kgpu::SP_THREADS * kgpu::SP_THREADS;
        KGPU_CUDA(h, cudaMalloc(&s.d_order, (size_t)cap * 4));
        KGPU_CUDA(h, cudaMalloc(&s.d_meta, (size_t)cap * 4));
        KGPU_CUDA(h, cudaMalloc(&s.d_rec, (size_t)cap * 112));
        s.slot_cap = cap;
    }
    const int nb = (int)((s.n + kgpu::ORD_BLOCK - 1) / kgpu::ORD_BLOCK);
    if (nb > s.ord_nb_cap) {
        if (s.d_ord_cnt) cudaFree(s.d_ord_cnt);
        if (s.d_ord_off) cudaFree(s.d_ord_off);
        s.d_ord_cnt = nullptr; s.d_ord_off = nullptr; s.ord_nb_cap = 0;
        KGPU_CUDA(h, cudaMalloc(&s.d_ord_cnt, (size_t)nb * 9 * 4));
        KGPU_CUDA(h, cudaMalloc(&s.d_ord_off, (size_t)nb * 9 * 4));
        s.ord_nb_cap = nb;
    }
    long long meta[kgpu::ORD_META] = {0};
    if (s.n > 0) {
        KGPU_CUDA(h, cudaMemsetAsync(s.d_order, 0xFF, (size_t)s.slot_cap * 4, s.stream));
        kgpu::order_count<<<nb, kgpu::ORD_BLOCK, 0, s.stream>>>(s.d_free, s.n, s.d_ord_cnt, nb);
        kgpu::order_scan<<<1, kgpu::ORD_BLOCK, 0, s.stream>>>(s.d_ord_cnt, nb, s.d_ord_off, s.d_ord_meta, kgpu::SP_THREADS, kOrderPad);
        kgpu::order_scatter<<<nb, kgpu::ORD_BLOCK, 0, s.stream>>>(s.d_free, s.n, s.d_ord_off, nb, s.d_order, s.d_slot_of);
        h->launches += 3;
        KGPU_CUDA(h, cudaGetLastError());
        KGPU_CUDA(h, cudaMemcpyAsync(meta, s.d_ord_meta, sizeof meta, cudaMemcpyDeviceToHost, s.stream));
        KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
    }
    for (int c = 0; c < 9; c++) s.class_count[c] = meta[c];
    s.n_slots = meta[9];
    // max free count per tile, from the class layout: class c's nodes sit in [start, start + count)
    s.tile_class.assign((size_t)(s.n_slots / kgpu::SP_THREADS), 0);
    int64_t start = 0;
    for (int c = 8; c >= 0; c--) {
        const int64_t cnt = s.class_count[c];
        if (cnt > 0)
            for (int64_t t = start / kgpu::SP_THREADS; t <= (start + cnt - 1) / kgpu::SP_THREADS; t++)
                s.tile_class[(size_t)t] = std::max<uint8_t>(s.tile_class[(size_t)t], (uint8_t)c);
        start += (cnt + kOrderPad - 1) / kOrderPad * kOrderPad;
    }
    s.work_P = -1;
    s.order_dirty = false;
    s.stale_nodes = 0;
    s.compact_dirty = true;            // records are stored in slot order
    return KGPU_OK;
}

// Bring the K1s node cache of shard s up to date (order, compacted records); everything runs on s.stream and
// has completed when this returns, so a following launch on any stream sees it.
int ensure_node_cache(kgpu_ctx *h, kgpu_shard &s) {
    static const int64_t resort_div = [] { const char *e = getenv("KGPU_RESORT_DIV"); int v = e ? atoi(e) : 0; return (int64_t)(v > 0 ? v : 200); }();
    // Nodes whose free count changed keep their slot, so their warp enumerates for the LARGEST count among its 32
    // lanes: measured on C2, 1 % of the nodes changed -> the next step takes 1.03 ms instead of 0.48; a re-sort +
    // recompaction costs 0.18 ms.  So re-sort as soon as more than n / 200 nodes have changed (KGPU_RESORT_DIV).
    if (!s.order_dirty && s.stale_nodes * resort_div > s.n) s.order_dirty = true;
    if (!s.order_dirty && !s.compact_dirty) return KGPU_OK;
    if (s.order_dirty) {
        const int rc = build_order(h, s);
        if (rc != KGPU_OK) return rc;
    }
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    if (s.compact_dirty && s.n_slots > 0) {
        kgpu::Weights W;
        memcpy(W.w, h->W, sizeof W.w);
        kgpu::compact_nodes<<<(unsigned)(s.n_slots / kgpu::SP_THREADS), kgpu::SP_THREADS, 0, s.stream>>>(
            reinterpret_cast<const int4 *>(s.d_topo), s.d_free, s.n_slots, W, s.d_order, s.d_slot_of, s.d_rec, s.d_meta);
        h->launches++;
        KGPU_CUDA(h, cudaGetLastError());
    }
    s.compact_dirty = false;
    KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
    return KGPU_OK;
}

int ensure_upd_capacity(kgpu_ctx *h, kgpu_shard &s, int64_t m) {
    if (m <= s.upd_cap) return KGPU_OK;
    if (s.d_upd) cudaFree(s.d_upd);
    if (s.d_fit_patch) cudaFree(s.d_fit_patch);
    s.d_upd = nullptr; s.d_fit_patch = nullptr; s.upd_cap = 0;
    const int64_t cap = std::max<int64_t>(256, m + m / 2);
    KGPU_CUDA(h, cudaMalloc(&s.d_upd, (size_t)cap * 8));
    KGPU_CUDA(h, cudaMalloc(&s.d_fit_patch, (size_t)cap * 36));
    s.upd_cap = cap;
    return KGPU_OK;
}

// (node, k) fit table of shard s: fit_nodes over every node, one copy to the pinned host array.
int build_fit_table(kgpu_ctx *h, kgpu_shard &s) {
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    if (s.n > s.fit_cap) {
        if (s.d_fit) cudaFree(s.d_fit);
        if (s.h_fit) cudaFreeHost(s.h_fit);
        s.d_fit = nullptr; s.h_fit = nullptr; s.fit_cap = 0;
        KGPU_CUDA(h, cudaMalloc(&s.d_fit, (size_t)s.n * 36));
        KGPU_CUDA(h, cudaMallocHost(&s.h_fit, (size_t)s.n * 36));
        s.fit_cap = s.n;
    }
    if (s.n > 0) {
        kgpu::Weights W;
        memcpy(W.w, h->W, sizeof W.w);
        kgpu::fit_nodes<<<(unsigned)((s.n + 127) / 128), 128, 0, s.stream>>>(reinterpret_cast<const int4 *>(s.d_topo), s.d_free,
                                                                          nullptr, s.n, W, PC, s.d_fit);
        h->launches++;
        KGPU_CUDA(h, cudaGetLastError());
        KGPU_CUDA(h, cudaMemcpyAsync(s.h_fit, s.d_fit, (size_t)s.n * 36, cudaMemcpyDeviceToHost, s.stream));
        KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
    }
    s.fit_valid = true;
    return KGPU_OK;
}

// State change of m nodes of shard s (local indices idx, already distinct): masks != nullptr scatters the new
// free masks; either way the compacted records of exactly those nodes are refreshed, and the fit table's rows
// if it is valid.  One H2D copy, one or two kernels, one small D2H copy (fit rows), synchronous.
int apply_node_updates(kgpu_ctx *h, kgpu_shard &s, const int32_t *idx, const int32_t *masks, int64_t m) {
    if (m <= 0) return KGPU_OK;
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    int rc = ensure_upd_capacity(h, s, m);
    if (rc != KGPU_OK) return rc;
    KGPU_CUDA(h, cudaMemcpyAsync(s.d_upd, idx, (size_t)m * 4, cudaMemcpyHostToDevice, s.stream));
    if (masks) KGPU_CUDA(h, cudaMemcpyAsync(s.d_upd + s.upd_cap, masks, (size_t)m * 4, cudaMemcpyHostToDevice, s.stream));
    kgpu::Weights W;
    memcpy(W.w, h->W, sizeof W.w);
    const int4 *topo4 = reinterpret_cast<const int4 *>(s.d_topo);
    const bool cache_live = !s.order_dirty && !s.compact_dirty;
    if (cache_live) {
        kgpu::compact_nodes<<<(unsigned)((m + kgpu::SP_THREADS - 1) / kgpu::SP_THREADS), kgpu::SP_THREADS, 0, s.stream>>>(
            topo4, s.d_free, m, W, s.d_order, s.d_slot_of, s.d_rec, s.d_meta, s.d_upd, masks ? s.d_upd + s.upd_cap : nullptr);
        h->launches++;
        s.stale_nodes += m;
    } else if (masks) {
        kgpu::scatter_masks<<<(unsigned)((m + 255) / 256), 256, 0, s.stream>>>(s.d_upd, s.d_upd + s.upd_cap, m, s.d_free);
        h->launches++;
    }
    KGPU_CUDA(h, cudaGetLastError());
    if (s.fit_valid) {
        // rows of the changed nodes -> the persistent patch buffer (grown with the index scratch) -> the host table
        kgpu::fit_nodes<<<(unsigned)((m + 127) / 128), 128, 0, s.stream>>>(topo4, s.d_free, s.d_upd, m, W, PC, s.d_fit_patch);
        h->launches++;
        KGPU_CUDA(h, cudaGetLastError());
        std::vector<uint32_t> rows((size_t)m * 9);
        KGPU_CUDA(h, cudaMemcpyAsync(rows.data(), s.d_fit_patch, (size_t)m * 36, cudaMemcpyDeviceToHost, s.stream));
        KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
        for (int k = 0; k < 9; k++)
            for (int64_t i = 0; i < m; i++) s.h_fit[(int64_t)k * s.n + idx[i]] = rows[(size_t)((int64_t)k * m + i)];
    } else {
        KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
    }
    return KGPU_OK;
}

// Enqueue K1 for P pods on shard s: d_keys[p] = best placement over this shard's nodes.
// has_mem: 1 / 0 = the host knows whether some pod carries min_mem > 0; -1 = unknown (device buffers).
int launch_score(kgpu_ctx *h, kgpu_shard &s, const int32_t *d_pods, int64_t P, unsigned long long *d_keys,
                 cudaStream_t st, int has_mem, bool keys_are_clean = false) {
    if (P <= 0) return KGPU_OK;
    if ((reinterpret_cast<uintptr_t>(d_pods) & 15u) != 0)
        return fail(h, KGPU_ERR_INVALID, "pods pointer must be 16-byte aligned");
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    kgpu::Weights W;
    memcpy(W.w, h->W, sizeof W.w);
    const int4 *topo4 = reinterpret_cast<const int4 *>(s.d_topo);
    const int4 *pods4 = reinterpret_cast<const int4 *>(d_pods);

    const int4 *mem4 = reinterpret_cast<const int4 *>(s.d_mem);
    if (!keys_are_clean) KGPU_CUDA(h, cudaMemsetAsync(d_keys, 0xFF, (size_t)P * 8, st));
    if (s.n == 0) return KGPU_OK;

    // AUTO (the default): whatever is cheapest for THIS batch with identical keys -- pods without a memory requirement
    // are memoisable by k under snapshot scoring (best[k] over all nodes once, then a gather: memo_best_by_k /
    // memo_gather), pods with one go to K1m (the sparse MEM instantiation).  The per-pair kernels are the explicitly
    // named variants (KGPU_VARIANT_SPARSE = what bench.py's headline times, per north_star).
    const bool automode = h->variant == KGPU_VARIANT_AUTO;
    const bool memo = automode || h->variant == KGPU_VARIANT_MEMO_BY_K;
    const bool wpp = h->variant == KGPU_VARIANT_WARP_PER_PAIR;
    const bool sparse_main = h->variant == KGPU_VARIANT_SPARSE;
    const bool sparse = sparse_main || (automode && has_mem != 0);
    if (sparse) {                      // order + compacted records current?  (no-op unless nodes / masks / weights changed)
        const int rc = ensure_node_cache(h, s);
        if (rc != KGPU_OK) return rc;
    }
    if (!wpp && has_mem != 0) {   // which pods go to K1m?  (flag read by its blocks; skipped when the host knows there are none)
        KGPU_CUDA(h, cudaMemsetAsync(s.d_flag, 0, sizeof(int), st));
        kgpu::any_mem_pod<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(pods4, P, s.d_flag);
        h->launches++;
    }
    if (memo) {
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
            // The work list (sparse_work.h: items of about equal work, heaviest first; runs of tiles when the batch is
            // one chunk) instead of the plain grid.  Measured on C2 at full size: 0.485 -> 0.433 ms (the plain grid's
            // equal pod ranges leave a 13 % tail); KGPU_SP_WORKLIST=0 forces the plain grid.
            static const int worklist_mode = [] { const char *e = getenv("KGPU_SP_WORKLIST"); return e ? atoi(e) : -1; }();
            const bool use_work = worklist_mode != 0;
            static const int tma_mode = [] { const char *e = getenv("KGPU_SP_TMA"); return e ? atoi(e) : 1; }();
            const int4 *d_work = nullptr;
            if (use_work) {
                if (s.work_P != P) {
                    kgpu::SparseWorkParams prm;
                    static const int max_run = [] { const char *e = getenv("KGPU_SP_MAXRUN"); return e ? atoi(e) : 0; }();
                    if (max_run > 0) prm.max_run = max_run;
                    static const int env_tail = [] { const char *e = getenv("KGPU_SP_TAIL"); return e ? atoi(e) : -1; }();
                    static const int env_waves = [] { const char *e = getenv("KGPU_SP_WAVES"); return e ? atoi(e) : -1; }();
                    static const int env_floor = [] { const char *e = getenv("KGPU_SP_FLOOR"); return e ? atoi(e) : -1; }();
                    if (env_tail >= 0) prm.tail_percent = env_tail;
                    if (env_waves > 0) prm.waves = env_waves;
                    if (env_floor > 0) prm.floor = env_floor;
                    const int64_t res_list = (tma_mode != 0 && P <= kgpu::SP_TMA_PODS) ? (int64_t)s.sm_count * kgpu::sp_tma_blocks(P)
                                             : P <= kgpu::kSparseChunk               ? (int64_t)s.sm_count * KGPU_SP_STREAM_MINBLOCKS
                                                                                     : resident;
                    kgpu::build_sparse_work(s.tile_class, P, res_list, s.h_work, prm);
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
#define KGPU_LAUNCH_SPARSE(MEMF, BK, ST)                                                                      \
    kgpu::score_pairs_sparse<true, MEMF, BK, ST><<<grid, kgpu::SP_THREADS, 0, st>>>(                          \
        s.d_rec, s.d_meta, s.d_mem, s.d_order, s.d_flag, s.node_id_base, pods4, P, (int)per, d_work, PC, d_keys)
#define KGPU_LAUNCH_SPARSE_TMA(BK, MB)                                                                        \
    kgpu::score_pairs_sparse<true, false, BK, true, MB><<<grid, kgpu::SP_THREADS, kgpu::SP_TMA_DYN_SMEM, st>>>(   \
        s.d_rec, s.d_meta, s.d_mem, s.d_order, s.d_flag, s.node_id_base, pods4, P, (int)per, d_work, PC, d_keys)
            // few pods (the batch is one chunk): the work list holds runs of tiles -> the STREAM instantiation (next
            // tile's record prefetched into registers while the current one is scored); at most 64 pods: the TMA one
            // (tiles staged by cp.async.bulk into shared memory, 7 or 8 blocks per SM; KGPU_SP_TMA=0 turns it off)
            const bool stream_build = use_work && P <= kgpu::kSparseChunk;
            const bool tma_build = stream_build && tma_mode != 0 && P <= kgpu::SP_TMA_PODS;
            if (tma_build && !s.tma_attr_done) {   // static + dynamic shared memory may pass 48 KB with more stages: opt in, once per device
                const int bytes = (int)kgpu::SP_TMA_DYN_SMEM;
                cudaError_t e = cudaFuncSetAttribute(kgpu::score_pairs_sparse<true, false, true, true, KGPU_SP_TMA_MINBLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(kgpu::score_pairs_sparse<true, false, false, true, KGPU_SP_TMA_MINBLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(kgpu::score_pairs_sparse<true, false, true, true, KGPU_SP_TMA_MINBLOCKS_FEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(kgpu::score_pairs_sparse<true, false, false, true, KGPU_SP_TMA_MINBLOCKS_FEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
                if (e != cudaSuccess) return fail(h, KGPU_ERR_CUDA, "cudaFuncSetAttribute(TMA K1s): %s", cudaGetErrorString(e));
                s.tma_attr_done = true;
            }
            if (sparse_main) {
                if (tma_build) {
                    if (P <= kgpu::SP_TMA_FEW_PODS) { if (byte_keys) KGPU_LAUNCH_SPARSE_TMA(true, KGPU_SP_TMA_MINBLOCKS_FEW); else KGPU_LAUNCH_SPARSE_TMA(false, KGPU_SP_TMA_MINBLOCKS_FEW); }
                    else                            { if (byte_keys) KGPU_LAUNCH_SPARSE_TMA(true, KGPU_SP_TMA_MINBLOCKS); else KGPU_LAUNCH_SPARSE_TMA(false, KGPU_SP_TMA_MINBLOCKS); }
                }
                else if (stream_build) { if (byte_keys) KGPU_LAUNCH_SPARSE(false, true, true); else KGPU_LAUNCH_SPARSE(false, false, true); }
                else                   { if (byte_keys) KGPU_LAUNCH_SPARSE(false, true, false); else KGPU_LAUNCH_SPARSE(false, false, false); }
                h->launches++;
            }
            if (has_mem != 0) {
                if (byte_keys) KGPU_LAUNCH_SPARSE(true, true, false); else KGPU_LAUNCH_SPARSE(true, false, false);
                h->launches++;
            }
#undef KGPU_LAUNCH_SPARSE
#undef KGPU_LAUNCH_SPARSE_TMA
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
    if (s.d_rec) cudaFree(s.d_rec);
    if (s.d_meta) cudaFree(s.d_meta);
    if (s.d_slot_of) cudaFree(s.d_slot_of);
    if (s.d_ord_cnt) cudaFree(s.d_ord_cnt);
    if (s.d_ord_off) cudaFree(s.d_ord_off);
    if (s.d_ord_meta) cudaFree(s.d_ord_meta);
    if (s.d_bad) cudaFree(s.d_bad);
    if (s.d_upd) cudaFree(s.d_upd);
    if (s.d_fit_patch) cudaFree(s.d_fit_patch);
    if (s.d_fit) cudaFree(s.d_fit);
    if (s.h_fit) cudaFreeHost(s.h_fit);
    if (s.d_free_scratch) cudaFree(s.d_free_scratch);
    if (s.d_query) cudaFree(s.d_query);
    if (s.d_qout) cudaFree(s.d_qout);
    if (s.d_pods) cudaFree(s.d_pods);
    if (s.d_keys) cudaFree(s.d_keys);
    if (s.d_gather) cudaFree(s.d_gather);
    if (s.d_bestk) cudaFree(s.d_bestk);
    if (s.d_nodebest) cudaFree(s.d_nodebest);
    if (s.d_tilebest) cudaFree(s.d_tilebest);
    if (s.d_half) cudaFree(s.d_half);
    if (s.d_work) cudaFree(s.d_work);
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = kgpu_shard();
}

}  // namespace

extern "C" {

const char *kgpu_version(void) { return "0.2.0"; }

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
        step(cudaMalloc(&s.d_ord_meta, kgpu::ORD_META * sizeof(long long)), "cudaMalloc");
        step(cudaMalloc(&s.d_bad, sizeof(unsigned long long)), "cudaMalloc");
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
        cudaFree(h->xch.final_keys);
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
    for (auto &s : h->shards) { s.compact_dirty = true; s.fit_valid = false; }
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
    if (variant < KGPU_VARIANT_AUTO || variant > KGPU_VARIANT_SPARSE)
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
    const auto t_begin = std::chrono::steady_clock::now();
    const int64_t G = (int64_t)h->shards.size();
    const int64_t per = (n + G - 1) / G;     // contiguous ranges: shard g holds [g*per, ...)
    if (per > 0x7FFFFFFFLL - 4096) return fail(h, KGPU_ERR_INVALID, "kgpu_upload_nodes: more than 2^31 nodes per device");
    // A failure below leaves the handle EMPTY (n = 0 everywhere), never half old / half new.
    auto empty_all = [&]() {
        for (auto &s : h->shards) { s.n = 0; s.n_slots = 0; s.order_dirty = true; s.compact_dirty = true; s.fit_valid = false; s.work_P = -1; }
        h->n_total = 0;
    };
#define KGPU_UP(expr)                                                                                 \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            empty_all();                                                                              \
            return fail(h, e__ == cudaErrorMemoryAllocation ? KGPU_ERR_NOMEM : KGPU_ERR_CUDA,          \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);  \
        }                                                                                             \
    } while (0)
    int64_t off = 0;
    for (auto &s : h->shards) {
        const int64_t cnt = std::max<int64_t>(0, std::min<int64_t>(per, n - off));
        KGPU_UP(cudaSetDevice(s.dev));
        if (cnt > s.cap) {
            if (s.d_topo) cudaFree(s.d_topo);
            if (s.d_free) cudaFree(s.d_free);
            if (s.d_mem) cudaFree(s.d_mem);
            if (s.d_slot_of) cudaFree(s.d_slot_of);
            s.d_topo = nullptr; s.d_free = nullptr; s.d_mem = nullptr; s.d_slot_of = nullptr; s.cap = 0; s.n = 0;
            KGPU_UP(cudaMalloc(&s.d_topo, (size_t)cnt * 256));
            KGPU_UP(cudaMalloc(&s.d_free, (size_t)cnt * 4));
            KGPU_UP(cudaMalloc(&s.d_mem, (size_t)cnt * 32));
            KGPU_UP(cudaMalloc(&s.d_slot_of, (size_t)cnt * 4));
            s.cap = cnt;
        }
        if (cnt > 0) {
            // per-GPU memory is unconstrained (0x7F7F7F7F MiB) until kgpu_upload_gpu_memory says otherwise
            KGPU_UP(cudaMemsetAsync(s.d_mem, 0x7F, (size_t)cnt * 32, s.stream));
            KGPU_UP(cudaMemcpyAsync(s.d_topo, topo + off * 64, (size_t)cnt * 256, cudaMemcpyHostToDevice, s.stream));
            KGPU_UP(cudaMemcpyAsync(s.d_free, free_mask + off, (size_t)cnt * 4, cudaMemcpyHostToDevice, s.stream));
            // value-domain check on the device (levels 0..15): one 8-byte read-back instead of an O(64 N) host pass
            KGPU_UP(cudaMemsetAsync(s.d_bad, 0xFF, sizeof(unsigned long long), s.stream));
            kgpu::validate_topo_dev<<<(unsigned)((cnt * 16 + 255) / 256), 256, 0, s.stream>>>(
                reinterpret_cast<const int4 *>(s.d_topo), cnt * 16, s.d_bad);
            h->launches++;
            KGPU_UP(cudaGetLastError());
        }
        s.n = cnt;
        s.node_id_base = node_id_base + off;
        s.order_dirty = true;
        s.compact_dirty = true;
        s.fit_valid = false;
        off += cnt;
    }
    h->n_total = n;
    // all devices copy concurrently; then the verdicts, then the K1s order + compacted records (device side)
    off = 0;
    for (auto &s : h->shards) {
        if (s.n > 0) {
            unsigned long long bad = ~0ull;
            KGPU_UP(cudaSetDevice(s.dev));
            KGPU_UP(cudaMemcpyAsync(&bad, s.d_bad, sizeof bad, cudaMemcpyDeviceToHost, s.stream));
            KGPU_UP(cudaStreamSynchronize(s.stream));
            if (bad != ~0ull) {
                const long long el = (long long)bad;
                empty_all();
                return fail(h, KGPU_ERR_INVALID, "topo[%lld][%d] = %d outside the link-level domain 0..15",
                            (long long)(off + el / 64), (int)(el % 64), topo[(off + el / 64) * 64 + el % 64]);
            }
        }
        off += s.n;
    }
#undef KGPU_UP
    for (auto &s : h->shards) {
        const int rc = ensure_node_cache(h, s);
        if (rc != KGPU_OK) { empty_all(); return rc; }
    }
    h->last_upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
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
    KGPU_CUDA(h, cudaMemcpyAsync(s->d_topo + local * 64, topo, 256, cudaMemcpyHostToDevice, s->stream));
    const int32_t li = (int32_t)local;
    return apply_node_updates(h, *s, &li, &free_mask, 1);      // new mask + this node's record (and fit row) only
}

int kgpu_set_free_masks(kgpu_t *h, const int64_t *idx, const int32_t *free_mask, int64_t n) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_set_free_masks: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (n < 0 || (n > 0 && (!idx || !free_mask))) return fail(h, KGPU_ERR_INVALID, "kgpu_set_free_masks: bad arguments");
    for (int64_t i = 0; i < n; i++)
        if (idx[i] < 0 || idx[i] >= h->n_total)
            return fail(h, KGPU_ERR_INVALID, "kgpu_set_free_masks: index %lld out of range", (long long)idx[i]);
    // per shard: local indices, the LAST entry wins when a node is listed more than once
    int64_t off = 0;
    for (auto &s : h->shards) {
        std::vector<int32_t> li, lm;
        if (n == 1) {
            if (idx[0] >= off && idx[0] < off + s.n) { li.push_back((int32_t)(idx[0] - off)); lm.push_back(free_mask[0]); }
        } else {
            std::vector<std::pair<int32_t, int64_t>> ent;     // (local index, position in the call)
            for (int64_t i = 0; i < n; i++)
                if (idx[i] >= off && idx[i] < off + s.n) ent.emplace_back((int32_t)(idx[i] - off), i);
            std::sort(ent.begin(), ent.end());
            for (size_t e = 0; e < ent.size(); e++)
                if (e + 1 == ent.size() || ent[e + 1].first != ent[e].first) { li.push_back(ent[e].first); lm.push_back(free_mask[ent[e].second]); }
        }
        off += s.n;
        if (li.empty()) continue;
        const int rc = apply_node_updates(h, s, li.data(), lm.data(), (int64_t)li.size());
        if (rc != KGPU_OK) return rc;
    }
    return KGPU_OK;
}

int kgpu_set_free_mask(kgpu_t *h, int64_t idx, int32_t free_mask) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_set_free_mask: NULL handle");
    {
        std::lock_guard<std::mutex> g(h->mu);
        if (idx < 0 || idx >= h->n_total) return fail(h, KGPU_ERR_INVALID, "kgpu_set_free_mask: index %lld out of range", (long long)idx);
    }
    return kgpu_set_free_masks(h, &idx, &free_mask, 1);
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
    return kgpu_place_batch_ex(h, pods, P, out_keys, 0);
}

int kgpu_place_batch_ex(kgpu_t *h, const int32_t *pods, int64_t P, uint64_t *out_keys, int flags) {
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
    const bool dry = (flags & KGPU_PLACE_DRY_RUN) != 0;
    int32_t *d_free = s.d_free;
    if (dry && s.n > 0) {                  // place on a scratch copy of the masks: conflict-free proposals, nothing taken
        if (s.n > s.free_scratch_cap) {
            if (s.d_free_scratch) cudaFree(s.d_free_scratch);
            s.d_free_scratch = nullptr; s.free_scratch_cap = 0;
            KGPU_CUDA(h, cudaMalloc(&s.d_free_scratch, (size_t)s.n * 4));
            s.free_scratch_cap = s.n;
        }
        KGPU_CUDA(h, cudaMemcpyAsync(s.d_free_scratch, s.d_free, (size_t)s.n * 4, cudaMemcpyDeviceToDevice, s.stream));
        d_free = s.d_free_scratch;
    }
    if (s.n == 0) {
        KGPU_CUDA(h, cudaMemsetAsync(s.d_keys, 0xFF, (size_t)P * 8, s.stream));
    } else {
        const int64_t T = (s.n + kgpu::PLACE_TILE - 1) / kgpu::PLACE_TILE, Npad = T * kgpu::PLACE_TILE;
        if (Npad * views.n > s.place_cap) {
            if (s.d_nodebest) cudaFree(s.d_nodebest);
            if (s.d_tilebest) cudaFree(s.d_tilebest);
            s.d_nodebest = nullptr; s.d_tilebest = nullptr; s.place_cap = 0;
            KGPU_CUDA(h, cudaMalloc(&s.d_nodebest, (size_t)Npad * views.n * 9 * 4));
            if (s.d_half) cudaFree(s.d_half);
            s.d_half = nullptr;
            KGPU_CUDA(h, cudaMalloc(&s.d_half, (size_t)Npad * kgpu::PLACE_HALF * 4));
            KGPU_CUDA(h, cudaMalloc(&s.d_tilebest, (size_t)T * views.n * 9 * 8));
            s.place_cap = Npad * views.n;
        }
        kgpu::Weights W;
        memcpy(W.w, h->W, sizeof W.w);
        KGPU_CUDA(h, cudaEventRecord(s.ev0, s.stream));
        kgpu::place_init<<<dim3((unsigned)T, (unsigned)views.n), kgpu::PLACE_TILE, 0, s.stream>>>(
            reinterpret_cast<const int4 *>(s.d_topo), d_free, s.d_mem, s.n, Npad, s.node_id_base, W, PC, views, s.d_nodebest,
            s.d_tilebest, T);
        kgpu::place_half_tables<<<(unsigned)std::min<int64_t>((s.n + 3) / 4, (int64_t)s.sm_count * 16), 128, 0, s.stream>>>(
            s.d_topo, s.n, W, s.d_half);
        h->launches++;
        static const cudaError_t place_attr = cudaFuncSetAttribute(kgpu::place_sequential, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                   (int)kgpu::PLACE_DYN_SMEM);
        KGPU_CUDA(h, place_attr);
        kgpu::place_sequential<<<1, kgpu::PLACE_THREADS, kgpu::PLACE_DYN_SMEM, s.stream>>>(s.d_topo, d_free, s.d_mem, s.n, Npad, s.node_id_base,
                                                                        reinterpret_cast<const int4 *>(s.d_pods), P, W, views,
                                                                        s.d_nodebest, s.d_tilebest, T, s.d_half, s.d_keys);
        h->launches += 2;
        if (!dry) {
            s.order_dirty = true;        // the free masks have changed on the device: re-sort + recompact lazily
            s.fit_valid = false;
        }
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

int kgpu_build_fit_table(kgpu_t *h) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_build_fit_table: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    for (auto &s : h->shards) {
        const int rc = build_fit_table(h, s);
        if (rc != KGPU_OK) return rc;
    }
    return KGPU_OK;
}

int kgpu_fit_lookup(kgpu_t *h, int64_t node_idx, int32_t k, uint32_t *out_node_key) {
    if (!h || !out_node_key) return fail(h, KGPU_ERR_INVALID, "kgpu_fit_lookup: NULL argument");
    std::lock_guard<std::mutex> g(h->mu);
    if (node_idx < 0 || node_idx >= h->n_total) return fail(h, KGPU_ERR_INVALID, "kgpu_fit_lookup: node index %lld out of range", (long long)node_idx);
    int64_t local = 0;
    kgpu_shard *s = shard_of(h, node_idx, &local);
    if (!s->fit_valid) {                          // built lazily, kept current by the state-change calls
        const int rc = build_fit_table(h, *s);
        if (rc != KGPU_OK) return rc;
    }
    *out_node_key = (k < 0 || k > 8) ? UINT32_MAX : s->h_fit[(int64_t)k * s->n + local];
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
    // Pairs without a memory requirement are answered from the (node, k) fit table's host copy: no launch,
    // no copy per call (PodFitsDevice is called once per (node, pod) by the core: gpu_scheduler.go:34-44).
    // The others go to the device in one packed copy, routed to the shard that holds their node.
    kgpu::Weights W;
    memcpy(W.w, h->W, sizeof W.w);
    int64_t off = 0;
    for (auto &s : h->shards) {
        std::vector<int4> q;
        std::vector<int64_t> pos;
        for (int64_t i = 0; i < n; i++) {
            if (node_idx[i] < off || node_idx[i] >= off + s.n) continue;
            const int64_t local = node_idx[i] - off;
            if (!min_mem_mib || min_mem_mib[i] <= 0) {
                if (!s.fit_valid) {
                    const int rc = build_fit_table(h, s);
                    if (rc != KGPU_OK) return rc;
                }
                out_node_keys[i] = (k[i] < 0 || k[i] > 8) ? UINT32_MAX : s.h_fit[(int64_t)k[i] * s.n + local];
            } else {
                q.push_back(make_int4((int)(local & 0xFFFFFFFFLL), (int)(local >> 32), k[i], min_mem_mib[i]));
                pos.push_back(i);
            }
        }
        off += s.n;
        if (q.empty()) continue;
        const int64_t m = (int64_t)q.size();
        KGPU_CUDA(h, cudaSetDevice(s.dev));
        if (m > s.query_cap) {
            if (s.d_free_scratch) cudaFree(s.d_free_scratch);
    if (s.d_query) cudaFree(s.d_query);
            if (s.d_qout) cudaFree(s.d_qout);
            s.d_query = nullptr; s.d_qout = nullptr; s.query_cap = 0;
            const int64_t cap = std::max<int64_t>(256, m + m / 2);
            KGPU_CUDA(h, cudaMalloc(&s.d_query, (size_t)cap * 16));
            KGPU_CUDA(h, cudaMalloc(&s.d_qout, (size_t)cap * 4));
            s.query_cap = cap;
        }
        KGPU_CUDA(h, cudaMemcpyAsync(s.d_query, q.data(), (size_t)m * 16, cudaMemcpyHostToDevice, s.stream));
        kgpu::score_pair_list<<<(unsigned)((m + 127) / 128), 128, 0, s.stream>>>(
            reinterpret_cast<const int4 *>(s.d_topo), s.d_free, s.d_mem, s.n, s.d_query, m, W, PC, s.d_qout);
        h->launches++;
        KGPU_CUDA(h, cudaGetLastError());
        std::vector<uint32_t> res((size_t)m);
        KGPU_CUDA(h, cudaMemcpyAsync(res.data(), s.d_qout, (size_t)m * 4, cudaMemcpyDeviceToHost, s.stream));
        KGPU_CUDA(h, cudaStreamSynchronize(s.stream));
        for (int64_t j = 0; j < m; j++) out_node_keys[pos[(size_t)j]] = res[(size_t)j];
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
// one allocation per rank, mapped by every peer: slots[2][world][max_pods] uint64 | flags[PEER_MAX_WORLD] uint32 | ticket | error
size_t xch_slot_bytes(int world, int64_t max_pods) { return (size_t)2 * world * max_pods * 8; }
size_t xch_bytes(int world, int64_t max_pods) { return xch_slot_bytes(world, max_pods) + kgpu::PEER_MAX_WORLD * 4 + 16; }
unsigned long long *xch_slots(void *base, int world, int64_t max_pods, int buf) {
    return reinterpret_cast<unsigned long long *>(base) + (int64_t)buf * world * max_pods;
}
uint32_t *xch_flags(void *base, int world, int64_t max_pods) {
    return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(base) + xch_slot_bytes(world, max_pods));
}
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
    KGPU_CUDA(h, cudaMalloc(&h->xch.base, xch_bytes(world, max_pods)));
    KGPU_CUDA(h, cudaMalloc(&h->xch.local, (size_t)max_pods * 8));
    KGPU_CUDA(h, cudaMalloc(&h->xch.final_keys, (size_t)max_pods * 8 * 2));
    KGPU_CUDA(h, cudaMemset(h->xch.base, 0xFF, xch_slot_bytes(world, max_pods)));
    KGPU_CUDA(h, cudaMemset(h->xch.local, 0xFF, (size_t)max_pods * 8));                                   // the exchange kernel keeps it that way
    KGPU_CUDA(h, cudaMemset(xch_flags(h->xch.base, world, max_pods), 0, kgpu::PEER_MAX_WORLD * 4 + 16));  // flags, ticket, error
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

namespace {
// K1 (P > 0) + the exchange kernel on `st`; *d_final (may be NULL) receives where the global keys will be
int exchange_step(kgpu_ctx *h, const int32_t *d_pods, int64_t P, const uint64_t **d_final_keys, cudaStream_t st, int batch_flags) {
    kgpu_shard &s = h->shards[0];
    KGPU_CUDA(h, cudaSetDevice(s.dev));
    const int world = h->xch.world;
    const int64_t mp = h->xch.max_pods;
    int *d_err = reinterpret_cast<int *>(xch_flags(h->xch.base, world, mp) + kgpu::PEER_MAX_WORLD + 1);
    if (h->xch.epoch > 0 && (h->xch.epoch & 1023u) == 0) {     // now and then: did a wait time out?  (costs a sync, so rarely)
        int err = 0;
        KGPU_CUDA(h, cudaMemcpyAsync(&err, d_err, sizeof err, cudaMemcpyDeviceToHost, st));
        KGPU_CUDA(h, cudaStreamSynchronize(st));
        if (err) return fail(h, KGPU_ERR_COMM, "key exchange: a rank waited more than ~10 s for its peers (ranks out of step or a peer died)");
    }
    const uint32_t epoch = ++h->xch.epoch;
    const int buf = (int)(epoch & 1u);
    unsigned long long *fin = h->xch.final_keys + (int64_t)buf * mp;
    if (d_final_keys) *d_final_keys = reinterpret_cast<const uint64_t *>(fin);
    if (P > 0) {
        // K1 accumulates into xch.local (kept at NO_FIT between steps by the exchange kernel)
        const int rc = launch_score(h, s, d_pods, P, h->xch.local, st, (batch_flags & KGPU_BATCH_NO_MIN_MEM) ? 0 : -1, true);
        if (rc != KGPU_OK) return rc;
    }
    kgpu::PeerTable tab;
    memset(&tab, 0, sizeof tab);
    for (int r = 0; r < world; r++) {
        tab.slots[r] = xch_slots(h->xch.peer_base[r], world, mp, buf);
        tab.flags[r] = xch_flags(h->xch.peer_base[r], world, mp);
    }
    unsigned int *ticket = reinterpret_cast<unsigned int *>(xch_flags(h->xch.base, world, mp) + kgpu::PEER_MAX_WORLD);
    kgpu::gather_and_min<<<(unsigned)std::max<int64_t>(1, (P + 255) / 256), 256, 0, st>>>(h->xch.local, P, mp, tab, h->xch.rank, world, epoch, ticket,
                                                                                        fin, d_err);
    h->launches++;
    KGPU_CUDA(h, cudaGetLastError());
    return KGPU_OK;
}
}  // namespace

int kgpu_score_batch_exchange(kgpu_t *h, const int32_t *d_pods, int64_t P, const uint64_t **d_final_keys, void *stream,
                              int batch_flags) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch_exchange: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->xch.connected) return fail(h, KGPU_ERR_STATE, "kgpu_score_batch_exchange: exchange not connected");
    if (P < 0 || P > h->xch.max_pods || (P > 0 && !d_pods) || !d_final_keys)
        return fail(h, KGPU_ERR_INVALID, "kgpu_score_batch_exchange: bad arguments (P <= max_pods of kgpu_exchange_init)");
    return exchange_step(h, d_pods, P, d_final_keys, (cudaStream_t)stream, batch_flags);
}

int kgpu_exchange_barrier(kgpu_t *h, void *stream) {
    if (!h) return fail(h, KGPU_ERR_INVALID, "kgpu_exchange_barrier: NULL handle");
    std::lock_guard<std::mutex> g(h->mu);
    if (!h->xch.connected) return fail(h, KGPU_ERR_STATE, "kgpu_exchange_barrier: exchange not connected");
    return exchange_step(h, nullptr, 0, nullptr, (cudaStream_t)stream, 0);
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

double kgpu_last_upload_ms(kgpu_t *h) {
    if (!h) return 0.0;
    std::lock_guard<std::mutex> g(h->mu);
    return h->last_upload_ms;
}

double kgpu_last_kernel_ms(kgpu_t *h) {
    if (!h) return 0.0;
    std::lock_guard<std::mutex> g(h->mu);
    return h->last_kernel_ms;
}

}  // extern "C"
