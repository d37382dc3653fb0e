// emu_kernels.cc -- runs the real kernel sources (kubegpu_b200/csrc/*.cuh) on the CPU through cuda_emu.h.
// TEST INFRASTRUCTURE ONLY.  The small host plans below restate what kgpu.cu does around the launches
// (node order by free count, pod splits); the kernels themselves are compiled from the product sources.
#include "cuda_emu.h"

#include <algorithm>
#include <cstdlib>

#include "score_pairs.cuh"
#include "score_pairs_sparse.cuh"
#include "place_sequential.cuh"
#include "sparse_work.h"
#include "peer_exchange.cuh"
#include "node_state.cuh"


namespace {

const kgpu::PipeConsts kPC = {1u, 0xFFFFFFFFu};

kgpu::Weights weights_of(const int32_t *W) {
    kgpu::Weights w;
    std::memcpy(w.w, W, sizeof w.w);
    return w;
}

int per_split(int64_t P, int splits) {
    int64_t per = (P + splits - 1) / splits;
    return (int)((per + 31) / 32 * 32);
}

template <class T>
T *aligned_array(size_t n) {
    void *p = nullptr;
    if (posix_memalign(&p, 64, std::max<size_t>(64, n * sizeof(T))) != 0) abort();
    std::memset(p, 0, std::max<size_t>(64, n * sizeof(T)));
    return (T *)p;
}

}  // namespace

extern "C" {

// The K1s node cache as kgpu.cu builds it (build_order + ensure_node_cache): device-side counting sort of the
// nodes by free count, then the compacted slot-ordered records.
struct EmuCache {
    std::vector<int32_t> order, slot_of;
    int4 *rec = nullptr;
    uint32_t *meta = nullptr;
    int64_t n_slots = 0;
    long long class_count[9] = {0};
    ~EmuCache() { free(rec); free(meta); }
};

void emu_build_order(const int32_t *free_mask, int64_t n, EmuCache &c) {
    const int nb = (int)((n + kgpu::ORD_BLOCK - 1) / kgpu::ORD_BLOCK);
    const int64_t cap = (n + 11 * kgpu::SP_THREADS) / kgpu::SP_THREADS * kgpu::SP_THREADS;
    c.order.assign((size_t)cap, -1);
    c.slot_of.assign((size_t)std::max<int64_t>(1, n), -1);
    std::vector<int32_t> cnt((size_t)nb * 9), off((size_t)nb * 9);
    long long meta[kgpu::ORD_META] = {0};
    emu::launch(dim3((unsigned)nb), dim3(kgpu::ORD_BLOCK), [&] { kgpu::order_count(free_mask, n, cnt.data(), nb); });
    emu::launch(dim3(1), dim3(kgpu::ORD_BLOCK), [&] { kgpu::order_scan(cnt.data(), nb, off.data(), meta, kgpu::SP_THREADS, 32); });
    emu::launch(dim3((unsigned)nb), dim3(kgpu::ORD_BLOCK),
                [&] { kgpu::order_scatter(free_mask, n, off.data(), nb, c.order.data(), c.slot_of.data()); });
    for (int k = 0; k < 9; k++) c.class_count[k] = meta[k];
    c.n_slots = meta[9];
}

void emu_build_cache(const int4 *topo4, int32_t *free_mask, int64_t n, const kgpu::Weights &Ws, EmuCache &c) {
    emu_build_order(free_mask, n, c);
    c.rec = aligned_array<int4>((size_t)c.order.size() * 7);
    c.meta = aligned_array<uint32_t>(c.order.size());
    emu::launch(dim3((unsigned)(c.n_slots / kgpu::SP_THREADS)), dim3(kgpu::SP_THREADS),
                [&] { kgpu::compact_nodes(topo4, free_mask, c.n_slots, Ws, c.order.data(), c.slot_of.data(), c.rec, c.meta); });
}

void emu_launch_sparse(const EmuCache &c, const int32_t *free_mask, const int32_t *mem, int64_t node_id_base, const int4 *pods4,
                       const int32_t *pods, int64_t P, const int32_t *W, int splits, unsigned long long *keys) {
    int flag = 0;
    for (int64_t p = 0; p < P; p++) flag |= pods[4 * p + 3] > 0;
    dim3 grid((unsigned)(c.n_slots / kgpu::SP_THREADS), (unsigned)std::max(1, splits));
    const int per = per_split(P, std::max(1, splits));
    // splits < 0: the work list of sparse_work.h for -splits resident blocks, as kgpu.cu builds it
    std::vector<kgpu::SparseWorkItem> items;
    const int4 *work = nullptr;
    if (splits < 0) {
        std::vector<uint8_t> tile_class((size_t)(c.n_slots / kgpu::SP_THREADS), 0);
        for (int64_t sl = 0; sl < c.n_slots; sl++)
            if (c.order[(size_t)sl] >= 0)
                tile_class[(size_t)(sl / kgpu::SP_THREADS)] = std::max<uint8_t>(tile_class[(size_t)(sl / kgpu::SP_THREADS)],
                                                                            (uint8_t)__builtin_popcount((unsigned)free_mask[c.order[(size_t)sl]] & 0xFFu));
        kgpu::build_sparse_work(tile_class, P, -splits, items);
        static_assert(sizeof(kgpu::SparseWorkItem) == sizeof(int4), "work item layout");
        work = reinterpret_cast<const int4 *>(items.data());
        grid = dim3((unsigned)items.size(), 1);
    }
    bool byte_keys = true;                               // kgpu.cu: every cost < 2^16
    for (int i = 0; i < 16; i++) byte_keys = byte_keys && W[i] <= 2340;
#define EMU_SPARSE(MEMF, BK, ST)                                                                               \
    emu::launch(grid, dim3(kgpu::SP_THREADS), [&] {                                                            \
        kgpu::score_pairs_sparse<true, MEMF, BK, ST>(c.rec, c.meta, mem, c.order.data(), &flag, node_id_base, pods4, P, per, work, kPC, keys); \
    })
    const bool stream_build = splits < 0 && P <= kgpu::kSparseChunk;      // kgpu.cu: runs of tiles -> the STREAM instantiation
    const bool tma_build = stream_build && P <= kgpu::SP_TMA_PODS;        //          at most 64 pods -> the TMA one
    if (tma_build) {
        if (byte_keys) emu::launch(grid, dim3(kgpu::SP_THREADS), [&] { kgpu::score_pairs_sparse<true, false, true, true, 8>(c.rec, c.meta, mem, c.order.data(), &flag, node_id_base, pods4, P, per, work, kPC, keys); });
        else emu::launch(grid, dim3(kgpu::SP_THREADS), [&] { kgpu::score_pairs_sparse<true, false, false, true, 8>(c.rec, c.meta, mem, c.order.data(), &flag, node_id_base, pods4, P, per, work, kPC, keys); });
    } else if (stream_build) { if (byte_keys) EMU_SPARSE(false, true, true); else EMU_SPARSE(false, false, true); }
    else              { if (byte_keys) EMU_SPARSE(false, true, false); else EMU_SPARSE(false, false, false); }
    if (flag) { if (byte_keys) EMU_SPARSE(true, true, false); else EMU_SPARSE(true, false, false); }
#undef EMU_SPARSE
}

// K1s (+ its MEM instantiation) exactly as kgpu.cu launches them: device-built order, compact cache, grid.
void emu_score_sparse(const int32_t *topo, const int32_t *free_mask_in, const int32_t *gpu_mem /*nullable*/, int64_t n,
                      int64_t node_id_base, const int32_t *pods, int64_t P, const int32_t *W, int splits,
                      unsigned long long *keys) {
    std::memset(keys, 0xFF, (size_t)P * 8);
    if (n == 0 || P == 0) return;
    std::vector<int32_t> free_mask(free_mask_in, free_mask_in + n);
    int4 *topo4 = aligned_array<int4>((size_t)n * 16);
    std::memcpy(topo4, topo, (size_t)n * 256);
    int4 *pods4 = aligned_array<int4>((size_t)P);
    std::memcpy(pods4, pods, (size_t)P * 16);
    int32_t *mem = aligned_array<int32_t>((size_t)n * 8);
    if (gpu_mem) std::memcpy(mem, gpu_mem, (size_t)n * 32); else std::memset(mem, 0x7F, (size_t)n * 32);
    const kgpu::Weights Ws = weights_of(W);
    EmuCache c;
    emu_build_cache(topo4, free_mask.data(), n, Ws, c);
    emu_launch_sparse(c, free_mask.data(), mem, node_id_base, pods4, pods, P, W, splits, keys);
    free(topo4); free(pods4); free(mem);
}

// kgpu_upload_nodes + kgpu_set_free_masks + kgpu_score_batch: the cache is built for free_mask0, then the m listed
// nodes get new masks through the incremental path (compact_nodes with a list; the order stays as it was, i.e.
// stale), then K1s runs.  Must equal a fresh upload with the final masks.
void emu_score_sparse_after_updates(const int32_t *topo, const int32_t *free_mask0, int64_t n, const int32_t *upd_idx,
                                    const int32_t *upd_mask, int64_t m, const int32_t *pods, int64_t P, const int32_t *W,
                                    int splits, unsigned long long *keys, int32_t *free_mask_out) {
    std::memset(keys, 0xFF, (size_t)P * 8);
    std::vector<int32_t> free_mask(free_mask0, free_mask0 + n);
    int4 *topo4 = aligned_array<int4>((size_t)n * 16);
    std::memcpy(topo4, topo, (size_t)n * 256);
    int4 *pods4 = aligned_array<int4>((size_t)P);
    std::memcpy(pods4, pods, (size_t)P * 16);
    int32_t *mem = aligned_array<int32_t>((size_t)n * 8);
    std::memset(mem, 0x7F, (size_t)n * 32);
    const kgpu::Weights Ws = weights_of(W);
    EmuCache c;
    emu_build_cache(topo4, free_mask.data(), n, Ws, c);
    if (m > 0)
        emu::launch(dim3((unsigned)((m + kgpu::SP_THREADS - 1) / kgpu::SP_THREADS)), dim3(kgpu::SP_THREADS), [&] {
            kgpu::compact_nodes(topo4, free_mask.data(), m, Ws, c.order.data(), c.slot_of.data(), c.rec, c.meta, upd_idx, upd_mask);
        });
    emu_launch_sparse(c, free_mask.data(), mem, 0, pods4, pods, P, W, splits, keys);
    std::memcpy(free_mask_out, free_mask.data(), (size_t)n * 4);
    free(topo4); free(pods4); free(mem);
}

// The device-built order alone: order[cap] (cap = n + 9*32 + 2*tile rounded down to tiles), slot_of[n], class counts.
int64_t emu_order(const int32_t *free_mask, int64_t n, int32_t *order_out, int64_t cap, int32_t *slot_of_out, long long *class_count) {
    EmuCache c;
    emu_build_order(free_mask, n, c);
    for (int64_t i = 0; i < cap; i++) order_out[i] = i < (int64_t)c.order.size() ? c.order[(size_t)i] : -1;
    std::memcpy(slot_of_out, c.slot_of.data(), (size_t)n * 4);
    for (int k = 0; k < 9; k++) class_count[k] = c.class_count[k];
    return c.n_slots;
}

// fit_nodes (the (node, k) fit table), optionally over a list; out[k * m + i]
void emu_fit_nodes(const int32_t *topo, const int32_t *free_mask, int64_t n, const int32_t *list /*nullable*/, int64_t m,
                   const int32_t *W, uint32_t *out) {
    int4 *topo4 = aligned_array<int4>((size_t)n * 16);
    std::memcpy(topo4, topo, (size_t)n * 256);
    const kgpu::Weights Ws = weights_of(W);
    emu::launch(dim3((unsigned)((m + 127) / 128)), dim3(128), [&] { kgpu::fit_nodes(topo4, free_mask, list, m, Ws, kPC, out); });
    free(topo4);
}

// validate_topo_dev: index of the first value outside 0..15, or -1
long long emu_validate_topo(const int32_t *topo, int64_t n) {
    int4 *topo4 = aligned_array<int4>((size_t)n * 16);
    std::memcpy(topo4, topo, (size_t)n * 256);
    unsigned long long bad = ~0ull;
    emu::launch(dim3((unsigned)((n * 16 + 255) / 256)), dim3(256), [&] { kgpu::validate_topo_dev(topo4, n * 16, &bad); });
    free(topo4);
    return bad == ~0ull ? -1 : (long long)bad;
}

// The peer-exchange kernel with world = 1 (the rank stores into its own slot array and passes its own barrier):
// exercises the stores, the ticket, the flag protocol and the final minimum, not the cross-GPU part.
void emu_gather_and_min(unsigned long long *local, int64_t P, int64_t max_pods, unsigned long long *slots, uint32_t *flags,
                        uint32_t epoch, unsigned int *ticket, unsigned long long *final_keys, int *error) {
    kgpu::PeerTable tab;
    std::memset(&tab, 0, sizeof tab);
    tab.slots[0] = slots;
    tab.flags[0] = flags;
    emu::launch(dim3((unsigned)std::max<int64_t>(1, (P + 255) / 256)), dim3(256),
                [&] { kgpu::gather_and_min(local, P, max_pods, tab, 0, 1, epoch, ticket, final_keys, error); });
}

// The host-side work-list builder alone (sparse_work.h): items as int32[.][4] {tile, pod_begin, pod_end, ntiles} and
// their model weights, returns the count.
int64_t emu_sparse_work(const uint8_t *tile_class, int64_t tiles, int64_t P, int64_t resident, int32_t *out, int64_t cap,
                        long long *weight_out) {
    std::vector<kgpu::SparseWorkItem> items;
    std::vector<int64_t> weight;
    kgpu::build_sparse_work(std::vector<uint8_t>(tile_class, tile_class + tiles), P, resident, items, kgpu::SparseWorkParams(), &weight);
    for (size_t i = 0; i < items.size() && (int64_t)i < cap; i++) {
        std::memcpy(out + 4 * i, &items[i], 16);
        weight_out[i] = weight[i];
    }
    return (int64_t)items.size();
}

// Dense K1 (+ K1m) as kgpu.cu launches them.
void emu_score_dense(const int32_t *topo, const int32_t *free_mask, const int32_t *gpu_mem /*nullable*/, int64_t n,
                     int64_t node_id_base, const int32_t *pods, int64_t P, const int32_t *W, int splits,
                     unsigned long long *keys) {
    std::memset(keys, 0xFF, (size_t)P * 8);
    if (n == 0 || P == 0) return;
    int4 *topo4 = aligned_array<int4>((size_t)n * 16);
    std::memcpy(topo4, topo, (size_t)n * 256);
    int4 *pods4 = aligned_array<int4>((size_t)P);
    std::memcpy(pods4, pods, (size_t)P * 16);
    int4 *mem4 = aligned_array<int4>((size_t)n * 2);
    if (gpu_mem) std::memcpy(mem4, gpu_mem, (size_t)n * 32); else std::memset(mem4, 0x7F, (size_t)n * 32);
    const kgpu::Weights Ws = weights_of(W);
    int flag = 0;
    for (int64_t p = 0; p < P; p++) flag |= pods[4 * p + 3] > 0;
    const dim3 grid((unsigned)((n + kgpu::LPN_THREADS - 1) / kgpu::LPN_THREADS), (unsigned)std::max(1, splits));
    const int per = per_split(P, std::max(1, splits));
    emu::launch(grid, dim3(kgpu::LPN_THREADS), [&] {
        kgpu::score_pairs_lane_per_node<true, false>(topo4, free_mask, mem4, &flag, n, node_id_base, pods4, P, per, Ws, kPC, keys);
    });
    if (flag)
        emu::launch(grid, dim3(kgpu::LPN_THREADS), [&] {
            kgpu::score_pairs_lane_per_node<true, true>(topo4, free_mask, mem4, &flag, n, node_id_base, pods4, P, per, Ws, kPC, keys);
        });
    free(topo4); free(pods4); free(mem4);
}

// The other K1 variants as kgpu.cu launches them: 1 = warp per pair (north_star mapping), 3 = memo by k,
// 4 = tile memo.  Memory-constrained pods go to K1m for variants 3 and 4, inline for variant 1.
void emu_score_variant(int variant, const int32_t *topo, const int32_t *free_mask, const int32_t *gpu_mem, int64_t n,
                       int64_t node_id_base, const int32_t *pods, int64_t P, const int32_t *W, int splits,
                       unsigned long long *keys) {
    std::memset(keys, 0xFF, (size_t)P * 8);
    if (n == 0 || P == 0) return;
    for (int k = 0; k <= 8; k++) {                       // what upload_subset_tables() puts in constant memory
        int c = 0;
        for (unsigned S = 0; S < 256; S++)
            if (__builtin_popcount(S) == k) kgpu::c_subsets[k][c++] = (uint8_t)S;
        kgpu::c_nsub[k] = (uint8_t)c;
    }
    int4 *topo4 = aligned_array<int4>((size_t)n * 16);
    std::memcpy(topo4, topo, (size_t)n * 256);
    int4 *pods4 = aligned_array<int4>((size_t)P);
    std::memcpy(pods4, pods, (size_t)P * 16);
    int32_t *mem = aligned_array<int32_t>((size_t)n * 8);
    if (gpu_mem) std::memcpy(mem, gpu_mem, (size_t)n * 32); else std::memset(mem, 0x7F, (size_t)n * 32);
    const int4 *mem4 = reinterpret_cast<const int4 *>(mem);
    const kgpu::Weights Ws = weights_of(W);
    int flag = 0;
    emu::launch(dim3((unsigned)((P + 255) / 256)), dim3(256), [&] { kgpu::any_mem_pod(pods4, P, &flag); });
    const int per = per_split(P, std::max(1, splits));
    if (variant == 1) {
        const dim3 grid((unsigned)((n + kgpu::WPP_TILE - 1) / kgpu::WPP_TILE), (unsigned)std::max(1, splits));
        emu::launch(grid, dim3(kgpu::WPP_THREADS), [&] {
            kgpu::score_pairs_warp_per_pair(topo4, free_mask, mem, n, node_id_base, pods4, P, per, Ws, kPC, keys);
        });
    } else {
        const dim3 grid((unsigned)((n + kgpu::LPN_THREADS - 1) / kgpu::LPN_THREADS), (unsigned)std::max(1, splits));
        if (variant == 3) {
            unsigned long long bestk[9];
            std::memset(bestk, 0xFF, sizeof bestk);
            emu::launch(dim3(std::min<unsigned>(grid.x, 3)), dim3(kgpu::LPN_THREADS),
                        [&] { kgpu::memo_best_by_k(topo4, free_mask, n, node_id_base, Ws, kPC, bestk); });
            emu::launch(dim3((unsigned)((P + 255) / 256)), dim3(256), [&] { kgpu::memo_gather(pods4, P, bestk, keys); });
        } else {
            emu::launch(grid, dim3(kgpu::LPN_THREADS), [&] {
                kgpu::score_pairs_lane_per_node<false, false>(topo4, free_mask, mem4, &flag, n, node_id_base, pods4, P, per, Ws, kPC, keys);
            });
        }
        if (flag)
            emu::launch(grid, dim3(kgpu::LPN_THREADS), [&] {
                kgpu::score_pairs_lane_per_node<true, true>(topo4, free_mask, mem4, &flag, n, node_id_base, pods4, P, per, Ws, kPC, keys);
            });
    }
    free(topo4); free(pods4); free(mem);
}

// K2
void emu_reduce_shards(const unsigned long long *gathered, int G, int64_t P, unsigned long long *out) {
    emu::launch(dim3((unsigned)((P + 255) / 256)), dim3(256), [&] { kgpu::reduce_shards(gathered, G, P, out); });
}

// kgpu_score_pairs (the min_mem path: packed queries)
void emu_score_pair_list(const int32_t *topo, const int32_t *free_mask, const int32_t *gpu_mem, int64_t n,
                         const long long *node_idx, const int32_t *ks, const int32_t *min_mem, int64_t m, const int32_t *W,
                         uint32_t *out) {
    int4 *topo4 = aligned_array<int4>((size_t)n * 16);
    std::memcpy(topo4, topo, (size_t)n * 256);
    int32_t *mem = aligned_array<int32_t>((size_t)n * 8);
    if (gpu_mem) std::memcpy(mem, gpu_mem, (size_t)n * 32); else std::memset(mem, 0x7F, (size_t)n * 32);
    int4 *q = aligned_array<int4>((size_t)m);
    for (int64_t i = 0; i < m; i++)
        q[i] = make_int4((int)(node_idx[i] & 0xFFFFFFFFLL), (int)(node_idx[i] >> 32), ks[i], min_mem ? min_mem[i] : 0);
    const kgpu::Weights Ws = weights_of(W);
    emu::launch(dim3((unsigned)((m + 127) / 128)), dim3(128),
                [&] { kgpu::score_pair_list(topo4, free_mask, mem, n, q, m, Ws, kPC, out); });
    free(topo4); free(mem); free(q);
}

// K3: place_init + place_sequential as kgpu_place_batch launches them (views from the batch's distinct
// min_mem values); free_mask is updated in place like the device copy.  Returns 0, or -1 for too many views.
int emu_place_batch(const int32_t *topo, int32_t *free_mask, const int32_t *gpu_mem /*nullable*/, int64_t n,
                    int64_t node_id_base, const int32_t *pods, int64_t P, const int32_t *W, unsigned long long *keys) {
    std::memset(keys, 0xFF, (size_t)P * 8);
    if (n == 0 || P == 0) return 0;
    kgpu::PlaceViews views;
    std::memset(&views, 0, sizeof views);
    views.n = 1;
    for (int64_t p = 0; p < P; p++) {
        const int32_t need = pods[4 * p + 3];
        if (need <= 0 || pods[4 * p] < 0 || pods[4 * p] > 8) continue;
        bool seen = false;
        for (int j = 1; j < views.n; j++) seen = seen || views.min_mem[j] == need;
        if (seen) continue;
        if (views.n == kgpu::PLACE_MAX_VIEWS) return -1;
        views.min_mem[views.n++] = need;
    }
    for (int k = 0; k <= 8; k++) {
        int c = 0;
        for (unsigned S = 0; S < 256; S++)
            if (__builtin_popcount(S) == k) kgpu::c_subsets[k][c++] = (uint8_t)S;
        kgpu::c_nsub[k] = (uint8_t)c;
    }
    int4 *topo4 = aligned_array<int4>((size_t)n * 16);
    std::memcpy(topo4, topo, (size_t)n * 256);
    int4 *pods4 = aligned_array<int4>((size_t)P);
    std::memcpy(pods4, pods, (size_t)P * 16);
    int32_t *mem = aligned_array<int32_t>((size_t)n * 8);
    if (gpu_mem) std::memcpy(mem, gpu_mem, (size_t)n * 32); else std::memset(mem, 0x7F, (size_t)n * 32);
    const int64_t T = (n + kgpu::PLACE_TILE - 1) / kgpu::PLACE_TILE, Npad = T * kgpu::PLACE_TILE;
    uint32_t *nodebest = aligned_array<uint32_t>((size_t)Npad * 9 * views.n);
    unsigned long long *tilebest = aligned_array<unsigned long long>((size_t)T * 9 * views.n);
    const kgpu::Weights Ws = weights_of(W);
    emu::launch(dim3((unsigned)T, (unsigned)views.n), dim3(kgpu::PLACE_TILE),
                [&] { kgpu::place_init(topo4, free_mask, mem, n, Npad, node_id_base, Ws, kPC, views, nodebest, tilebest, T); });
    int32_t *half_all = aligned_array<int32_t>((size_t)Npad * kgpu::PLACE_HALF);
    emu::launch(dim3((unsigned)std::min<int64_t>((n + 3) / 4, 8)), dim3(128),
                [&] { kgpu::place_half_tables(reinterpret_cast<const int32_t *>(topo4), n, Ws, half_all); });
    emu::launch(dim3(1), dim3(kgpu::PLACE_THREADS), [&] {
        kgpu::place_sequential(reinterpret_cast<const int32_t *>(topo4), free_mask, mem, n, Npad, node_id_base, pods4, P, Ws, views,
                               nodebest, tilebest, T, half_all, keys);
    });
    free(topo4); free(pods4); free(mem); free(nodebest); free(tilebest); free(half_all);
    return 0;
}

}  // extern "C"
