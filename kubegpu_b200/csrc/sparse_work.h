// sparse_work.h -- host side of K1s: the work list.  Plain C++ (no CUDA), shared by kgpu.cu and the
// CPU emulation driver of the tests.
//
// A K1s block stages one 128-slot tile of the node order and scores a RANGE of the batch's pods against it.
// What a pod costs depends on the tile's class F (the largest number of free GPUs among its nodes): the
// bucket loops enumerate C(F,k) subsets, about 1 warp instruction per pod for F = 1 and about 110 for F = 8
// (scripts/k1s_issue_model.py).  Equal pod ranges for every tile therefore give blocks whose run times differ
// by two orders of magnitude, and with few tiles (a shard of a multi-GPU run) either too few blocks or
// ranges so short that staging and the per-chunk pod sort dominate.  The work list cuts every tile's pods
// into ranges of about equal WORK instead, heaviest items first.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace kgpu {

struct SparseWorkItem {
    int32_t tile, pod_begin, pod_end, weight;     // 16 bytes: one LDG.128 per block
};

// warp instructions per pod for a tile of class F (k uniform over 1..8, byte-key build) + ~2 for the
// per-chunk pod sort and the block flush; per-block fixed cost (block launch, staging, first sort): ~600
constexpr int32_t kSparsePodCost[9] = {2, 3, 4, 6, 9, 16, 31, 58, 114};
constexpr int64_t kSparseFixedCost = 600;

struct SparseWorkParams {
    int64_t waves = 3;          // big items per resident block
    int64_t floor = 8;          // an item carries at least floor * kSparseFixedCost of work (<= 1/floor overhead)
    int64_t tail_percent = 25;  // the last quarter of every tile's pods ...
    int64_t tail_div = 4;       // ... goes into ranges a quarter as long: they fill the end of the launch
};

// Every tile's pods cut into ranges of about `target` work (heaviest items first; the grid runs them in
// index order), the tail of every tile in shorter ranges.  Ranges are multiples of 32 pods.
inline void build_sparse_work(const std::vector<uint8_t> &tile_class, int64_t P, int64_t resident_blocks,
                              std::vector<SparseWorkItem> &out, const SparseWorkParams prm = SparseWorkParams()) {
    out.clear();
    if (P <= 0 || tile_class.empty()) return;
    int64_t total = 0;
    for (uint8_t f : tile_class) total += kSparseFixedCost + P * kSparsePodCost[std::min<int>(f, 8)];
    const int64_t target = std::max<int64_t>(total / std::max<int64_t>(1, prm.waves * resident_blocks), prm.floor * kSparseFixedCost);
    for (size_t t = 0; t < tile_class.size(); t++) {
        const int64_t c = kSparsePodCost[std::min<int>(tile_class[t], 8)];
        const int64_t per = std::max<int64_t>(32, (target - kSparseFixedCost) / c / 32 * 32);
        const int64_t small = std::max<int64_t>(32, per / std::max<int64_t>(1, prm.tail_div) / 32 * 32);
        const int64_t big_end = P * (100 - prm.tail_percent) / 100 / per * per;
        int64_t b = 0;
        while (b < P) {
            const int64_t e = std::min(P, b + (b < big_end ? per : small));
            out.push_back(SparseWorkItem{(int32_t)t, (int32_t)b, (int32_t)e, (int32_t)std::min<int64_t>(0x7FFFFFFF, (e - b) * c)});
            b = e;
        }
    }
    std::stable_sort(out.begin(), out.end(), [](const SparseWorkItem &a, const SparseWorkItem &b) { return a.weight > b.weight; });
}

}  // namespace kgpu
