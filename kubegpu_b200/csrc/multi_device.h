// multi_device.h -- single-process multi-GPU exchange for ndev > 1 handles:
// one ncclAllGather of each shard's per-pod best keys over NVLink.
// NCCL is loaded lazily with dlopen so that single-device handles (and processes
// that already carry torch's NCCL) never depend on it at link time.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

namespace kgpu {

class MultiDevice {
public:
    static MultiDevice *create(const std::vector<int> &devs, std::string *why);
    ~MultiDevice();
    // recv[i] (on device i) receives send[0..G) concatenated, `count` uint64 each.
    bool all_gather_u64(const std::vector<const void *> &send, const std::vector<void *> &recv, size_t count,
                        const std::vector<cudaStream_t> &streams, std::string *why);

private:
    MultiDevice() = default;
    struct Impl;
    Impl *impl_ = nullptr;
};

}  // namespace kgpu
