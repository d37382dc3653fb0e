// gpus_info.h -- the node agent's wire format on the scheduler side.
//
// C++ mirror of what the reference's node agent does between the NVML probe and the resource
// names the scheduler sees, so that AddNode can be fed either from advertised names (as the
// reference does) or directly from the GPU inventory JSON with the real link matrix:
//   nvgputypes::GpusInfo / ParseGpusInfo   <- nvidiagpuplugin/gpu/nvgputypes/types.go:9-43 (json tags)
//   nvidia::DiscoverTopology               <- nvidiagpuplugin/gpu/nvidia/nvidia_gpu_manager.go:63-91,132-180
//   nvidia::UpdateNodeInfo                 <- nvidia_gpu_manager.go:191-214
//   nvidia::LinkMatrix                     <- inverse of nvidiagpuplugin/gpu/nvml/nvml.go:37-49,69-78
//   nvidia::VisibleDevices                 <- nvidia_gpu_manager.go:216-241 (Allocate -> NVIDIA_VISIBLE_DEVICES)
#pragma once
#include <string>
#include <vector>

#include "device_scheduler.h"

namespace nvgputypes {

struct TopologyInfo {
    std::string BusID;
    int32_t Link = 0;
};

struct GpuInfo {
    std::string ID, Model, Path, BusID;   // UUID, Model, Path, PCI.BusID
    int64_t MemoryGlobal = 0;             // Memory.Global
    int64_t Bandwidth = 0;                // PCI.Bandwidth
    std::vector<TopologyInfo> Topology;
    // bookkeeping fields of the reference struct (json:"-")
    bool Found = false, TopoDone = false, InUse = false;
    int Index = 0;
    std::string Name;
};

struct GpusInfo {
    std::string Driver, CUDA;
    std::vector<GpuInfo> Gpus;
};

// encoding/json semantics for this schema: unknown keys ignored, null arrays empty.
// Returns "" or an error text.
std::string ParseGpusInfo(const std::string &json, GpusInfo *out);

}  // namespace nvgputypes

namespace nvidia {

namespace types = kubedevice::types;

// Naming pass: fills Found/Index/Name ("gpugrp1/<m>/gpugrp0/<n>/gpu/<UUID>") exactly as
// UpdateGPUInfo + topologyDiscovery({6,5,4},0) + topologyDiscovery({6,5,4,3,2,1},1) do,
// including the missing TopoDone check (a GPU pulled in twice is prefixed twice).
// useNVML=false applies the nvidia-docker unit conversion (MiB -> bytes, MB/s -> B/s).
void DiscoverTopology(nvgputypes::GpusInfo *info, bool useNVML);

// UpdateNodeInfo: nvidia.com/gpu counts + per-GPU <Name>/cards = 1 and <Name>/memory.
void UpdateNodeInfo(const nvgputypes::GpusInfo &named, types::NodeInfo *nodeInfo);

// Dense link-level matrix (row-major n x n, n = number of GPUs) from the Topology lists.
std::vector<int32_t> LinkMatrix(const nvgputypes::GpusInfo &info);

// Allocate(): the UUIDs named by a container's AllocateFrom, joined by ','.
std::string VisibleDevices(const types::ContainerInfo &cont);

}  // namespace nvidia
