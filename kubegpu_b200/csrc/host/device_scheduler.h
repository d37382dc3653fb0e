// device_scheduler.h -- C++ mirror of the reference's scheduler-side plugin surface.
//
// The reference host language is Go (no Go toolchain in this image), so the host layer
// above the C ABI (include/kgpu.h) is written in C++ with the same names, argument
// meaning and error behaviour as the Go code it stands in for:
//
//   kubedevice::types::*            <- github.com/Microsoft/KubeDevice-API/pkg/types as USED by
//                                      the reference (SURVEY.md 8(b) "Types crossing the boundary")
//   gpuplugintypes::SortedTreeNode  <- gpuplugintypes/types.go:9-13, typeutils.go:10-93
//   gpuschedulerplugin::*           <- gpuschedulerplugin/gpu.go, gpu_scheduler.go
//   CreateDeviceSchedulerPlugin()   <- gpuschedulerplugin/plugin/gpuscheduler.go:8
//
// The tree cache / request translation is kept because the external core scheduler
// consumes the rewritten DevRequests (UsingGroupScheduler() == true in the reference).
// What is NEW is the score: PodFitsDevice's third return value, hard-coded 0.0 in the
// reference (gpu_scheduler.go:34-44), comes from the GPU scorer (K1), and ScoreBatch()
// is the batched side-door that amortises the cgo/launch cost over a whole scheduling
// cycle (SURVEY.md 7 "Hard parts": cgo).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

struct kgpu_ctx;

namespace kubedevice {
namespace types {

using ResourceName = std::string;
using ResourceList = std::map<ResourceName, int64_t>;          // ordered: SortedStringKeys for free
using ResourceLocation = std::map<ResourceName, ResourceName>;

extern const char *const DeviceGroupPrefix;                     // "resource/group"

struct ContainerInfo {
    ResourceList Requests, KubeRequests, DevRequests;
    ResourceLocation AllocateFrom;
};

struct PodInfo {
    std::string Name;
    ResourceList Requests;
    std::map<std::string, ContainerInfo> InitContainers, RunningContainers;
};

struct NodeInfo {
    std::string Name;   // present in KubeDevice-API's NodeInfo; the reference never reads it
    ResourceList Capacity, Allocatable, KubeCap, KubeAlloc;
};

void AddGroupResource(ResourceList &list, const std::string &key, int64_t val);   // gpu.go:50

}  // namespace types

namespace devicescheduler {
struct PredicateFailureReason {
    std::string reason;
};
}  // namespace devicescheduler
}  // namespace kubedevice

namespace gpuplugintypes {

extern const char *const ResourceGPU;   // "nvidia.com/gpu"  (gpuplugintypes/types.go:6)

struct SortedTreeNode {                 // gpuplugintypes/types.go:9-13
    int Val = 0;
    double Score = 0.0;
    std::vector<std::unique_ptr<SortedTreeNode>> Child;
};

SortedTreeNode *AddToSortedTreeNode(SortedTreeNode *node, int valToAdd);
SortedTreeNode *AddToSortedTreeNodeWithScore(SortedTreeNode *node, int valToAdd, double score);
void AddNodeToSortedTreeNode(SortedTreeNode *node, std::unique_ptr<SortedTreeNode> nodeToAdd);
bool CompareTreeNode(const SortedTreeNode *a, const SortedTreeNode *b);
std::string FormatTreeNode(const SortedTreeNode *node, int level = 0);   // PrintTreeNode's text

}  // namespace gpuplugintypes

namespace gpuschedulerplugin {

namespace types = kubedevice::types;
using gpuplugintypes::SortedTreeNode;

extern const char *const GPUTopologyGeneration;   // "gpu/gpu-generate-topology" (gpu_scheduler.go:14)
// Extension (no reference counterpart): pod-level request "every GPU I get must have at least this
// many MiB" -- checked against the per-GPU `<gpu>/memory` the node agent advertises.
extern const char *const GPUMinMemoryMiB;         // "gpu/gpu-min-memory-mib"

// ---- gpu.go, function for function ------------------------------------------------
types::ResourceList TranslateGPUResources(int64_t neededGPUs, const types::ResourceList &nodeResources,
                                          types::ResourceList containerRequests);
void SetGPUReqs(types::ContainerInfo &cont);
std::unique_ptr<SortedTreeNode> addToNode(std::unique_ptr<SortedTreeNode> node, const types::ResourceList &nodeResources,
                                          const std::string &partitionPrefix, const std::string &suffix, int partitionLevel);
double computeTreeScore(const SortedTreeNode *node);

// NodeCacheMap / NodeLocationMap (package globals in gpu.go:168-169) as an object.
class TreeCache {
public:
    struct Entry {
        std::unique_ptr<SortedTreeNode> tree;
        std::map<std::string, bool> ListOfNodes;
        double TreeScore = 0.0;
    };
    void AddResourcesToNodeTreeCache(const std::string &nodeName, const types::ResourceList &nodeResources);
    void RemoveNodeFromNodeTreeCache(const std::string &nodeName);
    // Deterministic where Go's map iteration is not: ties go to the shape that
    // compares smaller (same rule as the oracle).
    const SortedTreeNode *findBestTreeInCache(int num) const;
    const std::vector<std::unique_ptr<Entry>> &entries() const { return cache_; }
    const SortedTreeNode *location(const std::string &nodeName) const;

private:
    void removeNodeFromCache(const std::string &nodeName, const SortedTreeNode *loc);
    std::vector<std::unique_ptr<Entry>> cache_;
    std::map<std::string, const SortedTreeNode *> location_;
};

bool ConvertToBestGPURequests(const TreeCache &cache, types::PodInfo &podInfo);
// returns the error text ("" = nil) and `found`
std::string TranslatePodGPUResources(const TreeCache &cache, const types::NodeInfo &nodeInfo, types::PodInfo &podInfo,
                                     bool *found);

// ---- one GPU placement, as the kernels report it -------------------------------------
struct Placement {
    bool fits = false;
    uint32_t cost = 0;
    std::string nodeName;
    uint32_t gpuMask = 0;       // bit i = GPU slot i of that node
    uint64_t key = UINT64_MAX;  // raw (cost<<40 | node_id<<8 | mask)
    bool committed = false;     // the GPUs are already taken on the device (PlaceBatch) / by TakePodResources
};

// ---- gpu_scheduler.go: the DeviceScheduler boundary ------------------------------------
class NvidiaGPUScheduler {
public:
    // devices empty => host logic only (tree cache / translation); every scoring call
    // then fails loudly -- there is no CPU scorer in the product.
    // groupSchedulerMode = false (default): the plugin's own (node, GPU set) is authoritative -- PodAllocate fills
    //   AllocateFrom itself and UsingGroupScheduler() returns false, so the core does not run its group allocator
    //   over the rewritten DevRequests a second time ("return false iff it fills AllocateFrom").
    // groupSchedulerMode = true: the reference's contract (gpu_scheduler.go:69-71 returns true): PodAllocate only
    //   rewrites DevRequests (ConvertToBestGPURequests), AllocateFrom is left to the core's group allocator, and
    //   the GPU's score is advisory.
    explicit NvidiaGPUScheduler(const std::vector<int> &devices = {0}, bool groupSchedulerMode = false);
    ~NvidiaGPUScheduler();
    NvidiaGPUScheduler(const NvidiaGPUScheduler &) = delete;
    NvidiaGPUScheduler &operator=(const NvidiaGPUScheduler &) = delete;

    // The reference's eight methods (gpu_scheduler.go:21-71), same meaning.
    void AddNode(const std::string &nodeName, types::NodeInfo *nodeInfo);
    void RemoveNode(const std::string &nodeName);
    bool PodFitsDevice(types::NodeInfo *nodeInfo, types::PodInfo *podInfo, bool fillAllocateFrom,
                       std::vector<kubedevice::devicescheduler::PredicateFailureReason> *reasons, double *score);
    std::string PodAllocate(types::NodeInfo *nodeInfo, types::PodInfo *podInfo);              // "" = nil error
    std::string TakePodResources(types::NodeInfo *nodeInfo, types::PodInfo *podInfo);
    std::string ReturnPodResources(types::NodeInfo *nodeInfo, types::PodInfo *podInfo);
    std::string GetName() const { return "nvidiagpu"; }
    bool UsingGroupScheduler() const { return groupSchedulerMode_; }

    // Extensions (no reference counterpart).
    // Real NVML link matrix for a node (int32[8][8] row-major, levels 0..15) instead of the
    // one derived from its 2-level group names.
    std::string SetNodeTopology(const std::string &nodeName, const int32_t topo[64]);
    // Node from the node agent's GPU inventory JSON (nvgputypes.GpusInfo): names exactly as the
    // agent would advertise them, then AddNode, then the REAL link matrix (not the one implied by
    // the two group levels).  Fills *nodeInfo like the agent's UpdateNodeInfo.  "" or error text.
    std::string AddNodeFromGpusInfo(const std::string &nodeName, const std::string &gpusInfoJSON, bool useNVML,
                                    types::NodeInfo *nodeInfo);
    // Score a whole scheduling cycle in one kernel launch: best node + GPU set per pod.
    std::string ScoreBatch(const std::vector<const types::PodInfo *> &pods, std::vector<Placement> *out);
    // Place a whole cycle IN ORDER on the device: each pod takes its GPUs before the next one is
    // scored (K3).  The placements are remembered like ScoreBatch's and the usage is recorded, so a
    // later PodAllocate / ReturnPodResources works on them.
    std::string PlaceBatch(const std::vector<const types::PodInfo *> &pods, std::vector<Placement> *out);
    // Conflict-free PROPOSALS for a whole cycle: like PlaceBatch (each pod sees the GPUs the pods before it would
    // take) but on a scratch copy of the device state -- nothing is taken until TakePodResources commits a pod.
    // This is the conflict resolution ScoreBatch lacks: its snapshot scores hand every pod of equal k the same GPUs.
    std::string ProposeBatch(const std::vector<const types::PodInfo *> &pods, std::vector<Placement> *out);
    std::string LastError() const { return lastError_; }
    const TreeCache &cache() const { return cache_; }
    bool hasDevice() const { return handle_ != nullptr; }

    struct NodeRecord {                       // what AddNode learned about a node
        std::string name;
        int64_t index = -1;                   // slot in the device-side node array
        int nGpus = 0;
        std::vector<std::string> gpuNames;    // slot -> "gpugrp1/a/gpugrp0/b/gpu/<id>"
        int32_t topo[64] = {0};
        int32_t memMiB[8] = {0};              // per slot, from the advertised <gpu>/memory (0 = unknown)
        bool hasMem = false;
        uint32_t presentMask = 0, usedMask = 0;
        bool explicitTopo = false;
        bool removed = false;
    };
    const NodeRecord *node(const std::string &nodeName) const;
    static int64_t PodGPUCount(const types::PodInfo &podInfo);     // k of gpu.go:295-303

private:
    std::string syncNode(const NodeRecord &rec);
    std::string flushNodes();
    std::string scoreOne(const NodeRecord &rec, int k, int32_t minMemMiB, uint32_t *nodeKey);
    enum class BatchMode { Snapshot, Sequential, DryRun };
    std::string runBatch(const std::vector<const types::PodInfo *> &pods, std::vector<Placement> *out, BatchMode mode);
    const NodeRecord *recordOf(const types::NodeInfo *nodeInfo) const;
    std::string pushMask(NodeRecord &rec);
    TreeCache cache_;
    std::map<std::string, NodeRecord> nodes_;
    std::vector<std::string> indexToName_;
    std::map<std::string, Placement> lastPlacement_;   // pod name -> ScoreBatch result
    std::map<const types::NodeInfo *, std::string> infoToName_;   // AddNode's NodeInfo* -> node name
    kgpu_ctx *handle_ = nullptr;
    bool dirty_ = false;                                // node COUNT changed: the device array must be re-uploaded
    std::vector<std::string> changed_;                  // existing nodes whose matrix / presence changed (kgpu_update_node)
    bool groupSchedulerMode_ = false;
    std::string lastError_;
};

// gpuschedulerplugin/plugin/gpuscheduler.go:8 -- the factory symbol the core looks up.
std::unique_ptr<NvidiaGPUScheduler> CreateDeviceSchedulerPlugin(std::string *err);

}  // namespace gpuschedulerplugin
