"""Oracle A -- CPU restatement of the reference's tree-score / greedy-fill path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``kubegpu_b200/`` may import this
module; only ``tests/``, ``bench.py``'s CPU-baseline legs and
``__graft_entry__.smoke()`` do, and only as the checker.

It follows the Go code of microsoft/KubeGPU @ 73e59ce function by function
(citations are paths under /root/reference).  It is pinned by the reference's
own golden vectors (tests/test_oracle_a_goldens.py):
  * gpuplugintypes/typeutils_test.go:8-29         sorted insert order
  * gpuschedulerplugin/gpu_test.go:14-33,61-85    k=3 against {T1,T2,empty}
  * gpuschedulerplugin/gpu_test.go:89-109         k=3 against {T1,empty}
  * nvidiagpuplugin/gpu/nvidia/nvidia_gpu_manager_test.go:16,120-130  8xTITAN X
  * nvidiagpuplugin/gpu/nvidia/nvidia_gpu_manager_test.go:17,140-145  4xK80

Parity UNPINNED pieces (third-party, absent from /root/reference, no test in
the reference exercises them): ``translate_resource`` restates
``github.com/Microsoft/KubeDevice-API/pkg/resource.TranslateResource`` (no
pinned version: the reference has no go.mod) from the behaviour implied at
gpuschedulerplugin/gpu_scheduler.go:22-24 -- see its docstring.

One deliberate addition: Go iterates ``NodeCacheMap`` in random order
(gpuschedulerplugin/gpu.go:235), so ties in ``find_best_tree_in_cache`` are
nondeterministic in the reference; here ties go to the lexicographically
smaller ``shape()``.
"""
from __future__ import annotations

import json
import math
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# --- KubeDevice-API types as *used* by the reference ---------
DEVICE_GROUP_PREFIX = "resource/group"      # inferred: gpu_test.go:15 vs gpu.go:286
RESOURCE_GPU = "nvidia.com/gpu"             # gpuplugintypes/types.go:6
GPU_TOPOLOGY_GENERATION = "gpu/gpu-generate-topology"   # gpu_scheduler.go:14

ResourceList = Dict[str, int]
ResourceLocation = Dict[str, str]


@dataclass
class ContainerInfo:
    Requests: ResourceList = field(default_factory=dict)
    KubeRequests: ResourceList = field(default_factory=dict)
    DevRequests: ResourceList = field(default_factory=dict)
    AllocateFrom: ResourceLocation = field(default_factory=dict)


@dataclass
class PodInfo:
    Name: str = ""
    Requests: ResourceList = field(default_factory=dict)
    InitContainers: Dict[str, ContainerInfo] = field(default_factory=dict)
    RunningContainers: Dict[str, ContainerInfo] = field(default_factory=dict)


@dataclass
class NodeInfo:
    Capacity: ResourceList = field(default_factory=dict)
    Allocatable: ResourceList = field(default_factory=dict)
    KubeCap: ResourceList = field(default_factory=dict)
    KubeAlloc: ResourceList = field(default_factory=dict)


def add_group_resource(lst: ResourceList, key: str, val: int) -> None:
    """types.AddGroupResource: prepends the group prefix (gpu.go:50)."""
    lst[DEVICE_GROUP_PREFIX + "/" + key] = val


# --- gpuplugintypes -----------------------------------------------------------
class SortedTreeNode:
    """gpuplugintypes/types.go:9-13."""
    __slots__ = ("Val", "Score", "Child")

    def __init__(self, val: int, score: float = 0.0):
        self.Val = val
        self.Score = score
        self.Child: List["SortedTreeNode"] = []

    def shape(self):
        """Nested tuple of Vals (used for the deterministic tie-break)."""
        return (self.Val, tuple(c.shape() for c in self.Child))


def _find_node_insertion_point(node: SortedTreeNode, val_to_add: int, score: float) -> int:
    """gpuplugintypes/typeutils.go:10-23 (strict '<': equal keys keep insertion order)."""
    point = len(node.Child)
    for index, child in enumerate(node.Child):
        if child.Val < val_to_add or (child.Val == val_to_add and child.Score < score):
            point = index
            break
    node.Child.insert(point, None)  # type: ignore[arg-type]
    return point


def add_to_sorted_tree_node_with_score(node, val_to_add, score) -> SortedTreeNode:
    """typeutils.go:27-31."""
    point = _find_node_insertion_point(node, val_to_add, score)
    node.Child[point] = SortedTreeNode(val_to_add, score)
    return node.Child[point]


def add_node_to_sorted_tree_node(node, node_to_add) -> None:
    """typeutils.go:33-36."""
    point = _find_node_insertion_point(node, node_to_add.Val, node_to_add.Score)
    node.Child[point] = node_to_add


def add_to_sorted_tree_node(node, val_to_add) -> SortedTreeNode:
    """typeutils.go:38-40."""
    return add_to_sorted_tree_node_with_score(node, val_to_add, 0.0)


def compare_tree_node(n1: Optional[SortedTreeNode], n2: Optional[SortedTreeNode]) -> bool:
    """typeutils.go:75-93 -- Val and child count recursively; Score ignored."""
    if n1 is None and n2 is None:
        return True
    if n1 is None or n2 is None:
        return False
    if n1.Val != n2.Val or len(n1.Child) != len(n2.Child):
        return False
    return all(compare_tree_node(a, b) for a, b in zip(n1.Child, n2.Child))


def format_tree_node(node: SortedTreeNode, level: int = 0) -> str:
    """typeutils.go:42-63 (PrintTreeNode / logTreeNode text form)."""
    out = " " * (3 * level) + "%d\n" % node.Val
    for c in node.Child:
        out += format_tree_node(c, level + 1)
    return out


# --- gpuschedulerplugin/gpu.go ------------------------------------------------
def _go_div(num: int, den: int) -> float:
    """float64(num)/float64(den) with Go semantics (x/0 -> +-Inf, 0/0 -> NaN)."""
    if den != 0:
        return float(num) / float(den)
    if num == 0:
        return math.nan
    return math.inf if num > 0 else -math.inf


def compute_tree_score_at_level(node: SortedTreeNode, level: int, num_child: int) -> float:
    """gpu.go:180-186 -- exact operation order (DFS, running +=)."""
    score = _go_div(node.Val * level, num_child)
    for child in node.Child:
        score += compute_tree_score_at_level(child, level + 1, len(node.Child))
    return score


def compute_tree_score(node: SortedTreeNode) -> float:
    """gpu.go:188-190."""
    return compute_tree_score_at_level(node, 0, len(node.Child))


def add_to_node(node: Optional[SortedTreeNode], node_resources: ResourceList,
                partition_prefix: str, suffix: str, partition_level: int) -> SortedTreeNode:
    """gpu.go:129-161."""
    child_map: Dict[str, ResourceList] = {}
    rx = re.compile(r".*/" + partition_prefix + str(partition_level) + r"/(.*?)/.*/" + suffix)
    total_len = 0
    for key in sorted(node_resources):                      # utils.SortedStringKeys, gpu.go:133
        m = rx.search(key)                                   # FindStringSubmatch: unanchored
        if m is not None:
            child_map.setdefault(m.group(1), {})[key] = node_resources[key]
            total_len += 1
    if node is None:
        node = SortedTreeNode(total_len)
    for sub_key in sorted(child_map):                        # gpu.go:149
        sub_maps = child_map[sub_key]
        child = SortedTreeNode(len(sub_maps))
        if partition_level > 0:
            add_to_node(child, sub_maps, partition_prefix, suffix, partition_level - 1)
            child.Score = compute_tree_score(child)          # gpu.go:155
        add_node_to_sorted_tree_node(node, child)
    return node


@dataclass
class TreeInfo:
    ListOfNodes: Dict[str, bool]
    TreeScore: float


class TreeCache:
    """The package-global NodeCacheMap / NodeLocationMap (gpu.go:168-169) as an object."""

    def __init__(self) -> None:
        self.node_cache: List[Tuple[SortedTreeNode, TreeInfo]] = []   # keyed by identity
        self.node_location: Dict[str, SortedTreeNode] = {}

    def _info(self, tree: SortedTreeNode) -> Optional[TreeInfo]:
        for t, info in self.node_cache:
            if t is tree:
                return info
        return None

    def _remove_node_from_cache(self, node_name: str, loc: Optional[SortedTreeNode]) -> None:
        """gpu.go:171-178."""
        if loc is None:
            return
        info = self._info(loc)
        if info is None:
            return
        info.ListOfNodes.pop(node_name, None)
        if not info.ListOfNodes:
            self.node_cache = [(t, i) for t, i in self.node_cache if t is not loc]

    def add_resources_to_node_tree_cache(self, node_name: str, node_resources: Optional[ResourceList]) -> None:
        """gpu.go:192-224."""
        if not node_resources:
            return
        node = add_to_node(None, node_resources, "gpugrp", "cards", 1)
        loc = self.node_location.get(node_name)
        if compare_tree_node(node, loc):
            return
        self._remove_node_from_cache(node_name, loc)
        found = False
        for tree, info in self.node_cache:
            if compare_tree_node(node, tree):
                info.ListOfNodes[node_name] = True
                loc = tree
                found = True
                break
        if not found:
            self.node_cache.append((node, TreeInfo({node_name: True}, compute_tree_score(node))))
            loc = node
        self.node_location[node_name] = loc

    def remove_node_from_node_tree_cache(self, node_name: str) -> None:
        """gpu.go:226-230."""
        self._remove_node_from_cache(node_name, self.node_location.get(node_name))
        self.node_location.pop(node_name, None)

    def find_best_tree_in_cache(self, num: int) -> Optional[SortedTreeNode]:
        """gpu.go:232-245; strict '>' from 0.0 so NaN / <=0 never win.  Ties: see module doc."""
        best, best_score = None, 0.0
        for tree, info in self.node_cache:
            if tree.Val >= num:
                if info.TreeScore > best_score or (
                        best is not None and info.TreeScore == best_score and tree.shape() < best.shape()):
                    best, best_score = tree, info.TreeScore
        return best


def assign_gpus(node: SortedTreeNode, prefix: str, resource_grp: str, resource: str,
                suffix: str, level: int, num_left: List[int]) -> ResourceList:
    """gpu.go:247-271.  ``num_left`` is a one-element list standing in for *int."""
    res: ResourceList = {}
    if level == 0:
        to_take = node.Val
        if num_left[0] <= node.Val:
            to_take = num_left[0]
        for i in range(to_take):
            res[prefix + "/" + resource + "/" + str(i) + "/" + suffix] = 1
        num_left[0] -= to_take
    else:
        for i, child in enumerate(node.Child):
            new_prefix = prefix + str(level - 1) + "/" + str(i)
            if level - 1 != 0:
                new_prefix += "/" + resource_grp
            res.update(assign_gpus(child, new_prefix, resource_grp, resource, suffix, level - 1, num_left))
    return res


_RX_GPU_ANY = re.compile(r".*/gpu/.*")


def translate_to_tree(node: SortedTreeNode, cont: ContainerInfo) -> None:
    """gpu.go:273-291."""
    cont.DevRequests = {k: v for k, v in cont.DevRequests.items() if _RX_GPU_ANY.search(k) is None}
    num = [int(cont.Requests.get(RESOURCE_GPU, 0))]
    res = assign_gpus(node, DEVICE_GROUP_PREFIX + "/gpugrp", "gpugrp", "gpu", "cards", 2, num)
    cont.DevRequests.update(res)


def convert_to_best_gpu_requests(cache: TreeCache, pod: PodInfo) -> bool:
    """gpu.go:294-324."""
    num_gpus = 0
    for cont in pod.RunningContainers.values():
        num_gpus += cont.Requests.get(RESOURCE_GPU, 0)
    for cont in pod.InitContainers.values():
        if cont.Requests.get(RESOURCE_GPU, 0) > num_gpus:
            num_gpus = cont.Requests.get(RESOURCE_GPU, 0)
    best = cache.find_best_tree_in_cache(int(num_gpus))
    if best is None:
        return False
    for key in sorted(pod.RunningContainers):
        translate_to_tree(best, pod.RunningContainers[key])
    for key in sorted(pod.InitContainers):
        translate_to_tree(best, pod.InitContainers[key])
    return True


def translate_resource(node_resources: ResourceList, container_requests: ResourceList,
                       this_stage: str, next_stage: str) -> Tuple[bool, ResourceList]:
    """UNPINNED restatement of KubeDevice-API ``resource.TranslateResource``.

    Call sites: gpu.go:55,58.  The source is not in /root/reference and no
    reference test reaches it.  Behaviour implied by gpu_scheduler.go:22-24 and
    the comment at gpu_scheduler.go:13 ("everything in its own group"): when the
    node advertises names containing ``/<this_stage>/`` and a request names
    ``/<next_stage>/<id>/`` without a ``/<this_stage>/`` level in front of it,
    insert ``<this_stage>/<n>/`` before ``<next_stage>/<id>``, one fresh group
    index n per distinct <id> (numbered after the largest integer group index
    already present in the requests), visiting requests in sorted-key order.
    """
    rx_node = re.compile(r".*/" + re.escape(this_stage) + r"/(.*?)/" + re.escape(next_stage) + r"/.*")
    if not any(rx_node.search(k) for k in node_resources):
        return False, container_requests
    rx_has = re.compile(r"(.*)/" + re.escape(this_stage) + r"/(.*?)/" + re.escape(next_stage) + r"/(.*?)/(.*)")
    rx_need = re.compile(r"(.*)/" + re.escape(next_stage) + r"/(.*?)/(.*)")
    max_idx = -1
    for k in container_requests:
        m = rx_has.search(k)
        if m:
            try:
                max_idx = max(max_idx, int(m.group(2)))
            except ValueError:
                pass
    group_of: Dict[str, int] = {}
    out: ResourceList = {}
    modified = False
    for k in sorted(container_requests):
        v = container_requests[k]
        if rx_has.search(k):
            out[k] = v
            continue
        m = rx_need.match(k)
        if not m:
            out[k] = v
            continue
        ident = m.group(2)
        if ident not in group_of:
            max_idx += 1
            group_of[ident] = max_idx
        out[m.group(1) + "/" + this_stage + "/" + str(group_of[ident]) + "/" + next_stage + "/" + ident + "/" + m.group(3)] = v
        modified = True
    return modified, out


def translate_gpu_resources(needed_gpus: int, node_resources: ResourceList,
                            container_requests: ResourceList) -> ResourceList:
    """gpu.go:16-66."""
    rx = re.compile(DEVICE_GROUP_PREFIX + r".*/gpu/(.*?)/cards")
    if not any(rx.search(k) for k in node_resources):
        return container_requests
    have, max_index = 0, -1
    for k in container_requests:
        m = rx.search(k)
        if m:
            have += 1
            try:
                max_index = max(max_index, _atoi(m.group(1)))
            except ValueError:
                pass
    for i in range(int(needed_gpus - have)):
        add_group_resource(container_requests, "gpu/" + str(max_index + i + 1) + "/cards", 1)
    _, container_requests = translate_resource(node_resources, container_requests, "gpugrp0", "gpu")
    _, container_requests = translate_resource(node_resources, container_requests, "gpugrp1", "gpugrp0")
    return container_requests


def _atoi(s: str) -> int:
    """strconv.Atoi: optional sign + decimal digits only."""
    if not re.fullmatch(r"[+-]?[0-9]+", s):
        raise ValueError(s)
    return int(s)


def translate_gpu_container_resources(alloc: ResourceList, cont: ContainerInfo) -> ResourceList:
    """gpu.go:75-78."""
    return translate_gpu_resources(cont.Requests.get(RESOURCE_GPU, 0), alloc, cont.DevRequests)


def set_gpu_reqs(cont: ContainerInfo) -> None:
    """gpu.go:80-92 (writes through the shared Requests map)."""
    ok, ok_k = RESOURCE_GPU in cont.Requests, RESOURCE_GPU in cont.KubeRequests
    if ok and ok_k:
        cont.Requests[RESOURCE_GPU] = max(cont.Requests[RESOURCE_GPU], cont.KubeRequests[RESOURCE_GPU])
    elif ok:
        pass
    elif ok_k:
        cont.Requests[RESOURCE_GPU] = cont.KubeRequests[RESOURCE_GPU]
    else:
        cont.Requests[RESOURCE_GPU] = 0


def translate_pod_gpu_resources(cache: TreeCache, node: NodeInfo, pod: PodInfo) -> Tuple[Optional[str], bool]:
    """gpu.go:94-127.  Returns (error-or-None, found)."""
    for cont in pod.InitContainers.values():
        set_gpu_reqs(cont)
    for cont in pod.RunningContainers.values():
        set_gpu_reqs(cont)
    ok = GPU_TOPOLOGY_GENERATION in pod.Requests
    req = pod.Requests.get(GPU_TOPOLOGY_GENERATION, 0)
    found = True
    if (not ok) or req == 1:
        found = convert_to_best_gpu_requests(cache, pod)
        if found:
            return None, True
    if (not found) or req == 0:
        for cont in pod.InitContainers.values():
            cont.DevRequests = translate_gpu_container_resources(node.Allocatable, cont)
        for cont in pod.RunningContainers.values():
            cont.DevRequests = translate_gpu_container_resources(node.Allocatable, cont)
        return None, True
    return "Invalid topology generation request", False


class NvidiaGPUScheduler:
    """gpuschedulerplugin/gpu_scheduler.go:17-71 -- the DeviceScheduler boundary."""

    def __init__(self) -> None:
        self.cache = TreeCache()

    def AddNode(self, node_name: str, node: NodeInfo) -> None:                     # :21-28
        node.Allocatable = translate_gpu_resources(
            node.KubeAlloc.get(RESOURCE_GPU, 0),
            {DEVICE_GROUP_PREFIX + "/gpugrp1/A/gpugrp0/B/gpu/GPU0/cards": 1},
            node.Allocatable)
        self.cache.add_resources_to_node_tree_cache(node_name, node.Allocatable)

    def RemoveNode(self, node_name: str) -> None:                                   # :30-32
        self.cache.remove_node_from_node_tree_cache(node_name)

    def PodFitsDevice(self, node: NodeInfo, pod: PodInfo, fill_allocate_from: bool):  # :34-44
        err, found = translate_pod_gpu_resources(self.cache, node, pod)
        if err is not None or not found:
            return False, None, 0.0
        return True, None, 0.0

    def PodAllocate(self, node: NodeInfo, pod: PodInfo) -> Optional[str]:           # :46-55
        err, found = translate_pod_gpu_resources(self.cache, node, pod)
        if err is not None:
            return err
        if not found:
            return "TranslatePodGPUResources fails as no translation is found"
        return None

    def TakePodResources(self, node, pod) -> None:                                  # :57-59
        return None

    def ReturnPodResources(self, node, pod) -> None:                                # :61-63
        return None

    def GetName(self) -> str:                                                       # :65-67
        return "nvidiagpu"

    def UsingGroupScheduler(self) -> bool:                                          # :69-71
        return True


# --- node side: nvidiagpuplugin ----------------------------------------------
@dataclass
class GpuInfo:
    """nvidiagpuplugin/gpu/nvgputypes/types.go:22-34."""
    ID: str = ""
    Model: str = ""
    Path: str = ""
    MemoryGlobal: int = 0
    BusID: str = ""
    Bandwidth: int = 0
    Topology: List[Tuple[str, int]] = field(default_factory=list)   # (BusID, Link)
    Found: bool = False
    Index: int = 0
    InUse: bool = False
    TopoDone: bool = False
    Name: str = ""


def parse_gpus_info(text: str) -> List[GpuInfo]:
    """json.Unmarshal into GpusInfo (nvgputypes/types.go:40-43); unknown keys dropped."""
    doc = json.loads(text)
    out = []
    for d in doc.get("Devices") or []:
        out.append(GpuInfo(
            ID=d.get("UUID", ""), Model=d.get("Model", ""), Path=d.get("Path", ""),
            MemoryGlobal=int((d.get("Memory") or {}).get("Global", 0)),
            BusID=(d.get("PCI") or {}).get("BusID", ""),
            Bandwidth=int((d.get("PCI") or {}).get("Bandwidth", 0)),
            Topology=[(t["BusID"], int(t["Link"])) for t in (d.get("Topology") or [])]))
    return out


class NvidiaGPUManager:
    """nvidia_gpu_manager.go:20-241 restricted to naming/advertising/Allocate."""

    def __init__(self, gpus: List[GpuInfo], use_nvml: bool = False) -> None:
        self.source = gpus
        self.use_nvml = use_nvml
        self.gpus: Dict[str, GpuInfo] = {}
        self.bus_id_to_id: Dict[str, str] = {}
        self.index_to_id: List[str] = []
        self.num_gpus = 0

    def _topology_discovery(self, links: List[int], level: int) -> None:
        """nvidia_gpu_manager.go:63-91, including the missing done-check at :80-87
        (a pulled-in GPU is prefixed again even if already TopoDone)."""
        for g in self.gpus.values():
            g.TopoDone = False
        link_id = 0
        for gid in self.index_to_id:
            g = self.gpus[gid]
            if not g.Found or g.TopoDone:
                continue
            prefix = "gpugrp" + str(level) + "/" + str(link_id)
            link_id += 1
            g.Name = prefix + "/" + g.Name
            g.TopoDone = True
            for bus, link in g.Topology:
                if link in links:
                    other = self.gpus.get(self.bus_id_to_id.get(bus, ""))
                    if other is not None and other.Found:
                        other.Name = prefix + "/" + other.Name
                        other.TopoDone = True

    def update_gpu_info(self) -> None:
        """nvidia_gpu_manager.go:94-183."""
        import copy
        found = [copy.deepcopy(g) for g in self.source]
        if not self.use_nvml:                        # :124-129 unit conversion
            for g in found:
                g.MemoryGlobal *= 1024 * 1024
                g.Bandwidth *= 1000 * 1000
        for g in self.gpus.values():
            g.Found = False
        self.bus_id_to_id = {}
        self.index_to_id = [""] * len(found)
        for index, g in enumerate(found):
            old = self.gpus.get(g.ID)
            if old is not None:
                g.InUse = old.InUse
            g.Found, g.Index, g.Name = True, index, "gpu/" + g.ID
            self.gpus[g.ID] = g
            self.bus_id_to_id[g.BusID] = g.ID
            self.index_to_id[index] = g.ID
        self.num_gpus = len(found)
        self._topology_discovery([6, 5, 4], 0)              # :178
        self._topology_discovery([6, 5, 4, 3, 2, 1], 1)     # :180

    def UpdateNodeInfo(self, node: NodeInfo) -> None:
        """nvidia_gpu_manager.go:191-214."""
        self.update_gpu_info()
        n = len(self.gpus)
        for lst in (node.Capacity, node.Allocatable, node.KubeCap, node.KubeAlloc):
            lst[RESOURCE_GPU] = n
        for g in self.gpus.values():
            if g.Found:
                for lst in (node.Capacity, node.Allocatable):
                    add_group_resource(lst, g.Name + "/memory", g.MemoryGlobal)
                    add_group_resource(lst, g.Name + "/cards", 1)

    def Allocate(self, pod: PodInfo, cont: ContainerInfo):
        """nvidia_gpu_manager.go:216-241 -> env NVIDIA_VISIBLE_DEVICES."""
        if not cont.AllocateFrom:
            return None
        rx = re.compile(DEVICE_GROUP_PREFIX + r"/gpugrp1/.*/gpugrp0/.*/gpu/(.*?)/cards")
        ids = []
        for res in cont.AllocateFrom.values():
            m = rx.search(res)
            if m:
                ids.append(m.group(1))
        return {"NVIDIA_VISIBLE_DEVICES": ",".join(ids)}


def link_matrix_from_gpus(gpus: List[GpuInfo]) -> List[List[int]]:
    """Dense n x n link-level matrix from the per-GPU Topology lists, i.e. the
    inverse of nvml.go:37-49,69-78 (entries keyed by BusID, j != i only; pairs
    absent from the list -- e.g. cross-socket in the TITAN X fixture -- stay 0)."""
    bus_index = {g.BusID: i for i, g in enumerate(gpus)}
    n = len(gpus)
    m = [[0] * n for _ in range(n)]
    for i, g in enumerate(gpus):
        for bus, link in g.Topology:
            j = bus_index.get(bus)
            if j is not None and j != i:
                m[i][j] = link
    return m


# --- helpers tying Oracle A to Oracle B ------------------------
def shape_to_resources(shape: List[List[int]]) -> ResourceList:
    """[[2,2],[2,2]] -> resource/group/gpugrp1/<a>/gpugrp0/<b>/gpu/<n>/cards names."""
    res: ResourceList = {}
    g0 = gpu = 0
    for a, grp1 in enumerate(shape):
        for cnt in grp1:
            for _ in range(cnt):
                res["%s/gpugrp1/%d/gpugrp0/%d/gpu/%d/cards" % (DEVICE_GROUP_PREFIX, a, g0, gpu)] = 1
                gpu += 1
            g0 += 1
    return res


def tree_to_matrix(tree: SortedTreeNode, same_grp0: int = 5, same_grp1: int = 3, cross: int = 1) -> List[int]:
    """8x8 link-level matrix (row-major list of 64) for a sorted 2-level tree; GPU
    index = position in the sorted tree (the order assign_gpus walks).  Levels
    default to the TITAN X fixture's (nvidia_gpu_manager_test.go:16)."""
    owner = []
    for a, c1 in enumerate(tree.Child):
        for b, c0 in enumerate(c1.Child):
            owner += [(a, b)] * c0.Val
    m = [0] * 64
    for i, (a1, b1) in enumerate(owner[:8]):
        for j, (a2, b2) in enumerate(owner[:8]):
            if i == j:
                continue
            m[i * 8 + j] = same_grp0 if (a1, b1) == (a2, b2) else (same_grp1 if a1 == a2 else cross)
    return m


def greedy_fill_mask(tree: SortedTreeNode, k: int) -> int:
    """Bitmask (GPU index = sorted-tree position) of what assign_gpus picks for k GPUs."""
    mask, pos, left = 0, 0, k
    for c1 in tree.Child:
        for c0 in c1.Child:
            take = min(left, c0.Val)
            for t in range(take):
                mask |= 1 << (pos + t)
            left -= take
            pos += c0.Val
    return mask
