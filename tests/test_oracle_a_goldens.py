"""Oracle A pinned against the reference's own golden vectors.

Each test names the reference test it restates; the expected values are the ones the
Go tests assert (or print deterministically)."""
import json
import math
import os

import pytest

from oracle import oracle_a as oa

P = oa.DEVICE_GROUP_PREFIX

NODE_RES1 = {  # gpuschedulerplugin/gpu_test.go:14-23
    P + "/gpugrp1/A/gpugrp0/0/gpu/0/cards": 1, P + "/gpugrp1/A/gpugrp0/0/gpu/1/cards": 1,
    P + "/gpugrp1/A/gpugrp0/1/gpu/2/cards": 1, P + "/gpugrp1/A/gpugrp0/1/gpu/3/cards": 1,
    P + "/gpugrp1/B/gpugrp0/2/gpu/4/cards": 1, P + "/gpugrp1/B/gpugrp0/2/gpu/5/cards": 1,
    P + "/gpugrp1/B/gpugrp0/3/gpu/6/cards": 1, P + "/gpugrp1/B/gpugrp0/3/gpu/7/cards": 1,
}
NODE_RES2 = {  # gpuschedulerplugin/gpu_test.go:24-33
    P + "/gpugrp1/A/gpugrp0/0/gpu/0/cards": 1, P + "/gpugrp1/A/gpugrp0/0/gpu/1/cards": 1,
    P + "/gpugrp1/A/gpugrp0/1/gpu/2/cards": 1, P + "/gpugrp1/A/gpugrp0/1/gpu/3/cards": 1,
    P + "/gpugrp1/B/gpugrp0/2/gpu/4/cards": 1, P + "/gpugrp1/B/gpugrp0/2/gpu/5/cards": 1,
    P + "/gpugrp1/B/gpugrp0/2/gpu/6/cards": 1, P + "/gpugrp1/B/gpugrp0/2/gpu/7/cards": 1,
}


def test_sorted_tree_node_insert_order():
    """gpuplugintypes/typeutils_test.go:7-34 TestSortedTreeNode."""
    root = oa.SortedTreeNode(10)
    child0 = oa.add_to_sorted_tree_node(root, 4)
    child1 = oa.add_to_sorted_tree_node(root, 8)
    oa.add_to_sorted_tree_node(child0, 3)
    oa.add_to_sorted_tree_node(child0, 1)
    oa.add_to_sorted_tree_node(child1, 1)
    oa.add_to_sorted_tree_node(child1, 4)
    oa.add_to_sorted_tree_node(child1, 3)
    assert root.shape() == (10, ((8, ((4, ()), (3, ()), (1, ()))), (4, ((3, ()), (1, ())))))
    expected = oa.SortedTreeNode(10)
    e8, e4 = oa.SortedTreeNode(8), oa.SortedTreeNode(4)
    e8.Child = [oa.SortedTreeNode(4), oa.SortedTreeNode(3), oa.SortedTreeNode(1)]
    e4.Child = [oa.SortedTreeNode(3), oa.SortedTreeNode(1)]
    expected.Child = [e8, e4]
    assert oa.compare_tree_node(root, expected)
    assert not oa.compare_tree_node(root, e8)
    assert oa.compare_tree_node(None, None) and not oa.compare_tree_node(root, None)


def test_tree_build_and_scores():
    """gpu_test.go:35-42: the two trees the test prints, and their scores
    (12 and 16: re-derived in SURVEY.md 4 / 8(a) row a4)."""
    t1 = oa.add_to_node(None, NODE_RES1, "gpugrp", "cards", 1)
    t2 = oa.add_to_node(None, NODE_RES2, "gpugrp", "cards", 1)
    assert t1.shape() == (8, ((4, ((2, ()), (2, ()))), (4, ((2, ()), (2, ())))))
    assert t2.shape() == (8, ((4, ((4, ()),)), (4, ((2, ()), (2, ())))))   # 4-group sorts first (Score 4 vs 2)
    assert oa.compute_tree_score(t1) == 12.0
    assert oa.compute_tree_score(t2) == 16.0
    assert oa.format_tree_node(t2) == "8\n   4\n      4\n   4\n      2\n      2\n"


@pytest.mark.parametrize("shape,score", [
    ([[8]], 24.0), ([[4], [4]], 20.0), ([[4, 4]], 16.0), ([[1]] * 8, 17.0),
    ([[3, 3, 2]], 13.333333333333334), ([[3], [3], [2]], 18.666666666666668),
    ([[1, 1, 1, 1], [3]], 11.5), ([[2, 2], [2, 2]], 12.0), ([[4], [2, 2]], 16.0),
])
def test_tree_score_known_answers(shape, score):
    """TreeScore KATs of SURVEY.md 8(a) row a4 (float64, operation order of gpu.go:180-190)."""
    tree = oa.add_to_node(None, oa.shape_to_resources(shape), "gpugrp", "cards", 1)
    assert oa.compute_tree_score(tree) == score


def test_empty_tree_is_nan_and_never_selected():
    cache = oa.TreeCache()
    cache.add_resources_to_node_tree_cache("D", {"ABCD": 4})         # gpu_test.go:46
    (tree, info), = cache.node_cache
    assert tree.Val == 0 and math.isnan(info.TreeScore)
    assert cache.find_best_tree_in_cache(0) is None
    cache.add_resources_to_node_tree_cache("E", {})                  # gpu.go:193-195: ignored
    cache.add_resources_to_node_tree_cache("F", None)
    assert len(cache.node_cache) == 1 and set(cache.node_location) == {"D"}


def _pod_wanting_3():
    return oa.PodInfo(RunningContainers={"A": oa.ContainerInfo(
        Requests={oa.RESOURCE_GPU: 3},
        DevRequests={P + "/gpugrp1/B/gpugrp0/3/gpu/6/cards": 1, P + "/gpugrp1/B/gpugrp0/3/gpu/7/cards": 1})})


def test_gpu_test_go_TestTree():
    """gpuschedulerplugin/gpu_test.go:43-112 -- the two reflect.DeepEqual assertions."""
    cache = oa.TreeCache()
    cache.add_resources_to_node_tree_cache("A", NODE_RES1)
    cache.add_resources_to_node_tree_cache("B", NODE_RES2)
    cache.add_resources_to_node_tree_cache("C", dict(NODE_RES1))
    cache.add_resources_to_node_tree_cache("D", {"ABCD": 4})
    assert len(cache.node_cache) == 3
    assert sorted(sorted(i.ListOfNodes) for _, i in cache.node_cache) == [["A", "C"], ["B"], ["D"]]
    cache.remove_node_from_node_tree_cache("A")
    assert sorted(sorted(i.ListOfNodes) for _, i in cache.node_cache) == [["B"], ["C"], ["D"]]

    pod = _pod_wanting_3()
    assert oa.convert_to_best_gpu_requests(cache, pod)
    cont = pod.RunningContainers["A"]
    assert cont.Requests == {oa.RESOURCE_GPU: 3}
    assert cont.DevRequests == {                                      # gpu_test.go:74-85
        P + "/gpugrp1/0/gpugrp0/0/gpu/0/cards": 1,
        P + "/gpugrp1/0/gpugrp0/0/gpu/1/cards": 1,
        P + "/gpugrp1/0/gpugrp0/0/gpu/2/cards": 1,
    }
    assert cont.KubeRequests == {} and cont.AllocateFrom == {}

    cache.remove_node_from_node_tree_cache("B")
    assert sorted(sorted(i.ListOfNodes) for _, i in cache.node_cache) == [["C"], ["D"]]
    assert oa.convert_to_best_gpu_requests(cache, pod)
    assert pod.RunningContainers["A"].DevRequests == {                # gpu_test.go:98-109
        P + "/gpugrp1/0/gpugrp0/0/gpu/0/cards": 1,
        P + "/gpugrp1/0/gpugrp0/0/gpu/1/cards": 1,
        P + "/gpugrp1/0/gpugrp0/1/gpu/0/cards": 1,
    }


def test_readding_same_shape_is_a_noop_and_shape_change_moves_node():
    cache = oa.TreeCache()
    cache.add_resources_to_node_tree_cache("A", NODE_RES1)
    loc = cache.node_location["A"]
    cache.add_resources_to_node_tree_cache("A", dict(NODE_RES1))
    assert cache.node_location["A"] is loc and len(cache.node_cache) == 1
    cache.add_resources_to_node_tree_cache("A", NODE_RES2)
    assert len(cache.node_cache) == 1 and cache.node_location["A"].shape() != loc.shape()


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return oa.parse_gpus_info(f.read())


def test_node_agent_titanx_names(golden_dir):
    """nvidia_gpu_manager_test.go:100-130: 8 x TITAN X -> gpugrp1/<i/4>/gpugrp0/<i/2>."""
    gpus = _load(golden_dir, "gpus_titanx.json")
    ngm = oa.NvidiaGPUManager(gpus, use_nvml=False)
    node = oa.NodeInfo()
    ngm.UpdateNodeInfo(node)
    expected = {oa.RESOURCE_GPU: 8}
    for i, g in enumerate(gpus):
        prefix = "/gpugrp1/%d/gpugrp0/%d" % (i // 4, i // 2)
        expected[P + prefix + "/gpu/" + g.ID + "/cards"] = 1
        expected[P + prefix + "/gpu/" + g.ID + "/memory"] = g.MemoryGlobal * 1024 * 1024
    assert node.Capacity == expected
    assert node.Allocatable == expected
    assert node.KubeCap == {oa.RESOURCE_GPU: 8} and node.KubeAlloc == {oa.RESOURCE_GPU: 8}


def test_node_agent_k80_names_and_allocate(golden_dir):
    """nvidia_gpu_manager_test.go:132-149: 4 x K80 without topology -> own groups;
    Allocate turns AllocateFrom into NVIDIA_VISIBLE_DEVICES (nvidia_gpu_manager.go:216-241)."""
    gpus = _load(golden_dir, "gpus_k80.json")
    ngm = oa.NvidiaGPUManager(gpus, use_nvml=False)
    node = oa.NodeInfo()
    ngm.UpdateNodeInfo(node)
    expected = {oa.RESOURCE_GPU: 4}
    for i, g in enumerate(gpus):
        prefix = "/gpugrp1/%d/gpugrp0/%d" % (i, i)
        expected[P + prefix + "/gpu/" + g.ID + "/cards"] = 1
        expected[P + prefix + "/gpu/" + g.ID + "/memory"] = g.MemoryGlobal * 1024 * 1024
    assert node.Capacity == expected
    cont = oa.ContainerInfo(AllocateFrom={
        P + "/gpu/%d/cards" % frm: P + "/gpugrp1/%d/gpugrp0/%d/gpu/%s/cards" % (to // 4, to // 2, gpus[to].ID)
        for frm, to in {4: 2, 3: 0, 5: 1}.items()})                     # setAllocFrom, test :38-46
    env = ngm.Allocate(oa.PodInfo(Name="TestPod"), cont)
    assert sorted(env["NVIDIA_VISIBLE_DEVICES"].split(",")) == sorted(gpus[t].ID for t in (2, 0, 1))
    assert ngm.Allocate(oa.PodInfo(), oa.ContainerInfo()) is None


def test_titanx_link_matrix_and_tree(golden_dir):
    """The reference's JSON fixture as a dense matrix (nvml.go:37-49,69-78 inverted) and
    the advertised names fed back through the scheduler side give shape [[2,2],[2,2]]."""
    gpus = _load(golden_dir, "gpus_titanx.json")
    m = oa.link_matrix_from_gpus(gpus)
    for i in range(8):
        for j in range(8):
            want = 0 if i == j or i // 4 != j // 4 else (5 if i // 2 == j // 2 else 3)
            assert m[i][j] == want
    node = oa.NodeInfo()
    oa.NvidiaGPUManager(gpus).UpdateNodeInfo(node)
    tree = oa.add_to_node(None, node.Allocatable, "gpugrp", "cards", 1)
    assert tree.shape() == (8, ((4, ((2, ()), (2, ()))), (4, ((2, ()), (2, ())))))


def test_topology_discovery_double_prefix_quirk():
    """nvidia_gpu_manager.go:80-87 has no TopoDone check: on a non-transitive matrix a
    GPU pulled in twice gets two prefixes."""
    gpus = [oa.GpuInfo(ID="G%d" % i, BusID="B%d" % i) for i in range(3)]
    gpus[0].Topology = [("B1", 5)]
    gpus[1].Topology = [("B0", 5), ("B2", 5)]
    gpus[2].Topology = [("B1", 5)]
    ngm = oa.NvidiaGPUManager(gpus)
    ngm.update_gpu_info()
    # level 0: G0 opens grp 0 and pulls G1; G2 opens grp 1 and pulls G1 again
    assert ngm.gpus["G1"].Name.count("gpugrp0/") == 2
    assert ngm.gpus["G0"].Name.startswith("gpugrp1/0/gpugrp0/0/gpu/")


def test_scheduler_boundary_semantics():
    """gpu_scheduler.go:34-71: score always 0.0, knob handling, error strings."""
    s = oa.NvidiaGPUScheduler()
    assert s.GetName() == "nvidiagpu" and s.UsingGroupScheduler() is True
    node = oa.NodeInfo(Allocatable=dict(NODE_RES2), KubeAlloc={oa.RESOURCE_GPU: 8})
    s.AddNode("B", node)
    pod = _pod_wanting_3()
    assert s.PodFitsDevice(node, pod, False) == (True, None, 0.0)
    assert len(pod.RunningContainers["A"].DevRequests) == 3
    assert s.PodAllocate(node, _pod_wanting_3()) is None
    bad = _pod_wanting_3()
    bad.Requests[oa.GPU_TOPOLOGY_GENERATION] = 7
    assert s.PodFitsDevice(node, bad, False) == (False, None, 0.0)
    assert s.PodAllocate(node, bad) == "Invalid topology generation request"
    assert s.TakePodResources(node, pod) is None and s.ReturnPodResources(node, pod) is None
    s.RemoveNode("B")
    # no tree left -> falls through to the flat (no-topology) translation and still "fits"
    flat = oa.PodInfo(RunningContainers={"A": oa.ContainerInfo(Requests={oa.RESOURCE_GPU: 2})})
    assert s.PodFitsDevice(node, flat, False) == (True, None, 0.0)
    got = flat.RunningContainers["A"].DevRequests
    assert sorted(got) == [P + "/gpugrp1/0/gpugrp0/0/gpu/0/cards", P + "/gpugrp1/1/gpugrp0/1/gpu/1/cards"]


def test_set_gpu_reqs_max_of_device_and_kube():
    """gpu.go:80-92."""
    c = oa.ContainerInfo(Requests={oa.RESOURCE_GPU: 2}, KubeRequests={oa.RESOURCE_GPU: 5})
    oa.set_gpu_reqs(c)
    assert c.Requests[oa.RESOURCE_GPU] == 5
    c = oa.ContainerInfo(KubeRequests={oa.RESOURCE_GPU: 3})
    oa.set_gpu_reqs(c)
    assert c.Requests[oa.RESOURCE_GPU] == 3
    c = oa.ContainerInfo()
    oa.set_gpu_reqs(c)
    assert c.Requests[oa.RESOURCE_GPU] == 0


def test_init_container_raises_k_and_containers_restart_at_first_leaf():
    """gpu.go:295-303 (k = sum of running, raised to any larger init request) and
    gpu.go:285 (each container restarts at the tree's first leaf)."""
    cache = oa.TreeCache()
    cache.add_resources_to_node_tree_cache("B", NODE_RES2)
    pod = oa.PodInfo(
        RunningContainers={"r1": oa.ContainerInfo(Requests={oa.RESOURCE_GPU: 1}),
                           "r2": oa.ContainerInfo(Requests={oa.RESOURCE_GPU: 2})},
        InitContainers={"i": oa.ContainerInfo(Requests={oa.RESOURCE_GPU: 9})})
    assert not oa.convert_to_best_gpu_requests(cache, pod)       # needs a tree with Val >= 9
    pod.InitContainers["i"].Requests[oa.RESOURCE_GPU] = 5
    assert oa.convert_to_best_gpu_requests(cache, pod)
    assert sorted(pod.RunningContainers["r1"].DevRequests) == [P + "/gpugrp1/0/gpugrp0/0/gpu/0/cards"]
    assert sorted(pod.RunningContainers["r2"].DevRequests) == [P + "/gpugrp1/0/gpugrp0/0/gpu/0/cards",
                                                               P + "/gpugrp1/0/gpugrp0/0/gpu/1/cards"]
    assert len(pod.InitContainers["i"].DevRequests) == 5
    assert P + "/gpugrp1/1/gpugrp0/0/gpu/0/cards" in pod.InitContainers["i"].DevRequests
