"""The C++ host mirror of the reference's DeviceScheduler plugin
(kubegpu_b200/csrc/host/) driven through its line-protocol CLI and compared, transcript
against transcript, with Oracle A (the Python restatement pinned by the reference's goldens).

CPU part (`--no-device`): tree cache, request translation, knob/error behaviour.
GPU part (marked gpu): ScoreBatch / PodFitsDevice score / PodAllocate / Take-Return against
Oracle B on the matrices the group names imply."""
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import oracle_a as oa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "kubegpu_b200", "lib", "kgpu_sched_cli")
P = oa.DEVICE_GROUP_PREFIX


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-s", "-C", ROOT, "kubegpu_b200/lib/kgpu_sched_cli"])
    return CLI


def run_cli(cli, script, device=False):
    args = [cli] + ([] if device else ["--no-device"])
    return subprocess.run(args, input=script, capture_output=True, text=True, check=True).stdout


# ---- the same interpreter over Oracle A ----------------------------------------------------
def _dump_pod(pod, out):
    for kind, cs in (("run", pod.RunningContainers), ("init", pod.InitContainers)):
        for name in sorted(cs):
            c = cs[name]
            out.append("  %s %s req=%d" % (kind, name, c.Requests.get(oa.RESOURCE_GPU, -1)))
            for k in sorted(c.DevRequests):
                out.append("    dev %s=%d" % (k, c.DevRequests[k]))
            for k in sorted(c.AllocateFrom):
                out.append("    from %s -> %s" % (k, c.AllocateFrom[k]))


def run_oracle(script):
    sched = oa.NvidiaGPUScheduler()
    nodes, pods, out = {}, {}, []
    for line in script.splitlines():
        toks = line.split()
        if not toks or toks[0].startswith("#"):
            continue
        out.append("> " + line)
        cmd = toks[0]
        if cmd == "addnode":
            ni = oa.NodeInfo(KubeAlloc={oa.RESOURCE_GPU: int(toks[2])})
            for t in toks[3:]:
                k, v = t.rsplit("=", 1)
                ni.Allocatable[k] = int(v)
            nodes[toks[1]] = ni
            sched.AddNode(toks[1], ni)
            out += ["  alloc %s=%d" % (k, ni.Allocatable[k]) for k in sorted(ni.Allocatable)]
        elif cmd == "addjson":
            gpus = oa.parse_gpus_info(open(toks[2]).read())
            mgr = oa.NvidiaGPUManager(gpus, use_nvml=(len(toks) > 3 and toks[3] == "nvml"))
            ni = oa.NodeInfo()
            mgr.UpdateNodeInfo(ni)
            nodes[toks[1]] = ni
            sched.AddNode(toks[1], ni)
            out.append("  err=")
            out += ["  alloc %s=%d" % (k, ni.Allocatable[k]) for k in sorted(ni.Allocatable)]
            # slot i = i-th advertised cards key in sorted order -> GPU index -> real link matrix
            slot_names = [k[len(P) + 1:-len("/cards")] for k in sorted(ni.Allocatable) if k.endswith("/cards")]
            by_name = {g.Name: i for i, g in enumerate(mgr.gpus[uid] for uid in mgr.index_to_id)}
            links = oa.link_matrix_from_gpus(gpus)
            topo = [[0] * 8 for _ in range(8)]
            for i, a in enumerate(slot_names):
                for j, b in enumerate(slot_names):
                    if i != j:
                        topo[i][j] = links[by_name[a]][by_name[b]]
            out += ["  topo " + " ".join(str(v) for v in row) for row in topo]
        elif cmd == "visible":
            pod = pods[toks[1]]
            for name in sorted(pod.RunningContainers):
                env = oa.NvidiaGPUManager([]).Allocate(pod, pod.RunningContainers[name])
                out.append("  %s NVIDIA_VISIBLE_DEVICES=%s" % (name, "" if env is None else ",".join(sorted_ids(pod.RunningContainers[name]))))
        elif cmd == "rmnode":
            sched.RemoveNode(toks[1])
        elif cmd == "pod":
            pod = oa.PodInfo(Name=toks[1])
            cur, i = None, 2
            while i < len(toks):
                t = toks[i]
                if t in ("run", "init"):
                    cur = oa.ContainerInfo()
                    (pod.RunningContainers if t == "run" else pod.InitContainers)[toks[i + 1]] = cur
                    i += 1
                elif t.startswith("topogen="):
                    pod.Requests[oa.GPU_TOPOLOGY_GENERATION] = int(t[8:])
                elif cur is not None and t.startswith("req="):
                    cur.Requests[oa.RESOURCE_GPU] = int(t[4:])
                elif cur is not None and t.startswith("kube="):
                    cur.KubeRequests[oa.RESOURCE_GPU] = int(t[5:])
                elif cur is not None and t.startswith("dev:"):
                    k, v = t[4:].rsplit("=", 1)
                    cur.DevRequests[k] = int(v)
                i += 1
            pods[toks[1]] = pod
        elif cmd in ("fits", "allocate"):
            if toks[1] not in nodes or toks[2] not in pods:
                out.append("  unknown node or pod")
                continue
            if cmd == "fits":
                fits, reasons, score = sched.PodFitsDevice(nodes[toks[1]], pods[toks[2]], False)
                out.append("  fits=%d reasons=%d score=%.17g" % (1 if fits else 0, 0, score))
            else:
                out.append("  err=%s" % (sched.PodAllocate(nodes[toks[1]], pods[toks[2]]) or ""))
            _dump_pod(pods[toks[2]], out)
        elif cmd == "cache":
            rows = []
            for tree, info in sched.cache.node_cache:
                rows.append(oa.format_tree_node(tree) + "score=%.17g nodes=%s" % (
                    info.TreeScore, "".join(n + "," for n in sorted(info.ListOfNodes))))
            out += sorted(rows)
        elif cmd == "best":
            t = sched.cache.find_best_tree_in_cache(int(toks[1]))
            out.append(oa.format_tree_node(t).rstrip("\n") if t else "  none")
        else:
            out.append("  unknown command")
    return "\n".join(out) + "\n"


def sorted_ids(cont):
    """The node agent's Allocate walks AllocateFrom (a Go map); the C++ mirror walks it in
    sorted-key order, so compare in that order."""
    import re
    rx = re.compile(P + r"/gpugrp1/.*/gpugrp0/.*/gpu/(.*?)/cards")
    return [rx.search(cont.AllocateFrom[k]).group(1) for k in sorted(cont.AllocateFrom) if rx.search(cont.AllocateFrom[k])]


def node_line(name, shape, kube=None, ids=None):
    res = oa.shape_to_resources(shape)
    if ids:
        res = {k.replace("/gpu/%d/" % i, "/gpu/%s/" % ids[i]): v for i, k in enumerate(sorted(res, key=lambda s: int(s.split("/gpu/")[1].split("/")[0])))
               for v in [res[k]]}
    n = sum(sum(g) for g in shape)
    return "addnode %s %d %s" % (name, n if kube is None else kube, " ".join("%s=%d" % kv for kv in sorted(res.items())))


GOLDEN_SCRIPT = "\n".join([
    node_line("A", [[2, 2], [2, 2]]), node_line("B", [[2, 2], [4]]), node_line("C", [[2, 2], [2, 2]]),
    "addnode D 0 ABCD=4", "cache", "rmnode A", "cache",
    "pod p3 run A req=3 dev:%s/gpugrp1/B/gpugrp0/3/gpu/6/cards=1 dev:%s/gpugrp1/B/gpugrp0/3/gpu/7/cards=1" % (P, P),
    "fits C p3", "rmnode B", "cache", "fits C p3", "best 3", "best 9",
]) + "\n"


def test_reference_TestTree_through_the_cpp_host(cli):
    """gpuschedulerplugin/gpu_test.go:43-112 replayed through the C++ mirror."""
    got = run_cli(cli, GOLDEN_SCRIPT)
    assert got == run_oracle(GOLDEN_SCRIPT)
    first, second = got.split("> fits C p3")[1:3]
    assert "gpugrp1/0/gpugrp0/0/gpu/2/cards=1" in first            # golden 1: gpu 0,1,2 of the 4-group
    assert "gpugrp1/0/gpugrp0/1/gpu/0/cards=1" in second and "gpu/2/" not in second.split("> best")[0]
    assert "score=nan nodes=D," in got


def test_knobs_errors_and_flat_translation(cli):
    script = "\n".join([
        node_line("N1", [[4], [2, 2]]),
        "pod bad topogen=7 run a req=2", "fits N1 bad", "allocate N1 bad",
        "pod flat topogen=0 run a req=2 run b req=1", "fits N1 flat", "allocate N1 flat",
        "pod auto topogen=1 run a req=2 kube=5 init i req=6", "fits N1 auto",
        "pod big run a req=9", "fits N1 big", "allocate N1 big",
        "rmnode N1", "pod late run a req=2", "fits N1 late", "allocate N1 late",
        "addnode F 3 %s/gpu/0/cards=1 %s/gpu/1/cards=1 %s/gpu/2/cards=1" % (P, P, P), "cache",
        "pod onflat run a req=2", "fits F onflat",
    ]) + "\n"
    got = run_cli(cli, script)
    assert got == run_oracle(script)
    assert "err=Invalid topology generation request" in got


@pytest.mark.parametrize("seed", range(6))
def test_random_scripts_match_oracle(cli, seed):
    rng = random.Random(seed)
    shapes = [[[8]], [[4], [4]], [[2, 2], [2, 2]], [[4, 4]], [[4], [2, 2]], [[1]] * 4, [[3, 3, 2]], [[3], [3], [2]],
              [[1, 1, 1, 1], [3]], [[2, 2, 1], [3]], [[1, 1, 1, 1, 1, 1], [2]], [[6], [1, 1]], [[5, 1], [2]]]
    lines, names = [], []
    for step in range(60):
        r = rng.random()
        if r < 0.35 or not names:
            name = "n%d" % rng.randrange(12)
            lines.append(node_line(name, rng.choice(shapes)))
            if name not in names:
                names.append(name)
        elif r < 0.45:
            lines.append("rmnode %s" % rng.choice(names))
        elif r < 0.55:
            lines.append("cache")
        elif r < 0.65:
            lines.append("best %d" % rng.randrange(0, 10))
        else:
            pod = "p%d" % step
            conts = " ".join("%s c%d req=%d" % (rng.choice(["run", "run", "init"]), i, rng.randrange(0, 5))
                             for i in range(rng.randrange(1, 4)))
            knob = rng.choice(["", "", "topogen=1 ", "topogen=0 ", "topogen=3 "])
            lines.append("pod %s %s%s" % (pod, knob, conts))
            lines.append("%s %s %s" % (rng.choice(["fits", "allocate"]), rng.choice(names), pod))
    script = "\n".join(lines) + "\n"
    assert run_cli(cli, script) == run_oracle(script)


def _inventory_json(rng, n, sockets, pair_level=5, socket_level=3, cross=None):
    """A GPU inventory in the reference's JSON schema: GPUs paired on switches, grouped on sockets."""
    import json
    buses = ["%04X:%02X:00.0" % (rng.randrange(1 << 16), i) for i in range(n)]
    devs = []
    for i in range(n):
        topo = []
        for j in range(n):
            if i == j:
                continue
            if i // 2 == j // 2:
                topo.append({"BusID": buses[j], "Link": pair_level})
            elif i * sockets // n == j * sockets // n:
                topo.append({"BusID": buses[j], "Link": socket_level})
            elif cross is not None:
                topo.append({"BusID": buses[j], "Link": cross})
        devs.append({"UUID": "GPU-%04x-%d" % (rng.randrange(1 << 16), i), "Path": "/dev/nvidia%d" % i, "Model": "B200",
                     "PCI": {"BusID": buses[i], "Bandwidth": 63000}, "Topology": topo or None,
                     "Memory": {"Global": 183359}, "Unknown": {"nested": [1, 2.5e3, "x\\y", None, True]}})
    return json.dumps({"Version": {"Driver": "580.159", "CUDA": "12.9"}, "Devices": devs})


def test_gpus_info_json_ingestion_matches_oracle(cli, golden_dir, tmp_path):
    """node agent JSON -> advertised names (the reference's TestAlloc goldens,
    nvidia_gpu_manager_test.go:120-145) -> AddNode -> real link matrix, C++ mirror vs Oracle A."""
    rng = random.Random(11)
    files = [os.path.join(golden_dir, "gpus_titanx.json"), os.path.join(golden_dir, "gpus_k80.json")]
    for i, (n, sockets, cross) in enumerate([(8, 2, 1), (8, 1, None), (4, 2, 2), (6, 2, None), (2, 1, None), (8, 4, 1)]):
        path = tmp_path / ("inv%d.json" % i)
        path.write_text(_inventory_json(rng, n, sockets, pair_level=rng.choice([4, 5, 6]), socket_level=rng.choice([1, 2, 3]), cross=cross))
        files.append(str(path))
    lines = []
    for i, f in enumerate(files):
        lines.append("addjson J%d %s%s" % (i, f, " nvml" if i % 3 == 2 else ""))
    lines += ["cache", "best 4", "pod p run a req=3", "fits J0 p", "rmnode J0", "best 8"]
    script = "\n".join(lines) + "\n"
    got = run_cli(cli, script)
    assert got == run_oracle(script)
    # the reference's own expectations for its two fixtures
    titan = got.split("> addjson J1")[0]
    for i in range(8):
        assert "alloc %s/gpugrp1/%d/gpugrp0/%d/gpu/GPU0%d/cards=1" % (P, i // 4, i // 2, i) in titan
        assert "alloc %s/gpugrp1/%d/gpugrp0/%d/gpu/GPU0%d/memory=%d" % (P, i // 4, i // 2, i, 12238 * 1024 * 1024) in titan
    k80 = got.split("> addjson J1")[1].split("> addjson J2")[0]
    assert "gpugrp1/3/gpugrp0/3/gpu/GPU-aa4a86d4-3e1b-f48d-a69f-6aadd5f94466/cards=1" in k80
    bad = tmp_path / "bad.json"
    bad.write_text('{"Devices": [ {"UUID": "x", ')
    assert "err=GpusInfo:" in run_cli(cli, "addjson B %s\n" % bad)


# ---- GPU part --------------------------------------------------------------------------------
@pytest.mark.gpu
def test_score_batch_allocate_take_return_on_gpu(cli, oracle_b):
    from kubegpu_b200 import synth
    ids = ["GPU-%02d" % i for i in range(8)]
    script = "\n".join([
        node_line("A", [[2, 2], [2, 2]], ids=ids), node_line("B", [[4], [2, 2]], ids=ids), node_line("C", [[1]] * 4),
        "pod p1 run a req=1", "pod p2 run a req=2", "pod p3 run a req=1 run b req=2", "pod p4 run a req=4",
        "pod p8 run a req=8", "pod p9 run a req=9", "pod p5 run a req=5",
        "scorebatch p1 p2 p3 p4 p8 p9 p5",
        "fits A p4", "fits B p4", "fits C p4", "fits C p2",
        "allocate B p3", "take p4", "scorebatch p4 p2", "take p4", "return p4", "scorebatch p4",
    ]) + "\n"
    got = run_cli(cli, script, device=True)
    W = oracle_b.DEFAULT_WEIGHTS
    tA, tB = synth.shape_matrix([[2, 2], [2, 2]]), synth.shape_matrix([[4], [2, 2]])
    tC = np.zeros(64, np.int32)
    tC.reshape(8, 8)[:4, :4] = 1 - np.eye(4, dtype=np.int32)          # four singletons: cross level 1
    topo, free = np.stack([tA, tB, tC]), np.array([0xFF, 0xFF, 0x0F], np.int32)
    names = ["A", "B", "C"]
    keys = oracle_b.score_batch(topo, free, synth.make_pods(np.array([1, 2, 3, 4, 8, 9, 5], np.int32)), W)
    block = got.split("> scorebatch p1 p2 p3 p4 p8 p9 p5\n")[1].split("> fits")[0].splitlines()
    assert block[0] == "  err="
    for line, pod, key in zip(block[1:], ["p1", "p2", "p3", "p4", "p8", "p9", "p5"], keys):
        u = oracle_b.unpack_key(key)
        want = ("  %s fits=0 cost=0 node= mask=0x00" % pod) if u is None else \
            "  %s fits=1 cost=%d node=%s mask=0x%02x" % (pod, u[0], names[u[1]], u[2])
        assert line == want
    # per-pair score = 1/(1+cost of this node's best subset); C cannot host 4 GPUs -> k=4 fits the 4 singletons
    def score_of(node, k):
        nk = oracle_b.node_key(topo[node], int(free[node]), k)
        return None if nk == oracle_b.NODE_NO_FIT else 1.0 / (1.0 + (nk >> 8))
    for node, pod, k in (("A", "p4", 4), ("B", "p4", 4), ("C", "p4", 4), ("C", "p2", 2)):
        line = got.split("> fits %s %s\n" % (node, pod))[1].splitlines()[0]
        s = score_of(names.index(node), k)
        assert line == ("  fits=0 reasons=1 score=0" if s is None else "  fits=1 reasons=0 score=%.17g" % s)
    # PodAllocate after ScoreBatch: p3 (k=3) was placed on B's 4-group, slots 0,1,2 -> real GPU ids
    alloc = got.split("> allocate B p3\n")[1].split("> take")[0]
    assert "err=\n" in alloc
    froms = [ln.strip() for ln in alloc.splitlines() if ln.strip().startswith("from")]
    assert len(froms) == 3 and {f.split("/gpu/")[-1].split("/")[0] for f in froms} == {"GPU-00", "GPU-01", "GPU-02"}
    assert all("-> %s/gpugrp1/" % P in f and f.endswith("/cards") for f in froms)
    # Take: p4 took B's 4-group (mask 0x0f) -> rescoring p4 moves it off those GPUs, p2 avoids them too
    after = got.split("> take p4\n")[1]
    assert after.splitlines()[0] == "  err="
    free2 = free.copy()
    free2[1] = 0xF0
    k2 = oracle_b.score_batch(topo, free2, synth.make_pods(np.array([4, 2], np.int32)), W)
    lines = after.split("> scorebatch p4 p2\n")[1].splitlines()
    for line, pod, key in zip(lines[1:3], ["p4", "p2"], k2):
        u = oracle_b.unpack_key(key)
        assert line == "  %s fits=1 cost=%d node=%s mask=0x%02x" % (pod, u[0], names[u[1]], u[2])


@pytest.mark.gpu
def test_json_node_scored_with_real_matrix_and_visible_devices(cli, golden_dir, oracle_b):
    """The TITAN X inventory ingested from JSON is scored with its real NVML matrix (cross-socket
    pairs are absent there -> level 0 -> W[0]=64), not the 5/3/1 matrix its group names imply;
    PodAllocate + the node agent's Allocate regex give NVIDIA_VISIBLE_DEVICES."""
    f = os.path.join(golden_dir, "gpus_titanx.json")
    script = "\n".join([
        "addjson T %s" % f, "pod p5 run a req=5", "pod p2 run a req=2 run b req=1",
        "scorebatch p5 p2", "fits T p5", "allocate T p2", "visible p2",
    ]) + "\n"
    got = run_cli(cli, script, device=True)
    gpus = oa.parse_gpus_info(open(f).read())
    M = np.array(oa.link_matrix_from_gpus(gpus), dtype=np.int32).reshape(64)      # slots are GPU00..GPU07 in order
    k5 = oracle_b.node_key(M, 0xFF, 5)
    k3 = oracle_b.node_key(M, 0xFF, 3)
    assert k5 >> 8 == 2 * 2 + 4 * 8 + 4 * 64         # one whole socket (2 pairs W[5]=2, 4 pairs W[3]=8) + 1 GPU across (4 x W[0]=64)
    lines = got.split("> scorebatch p5 p2\n")[1].splitlines()
    assert lines[1] == "  p5 fits=1 cost=%d node=T mask=0x%02x" % (k5 >> 8, k5 & 0xFF)
    assert lines[2] == "  p2 fits=1 cost=%d node=T mask=0x%02x" % (k3 >> 8, k3 & 0xFF)
    assert got.split("> fits T p5\n")[1].splitlines()[0] == "  fits=1 reasons=0 score=%.17g" % (1.0 / (1.0 + (k5 >> 8)))
    vis = got.split("> visible p2\n")[1].splitlines()
    assert vis[0] == "  a NVIDIA_VISIBLE_DEVICES=GPU00,GPU01" and vis[1] == "  b NVIDIA_VISIBLE_DEVICES=GPU02"


@pytest.mark.gpu
def test_host_memory_constraint_and_place_batch(cli, golden_dir, tmp_path, oracle_b):
    """Host layer over K1m and K3: the advertised per-GPU memory becomes the device's gpu_mem, the pod
    knob gpu/gpu-min-memory-mib becomes min_mem, and PlaceBatch places a cycle in order."""
    rng = random.Random(5)
    big = tmp_path / "big.json"
    big.write_text(_inventory_json(rng, 8, 2, pair_level=6, socket_level=3, cross=1))      # 183359 MiB GPUs
    titan = os.path.join(golden_dir, "gpus_titanx.json")                                  # 12238 MiB GPUs
    script = "\n".join([
        "addjson T %s" % titan, "addjson B %s" % big,
        "pod small run a req=2", "pod hungry minmem=40000 run a req=2", "pod greedy minmem=400000 run a req=1",
        "scorebatch small hungry greedy",
        "pod q1 run a req=4", "pod q2 run a req=4", "pod q3 run a req=4", "pod q4 run a req=4", "pod q5 run a req=4",
        "placebatch q1 q2 q3 q4 q5", "scorebatch small", "allocate B q1",
    ]) + "\n"
    got = run_cli(cli, script, device=True)
    lines = got.split("> scorebatch small hungry greedy\n")[1].splitlines()
    # node ids: T = 0, B = 1.  `small` takes the cheapest pair anywhere (T's level-5 pair costs W[5]=2, B's
    # level-6 pair costs W[6]=1 -> B); `hungry` must go to B; nobody has 400 GB
    assert lines[1] == "  small fits=1 cost=1 node=B mask=0x03"
    assert lines[2] == "  hungry fits=1 cost=1 node=B mask=0x03"
    assert lines[3] == "  greedy fits=0 cost=0 node= mask=0x00"
    # sequential: four 4-GPU pods fill B's two sockets then T's two; the fifth finds nothing
    gpusT = oa.parse_gpus_info(open(titan).read())
    gpusB = oa.parse_gpus_info(open(big).read())
    topo = np.zeros((2, 64), np.int32)
    topo[0] = np.array(oa.link_matrix_from_gpus(gpusT), np.int32).reshape(64)
    topo[1] = np.array(oa.link_matrix_from_gpus(gpusB), np.int32).reshape(64)
    from kubegpu_b200 import synth
    want, free_after = oracle_b.place_batch(topo, np.array([0xFF, 0xFF], np.int32), synth.make_pods(np.array([4] * 5, np.int32)))
    names = ["T", "B"]
    plines = got.split("> placebatch q1 q2 q3 q4 q5\n")[1].splitlines()
    assert plines[0] == "  err="
    for line, pod, key in zip(plines[1:6], ["q1", "q2", "q3", "q4", "q5"], want):
        u = oracle_b.unpack_key(key)
        exp = "  %s fits=0 cost=0 node= mask=0x00" % pod if u is None else "  %s fits=1 cost=%d node=%s mask=0x%02x" % (pod, u[0], names[u[1]], u[2])
        assert line == exp
    assert free_after.tolist() == [0, 0]
    after = got.split("> scorebatch small\n")[1].splitlines()
    assert after[1] == "  small fits=0 cost=0 node= mask=0x00"          # the cluster is full now
    alloc = got.split("> allocate B q1\n")[1]
    assert alloc.count("from ") == 4 and "/gpu/" in alloc


@pytest.mark.gpu
def test_take_is_idempotent_and_propose_batch_resolves_conflicts(cli, oracle_b):
    """ADVICE r1 (device_scheduler.cc): (1) PlaceBatch already took the GPUs on the device, the TakePodResources the
    DeviceScheduler contract makes next must succeed; (2) ScoreBatch is snapshot scoring -- pods of equal k get the
    same GPUs and only the first Take can succeed; ProposeBatch hands out conflict-free proposals instead (the
    sequential placement run on a scratch copy of the device state), every Take succeeds and the state after the
    Takes equals the sequential oracle's."""
    from kubegpu_b200 import synth
    ids = ["GPU-%02d" % i for i in range(8)]
    script = "\n".join([
        node_line("A", [[2, 2], [2, 2]], ids=ids), node_line("B", [[4], [2, 2]], ids=ids),
        "pod a1 run a req=2", "pod a2 run a req=2", "pod a3 run a req=2", "pod a4 run a req=4", "pod a5 run a req=4",
        "using",
        "scorebatch a1 a2", "take a1", "take a2",                    # snapshot: the same GPUs twice -> second Take refused
        "return a1", "return a2",
        "proposebatch a1 a2 a3 a4 a5", "scorebatch a4",              # proposals take nothing: a4 still sees a free cluster
        "take a1", "take a2", "take a3", "take a4", "take a5", "take a1",
        "scorebatch a4", "allocate A a1",
        "placebatch a3", "take a3",
    ]) + "\n"
    got = run_cli(cli, script, device=True)
    assert got.split("> using\n")[1].splitlines()[0] == "  UsingGroupScheduler=0 name=nvidiagpu"
    snap = got.split("> scorebatch a1 a2\n")[1].splitlines()
    assert snap[1].split("fits=1")[1] == snap[2].split("fits=1")[1]            # identical (node, mask)
    t = got.split("> take a1\n")[1].splitlines()
    assert t[0] == "  err=" and "already in use" in got.split("> take a2\n")[1].splitlines()[0]
    tA, tB = synth.shape_matrix([[2, 2], [2, 2]]), synth.shape_matrix([[4], [2, 2]])
    topo, free = np.stack([tA, tB]), np.array([0xFF, 0xFF], np.int32)
    pods = synth.make_pods(np.array([2, 2, 2, 4, 4], np.int32))
    want, wf = oracle_b.place_batch(topo, free, pods)
    names = ["A", "B"]
    prop = got.split("> proposebatch a1 a2 a3 a4 a5\n")[1].splitlines()
    assert prop[0] == "  err="
    for line, pod, key in zip(prop[1:6], ["a1", "a2", "a3", "a4", "a5"], want):
        u = oracle_b.unpack_key(key)
        assert line == "  %s fits=1 cost=%d node=%s mask=0x%02x" % (pod, u[0], names[u[1]], u[2])
    k4 = oracle_b.unpack_key(oracle_b.score_batch(topo, free, synth.make_pods(np.array([4], np.int32)))[0])
    assert got.split("> scorebatch a4\n")[1].splitlines()[1] == "  a4 fits=1 cost=%d node=%s mask=0x%02x" % (k4[0], names[k4[1]], k4[2])
    takes = got.split("> proposebatch")[1].split("> take ")[1:7]
    assert all(blk.splitlines()[1] == "  err=" for blk in takes)              # a1..a5, and a1 again (idempotent)
    after = oracle_b.score_batch(topo, wf, synth.make_pods(np.array([4], np.int32)))[0]
    line = got.split("> scorebatch a4\n")[2].splitlines()[1]
    u = oracle_b.unpack_key(after)
    assert line == ("  a4 fits=0 cost=0 node= mask=0x00" if u is None else "  a4 fits=1 cost=%d node=%s mask=0x%02x" % (u[0], names[u[1]], u[2]))
    alloc = got.split("> allocate A a1\n")[1].split("> placebatch")[0]
    u1 = oracle_b.unpack_key(want[0])
    froms = [ln.strip() for ln in alloc.splitlines() if ln.strip().startswith("from")]
    assert len(froms) == 2 and {f.split("/gpu/")[-1].split("/")[0] for f in froms} == {"GPU-%02d" % i for i in range(8) if (u1[2] >> i) & 1}
    tail = got.split("> placebatch a3\n")[1]
    assert tail.splitlines()[0] == "  err=" and tail.split("> take a3\n")[1].splitlines()[0] == "  err="


@pytest.mark.gpu
def test_group_scheduler_mode_is_the_reference_contract(cli):
    """--group-scheduler: UsingGroupScheduler() == true like gpu_scheduler.go:69-71, PodAllocate only rewrites
    DevRequests and leaves AllocateFrom to the core's group allocator."""
    script = "\n".join([node_line("A", [[2, 2], [2, 2]]), "pod p run a req=2", "using", "scorebatch p", "allocate A p"]) + "\n"
    import subprocess
    got = subprocess.run([cli, "--group-scheduler"], input=script, capture_output=True, text=True, check=True).stdout
    assert got.split("> using\n")[1].splitlines()[0] == "  UsingGroupScheduler=1 name=nvidiagpu"
    alloc = got.split("> allocate A p\n")[1]
    assert "err=\n" in alloc and "from " not in alloc and "/gpugrp1/0/gpugrp0/0/gpu/" in alloc
