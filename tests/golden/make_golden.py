#!/usr/bin/env python3
"""Regenerates the committed fixtures under tests/golden/.

1. gpus_titanx.json / gpus_k80.json -- the two GPU inventories the reference's node
   agent test feeds its fake backend (nvidiagpuplugin/gpu/nvidia/
   nvidia_gpu_manager_test.go:16-17), rebuilt here from their parameters (8 TITAN X on
   two sockets: PCIe pairs at link level 5, same socket 3, cross socket absent;
   4 K80 with no topology).  Only the fields the reference's GpusInfo schema keeps
   (nvgputypes/types.go:22-43) are emitted.  When /root/reference is present the
   script checks them field by field against the reference's own constants.
2. oracle_b_vectors.npz -- seeded synthetic inputs (kubegpu_b200.synth) with the keys
   Oracle B (oracle/oracle_b.c, plain version) produces for them; the GPU parity
   tests also compare against these committed outputs.
3. agree_set.json -- Oracle A greedy vs Oracle B optimum over every two-level shape
   of 8 GPUs x k=1..8: the counts and the full divergent list.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from kubegpu_b200 import synth  # noqa: E402
from oracle import oracle_a as oa  # noqa: E402
from oracle import oracle_b as ob  # noqa: E402


def titanx_doc():
    buses = ["0000:04:00.0", "0000:05:00.0", "0000:08:00.0", "0000:09:00.0",
             "0000:85:00.0", "0000:86:00.0", "0000:89:00.0", "0000:8A:00.0"]
    devs = []
    for i in range(8):
        topo = []
        for j in range(8):
            if j == i or j // 4 != i // 4:
                continue                      # cross-socket pairs are absent from the fixture
            topo.append({"BusID": buses[j], "Link": 5 if j // 2 == i // 2 else 3})
        devs.append({"UUID": "GPU0%d" % i, "Path": "/dev/nvidia%d" % i, "Model": "GeForce GTX TITAN X",
                     "PCI": {"BusID": buses[i], "Bandwidth": 15760}, "Topology": topo,
                     "Memory": {"Global": 12238}})
    return {"Version": {"Driver": "375.20", "CUDA": "8.0"}, "Devices": devs}


def k80_doc():
    uuids = ["GPU01", "GPU-dc6182bb-4760-894c-e144-592b0acd7657",
             "GPU-9f0b1fcf-222f-0701-a230-ad08406c0104", "GPU-aa4a86d4-3e1b-f48d-a69f-6aadd5f94466"]
    buses = ["777C:00:00.0", "9710:00:00.0", "B29F:00:00.0", "CF72:00:00.0"]
    devs = [{"UUID": uuids[i], "Path": "/dev/nvidia%d" % i, "Model": "Tesla K80",
             "PCI": {"BusID": buses[i], "Bandwidth": 15760}, "Topology": None,
             "Memory": {"Global": 11439}} for i in range(4)]
    return {"Version": {"Driver": "384.111", "CUDA": "9.0"}, "Devices": devs}


def check_against_reference(name, doc):
    path = "/root/reference/nvidiagpuplugin/gpu/nvidia/nvidia_gpu_manager_test.go"
    if not os.path.exists(path):
        print("  (reference not present: %s not cross-checked)" % name)
        return
    src = open(path).read()
    m = re.search(name + r"\s*=\s*`(.*?)`", src, re.S)
    ref = json.loads(m.group(1))
    assert ref["Version"] == doc["Version"]
    assert len(ref["Devices"]) == len(doc["Devices"])
    for r, d in zip(ref["Devices"], doc["Devices"]):
        assert r["UUID"] == d["UUID"] and r["Path"] == d["Path"] and r["Model"] == d["Model"]
        assert r["PCI"]["BusID"] == d["PCI"]["BusID"] and r["PCI"]["Bandwidth"] == d["PCI"]["Bandwidth"]
        assert r["Memory"]["Global"] == d["Memory"]["Global"]
        assert r["Topology"] == d["Topology"], (r["Topology"], d["Topology"])
    print("  %s matches the reference's test constant field by field" % name)


def two_level_shapes(n=8):
    """All multisets of multisets of positive ints summing to n (223 for n=8)."""
    def partitions(m, mx):
        if m == 0:
            yield ()
            return
        for first in range(min(m, mx), 0, -1):
            for rest in partitions(m - first, first):
                yield (first,) + rest
    groups = {m: list(partitions(m, m)) for m in range(1, n + 1)}

    def build(m, last):
        if m == 0:
            yield ()
            return
        for size in range(m, 0, -1):
            for g in groups[size]:
                key = (size, g)
                if last is not None and key > last:
                    continue
                for rest in build(m - size, key):
                    yield (g,) + rest
    return [tuple(s) for s in build(n, None)]


def agree_set():
    W = ob.DEFAULT_WEIGHTS
    shapes = two_level_shapes(8)
    total = agree = 0
    divergent = []
    for shp in shapes:
        tree = oa.add_to_node(None, oa.shape_to_resources([list(g) for g in shp]), "gpugrp", "cards", 1)
        M = oa.tree_to_matrix(tree)
        for k in range(1, 9):
            total += 1
            greedy = oa.greedy_fill_mask(tree, k)
            gcost = sum(int(W[M[i * 8 + j]]) for i in range(8) for j in range(i + 1, 8)
                        if (greedy >> i) & 1 and (greedy >> j) & 1)
            best = ob.node_key(M, 0xFF, k) >> 8
            if gcost == best:
                agree += 1
            else:
                divergent.append({"shape": [list(g) for g in shp], "sorted_shape": tree.shape(), "k": k,
                                  "greedy_cost": gcost, "optimal_cost": best})
    return {"shapes": len(shapes), "cases": total, "agree": agree, "divergent": divergent}


def main():
    for name, doc, fn in (("jsonString", titanx_doc(), "gpus_titanx.json"), ("jsonString2", k80_doc(), "gpus_k80.json")):
        check_against_reference(name, doc)
        with open(os.path.join(HERE, fn), "w") as f:
            json.dump(doc, f, indent=1)

    vec = {}
    for tag, (topo, free, pods) in {
        "c1": synth.gen_c1(),
        "c2": synth.gen_c2(N=3000, P=64),
        "c3": synth.gen_c3(N=2000, P=96),
        "c4": synth.gen_c4(N=2500, P=96),
    }.items():
        vec[tag + "_topo"], vec[tag + "_free"], vec[tag + "_pods"] = topo.astype(np.int8), free, pods
        vec[tag + "_keys"] = ob.score_batch(topo, free, pods)
    np.savez_compressed(os.path.join(HERE, "oracle_b_vectors.npz"), **vec)

    res = agree_set()
    with open(os.path.join(HERE, "agree_set.json"), "w") as f:
        json.dump(res, f, indent=1)
    print("agree-set: %d shapes, %d/%d cases agree, %d divergent" % (res["shapes"], res["agree"], res["cases"], len(res["divergent"])))
    ks = [d["k"] for d in res["divergent"]]
    print("divergent by k:", {k: ks.count(k) for k in sorted(set(ks))})


if __name__ == "__main__":
    main()
