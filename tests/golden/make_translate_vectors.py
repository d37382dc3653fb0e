#!/usr/bin/env python3
"""Derived test vectors for KubeDevice-API's ``resource.TranslateResource`` (absent from /root/reference: un-vendored
dependency without a pinned version; call sites gpuschedulerplugin/gpu.go:55,58).

No reference test reaches the function, so it cannot be pinned by a golden of its own.  What CAN be pinned is every
constraint the reference's own code puts on it.  Each case below names the reference lines it is derived from and
states its expectation either as

  "properties"  facts every implementation consistent with those lines must satisfy (PINNED by the reference), or
  "exact"       the full output map of this repository's restatement, which ALSO fixes what the reference leaves
                open (numbering of fresh groups, visiting order).  Flagged ``"pinned": false``.

tests/test_translate_resource_pin.py checks Oracle A (oracle/oracle_a.py) and the C++ host mirror against all of it.
Run:  python tests/golden/make_translate_vectors.py   (rewrites translate_resource_vectors.json from Oracle A and
re-checks the properties while doing so)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_a as oa  # noqa: E402

P = oa.DEVICE_GROUP_PREFIX
TWO_LEVEL_NODE = {P + "/gpugrp1/A/gpugrp0/B/gpu/GPU0/cards": 1}     # gpu_scheduler.go:22-24


def cases():
    flat3 = {P + "/gpu/%d/cards" % i: 1 for i in range(3)}
    uuid_flat = {}
    for u in ("GPU-aa", "GPU-bb"):
        uuid_flat[P + "/gpu/%s/cards" % u] = 1
        uuid_flat[P + "/gpu/%s/memory" % u] = 12_000_000_000
    grouped = dict(oa.shape_to_resources([[2, 2], [4]]))                      # gpu_test.go:24-33 layout
    titan = {P + "/gpugrp1/%d/gpugrp0/%d/gpu/GPU0%d/cards" % (i // 4, i // 2, i): 1 for i in range(8)}   # nvidia_gpu_manager_test.go:120-130
    return [
        {"name": "stage_not_advertised_is_identity",
         "derivation": "gpu.go:15 'translates GPU resources to max level advertised by the node': a node that does not "
                       "advertise <this_stage> leaves the requests alone (and reports 'not modified', gpu.go:55-59 OR the flags)",
         "call": ["translate_resource", {P + "/gpu/0/cards": 1}, flat3, "gpugrp0", "gpu"],
         "properties": ["identity", "not_modified"]},
        {"name": "fully_grouped_requests_pass_through",
         "derivation": "gpu_scheduler.go:21-28: AddNode runs the advertised names through TranslateGPUResources before the tree "
                       "cache sees them; the node agent advertises gpugrp1/<a>/gpugrp0/<b>/gpu/<id> (nvidia_gpu_manager.go:178-211), "
                       "and gpu_test.go:14-33 / nvidia_gpu_manager_test.go:120-130 expect exactly those groups in the tree: "
                       "names that already carry both levels must come back unchanged",
         "call": ["translate_gpu_resources", 8, TWO_LEVEL_NODE, titan],
         "properties": ["identity"]},
        {"name": "flat_requests_get_both_levels_each_gpu_its_own_groups",
         "derivation": "gpu_scheduler.go:13 'auto topology generation \"0\" means default (everything in its own group)' + "
                       "gpu_scheduler.go:20 'force translation to two levels' + gpu.go:131 (addToNode's regexp needs "
                       ".*/gpugrp1/<a>/.*/cards and .*/gpugrp0/<b>/.*/cards, gpugrp1 outside gpugrp0): every gpu/<i>/cards key "
                       "gains a gpugrp0 and a gpugrp1 level, distinct per GPU, values and count preserved",
         "call": ["translate_gpu_resources", 3, TWO_LEVEL_NODE, flat3],
         "properties": ["two_levels", "own_groups", "count_and_values_preserved", "gpu_ids_preserved"]},
        {"name": "stage_one_creates_the_missing_card_requests",
         "derivation": "gpu.go:31-53: a container asking for n GPUs with no card requests gets gpu/<0..n-1>/cards = 1 "
                       "(indices continue after the largest integer index present), then both group levels",
         "call": ["translate_gpu_resources", 2, TWO_LEVEL_NODE, {}],
         "properties": ["two_levels", "own_groups", "n_cards:2"]},
        {"name": "cards_and_memory_of_one_gpu_stay_together",
         "derivation": "nvidia_gpu_manager.go:204-211 advertises <name>/memory next to <name>/cards for every GPU and AddNode "
                       "translates the whole Allocatable list: both keys of one GPU id must land in the same groups, or the "
                       "tree (cards) and the memory resource would describe different GPUs",
         "call": ["translate_gpu_resources", 2, TWO_LEVEL_NODE, uuid_flat],
         "properties": ["two_levels", "own_groups", "same_groups_for_all_keys_of_a_gpu", "count_and_values_preserved"]},
        {"name": "add_node_on_flat_names_gives_the_singleton_tree",
         "derivation": "gpu_scheduler.go:21-28 + gpu.go:129-161: AddNode on three flat GPUs must produce a tree the cache can "
                       "score: 3 gpugrp1 children with one gpugrp0 child of 1 card each -- the shape the node agent itself "
                       "gives GPUs without topology (nvidia_gpu_manager_test.go:140-145: gpugrp1/<i>/gpugrp0/<i>)",
         "call": ["add_node_tree_shape", 3, flat3],
         "properties": ["tree_shape:[3, [[1, [[1, []]]], [1, [[1, []]]], [1, [[1, []]]]]]"]},
        {"name": "mixed_grouped_and_flat_requests",
         "derivation": "gpu.go:31-59 with a container that already holds one grouped card request and needs two more: the "
                       "grouped one is untouched, the new ones get fresh groups that do not collide with it",
         "call": ["translate_gpu_resources", 3, TWO_LEVEL_NODE, {P + "/gpugrp1/0/gpugrp0/0/gpu/0/cards": 1}],
         "properties": ["two_levels", "own_groups", "n_cards:3", "keeps:" + P + "/gpugrp1/0/gpugrp0/0/gpu/0/cards"]},
        {"name": "reference_fixture_T1_through_add_node",
         "derivation": "gpu_test.go:14-23 through AddNode: unchanged names, tree [[2,2],[2,2]] with TreeScore 12.0 (SURVEY 8(a) a4)",
         "call": ["add_node_tree_shape", 8, dict(oa.shape_to_resources([[2, 2], [2, 2]]))],
         "properties": ["tree_shape:[8, [[4, [[2, []], [2, []]]], [4, [[2, []], [2, []]]]]]"]},
        {"name": "grouped_shape_passes_stage_by_stage", "derivation": "as fully_grouped_requests_pass_through, per stage",
         "call": ["translate_resource", TWO_LEVEL_NODE, grouped, "gpugrp0", "gpu"], "properties": ["identity", "not_modified"]},
    ]


def run(call):
    fn = call[0]
    if fn == "translate_resource":
        mod, out = oa.translate_resource(dict(call[1]), dict(call[2]), call[3], call[4])
        return {"modified": bool(mod), "out": dict(out)}
    if fn == "translate_gpu_resources":
        return {"out": dict(oa.translate_gpu_resources(call[1], dict(call[2]), dict(call[3])))}
    if fn == "add_node_tree_shape":
        s = oa.NvidiaGPUScheduler()
        ni = oa.NodeInfo(Allocatable=dict(call[2]), KubeAlloc={oa.RESOURCE_GPU: call[1]})
        s.AddNode("n", ni)
        tree = oa.add_to_node(None, ni.Allocatable, "gpugrp", "cards", 1)
        return {"out": dict(ni.Allocatable), "tree_shape": json.loads(json.dumps(tree.shape()))}
    raise ValueError(fn)


def check_properties(case, res):
    import re
    inp = case["call"][2] if case["call"][0] != "translate_gpu_resources" else case["call"][3]
    out = res["out"]
    rx = re.compile(r"^(.*)/gpugrp1/([^/]*)/gpugrp0/([^/]*)/gpu/([^/]*)/(cards|memory)$")
    for prop in case["properties"]:
        if prop == "identity":
            assert out == inp, (case["name"], out)
        elif prop == "not_modified":
            assert res["modified"] is False
        elif prop == "two_levels":
            assert all(rx.match(k) for k in out if "/gpu/" in k), (case["name"], sorted(out))
        elif prop == "own_groups":
            gpus = {}
            for k in out:
                m = rx.match(k)
                if m:
                    gpus.setdefault(m.group(4), set()).add((m.group(2), m.group(3)))
            pairs = [next(iter(v)) for v in gpus.values()]
            assert len({p[1] for p in pairs if True}) >= 1
            fresh = [g for g in gpus if not any(g == rx.match(k).group(4) for k in inp if rx.match(k))]
            assert len({gpus[g].__iter__().__next__() for g in fresh}) == len(fresh)              # distinct (grp1, grp0) per new GPU
            assert len({next(iter(gpus[g]))[0] for g in fresh}) == len(fresh)                     # ... and distinct grp1: its own group at BOTH levels
        elif prop == "same_groups_for_all_keys_of_a_gpu":
            gpus = {}
            for k in out:
                m = rx.match(k)
                gpus.setdefault(m.group(4), set()).add((m.group(2), m.group(3)))
            assert all(len(v) == 1 for v in gpus.values())
        elif prop == "count_and_values_preserved":
            assert sorted(out.values()) == sorted(inp.values()) and len(out) == len(inp)
        elif prop == "gpu_ids_preserved":
            assert sorted(rx.match(k).group(4) for k in out) == sorted(k.split("/gpu/")[1].split("/")[0] for k in inp)
        elif prop.startswith("n_cards:"):
            assert sum(1 for k in out if k.endswith("/cards")) == int(prop.split(":")[1])
        elif prop.startswith("keeps:"):
            assert prop.split(":", 1)[1] in out
        elif prop.startswith("tree_shape:"):
            assert res["tree_shape"] == json.loads(prop.split(":", 1)[1]), res["tree_shape"]
        else:
            raise ValueError(prop)


def main():
    vec = []
    for c in cases():
        res = run(c["call"])
        check_properties(c, res)
        c = dict(c)
        c["exact"] = res
        c["pinned"] = {"properties": True, "exact": c["properties"] in (["identity"], ["identity", "not_modified"])}
        vec.append(c)
    with open(os.path.join(HERE, "translate_resource_vectors.json"), "w") as f:
        json.dump(vec, f, indent=1, sort_keys=True)
    print("%d cases written; properties hold for Oracle A" % len(vec))


if __name__ == "__main__":
    main()
