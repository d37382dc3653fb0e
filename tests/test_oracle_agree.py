"""Ties Oracle A (reference's greedy tree fill) to Oracle B (min pairwise link cost):
SURVEY.md 8(c) "How A and B are tied together"."""
import json
import os
import sys

import pytest

from oracle import oracle_a as oa

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402


def test_agree_set_matches_committed_and_survey(oracle_b, golden_dir):
    res = make_golden.agree_set()
    with open(os.path.join(golden_dir, "agree_set.json")) as f:
        committed = json.load(f)
    assert res["shapes"] == committed["shapes"] == 223
    assert res["cases"] == 1784 and res["agree"] == committed["agree"] == 1759
    assert json.loads(json.dumps(res["divergent"])) == committed["divergent"]
    ks = [d["k"] for d in res["divergent"]]
    assert ks.count(2) == 13 and all(d["greedy_cost"] > d["optimal_cost"] for d in res["divergent"])
    # restricted to BASELINE's k in {1,2,4,8}: 879 of 892, all divergences at k=2
    base = [d for d in res["divergent"] if d["k"] in (1, 2, 4, 8)]
    assert len(base) == 13 and {d["k"] for d in base} == {2}


@pytest.mark.parametrize("shape", [[[2, 2], [2, 2]], [[4], [2, 2]], [[8]], [[4], [4]], [[4, 4]], [[1]] * 8])
def test_reference_fixture_shapes_are_in_the_agree_set(oracle_b, shape):
    """Every shape in the reference's fixtures and in config C2: greedy == optimal for all k."""
    tree = oa.add_to_node(None, oa.shape_to_resources(shape), "gpugrp", "cards", 1)
    M = oa.tree_to_matrix(tree)
    W = oracle_b.DEFAULT_WEIGHTS
    for k in range(1, 9):
        g = oa.greedy_fill_mask(tree, k)
        gcost = sum(int(W[M[i * 8 + j]]) for i in range(8) for j in range(i + 1, 8) if (g >> i) & 1 and (g >> j) & 1)
        key = oracle_b.node_key(M, 0xFF, k)
        assert gcost == key >> 8
        # with everything free the greedy mask is the lowest-index optimum as well
        assert g == key & 0xFF


def test_golden_placement_equals_oracle_b_choice(oracle_b):
    """The reference's TestTree golden (k=3 on [[4],[2,2]] -> grp0/0 gpu 0,1,2) is exactly
    Oracle B's mask 0b111 on the matrix of that tree."""
    tree = oa.add_to_node(None, oa.shape_to_resources([[2, 2], [4]]), "gpugrp", "cards", 1)
    assert tree.shape() == (8, ((4, ((4, ()),)), (4, ((2, ()), (2, ())))))
    key = oracle_b.node_key(oa.tree_to_matrix(tree), 0xFF, 3)
    assert key & 0xFF == oa.greedy_fill_mask(tree, 3) == 0b111
