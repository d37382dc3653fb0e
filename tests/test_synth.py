"""Synthetic generator: deterministic, shardable, right distributions."""
import hashlib

import numpy as np

from kubegpu_b200 import synth


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()[:16]


def test_counter_based_and_shardable():
    topo, free, pods = synth.gen_c2(N=1000, P=50)
    t2, f2, _ = synth.gen_c2(N=300, P=0, node_start=400)
    assert (topo[400:700] == t2).all() and (free[400:700] == f2).all()
    t4, f4, _ = synth.gen_c4(N=64, P=0, node_start=100)
    ta, fa, _ = synth.gen_c4(N=200, P=0)
    assert (ta[100:164] == t4).all() and (fa[100:164] == f4).all()
    assert _digest(*synth.gen_c2(N=1000, P=50)) == _digest(topo, free, pods)


def test_domains():
    topo, free, pods = synth.gen_c2(N=4000, P=400)
    assert topo.dtype == np.int32 and topo.shape == (4000, 64) and pods.shape == (400, 4)
    assert set(np.unique(pods[:, 0])) == {1, 2, 4, 8} and (pods[:, 1] == np.arange(400)).all()
    assert free.min() >= 0 and free.max() <= 255 and set(np.unique(topo)) <= {0, 1, 3, 5}
    m = topo.reshape(-1, 8, 8)
    assert (m == m.transpose(0, 2, 1)).all() and (m[:, np.arange(8), np.arange(8)] == 0).all()
    _, _, pods3 = synth.gen_c3(N=10, P=800)
    assert set(np.unique(pods3[:, 0])) == set(range(1, 9))
    t4, f4, _ = synth.gen_c4(N=3000, P=8)
    m4 = t4.reshape(-1, 8, 8)
    assert (m4 == m4.transpose(0, 2, 1)).all() and t4.max() == 12 and t4.min() == 0
    small = (m4[:, 0, 7] == 0)                      # n_gpus == 4 nodes: GPUs 4..7 absent
    assert 0.3 < small.mean() < 0.7 and (f4[small] < 16).all()
    nvl = (t4[~small] >= 7).sum() / (t4[~small] > 0).sum()
    assert 0.25 < nvl < 0.35


def test_c1_fixture_shapes():
    topo, free, pods = synth.gen_c1()
    assert topo.shape == (16, 64) and pods[:, 0].tolist() == [1, 2, 3, 4]
    assert free[:4].tolist() == [0xFF, 0xFF, 0x0F, 0x00]
