"""Python face of Oracle B (oracle/oracle_b.c) + an independent pure-Python twin.

TEST INFRASTRUCTURE ONLY -- see the header of oracle_b.c.  Nothing under
``kubegpu_b200/`` may import this module.

``score_batch`` / ``score_batch_fast`` call the C restatement through ctypes
(built by ``oracle/Makefile`` into ``oracle/_build/liboracle_b.so``).
``node_key_py`` / ``score_batch_py`` restate the same definition (SURVEY.md
8(c) "Oracle B") a second time in plain Python with ``itertools.combinations``
instead of Gosper's hack, so the C file is checked by something that shares no
code with it (tests/test_oracle_b.py).
"""
from __future__ import annotations

import ctypes
import itertools
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_b.so")
NO_FIT = np.uint64(0xFFFFFFFFFFFFFFFF)
NODE_NO_FIT = 0xFFFFFFFF
DEFAULT_WEIGHTS = np.array([64, 32, 16, 8, 4, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=np.int32)

_lib: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> str:
    """Compile oracle_b.c with gcc (no GPU, no reference sources involved)."""
    src = os.path.join(_HERE, "oracle_b.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i32p, u64p = ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint64)
        L.kgpu_oracle_node_key.restype = ctypes.c_uint32
        L.kgpu_oracle_node_key.argtypes = [i32p, ctypes.c_int32, ctypes.c_int, i32p]
        L.kgpu_oracle_subset_cost.restype = ctypes.c_uint32
        L.kgpu_oracle_subset_cost.argtypes = [i32p, ctypes.c_uint, i32p]
        L.kgpu_oracle_score_batch.restype = None
        L.kgpu_oracle_score_batch.argtypes = [i32p, i32p, ctypes.c_int64, ctypes.c_int64, i32p,
                                              ctypes.c_int64, i32p, u64p]
        L.kgpu_oracle_score_batch_fast.restype = None
        L.kgpu_oracle_score_batch_fast.argtypes = [i32p, i32p, ctypes.c_int64, ctypes.c_int64, i32p,
                                                   ctypes.c_int64, i32p, u64p, ctypes.c_int]
        L.kgpu_oracle_score_batch_mem.restype = None
        L.kgpu_oracle_score_batch_mem.argtypes = [i32p, i32p, i32p, ctypes.c_int64, ctypes.c_int64, i32p,
                                                  ctypes.c_int64, i32p, u64p]
        L.kgpu_oracle_score_batch_fast_mem.restype = None
        L.kgpu_oracle_score_batch_fast_mem.argtypes = [i32p, i32p, i32p, ctypes.c_int64, ctypes.c_int64, i32p,
                                                       ctypes.c_int64, i32p, u64p, ctypes.c_int]
        L.kgpu_oracle_score_batch_memo.restype = None
        L.kgpu_oracle_score_batch_memo.argtypes = [i32p, i32p, ctypes.c_int64, ctypes.c_int64, i32p,
                                                   ctypes.c_int64, i32p, u64p, ctypes.c_int]
        for fn in (L.kgpu_oracle_place_batch, L.kgpu_oracle_place_batch_plain, L.kgpu_oracle_place_batch_tiled):
            fn.restype = None
            fn.argtypes = [i32p, i32p, ctypes.c_int64, ctypes.c_int64, i32p, ctypes.c_int64, i32p, u64p]
        L.kgpu_oracle_place_batch_mem.restype = None
        L.kgpu_oracle_place_batch_mem.argtypes = [i32p, i32p, i32p, ctypes.c_int64, ctypes.c_int64, i32p, ctypes.c_int64, i32p, u64p]
        L.kgpu_oracle_reduce_shards.restype = None
        L.kgpu_oracle_reduce_shards.argtypes = [u64p, ctypes.c_int, ctypes.c_int64, u64p]
        _lib = L
    return _lib


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def node_key(M, free_mask: int, k: int, W=DEFAULT_WEIGHTS) -> int:
    M, W = _i32(M).reshape(64), _i32(W)
    return int(lib().kgpu_oracle_node_key(_p(M, ctypes.c_int32), int(free_mask), int(k), _p(W, ctypes.c_int32)))


def score_batch(topo, free_mask, pods, W=DEFAULT_WEIGHTS, node_id_base: int = 0,
                fast: bool = False, nthreads: int = 1, mem=None) -> np.ndarray:
    """keys[P] (uint64) for nodes topo[N,64]/free_mask[N] and pods[P,4].  With mem[N,8] (MiB per
    GPU) the pods' min_mem field (pods[:,3]) is honoured; without it it is ignored."""
    topo, free_mask, pods, W = _i32(topo), _i32(free_mask), _i32(pods), _i32(W)
    N, P = free_mask.shape[0], pods.shape[0]
    assert topo.size == 64 * N and pods.size == 4 * P and W.size == 16
    out = np.empty(P, dtype=np.uint64)
    if mem is not None:
        mem = _i32(mem)
        assert mem.size == 8 * N
        args = [_p(topo, ctypes.c_int32), _p(free_mask, ctypes.c_int32), _p(mem, ctypes.c_int32), N, int(node_id_base),
                _p(pods, ctypes.c_int32), P, _p(W, ctypes.c_int32), _p(out, ctypes.c_uint64)]
        if fast:
            lib().kgpu_oracle_score_batch_fast_mem(*args, int(nthreads))
        else:
            lib().kgpu_oracle_score_batch_mem(*args)
        return out
    args = [_p(topo, ctypes.c_int32), _p(free_mask, ctypes.c_int32), N, int(node_id_base),
            _p(pods, ctypes.c_int32), P, _p(W, ctypes.c_int32), _p(out, ctypes.c_uint64)]
    if fast:
        lib().kgpu_oracle_score_batch_fast(*args, int(nthreads))
    else:
        lib().kgpu_oracle_score_batch(*args)
    return out


def score_batch_memo(topo, free_mask, pods, W=DEFAULT_WEIGHTS, node_id_base: int = 0, nthreads: int = 1) -> np.ndarray:
    """Snapshot scoring memoised by k (the CPU twin of the GPU's memo_by_k variant): best[k] over all nodes once,
    every pod reads best[k_p].  Pods with min_mem > 0 are not memoisable (NO_FIT here)."""
    topo, free_mask, pods, W = _i32(topo), _i32(free_mask), _i32(pods), _i32(W)
    N, P = free_mask.shape[0], pods.shape[0]
    out = np.empty(P, dtype=np.uint64)
    lib().kgpu_oracle_score_batch_memo(_p(topo, ctypes.c_int32), _p(free_mask, ctypes.c_int32), N, int(node_id_base),
                                       _p(pods, ctypes.c_int32), P, _p(W, ctypes.c_int32), _p(out, ctypes.c_uint64), int(nthreads))
    return out


def place_batch(topo, free_mask, pods, W=DEFAULT_WEIGHTS, node_id_base: int = 0, plain: bool = False, mem=None, tiled: bool = False):
    """K3 twin: sequential stateful placement.  Returns (keys[P], free_mask_after[N]).  With mem[N,8]
    the pods' min_mem (pods[:,3]) is honoured (plain loop).  tiled: the same two-level minima as the GPU
    kernel (the fair single-thread CPU baseline of the sequential path)."""
    topo, pods, W = _i32(topo), _i32(pods), _i32(W)
    free_after = np.array(free_mask, dtype=np.int32, copy=True)
    N, P = free_after.shape[0], pods.shape[0]
    out = np.empty(P, dtype=np.uint64)
    if mem is not None:
        mem = _i32(mem)
        assert mem.size == 8 * N
        lib().kgpu_oracle_place_batch_mem(_p(topo, ctypes.c_int32), _p(free_after, ctypes.c_int32), _p(mem, ctypes.c_int32), N,
                                          int(node_id_base), _p(pods, ctypes.c_int32), P, _p(W, ctypes.c_int32),
                                          _p(out, ctypes.c_uint64))
        return out, free_after
    fn = lib().kgpu_oracle_place_batch_plain if plain else lib().kgpu_oracle_place_batch_tiled if tiled else lib().kgpu_oracle_place_batch
    fn(_p(topo, ctypes.c_int32), _p(free_after, ctypes.c_int32), N, int(node_id_base), _p(pods, ctypes.c_int32), P,
       _p(W, ctypes.c_int32), _p(out, ctypes.c_uint64))
    return out, free_after


def reduce_shards(gathered) -> np.ndarray:
    g = np.ascontiguousarray(gathered, dtype=np.uint64)
    G, P = g.shape
    out = np.empty(P, dtype=np.uint64)
    lib().kgpu_oracle_reduce_shards(_p(g, ctypes.c_uint64), G, P, _p(out, ctypes.c_uint64))
    return out


# ---- independent pure-Python twin (small cases only) ---------------------------
def node_key_py(M: Sequence[int], free_mask: int, k: int, W: Sequence[int] = DEFAULT_WEIGHTS) -> int:
    fm = int(free_mask) & 0xFF
    if k < 0 or k > 8:
        return NODE_NO_FIT
    free_bits = [i for i in range(8) if (fm >> i) & 1]
    best = NODE_NO_FIT
    for combo in itertools.combinations(free_bits, k):
        cost = sum(int(W[int(M[i * 8 + j]) & 15]) for i, j in itertools.combinations(combo, 2))
        key = (cost << 8) | sum(1 << i for i in combo)
        best = min(best, key)
    return best


def score_batch_py(topo, free_mask, pods, W=DEFAULT_WEIGHTS, node_id_base: int = 0, mem=None) -> np.ndarray:
    topo = np.asarray(topo).reshape(-1, 64)
    out = np.full(len(pods), NO_FIT, dtype=np.uint64)
    for p, pod in enumerate(np.asarray(pods).reshape(-1, 4)):
        best = int(NO_FIT)
        for n in range(topo.shape[0]):
            fm = int(free_mask[n])
            if mem is not None and int(pod[3]) > 0:
                fm &= sum(1 << i for i in range(8) if int(np.asarray(mem).reshape(-1, 8)[n][i]) >= int(pod[3]))
            nk = node_key_py(topo[n], fm, int(pod[0]), W)
            if nk != NODE_NO_FIT:
                best = min(best, ((nk >> 8) << 40) | ((node_id_base + n) << 8) | (nk & 0xFF))
        out[p] = np.uint64(best)
    return out


def unpack_key(key: int):
    """(cost, node_id, mask) or None for NO_FIT."""
    key = int(key)
    if key == int(NO_FIT):
        return None
    return key >> 40, (key >> 8) & 0xFFFFFFFF, key & 0xFF
