"""Memory-aware placement (per-GPU memory as a pod constraint): the oracle's
definition on CPU, and every K1 variant + K1m through the C ABI on the GPU."""
import numpy as np
import pytest

from kubegpu_b200 import _lib, synth


def test_oracle_mem_semantics(oracle_b):
    ob = oracle_b
    t = synth.shape_matrix([[2, 2], [2, 2]])
    topo = np.stack([t, t])
    free = np.array([0xFF, 0xFF], np.int32)
    mem = np.array([[16, 16, 80, 80, 16, 16, 80, 80], [80] * 8], np.int32) * 1024
    pods = np.array([[2, 0, 0, 0], [2, 1, 0, 40_000], [4, 2, 0, 40_000], [2, 3, 0, 90_000], [2, 4, 0, 16_384]], np.int32)
    keys = ob.score_batch(topo, free, pods, mem=mem)
    assert ob.unpack_key(keys[0]) == (2, 0, 0x03)              # no requirement: tight pair on node 0
    assert ob.unpack_key(keys[1]) == (2, 0, 0x0C)              # needs 40 GB: node 0's big pair {2,3}
    assert ob.unpack_key(keys[2]) == (36, 1, 0x0F)             # 4 big GPUs: node 0 would have to cross sockets (2*8+... ) -> node 1's socket
    assert keys[3] == ob.NO_FIT                                # nobody has 90 GB
    assert ob.unpack_key(keys[4]) == (2, 0, 0x03)              # exactly 16 GiB is enough (>=)
    # without the mem array the requirement is ignored
    assert (ob.score_batch(topo, free, pods)[:5] == ob.score_batch(topo, free, pods * np.array([1, 1, 1, 0], np.int32))).all()


@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_mem_c_vs_python_twin(oracle_b, seed):
    rng = np.random.default_rng(seed)
    N, P = 30, 40
    topo = rng.integers(0, 16, size=(N, 64)).astype(np.int32)
    free = rng.integers(0, 256, size=N).astype(np.int32)
    mem = rng.choice(np.array([8, 16, 32, 80], np.int32) * 1024, size=(N, 8)).astype(np.int32)
    pods = synth.make_pods(rng.integers(0, 9, size=P).astype(np.int32))
    pods[:, 3] = rng.choice(np.array([0, -5, 10_000, 20_000, 40_000, 100_000], np.int32), size=P)
    a = oracle_b.score_batch(topo, free, pods, mem=mem)
    b = oracle_b.score_batch_py(topo, free, pods, mem=mem)
    c = oracle_b.score_batch(topo, free, pods, mem=mem, fast=True, nthreads=3)
    assert (a == b).all() and (a == c).all()


def test_synth_c6_shapes():
    topo, free, mem, pods = synth.gen_c6(N=2000, P=300)
    assert mem.shape == (2000, 8) and mem.dtype == np.int32 and set(np.unique(mem)) <= set(synth.GPU_MEM_CLASSES_MIB)
    uniform = (mem == mem[:, :1]).all(axis=1).mean()
    assert 0.7 < uniform < 0.95
    assert set(np.unique(pods[:, 3])) <= set(synth.POD_MIN_MEM_CHOICES_MIB) and (pods[:, 3] > 0).mean() > 0.3
    assert (synth.gen_gpu_memory(100, node_start=500) == synth.gen_gpu_memory(700)[500:600]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [_lib.VARIANT_AUTO, _lib.VARIANT_LANE_PER_NODE, _lib.VARIANT_WARP_PER_PAIR, _lib.VARIANT_MEMO_BY_K,
                                     _lib.VARIANT_TILE_MEMO, _lib.VARIANT_SPARSE])
def test_gpu_memory_aware_parity(oracle_b, variant):
    from kubegpu_b200.scorer import KgpuError, Scorer
    topo, free, mem, pods = synth.gen_c6(N=30_001, P=700)
    want = oracle_b.score_batch(topo, free, pods, mem=mem, node_id_base=5, fast=True, nthreads=8)
    with Scorer((0,)) as s:
        s.set_variant(variant)
        s.upload_nodes(topo, free, node_id_base=5)
        # before the memory is uploaded every GPU is unconstrained: min_mem changes nothing
        assert (s.score_batch(pods) == oracle_b.score_batch(topo, free, pods, node_id_base=5, fast=True, nthreads=8)).all()
        s.upload_gpu_memory(mem)
        got = s.score_batch(pods)
        assert (got == want).all()
        assert (got != oracle_b.score_batch(topo, free, pods, node_id_base=5, fast=True, nthreads=8)).any()
        # only memory-constrained pods / only unconstrained pods / device-buffer entry point
        only = pods[pods[:, 3] > 0]
        assert (s.score_batch(only) == oracle_b.score_batch(topo, free, only, mem=mem, node_id_base=5, fast=True, nthreads=8)).all()
        import torch
        d_pods = torch.from_numpy(pods).cuda()
        d_keys = torch.empty(len(pods), dtype=torch.int64, device="cuda")
        s.score_batch_device(d_pods.data_ptr(), len(pods), d_keys.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert (d_keys.cpu().numpy().view(np.uint64) == want).all()
        nomem = pods.copy()
        nomem[:, 3] = 0
        d_pods.copy_(torch.from_numpy(nomem))
        s.score_batch_device(d_pods.data_ptr(), len(pods), d_keys.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert (d_keys.cpu().numpy().view(np.uint64) == oracle_b.score_batch(topo, free, nomem, node_id_base=5, fast=True, nthreads=8)).all()
        d_keys.zero_()                     # and with the caller's promise that no pod has min_mem
        s.score_batch_device(d_pods.data_ptr(), len(pods), d_keys.data_ptr(), torch.cuda.current_stream().cuda_stream,
                             _lib.BATCH_NO_MIN_MEM)
        torch.cuda.synchronize()
        assert (d_keys.cpu().numpy().view(np.uint64) == oracle_b.score_batch(topo, free, nomem, node_id_base=5, fast=True, nthreads=8)).all()
        # per-pair queries honour the requirement too
        idx = np.array([0, 7, 100, 12_345, 30_000], dtype=np.int64)
        ks = np.array([1, 2, 3, 4, 2], dtype=np.int32)
        mm = np.array([0, 20_000, 40_000, 100_000, 8_000], dtype=np.int32)
        gotp = s.score_pairs(idx, ks, mm)
        for i, k, m, g in zip(idx, ks, mm, gotp):
            fm = int(free[i]) & sum(1 << b for b in range(8) if mem[i][b] >= m) if m > 0 else int(free[i])
            assert int(g) == oracle_b.node_key(topo[i], fm, int(k))
        # one node's memory changes
        mem2 = mem.copy()
        mem2[12_345] = 184_320
        s.update_gpu_memory(12_345, mem2[12_345])
        assert (s.score_batch(pods) == oracle_b.score_batch(topo, free, pods, mem=mem2, node_id_base=5, fast=True, nthreads=8)).all()
        with pytest.raises(ValueError):
            s.upload_gpu_memory(mem[:10])          # the wrapper checks the shape before the C call
        bad = mem.copy()
        bad[3, 3] = -1
        with pytest.raises(KgpuError):
            s.upload_gpu_memory(bad)
        many = pods[:16].copy()
        many[:8, 0], many[:8, 3] = 1, np.arange(1, 9) * 1000
        with pytest.raises(KgpuError):
            s.place_batch(many)            # sequential path: at most 7 distinct requirements per batch


@pytest.mark.gpu
def test_gpu_memory_aware_multi_device(oracle_b):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from kubegpu_b200.scorer import Scorer
    topo, free, mem, pods = synth.gen_c6(N=20_000, P=400)
    with Scorer((0, 1)) as s:
        s.upload_nodes(topo, free)
        s.upload_gpu_memory(mem)
        assert (s.score_batch(pods) == oracle_b.score_batch(topo, free, pods, mem=mem, fast=True, nthreads=8)).all()
