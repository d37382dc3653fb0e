#!/usr/bin/env python3
"""bench.py -- placements/sec of the topology-aware GPU placement scorer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "C2"): 100,000 nodes x 10,000 pods,
8 GPUs per node, k in {1,2,4,8}, seeded synthetic.  One "step" = one pass of the hot
path over the whole pod batch: every pod scored against every node, best
(cost, node, mask) per pod.  With N > 1 GPUs the node list is sharded contiguously
over the ranks (strong scaling: the cluster and the pod batch stay fixed), each rank
scores its shard, one NCCL all-gather exchanges the per-pod bests and K2 picks the
final key on every rank.

Prints ONE JSON line on rank 0.  Keys beyond the base contract:
  roofline      algorithmic HBM bytes (260 B per pair + 24 B per pod) / K1 time vs the
                measured copy peak; > 1 is expected and explained in DESIGN.md: a staged
                node is reused for every pod of the block, so DRAM traffic (`traffic`,
                from the committed ncu capture) is far below the algorithmic bytes and the
                kernel is integer-issue bound.
  cpu_baseline  Oracle B (tuned C port, all host cores) on a bounded pod sample.
  variants      the north_star warp-per-pair mapping and the two memoising shortcuts (per-tile
                hoisting, global memoise-by-k), for context only (never the headline).
`--impl reference` times the CPU port of the path instead (the reference itself is Go
with un-vendored dependencies and cannot be built here: DESIGN.md "Oracle").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_NODES = 100_000
N_PODS = 10_000
METRIC = "placements/sec on 100k-node x 10k-pod synthetic"
UNIT = "placements/s"
WORKLOAD = "C2: 100k nodes x 10k pods, 8 GPUs/node, k in {1,2,4,8}, seed 0xB2000001"


def algorithmic_bytes(n_nodes: int, n_pods: int) -> float:
    """260 B per (pod,node) pair + 24 B per pod."""
    return 260.0 * n_nodes * n_pods + 24.0 * n_pods


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic():
    """dram bytes per K1 launch from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch")
    except Exception:
        return None


# ---------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, cuda_index: int):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.handle = None
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.error = repr(e)

    def run(self):
        if self.handle is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": getattr(self, "error", "no samples")}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------
def cpu_baseline(topo, free, pods, target_s: float = 12.0):
    """Oracle B (tuned port, all cores) on a bounded sample: the first S pods of the
    workload against ALL nodes.  Returns the cpu_baseline object."""
    from oracle import oracle_b
    cores = usable_cores()
    S, dt = _sized_cpu_run(oracle_b, topo, free, pods, cores, target_s)
    out = {"value": S / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": "first %d of %d pods x all %d nodes, oracle/oracle_b.c tuned variant, %d threads, %.1f s"
                     % (S, len(pods), len(free), cores, dt)}
    # SURVEY.md 8(d) extras: the same port on ONE thread, and the reference's real per-call path
    # (PodFitsDevice = regex + map + tree + greedy per (node, pod) call) restated in Python.
    t0 = time.perf_counter()
    oracle_b.score_batch(topo, free, pods[:16], fast=True, nthreads=1)
    out["single_thread"] = {"value": 16 / (time.perf_counter() - t0), "unit": UNIT, "sample": "16 pods x all nodes"}
    try:
        out["reference_algorithm_per_call"] = oracle_a_per_call_cost()
    except Exception as e:     # never let the context line break the bench
        out["reference_algorithm_per_call"] = {"error": repr(e)}
    return out


def oracle_a_per_call_cost(n_nodes: int = 1000, n_pods: int = 100):
    """Restatement of the reference algorithm (NOT Go): PodFitsDevice once per (node, pod) pair on a
    1,000-node x 100-pod subsample, as SURVEY.md 8(d) asks; python, single thread."""
    from oracle import oracle_a as oa
    shapes = ([[8]], [[4], [4]], [[2, 2], [2, 2]], [[4, 4]])
    sched = oa.NvidiaGPUScheduler()
    nodes = []
    for i in range(n_nodes):
        ni = oa.NodeInfo(Allocatable=oa.shape_to_resources(shapes[i % 4]), KubeAlloc={oa.RESOURCE_GPU: 8})
        sched.AddNode("n%d" % i, ni)
        nodes.append(ni)
    t0 = time.perf_counter()
    calls = 0
    for p in range(n_pods):
        for ni in nodes:
            pod = oa.PodInfo(RunningContainers={"c": oa.ContainerInfo(Requests={oa.RESOURCE_GPU: (1, 2, 4, 8)[p % 4]})})
            sched.PodFitsDevice(ni, pod, False)
            calls += 1
    dt = time.perf_counter() - t0
    return {"us_per_PodFitsDevice_call": 1e6 * dt / calls, "calls": calls,
            "placements_per_s_if_100k_nodes": 1.0 / (dt / calls * 100_000),
            "note": "oracle/oracle_a.py (Python restatement of gpuschedulerplugin/gpu.go:94-324), 1 thread; "
                    "the Go original would be faster per call but does the same regex/map/tree work per pair"}


def usable_cores() -> int:
    """Threads the CPU arm may really use: the scheduler affinity capped by the cgroup CPU quota
    (the GPU boxes show 128 logical CPUs but a 24-CPU quota)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


def _sized_cpu_run(oracle_b, topo, free, pods, cores, target_s):
    """Grow the pod sample until one run takes >= target_s/2 (per-thread table building is a
    fixed cost, so a linear guess from a tiny probe undershoots).  Returns (S, seconds)."""
    S = min(len(pods), 2 * cores)
    while True:
        t0 = time.perf_counter()
        oracle_b.score_batch(topo, free, pods[:S], fast=True, nthreads=cores)
        dt = max(time.perf_counter() - t0, 1e-6)
        if dt >= target_s / 2 or S >= len(pods):
            return S, dt
        S = int(min(len(pods), max(S + cores, S * min(8.0, target_s / dt))))
        S = min(len(pods), (S + cores - 1) // cores * cores)


def run_reference(args):
    """--impl reference: the CPU port of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from kubegpu_b200 import synth
    from oracle import oracle_b
    topo, free, pods = synth.gen_c2(N_NODES, N_PODS)
    cores = usable_cores()
    oracle_b.lib()
    # size each step's pod sample so that warmup+steps finish in a few minutes (<= ~6 s per step)
    budget = min(6.0, 150.0 / max(1, args.steps + args.warmup))
    S, _ = _sized_cpu_run(oracle_b, topo, free, pods, cores, budget)
    for _ in range(args.warmup):
        oracle_b.score_batch(topo, free, pods[:S], fast=True, nthreads=cores)
    t0 = time.perf_counter()
    for i in range(args.steps):
        off = (i * S) % max(1, N_PODS - S + 1)
        oracle_b.score_batch(topo, free, pods[off:off + S], fast=True, nthreads=cores)
    dt = time.perf_counter() - t0
    value = args.steps * S / dt
    sample = "each step: %d of %d pods x all %d nodes, oracle/oracle_b.c tuned variant, %d threads" % (S, N_PODS, N_NODES, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps * (N_PODS / S),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Go with un-vendored deps (no toolchain here): CPU port of the path timed instead; "
                "ms_per_step is extrapolated to the full 10k-pod batch",
    }))


# ---------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kgpu", choices=["kgpu", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "allreduce", "push"],
                    help="multi-GPU key exchange: NCCL all-gather + K2 (default, measured in round 1) or the "
                         "(round-2 prep, untested) peer-memory push kernel of kgpu_score_batch_exchange")
    ap.add_argument("--graph", action="store_true",
                    help="(round-2 prep, untested) replay the step (memset + K1s + all-gather + K2) as one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from kubegpu_b200 import _lib, synth
    from kubegpu_b200.distributed import all_reduce_min_keys, shard_range
    from kubegpu_b200.scorer import Scorer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N>1)" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kubegpu_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"                      # no version banner on stdout: one JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    # ---- inputs: this rank's contiguous node shard, the full pod batch -----------------
    lo, hi = shard_range(N_NODES, world, rank)
    topo, free, pods = synth.gen_c2(hi - lo, N_PODS, node_start=lo)
    scorer = Scorer((local_rank,))
    scorer.set_variant(_lib.VARIANT_SPARSE)                   # headline kernel K1s
    scorer.upload_nodes(topo, free, node_id_base=lo)          # resident in HBM before timing

    d_pods = torch.from_numpy(pods).to(dev)
    d_local = torch.empty(N_PODS, dtype=torch.int64, device=dev)
    d_gather = torch.empty((world, N_PODS), dtype=torch.int64, device=dev) if world > 1 else None
    d_final = torch.empty(N_PODS, dtype=torch.int64, device=dev) if world > 1 else d_local
    h_pods = torch.from_numpy(pods).pin_memory()
    h_keys = torch.empty(N_PODS, dtype=torch.int64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.Stream(device=dev)          # everything (K1, NCCL, K2, copies) is enqueued here
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream

    push = args.exchange == "push"
    if push and args.graph:
        raise SystemExit("--exchange push cannot be captured in a graph (the epoch is a kernel argument)")
    if push:                                        # every rank maps every rank's result array (CUDA IPC)
        mine = scorer.exchange_init(world, rank, N_PODS)
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, mine)
        else:
            handles = [mine]
        scorer.exchange_connect(handles)
    wrapped = {}

    class _DeviceKeys:                              # zero-copy view of the handle-owned result array
        def __init__(self, ptr):
            self.__cuda_array_interface__ = {"shape": (N_PODS,), "typestr": "<i8", "data": (ptr, False), "version": 3}

    def step_device():
        """pods already in HBM -> final keys in HBM (all ranks hold the answer); returns the tensor holding them."""
        sptr = torch.cuda.current_stream().cuda_stream          # the capture stream while a graph is recorded
        if push:
            ptr = scorer.score_batch_exchange(d_pods.data_ptr(), N_PODS, sptr, _lib.BATCH_NO_MIN_MEM)
            if ptr not in wrapped:
                wrapped[ptr] = torch.as_tensor(_DeviceKeys(ptr), device=dev)
            return wrapped[ptr]
        scorer.score_batch_device(d_pods.data_ptr(), N_PODS, d_local.data_ptr(), sptr, _lib.BATCH_NO_MIN_MEM)
        if world > 1 and args.exchange == "allreduce":          # one collective, no K2 (round-2 experiment)
            return all_reduce_min_keys(d_local)
        if world > 1:
            dist.all_gather_into_tensor(d_gather.view(-1), d_local)
            scorer.reduce_shards_device(d_gather.data_ptr(), world, N_PODS, d_final.data_ptr(), sptr)
        return d_final

    def step_e2e():
        """host pods -> host keys through the public call."""
        if world == 1 and not push:
            scorer.score_batch_ptr(h_pods.data_ptr(), N_PODS, h_keys.data_ptr())   # kgpu_score_batch: H2D + K1 + D2H
        else:
            d_pods.copy_(h_pods, non_blocking=True)
            h_keys.copy_(step_device(), non_blocking=True)
            stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- K1-resident timing: `value` and the roofline ---------------------------------
    for _ in range(W):
        flush.zero_()
        step_device()
    graph = None
    if args.graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            step_device()
    barrier()
    launches0 = scorer.kernel_launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
           torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sync_token = torch.zeros(1, device=dev)
    for i in range(K):
        flush.zero_()                                  # L2 flush between timed iterations (outside the events)
        if world > 1:
            dist.all_reduce(sync_token)                # device-side rendezvous so no rank times another's flush
        ev[i][0].record(stream)
        if graph is not None:
            graph.replay()
            ev[i][1].record(stream)                    # (no K1-only split inside a graph replay)
        elif push:
            step_device()
            ev[i][1].record(stream)                    # (K1 and the push/sync kernel are one call)
        else:
            scorer.score_batch_device(d_pods.data_ptr(), N_PODS, d_local.data_ptr(), sptr, _lib.BATCH_NO_MIN_MEM)
            ev[i][1].record(stream)                    # K1 only: roofline numerator
            if world > 1 and args.exchange == "allreduce":
                all_reduce_min_keys(d_local)
            elif world > 1:
                dist.all_gather_into_tensor(d_gather.view(-1), d_local)
                scorer.reduce_shards_device(d_gather.data_ptr(), world, N_PODS, d_final.data_ptr(), sptr)
        ev[i][2].record(stream)
    barrier()
    clocks = sampler.finish()
    gpu_launches = scorer.kernel_launches - launches0
    step_ms = [a.elapsed_time(c) for a, _, c in ev]
    k1_ms = [a.elapsed_time(b) for a, b, _ in ev]
    total_ms = torch.tensor([sum(step_ms), sum(k1_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms, k1_total_ms = (float(x) for x in total_ms.tolist())
    ms_per_step = total_ms / K
    value = N_PODS / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer call ---------------------------------------
    for _ in range(W):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = N_PODS * K / float(e2e_s.item())
    final_keys = h_keys.numpy().view(np.uint64).copy()

    # ---- context lines: the other K1 variants (few steps, rank-local, N=1 only) ---------
    variants = {}
    if world == 1 and not args.no_variants and not push:
        for name, var, reps in (("lane_per_node_dense_all_C8k_subsets", _lib.VARIANT_LANE_PER_NODE, 5),
                                ("warp_per_pair_north_star_mapping", _lib.VARIANT_WARP_PER_PAIR, 3),
                                ("tile_memo_not_headline", _lib.VARIANT_TILE_MEMO, 10),
                                ("memo_by_k_not_headline", _lib.VARIANT_MEMO_BY_K, 10)):
            scorer.set_variant(var)
            step_device()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                step_device()
            b.record(stream)
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            same = bool((d_local.cpu().numpy().view(np.uint64) == final_keys).all())
            variants[name] = {"ms_per_step": ms, "value": N_PODS / (ms * 1e-3), "unit": UNIT,
                              "roofline_frac": algorithmic_bytes(N_NODES, N_PODS) / (ms * 1e-3) / 1e9 / measured_peak_gbs()[0],
                              "keys_identical_to_headline": same}
        scorer.set_variant(_lib.VARIANT_SPARSE)

    # ---- context: K3 stateful sequential placement (pods placed in order, masks updated) ----
    sequential = None
    if world == 1 and not args.no_variants:
        scorer.place_batch(pods[:256])                     # warm-up (changes the masks: re-upload below)
        seq_ms = []
        for _ in range(3):
            scorer.upload_nodes(topo, free, node_id_base=lo)
            scorer.place_batch(pods)
            seq_ms.append(scorer.last_kernel_ms)
        scorer.upload_nodes(topo, free, node_id_base=lo)
        ms = float(np.median(seq_ms))
        sequential = {"kernel": "place_init + place_sequential (kgpu_place_batch)", "ms_per_batch": ms,
                      "value": N_PODS / (ms * 1e-3), "unit": UNIT,
                      "note": "each pod sees the free masks left by the pods before it (no snapshot collapse); "
                              "bit-exact vs the CPU twin in tests/test_place_sequential.py"}

    # ---- context: memory-aware pods (config C6: per-GPU memory classes, pods with min_mem) ----
    memory_aware = None
    if world == 1 and not args.no_variants and not push:
        topo6, free6, mem6, pods6 = synth.gen_c6()
        with Scorer((local_rank,)) as s6:
            s6.set_variant(_lib.VARIANT_SPARSE)
            s6.upload_nodes(topo6, free6)
            s6.upload_gpu_memory(mem6)
            d_pods6 = torch.from_numpy(pods6).to(dev)
            d_keys6 = torch.empty(len(pods6), dtype=torch.int64, device=dev)
            for _ in range(3):
                s6.score_batch_device(d_pods6.data_ptr(), len(pods6), d_keys6.data_ptr(), sptr)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(5):
                s6.score_batch_device(d_pods6.data_ptr(), len(pods6), d_keys6.data_ptr(), sptr)
            b.record(stream)
            torch.cuda.synchronize()
            snap_ms = a.elapsed_time(b) / 5
            s6.place_batch(pods6[:256])
            s6.upload_nodes(topo6, free6)
            s6.upload_gpu_memory(mem6)
            s6.place_batch(pods6)
            seq6_ms = s6.last_kernel_ms
        memory_aware = {"workload": "C6: 100k heterogeneous nodes with per-GPU memory classes x 10k pods, k in 1..8, 4 of 7 pods with min_mem",
                        "snapshot_ms_per_batch": snap_ms, "snapshot_value": len(pods6) / (snap_ms * 1e-3),
                        "sequential_ms_per_batch": seq6_ms, "sequential_value": len(pods6) / (seq6_ms * 1e-3), "unit": UNIT,
                        "note": "K1s + K1m launches / place_init + place_sequential with one table set per distinct min_mem; "
                                "parity in tests/test_memory_aware.py and tests/test_place_sequential.py"}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        k1_s = (k1_total_ms / K) * 1e-3
        achieved = algorithmic_bytes(hi - lo, N_PODS) / k1_s / 1e9      # rank 0's shard (shards are equal +-1 node)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "nodes": N_NODES, "pods": N_PODS,
                       "parallelism": ("node list sharded over %d GPU(s), " % world + ("peer-memory push + flag barrier (1 kernel)" if push else "1 NCCL all-reduce(min)" if args.exchange == "allreduce" else "1 NCCL all-gather + K2")) if world > 1 else "1 GPU, no collective",
                       "kernel": "score_pairs_sparse (per pair: every k-subset of the node's free-GPU positions)",
                       "l2": "flushed between timed iterations (256 MiB write); node array is 26 MB < L2"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": committed_traffic(), "peak_source": peak_src,
                         "kernel_ms": k1_total_ms / K,
                         "note": "algorithmic bytes = 260 B/pair + 24 B/pod; frac > 1 because each staged node is "
                                 "reused for every pod of the block (kernel is integer-issue bound, see DESIGN.md)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_PODS * 16, "d2h_bytes_per_step": N_PODS * 8,
                    "call": "kgpu_score_batch(host pods, host keys)" if world == 1 else "pinned H2D + K1 + all-gather + K2 + D2H"},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "no_fit_pods": int((final_keys == np.uint64(_lib.NO_FIT)).sum()),
            "keys_sha256_12": __import__("hashlib").sha256(final_keys.tobytes()).hexdigest()[:12],
        }
        if variants:
            line["variants"] = variants
        if sequential:
            line["stateful_sequential"] = sequential
        if memory_aware:
            line["memory_aware"] = memory_aware
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(topo, free, pods)
        elif world == 1:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    scorer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
