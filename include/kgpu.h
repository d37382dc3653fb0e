/*
 * kgpu.h -- C ABI of libkgpu: the B200-native topology-aware GPU placement scorer
 * that replaces the scoring path of microsoft/KubeGPU's gpuschedulerplugin.
 *
 * This is the drop-in boundary.  A Go host binds these entry
 * points through cgo (`#include "kgpu.h"`); in this
 * repository they are exercised through ctypes (kubegpu_b200/_lib.py) and through
 * the C++ host mirror of the reference's DeviceScheduler interface
 * (kubegpu_b200/csrc/host/).  Plain pointers and sizes only -- no C++/torch types.
 *
 * What each group replaces in the reference (/root/reference, commit 73e59ce):
 *   kgpu_upload_nodes / kgpu_update_node / kgpu_remove_node
 *       <- NvidiaGPUScheduler.AddNode / RemoveNode   gpuschedulerplugin/gpu_scheduler.go:21-32
 *          (node cache: AddResourcesToNodeTreeCache   gpuschedulerplugin/gpu.go:192-230);
 *          the matrix is the one nvml.GetDevices builds nvidiagpuplugin/gpu/nvml/nvml.go:37-49
 *   kgpu_score_batch[_device]
 *       <- NvidiaGPUScheduler.PodFitsDevice (score)   gpuschedulerplugin/gpu_scheduler.go:34-44
 *          -> TranslatePodGPUResources                gpuschedulerplugin/gpu.go:94-127
 *          -> findBestTreeInCache + assignGPUs        gpuschedulerplugin/gpu.go:232-271
 *          batched over every (pod, node) pair instead of one cgo call per pair
 *   kgpu_fit_lookup / kgpu_build_fit_table / kgpu_score_pairs
 *       <- the same PodFitsDevice, called once per (node, pod) by the core: served from a host-side (node, k) table
 *   kgpu_set_free_mask[s] / kgpu_place_batch[_ex] / kgpu_get_free_masks
 *       <- TakePodResources / ReturnPodResources       gpuschedulerplugin/gpu_scheduler.go:57-63 (no-ops there)
 *   kgpu_reduce_shards_device
 *       <- (no reference counterpart; the reference is single-process) the final
 *          per-pod pick over the all-gathered shard results, SURVEY.md 8(e)
 *   kgpu_set_weights
 *       <- the link-level grouping tables {6,5,4} / {6,5,4,3,2,1}
 *          nvidiagpuplugin/gpu/nvidia/nvidia_gpu_manager.go:178-180
 *   kgpu_last_error
 *       <- Go `error` returns (gpu_scheduler.go:46-55, gpu.go:125-126)
 *
 * Data layouts (all little-endian, host or device as stated per function):
 *   topo       int32[N][64]   row-major 8x8 link-level matrix per node; only the
 *                             upper triangle (i<j) is read; values 0..15
 *                             (0 unknown, 1..6 NVML P2P level, 7..12 NVLink links).
 *   free_mask  int32[N]       bit i = GPU i present and free (low 8 bits).
 *   pods       int32[P][4]    {k, pod_id, flags, min_mem_mib}; k GPUs wanted, 0..8;
 *                             min_mem_mib > 0: only GPUs with at least that much memory are
 *                             eligible for this pod (see kgpu_upload_gpu_memory); pod_id/flags
 *                             are carried, not interpreted.
 *   gpu_mem    int32[N][8]    MiB per GPU slot (the node agent advertises it per GPU as
 *                             `.../memory`, nvidia_gpu_manager.go:204-211); optional.
 *   keys       uint64[P]      (cost << 40) | (node_id << 8) | gpu_mask, or
 *                             KGPU_NO_FIT.  cost = sum over GPU pairs i<j in the
 *                             mask of W[topo[i][j]]; the key is the minimum over
 *                             all nodes and all k-subsets of free GPUs, so ties go
 *                             to the lower node_id, then the lower mask.
 *
 * Threading: every call on one handle is serialised by an internal mutex.  State-changing calls (upload, update,
 * set_free_mask(s), place_batch) and the lazy rebuilds they trigger run on the handle's own stream and have COMPLETED
 * when the call returns, so a following launch on any stream sees them.  The *_device entry points only enqueue: all
 * such calls on one handle must go to ONE stream at a time (the handle's per-batch scratch -- the "batch has
 * memory-constrained pods" flag, the work list -- is not duplicated per stream); use one handle per stream otherwise.
 * Errors: functions return KGPU_OK (0) or a negative KGPU_ERR_*; the message is
 * kept per handle (kgpu_last_error(h)) and per thread (kgpu_last_error(NULL)).
 * "No node fits" is a result (KGPU_NO_FIT), never an error.  Nothing here falls
 * back to the CPU: without a usable CUDA device every call fails.
 */
#ifndef KGPU_H_
#define KGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGPU_OK 0
#define KGPU_ERR_INVALID (-1) /* bad argument / value out of domain          */
#define KGPU_ERR_CUDA (-2)    /* CUDA runtime error (message has the detail) */
#define KGPU_ERR_NOMEM (-3)   /* host or device allocation failed            */
#define KGPU_ERR_COMM (-4)    /* NCCL / peer-access failure (multi-device)   */
#define KGPU_ERR_STATE (-5)   /* call not valid in the handle's state        */

#define KGPU_NO_FIT UINT64_MAX
#define KGPU_MAX_GPUS_PER_NODE 8
#define KGPU_NUM_LEVELS 16
#define KGPU_MAX_WEIGHT 4095

/* Kernel variants of K1 `score_pairs` (all bit-identical in result). */
#define KGPU_VARIANT_AUTO 0          /* default: the cheapest path with identical keys for the batch at hand: pods without a
                                      * memory requirement are memoised by k (best[k] over all nodes once, then a gather),
                                      * pods with one are scored per pair by the memory-aware kernel (K1m)          */
#define KGPU_VARIANT_WARP_PER_PAIR 1 /* north_star mapping: warp per (pod,node), lane per subset */
#define KGPU_VARIANT_LANE_PER_NODE 2 /* lane per node, pair costs in registers, all C(8,k) subsets per pair */
#define KGPU_VARIANT_MEMO_BY_K 3     /* global best[k] computed once, pods look it up (what AUTO does; dense K1m for min_mem pods) */
#define KGPU_VARIANT_SPARSE 5        /* PER-PAIR work (north_star): lane per node, nodes ordered by free-GPU count, every pod
                                      * enumerates the k-subsets of the node's free positions; bench.py's headline      */
#define KGPU_VARIANT_TILE_MEMO 4     /* lane per node, per-k minima hoisted out of the pod loop (NOT the headline) */

#define KGPU_KEY_COST(key) ((uint32_t)((key) >> 40))
#define KGPU_KEY_NODE(key) ((uint32_t)(((key) >> 8) & 0xFFFFFFFFu))
#define KGPU_KEY_MASK(key) ((uint32_t)((key)&0xFFu))

typedef struct kgpu_ctx kgpu_t;

/* "major.minor.patch" of the library. */
const char *kgpu_version(void);

/* Create a scorer on CUDA device(s) dev_ids[0..ndev).  ndev == 1: one handle per
 * GPU (one process per GPU, or several handles in one process).  ndev > 1: the
 * handle shards the node array over the devices and combines their results
 * (single scheduler process driving all local GPUs). */
int kgpu_create(const int *dev_ids, int ndev, kgpu_t **out);
int kgpu_destroy(kgpu_t *h);

/* Message of the last failing call on h (or on this thread if h == NULL). */
const char *kgpu_last_error(kgpu_t *h);

/* Link-level -> cost table, 16 entries, each 0..KGPU_MAX_WEIGHT.
 * Default {64,32,16,8,4,2,1,0,...}: level 1 (cross-CPU) costs most, NVLink 0. */
int kgpu_set_weights(kgpu_t *h, const int32_t w[KGPU_NUM_LEVELS]);
int kgpu_get_weights(kgpu_t *h, int32_t w[KGPU_NUM_LEVELS]);

int kgpu_set_variant(kgpu_t *h, int variant);

/* Replace the node array (host pointers).  node_id_base is added to the local
 * index to form the node_id field of the keys (global id of this shard's node 0). */
int kgpu_upload_nodes(kgpu_t *h, const int32_t *topo, const int32_t *free_mask, int64_t n,
                      int64_t node_id_base);
/* Per-GPU memory (MiB) of every node, n = kgpu_num_nodes entries of 8.  Until this is called every
 * GPU has unlimited memory, i.e. the pods' min_mem_mib never excludes anything.  A later
 * kgpu_upload_nodes resets it to unlimited. */
int kgpu_upload_gpu_memory(kgpu_t *h, const int32_t *mem_mib, int64_t n);
int kgpu_update_gpu_memory(kgpu_t *h, int64_t idx, const int32_t mem_mib[8]);
/* Overwrite one node (AddNode on an existing name / usage update). */
int kgpu_update_node(kgpu_t *h, int64_t idx, const int32_t topo[64], int32_t free_mask);
int kgpu_set_free_mask(kgpu_t *h, int64_t idx, int32_t free_mask);
/* The same for n nodes in one call (a scheduling cycle's TakePodResources / ReturnPodResources,
 * gpu_scheduler.go:57-63): one host-to-device copy and one kernel that stores the masks and refreshes the
 * scorer's cached records of exactly those nodes.  A node listed twice gets its LAST mask. */
int kgpu_set_free_masks(kgpu_t *h, const int64_t *idx, const int32_t *free_mask, int64_t n);
/* RemoveNode: the slot stays, its GPUs become unschedulable (free_mask = 0). */
int kgpu_remove_node(kgpu_t *h, int64_t idx);
int64_t kgpu_num_nodes(kgpu_t *h);

/* Score P pods against every node: host buffers in, host buffers out
 * (H2D copy of pods + kernel(s) + D2H copy of keys, synchronous). */
int kgpu_score_batch(kgpu_t *h, const int32_t *pods, int64_t P, uint64_t *out_keys);

/* Same with device buffers on the handle's device (ndev == 1 handles only),
 * enqueued on `stream` (a cudaStream_t; NULL = the CUDA default stream, as in every CUDA
 * API), no sync.  The node array must not be changed until the work has completed.
 * d_pods must be 16-byte aligned. */
int kgpu_score_batch_device(kgpu_t *h, const int32_t *d_pods, int64_t P, uint64_t *d_keys,
                            void *stream);

/* Same, with promises about the batch the device-buffer path cannot check for itself:
 * KGPU_BATCH_NO_MIN_MEM = no pod of this batch has min_mem_mib > 0 (skips the flag kernel and the
 * K1m launch whose blocks would only read the flag and exit). */
#define KGPU_BATCH_NO_MIN_MEM 1
int kgpu_score_batch_device_ex(kgpu_t *h, const int32_t *d_pods, int64_t P, uint64_t *d_keys, void *stream,
                               int batch_flags);

/* Per-pair query (one PodFitsDevice(node, pod) call, or a list of them): for each i,
 * out_node_keys[i] = (cost << 8) | gpu_mask of the cheapest k[i]-subset of the free GPUs
 * of node node_idx[i] (local index as uploaded), or UINT32_MAX if it does not fit.
 * min_mem_mib may be NULL (no memory requirement) or hold one requirement per pair.
 * Host buffers, synchronous. */
int kgpu_score_pairs(kgpu_t *h, const int64_t *node_idx, const int32_t *k, const int32_t *min_mem_mib, int64_t n,
                     uint32_t *out_node_keys);

/* The (node, k) fit table: for every node and k = 0..8 the (cost << 8) | gpu_mask of its cheapest k-subset of
 * free GPUs (UINT32_MAX: does not fit), computed by one launch and kept as a host copy inside the handle.
 * kgpu_fit_lookup reads that copy -- no launch, no copy: this is what serves PodFitsDevice, which the core calls
 * once per (node, pod) pair (gpu_scheduler.go:34-44).  The table is built on first use and kept current by
 * kgpu_update_node / kgpu_set_free_mask(s) (they refresh the rows of the nodes they touch); kgpu_upload_nodes,
 * kgpu_set_weights and kgpu_place_batch invalidate it (rebuilt on the next lookup).  kgpu_build_fit_table
 * forces the build (e.g. at the start of a scheduling cycle).  kgpu_score_pairs uses the same table for pairs
 * without a memory requirement. */
int kgpu_build_fit_table(kgpu_t *h);
int kgpu_fit_lookup(kgpu_t *h, int64_t node_idx, int32_t k, uint32_t *out_node_key);

/* Stateful sequential placement (what TakePodResources would make of a scheduling cycle;
 * a no-op in the reference, gpu_scheduler.go:57-63).  Pods are placed IN ORDER; each one gets
 * the best (cost, node, mask) under the free masks left by the pods before it and then takes
 * those GPUs: the handle's device-side free masks are updated.  min_mem_mib is honoured like
 * in kgpu_score_batch; one batch may carry at most 7 distinct positive min_mem_mib values
 * (KGPU_ERR_INVALID otherwise).  Host buffers, synchronous, single-device handles only. */
int kgpu_place_batch(kgpu_t *h, const int32_t *pods, int64_t P, uint64_t *out_keys);
/* Same with flags.  KGPU_PLACE_DRY_RUN: the pods are placed in order on a scratch copy of the free masks, so
 * the keys are conflict-free PROPOSALS for the whole batch (no two pods share a GPU) while the handle's state
 * is untouched; commit the accepted ones with kgpu_set_free_masks (TakePodResources). */
#define KGPU_PLACE_DRY_RUN 1
int kgpu_place_batch_ex(kgpu_t *h, const int32_t *pods, int64_t P, uint64_t *out_keys, int flags);
/* Copy the current free masks (n = kgpu_num_nodes entries) back to the host. */
int kgpu_get_free_masks(kgpu_t *h, int32_t *out_free_mask, int64_t n);

/* K2: d_out[p] = min over g < G of d_gathered[g*P + p] (after an all-gather of
 * every shard's keys), enqueued on `stream`. */
int kgpu_reduce_shards_device(kgpu_t *h, const uint64_t *d_gathered, int G, int64_t P,
                              uint64_t *d_out, void *stream);

/* Peer-memory key exchange for the one-process-per-GPU launch (replaces the all-gather + kgpu_reduce_shards_device pair;
 * no reference counterpart).  Every rank:
 *   kgpu_exchange_init(h, world, rank, max_pods, handle)   allocate the rank's slot/flag memory, export it
 *   <all-gather the KGPU_IPC_HANDLE_BYTES-byte handles through the launcher, e.g. torch.distributed>
 *   kgpu_exchange_connect(h, handles)                      map every peer's memory (handles = [world][64])
 *   kgpu_score_batch_exchange(h, d_pods, P, &d_final, stream, flags)   per step, TWO launches: K1 on the local shard, then
 *       one kernel that stores the P bests into every rank's slot array over NVLink (plain coalesced stores), meets the
 *       other ranks at a flag barrier in peer memory and takes the per-pod minimum over the G slots.  *d_final (device
 *       memory owned by the handle, valid until the next-but-one call) then holds the global keys on every rank.
 *   kgpu_exchange_barrier(h, stream)                       the same kernel with P = 0: a device-side barrier of the ranks
 * All ranks must make the same sequence of exchange calls.  A rank that waits more than ~10 s for its peers raises an
 * error (reported as KGPU_ERR_COMM by a later call) instead of hanging the GPU.  Single-device handles only. */
#define KGPU_IPC_HANDLE_BYTES 64
int kgpu_exchange_init(kgpu_t *h, int world, int rank, int64_t max_pods, unsigned char *out_handle);
int kgpu_exchange_connect(kgpu_t *h, const unsigned char *handles);
int kgpu_score_batch_exchange(kgpu_t *h, const int32_t *d_pods, int64_t P, const uint64_t **d_final_keys,
                              void *stream, int batch_flags);
int kgpu_exchange_barrier(kgpu_t *h, void *stream);

/* Number of CUDA kernels this handle has launched so far (bench bookkeeping). */
int64_t kgpu_kernel_launches(kgpu_t *h);
/* Wall-clock duration (ms) of the most recent kgpu_upload_nodes: host-to-device copies, the device-side
 * value-domain check, the K1s order (counting sort by free-GPU count) and the compacted records. */
double kgpu_last_upload_ms(kgpu_t *h);
/* Device duration (ms, CUDA events on the launching stream) of the K1 launch(es)
 * of the most recent kgpu_score_batch call. */
double kgpu_last_kernel_ms(kgpu_t *h);

#ifdef __cplusplus
}
#endif
#endif /* KGPU_H_ */
