/*
 * oracle_b.c -- Oracle B: CPU twin of the subset-enumeration placement scorer.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kubegpu_b200/ (the product) may
 * include, link, import or execute this file.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it, and
 * there only as the checker / the reported CPU baseline.
 *
 * Parity status: the subset scorer named by BASELINE.json:north_star does not
 * exist as code in the reference (SURVEY.md "Read this first"); this file is
 * the definition fixed in SURVEY.md 8(c) "Oracle B".  It is tied to the
 * reference in two ways, both tested in tests/test_oracle_agree.py:
 *   - its inputs (the 8x8 link-level matrix, value domain 0..6 = NVML P2P
 *     levels, 0 = unknown/diagonal) follow
 *     nvidiagpuplugin/gpu/nvml/nvml.go:37-49,69-78 and
 *     nvidiagpuplugin/gpu/nvidia/nvidia_gpu_manager.go:159-180;
 *   - on matrices generated from a 2-level group shape its min cost equals the
 *     pairwise cost of the reference's greedy fill (gpuschedulerplugin/gpu.go:
 *     247-271, restated in oracle/oracle_a.py) on the documented agree-set.
 * "Bit-exact" for the CUDA path means bit-exact against THIS file.
 *
 * Definition (all integer):
 *   node n:  M = int32[8][8] link levels (only the upper triangle i<j is read,
 *            each value masked to 0..15), free = free_mask & 0xFF.
 *   pod:     k = pods[4*p+0]  (k==0: empty set, cost 0; k<0 or k>8: no fit).
 *   cost(S)  = sum_{i<j in S} W[M[i][j]]          W = int32[16], 0..4095
 *   nodekey  = min over S (popcount(S)==k, S subset of free) of (cost<<8 | S)
 *   podkey   = min over nodes of (cost<<40 | node_id<<8 | S)   (uint64)
 *   no feasible (node,S) anywhere -> UINT64_MAX.
 *   Subsets are enumerated in increasing integer order (Gosper's hack).
 *
 * Memory-aware extension (the node agent advertises per-GPU
 * `.../memory`, nvidia_gpu_manager.go:204-211): with mem = int32[N][8] (MiB per GPU)
 * and min_mem = pods[4*p+3] > 0, GPU i of node n is eligible for pod p only if
 * mem[n][i] >= min_mem, i.e. free is replaced by free & eligible(p, n).  mem == NULL or
 * min_mem <= 0: no constraint.  The *_mem entry points take that extra array.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define KGPU_NO_FIT UINT64_MAX
#define NODE_NO_FIT UINT32_MAX

/* ---- plain restatement: the checker ------------------------------------ */

uint32_t kgpu_oracle_subset_cost(const int32_t *M, unsigned S, const int32_t *W)
{
    uint32_t cost = 0;
    for (int i = 0; i < 8; i++) {
        if (!((S >> i) & 1u)) continue;
        for (int j = i + 1; j < 8; j++)
            if ((S >> j) & 1u) cost += (uint32_t)W[M[i * 8 + j] & 15];
    }
    return cost;
}

uint32_t kgpu_oracle_node_key(const int32_t *M, int32_t free_mask, int k, const int32_t *W)
{
    unsigned fm = (unsigned)free_mask & 0xFFu;
    if (k < 0 || k > 8) return NODE_NO_FIT;
    if (k == 0) return 0u;
    uint32_t best = NODE_NO_FIT;
    unsigned S = (1u << k) - 1u;
    while (S < 256u) {
        if ((S & ~fm) == 0u) {
            uint32_t key = (kgpu_oracle_subset_cost(M, S, W) << 8) | S;
            if (key < best) best = key;
        }
        unsigned c = S & (0u - S);          /* Gosper: next k-subset */
        unsigned r = S + c;
        S = (((r ^ S) >> 2) / c) | r;
    }
    return best;
}

/* free mask of node n as pod p sees it: GPUs with too little memory drop out */
static inline int32_t eff_free(int32_t free_mask, const int32_t *mem8, int32_t min_mem)
{
    if (!mem8 || min_mem <= 0) return free_mask;
    unsigned ok = 0;
    for (int i = 0; i < 8; i++)
        if (mem8[i] >= min_mem) ok |= 1u << i;
    return (int32_t)((unsigned)free_mask & ok);
}

static inline uint64_t pod_key(uint32_t nk, uint64_t node_id)
{
    return ((uint64_t)(nk >> 8) << 40) | (node_id << 8) | (uint64_t)(nk & 0xFFu);
}

/* out_keys[p] = best placement of pod p over nodes [0,N) whose global ids are
 * node_id_base + index.  Single thread, no precomputation. */
void kgpu_oracle_score_batch_mem(const int32_t *topo, const int32_t *free_mask, const int32_t *mem,
                                 int64_t N, int64_t node_id_base, const int32_t *pods, int64_t P,
                                 const int32_t *W, uint64_t *out_keys)
{
    for (int64_t p = 0; p < P; p++) {
        int k = pods[4 * p];
        uint64_t best = KGPU_NO_FIT;
        for (int64_t n = 0; n < N; n++) {
            int32_t fm = eff_free(free_mask[n], mem ? mem + 8 * n : NULL, pods[4 * p + 3]);
            uint32_t nk = kgpu_oracle_node_key(topo + 64 * n, fm, k, W);
            if (nk == NODE_NO_FIT) continue;
            uint64_t key = pod_key(nk, (uint64_t)(node_id_base + n));
            if (key < best) best = key;
        }
        out_keys[p] = best;
    }
}

void kgpu_oracle_score_batch(const int32_t *topo, const int32_t *free_mask, int64_t N,
                             int64_t node_id_base, const int32_t *pods, int64_t P,
                             const int32_t *W, uint64_t *out_keys)
{
    int32_t *nomem = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)(P > 0 ? P : 1));
    memcpy(nomem, pods, sizeof(int32_t) * 4 * (size_t)P);
    for (int64_t p = 0; p < P; p++) nomem[4 * p + 3] = 0;      /* the plain entry point ignores min_mem */
    kgpu_oracle_score_batch_mem(topo, free_mask, NULL, N, node_id_base, nomem, P, W, out_keys);
    free(nomem);
}

/* ---- tuned CPU variant: the reported CPU baseline ---------------------- */
/* Same results (tests assert equality with the plain version).  Per node a
 * 256-entry subset-cost table is built once (cost[S] = cost[S minus lowest
 * bit] + row sum) together with the lists of its feasible subsets per size, then
 * every pod of the thread's range enumerates the feasible k-subsets through the
 * table.  Threads split the pod range, so no merge is needed. */

static const uint8_t *subsets_of_size(int k, int *count)
{
    static uint8_t tbl[9][70];
    static int cnt[9];
    static int ready = 0;
    if (!ready) {
        for (int kk = 0; kk <= 8; kk++) {
            cnt[kk] = 0;
            for (unsigned S = 0; S < 256; S++)
                if (__builtin_popcount(S) == kk) tbl[kk][cnt[kk]++] = (uint8_t)S;
        }
        ready = 1;
    }
    *count = cnt[k];
    return tbl[k];
}

static void build_cost_table(const int32_t *M, const int32_t *W, uint32_t *cost)
{
    cost[0] = 0;
    for (unsigned S = 1; S < 256; S++) {
        int i = __builtin_ctz(S);
        unsigned rest = S & (S - 1);
        uint32_t c = cost[rest];
        for (unsigned t = rest; t; t &= t - 1)
            c += (uint32_t)W[M[i * 8 + __builtin_ctz(t)] & 15];
        cost[S] = c;
    }
}

struct fast_job {
    const int32_t *topo, *free_mask, *mem, *pods, *W;
    int64_t N, node_id_base, p0, p1;
    uint64_t *out;
};

/* Per node: the 256-entry subset-cost table, plus the node's feasible subsets (subsets of its free mask)
 * bucketed by size in increasing mask order -- the same sparsity the GPU headline kernel exploits: a pod
 * wanting k GPUs only visits the C(f,k) subsets of the f free GPUs.  A pod with a memory requirement sees
 * a smaller mask and filters the list. */
static void *fast_worker(void *arg)
{
    struct fast_job *j = (struct fast_job *)arg;
    uint32_t cost[256];
    uint8_t list[9][70];
    int cnt[9];
    for (int64_t p = j->p0; p < j->p1; p++) j->out[p] = KGPU_NO_FIT;
    for (int64_t n = 0; n < j->N; n++) {
        const unsigned fm0 = (unsigned)j->free_mask[n] & 0xFFu;
        uint64_t nid = (uint64_t)(j->node_id_base + n);
        build_cost_table(j->topo + 64 * n, j->W, cost);
        for (int k = 0; k <= 8; k++) cnt[k] = 0;
        for (unsigned S = 0; S < 256; S++)
            if ((S & ~fm0) == 0) { int k = __builtin_popcount(S); list[k][cnt[k]++] = (uint8_t)S; }
        for (int64_t p = j->p0; p < j->p1; p++) {
            int k = j->pods[4 * p];
            if (k < 0 || k > 8) continue;
            unsigned fm = fm0;
            if (j->mem && j->pods[4 * p + 3] > 0)
                fm = (unsigned)eff_free((int32_t)fm0, j->mem + 8 * n, j->pods[4 * p + 3]) & 0xFFu;
            uint32_t best = NODE_NO_FIT;
            const uint8_t *subs = list[k];
            for (int s = 0; s < cnt[k]; s++) {
                unsigned S = subs[s];
                if (S & ~fm) continue;                 /* only bites for memory-constrained pods */
                uint32_t key = (cost[S] << 8) | S;
                if (key < best) best = key;
            }
            if (best == NODE_NO_FIT) continue;
            uint64_t key = pod_key(best, nid);
            if (key < j->out[p]) j->out[p] = key;
        }
    }
    return NULL;
}

void kgpu_oracle_score_batch_fast_mem(const int32_t *topo, const int32_t *free_mask, const int32_t *mem,
                                      int64_t N, int64_t node_id_base, const int32_t *pods, int64_t P,
                                      const int32_t *W, uint64_t *out_keys, int nthreads)
{
    int dummy;
    (void)subsets_of_size(0, &dummy); /* build the table before threads start */
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > P) nthreads = P > 0 ? (int)P : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    struct fast_job *jobs = (struct fast_job *)malloc(sizeof(struct fast_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (struct fast_job){topo, free_mask, mem, pods, W, N, node_id_base,
                                    P * t / nthreads, P * (t + 1) / nthreads, out_keys};
        if (t > 0) pthread_create(&th[t], NULL, fast_worker, &jobs[t]);
    }
    fast_worker(&jobs[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

void kgpu_oracle_score_batch_fast(const int32_t *topo, const int32_t *free_mask, int64_t N,
                                  int64_t node_id_base, const int32_t *pods, int64_t P,
                                  const int32_t *W, uint64_t *out_keys, int nthreads)
{
    int32_t *nomem = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)(P > 0 ? P : 1));
    memcpy(nomem, pods, sizeof(int32_t) * 4 * (size_t)P);
    for (int64_t p = 0; p < P; p++) nomem[4 * p + 3] = 0;
    kgpu_oracle_score_batch_fast_mem(topo, free_mask, NULL, N, node_id_base, nomem, P, W, out_keys, nthreads);
    free(nomem);
}

/* K2 twin: column-min over G gathered key arrays. */
void kgpu_oracle_reduce_shards(const uint64_t *gathered, int G, int64_t P, uint64_t *out)
{
    for (int64_t p = 0; p < P; p++) {
        uint64_t best = KGPU_NO_FIT;
        for (int g = 0; g < G; g++)
            if (gathered[(int64_t)g * P + p] < best) best = gathered[(int64_t)g * P + p];
        out[p] = best;
    }
}

/* ---- K3 twin: stateful sequential placement ------------------- */
/* for p in order: key_p = kgpu_oracle_score_batch of pod p alone under the CURRENT free
 * masks; if it fits, the chosen GPUs leave free_mask[node].  free_mask is updated in place.
 * Plain version: literally that (small cases). */
void kgpu_oracle_place_batch_plain(const int32_t *topo, int32_t *free_mask, int64_t N,
                                   int64_t node_id_base, const int32_t *pods, int64_t P,
                                   const int32_t *W, uint64_t *out_keys)
{
    for (int64_t p = 0; p < P; p++) {
        kgpu_oracle_score_batch(topo, free_mask, N, node_id_base, pods + 4 * p, 1, W, out_keys + p);
        if (out_keys[p] == KGPU_NO_FIT) continue;
        int64_t n = (int64_t)((out_keys[p] >> 8) & 0xFFFFFFFFu) - node_id_base;
        free_mask[n] = (int32_t)(((uint32_t)free_mask[n] & 0xFFu) & ~(uint32_t)(out_keys[p] & 0xFFu));
    }
}

/* Memory-aware sequential placement: as above, each pod scored with kgpu_oracle_score_batch_mem
 * (GPUs with less than the pod's min_mem MiB do not count as free for that pod). */
void kgpu_oracle_place_batch_mem(const int32_t *topo, int32_t *free_mask, const int32_t *mem, int64_t N,
                                 int64_t node_id_base, const int32_t *pods, int64_t P,
                                 const int32_t *W, uint64_t *out_keys)
{
    for (int64_t p = 0; p < P; p++) {
        kgpu_oracle_score_batch_mem(topo, free_mask, mem, N, node_id_base, pods + 4 * p, 1, W, out_keys + p);
        if (out_keys[p] == KGPU_NO_FIT) continue;
        int64_t n = (int64_t)((out_keys[p] >> 8) & 0xFFFFFFFFu) - node_id_base;
        free_mask[n] = (int32_t)(((uint32_t)free_mask[n] & 0xFFu) & ~(uint32_t)(out_keys[p] & 0xFFu));
    }
}

/* Same results with a per-node cache of the 9 node keys (only the chosen node changes). */
void kgpu_oracle_place_batch(const int32_t *topo, int32_t *free_mask, int64_t N,
                             int64_t node_id_base, const int32_t *pods, int64_t P,
                             const int32_t *W, uint64_t *out_keys)
{
    uint32_t *nb = (uint32_t *)malloc(sizeof(uint32_t) * 9 * (size_t)(N > 0 ? N : 1));
    for (int64_t n = 0; n < N; n++)
        for (int k = 0; k <= 8; k++) nb[9 * n + k] = kgpu_oracle_node_key(topo + 64 * n, free_mask[n], k, W);
    for (int64_t p = 0; p < P; p++) {
        int k = pods[4 * p];
        out_keys[p] = KGPU_NO_FIT;
        if (k < 0 || k > 8) continue;
        uint32_t best = NODE_NO_FIT;
        int64_t bn = -1;
        for (int64_t n = 0; n < N; n++) {
            uint32_t nk = nb[9 * n + k];
            if (nk == NODE_NO_FIT) continue;
            if (bn < 0 || (nk >> 8) < (best >> 8)) { best = nk; bn = n; }   /* lower cost; ties keep lower node */
        }
        if (bn < 0) continue;
        out_keys[p] = pod_key(best, (uint64_t)(node_id_base + bn));
        free_mask[bn] = (int32_t)(((uint32_t)free_mask[bn] & 0xFFu) & ~(best & 0xFFu));
        for (int kk = 0; kk <= 8; kk++) nb[9 * bn + kk] = kgpu_oracle_node_key(topo + 64 * bn, free_mask[bn], kk, W);
    }
    free(nb);
}

/* ---- fair CPU twins for the bench ("a CPU twin with the same two-level minima") ------------- */

/* The nine node keys of one node from its 256-entry subset-cost table. */
static void node_keys_from_table(const uint32_t *cost, unsigned fm, uint32_t *nk9)
{
    for (int k = 0; k <= 8; k++) nk9[k] = NODE_NO_FIT;
    for (unsigned S = 0; S < 256; S++) {
        if (S & ~fm) continue;
        int k = __builtin_popcount(S);
        uint32_t key = (cost[S] << 8) | S;
        if (key < nk9[k]) nk9[k] = key;
    }
}

/* Sequential placement with the SAME data structure as the GPU kernel K3 (place_sequential.cuh): node keys
 * nb[9][N], minima per 128-node tile, minima per supertile of 32 tiles; per pod: scan the supertile minima,
 * take the winner, re-enumerate that one node, refresh its tile and its supertile.  One thread (the chain is
 * serial).  Same results as kgpu_oracle_place_batch (tests assert it). */
void kgpu_oracle_place_batch_tiled(const int32_t *topo, int32_t *free_mask, int64_t N,
                                   int64_t node_id_base, const int32_t *pods, int64_t P,
                                   const int32_t *W, uint64_t *out_keys)
{
    const int64_t TILE = 128, SUP = 32;
    const int64_t T = (N + TILE - 1) / TILE, ST = (T + SUP - 1) / SUP;
    uint32_t *nb = (uint32_t *)malloc(sizeof(uint32_t) * 9 * (size_t)(T * TILE > 0 ? T * TILE : 1));
    uint64_t *tb = (uint64_t *)malloc(sizeof(uint64_t) * 9 * (size_t)(T > 0 ? T : 1));
    uint64_t *sb = (uint64_t *)malloc(sizeof(uint64_t) * 9 * (size_t)(ST > 0 ? ST : 1));
    uint32_t cost[256], nk9[9];
    for (int64_t i = 0; i < 9 * T * TILE; i++) nb[i] = NODE_NO_FIT;
    for (int64_t n = 0; n < N; n++) {
        build_cost_table(topo + 64 * n, W, cost);
        node_keys_from_table(cost, (unsigned)free_mask[n] & 0xFFu, nk9);
        for (int k = 0; k <= 8; k++) nb[(int64_t)k * T * TILE + n] = nk9[k];
    }
    for (int k = 0; k <= 8; k++) {
        for (int64_t t = 0; t < T; t++) {
            uint64_t b = KGPU_NO_FIT;
            for (int64_t n = t * TILE; n < (t + 1) * TILE; n++) {
                uint32_t v = nb[(int64_t)k * T * TILE + n];
                if (v != NODE_NO_FIT) { uint64_t key = pod_key(v, (uint64_t)(node_id_base + n)); if (key < b) b = key; }
            }
            tb[(int64_t)k * T + t] = b;
        }
        for (int64_t s = 0; s < ST; s++) {
            uint64_t b = KGPU_NO_FIT;
            for (int64_t t = s * SUP; t < (s + 1) * SUP && t < T; t++)
                if (tb[(int64_t)k * T + t] < b) b = tb[(int64_t)k * T + t];
            sb[(int64_t)k * ST + s] = b;
        }
    }
    for (int64_t p = 0; p < P; p++) {
        int k = pods[4 * p];
        out_keys[p] = KGPU_NO_FIT;
        if (k < 0 || k > 8) continue;
        uint64_t win = KGPU_NO_FIT;
        for (int64_t s = 0; s < ST; s++)
            if (sb[(int64_t)k * ST + s] < win) win = sb[(int64_t)k * ST + s];
        out_keys[p] = win;
        if (win == KGPU_NO_FIT || k == 0) continue;
        const int64_t n = (int64_t)((win >> 8) & 0xFFFFFFFFu) - node_id_base;
        free_mask[n] = (int32_t)(((uint32_t)free_mask[n] & 0xFFu) & ~(uint32_t)(win & 0xFFu));
        build_cost_table(topo + 64 * n, W, cost);
        node_keys_from_table(cost, (unsigned)free_mask[n] & 0xFFu, nk9);
        const int64_t t = n / TILE, s = t / SUP;
        for (int kk = 0; kk <= 8; kk++) {
            nb[(int64_t)kk * T * TILE + n] = nk9[kk];
            uint64_t b = KGPU_NO_FIT;
            for (int64_t m = t * TILE; m < (t + 1) * TILE; m++) {
                uint32_t v = nb[(int64_t)kk * T * TILE + m];
                if (v != NODE_NO_FIT) { uint64_t key = pod_key(v, (uint64_t)(node_id_base + m)); if (key < b) b = key; }
            }
            tb[(int64_t)kk * T + t] = b;
            b = KGPU_NO_FIT;
            for (int64_t u = s * SUP; u < (s + 1) * SUP && u < T; u++)
                if (tb[(int64_t)kk * T + u] < b) b = tb[(int64_t)kk * T + u];
            sb[(int64_t)kk * ST + s] = b;
        }
    }
    free(nb); free(tb); free(sb);
}

/* Snapshot scoring memoised by k (what the GPU's memo_by_k variant does): with no per-pod constraint a pod's
 * key depends on the pod only through k, so best[k] over all nodes is computed once (threads split the nodes)
 * and every pod reads best[k_p].  Pods with min_mem > 0 are not memoisable and are left KGPU_NO_FIT here. */
struct memo_job {
    const int32_t *topo, *free_mask, *W;
    int64_t n0, n1, node_id_base;
    uint64_t best[9];
};

static void *memo_worker(void *arg)
{
    struct memo_job *j = (struct memo_job *)arg;
    uint32_t cost[256], nk9[9];
    for (int k = 0; k <= 8; k++) j->best[k] = KGPU_NO_FIT;
    for (int64_t n = j->n0; n < j->n1; n++) {
        build_cost_table(j->topo + 64 * n, j->W, cost);
        node_keys_from_table(cost, (unsigned)j->free_mask[n] & 0xFFu, nk9);
        for (int k = 0; k <= 8; k++)
            if (nk9[k] != NODE_NO_FIT) {
                uint64_t key = pod_key(nk9[k], (uint64_t)(j->node_id_base + n));
                if (key < j->best[k]) j->best[k] = key;
            }
    }
    return NULL;
}

void kgpu_oracle_score_batch_memo(const int32_t *topo, const int32_t *free_mask, int64_t N,
                                  int64_t node_id_base, const int32_t *pods, int64_t P,
                                  const int32_t *W, uint64_t *out_keys, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if ((int64_t)nthreads > N) nthreads = N > 0 ? (int)N : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    struct memo_job *jobs = (struct memo_job *)malloc(sizeof(struct memo_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        jobs[t].topo = topo; jobs[t].free_mask = free_mask; jobs[t].W = W;
        jobs[t].n0 = N * t / nthreads; jobs[t].n1 = N * (t + 1) / nthreads; jobs[t].node_id_base = node_id_base;
        if (t > 0) pthread_create(&th[t], NULL, memo_worker, &jobs[t]);
    }
    memo_worker(&jobs[0]);
    uint64_t best[9];
    for (int k = 0; k <= 8; k++) best[k] = jobs[0].best[k];
    for (int t = 1; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        for (int k = 0; k <= 8; k++)
            if (jobs[t].best[k] < best[k]) best[k] = jobs[t].best[k];
    }
    for (int64_t p = 0; p < P; p++) {
        int k = pods[4 * p];
        out_keys[p] = (k >= 0 && k <= 8 && pods[4 * p + 3] <= 0) ? best[k] : KGPU_NO_FIT;
    }
    free(th);
    free(jobs);
}
