/* cgo_sequence.c -- the exact call sequence of go/kgpuscheduler/kgpu_cgo.go, in plain C99.
 *
 * The reference's host language is Go and there is no Go toolchain in this image, so the cgo shim cannot be
 * compiled here.  cgo compiles its preamble with the C compiler and calls the functions exactly as C would, so a
 * C99 program that includes the same header and makes the same calls with the same argument shapes is the
 * closest runnable stand-in: create -> uploadNodes -> scoreBatch -> scorePair (PodFitsDevice) -> setFreeMask
 * (TakePodResources) -> scoreBatch -> setFreeMasks (batched Take) -> fitTable/fitLookup -> placeBatch ->
 * freeMasks -> a failing call + kgpu_last_error -> destroy.  tests/test_cabi_sequence.py builds it with
 * `gcc -std=c99 -pedantic -Wall -Wextra -Werror` (header hygiene), runs it and checks every output against
 * the oracle.
 *
 * usage: cgo_sequence <topo.bin> <free.bin> <pods.bin> <N> <P> <out.bin>
 * exit:  0 ok | 3 no CUDA device (the library has no CPU path) | 1 anything else
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kgpu.h"

static void *slurp(const char *path, size_t bytes) {
    FILE *f = fopen(path, "rb");
    void *buf = malloc(bytes ? bytes : 1);
    if (!f || !buf || fread(buf, 1, bytes, f) != bytes) {
        fprintf(stderr, "cannot read %lu bytes from %s\n", (unsigned long)bytes, path);
        exit(1);
    }
    fclose(f);
    return buf;
}

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc__ = (call);                                                               \
        if (rc__ != KGPU_OK) {                                                           \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, kgpu_last_error(h));           \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char **argv) {
    if (argc != 7) {
        fprintf(stderr, "usage: %s topo.bin free.bin pods.bin N P out.bin\n", argv[0]);
        return 1;
    }
    const int64_t N = atoll(argv[4]), P = atoll(argv[5]);
    int32_t *topo = (int32_t *)slurp(argv[1], (size_t)N * 256);
    int32_t *free_mask = (int32_t *)slurp(argv[2], (size_t)N * 4);
    int32_t *pods = (int32_t *)slurp(argv[3], (size_t)P * 16);
    const int NPAIR = 8, NBATCH = 5;

    /* create(devs []int32) */
    kgpu_t *h = NULL;
    int dev = 0;
    if (kgpu_create(&dev, 1, &h) != KGPU_OK) {
        fprintf(stderr, "kgpu_create: %s\n", kgpu_last_error(NULL));
        return 3;
    }
    printf("libkgpu %s\n", kgpu_version());

    uint64_t *keys1 = (uint64_t *)malloc((size_t)P * 8), *keys2 = (uint64_t *)malloc((size_t)P * 8);
    uint64_t *keys3 = (uint64_t *)malloc((size_t)P * 8), *keys4 = (uint64_t *)malloc((size_t)P * 8);
    int32_t *masks = (int32_t *)malloc((size_t)N * 4);
    uint32_t nk[8], fit[8];
    int64_t pair_node[8], take_idx[5];
    int32_t pair_k[8], take_mask[5];
    if (!keys1 || !keys2 || !keys3 || !keys4 || !masks) return 1;

    /* uploadNodes, scoreBatch */
    CHECK(kgpu_upload_nodes(h, topo, free_mask, N, 0));
    if (kgpu_num_nodes(h) != N) return 1;
    CHECK(kgpu_score_batch(h, pods, P, keys1));

    /* scorePair: PodFitsDevice(node chosen for pod i, pod i) one pair per call, like the Go method */
    for (int i = 0; i < NPAIR; i++) {
        pair_node[i] = keys1[i] == KGPU_NO_FIT ? i : (int64_t)KGPU_KEY_NODE(keys1[i]);
        pair_k[i] = pods[4 * i];
        CHECK(kgpu_score_pairs(h, &pair_node[i], &pair_k[i], NULL, 1, &nk[i]));
    }

    /* setFreeMask: TakePodResources of pod 0's placement */
    if (keys1[0] != KGPU_NO_FIT) {
        const int64_t n0 = (int64_t)KGPU_KEY_NODE(keys1[0]);
        free_mask[n0] &= ~(int32_t)KGPU_KEY_MASK(keys1[0]);
        CHECK(kgpu_set_free_mask(h, n0, free_mask[n0]));
    }
    CHECK(kgpu_score_batch(h, pods, P, keys2));

    /* setFreeMasks: a batch of Takes in one call (nodes 1, 3, 5, 7, 9 lose their two lowest free GPUs) */
    for (int i = 0; i < NBATCH; i++) {
        int64_t n = 1 + 2 * i;
        int32_t m = free_mask[n];
        m &= m - 1;
        m &= m - 1;
        take_idx[i] = n;
        take_mask[i] = m;
        free_mask[n] = m;
    }
    CHECK(kgpu_set_free_masks(h, take_idx, take_mask, NBATCH));
    CHECK(kgpu_score_batch(h, pods, P, keys3));

    /* per-cycle fit table: PodFitsDevice served from the host copy, no launch per call */
    CHECK(kgpu_build_fit_table(h));
    for (int i = 0; i < NPAIR; i++) CHECK(kgpu_fit_lookup(h, pair_node[i], pair_k[i], &fit[i]));

    /* placeBatch + freeMasks */
    CHECK(kgpu_place_batch(h, pods, P, keys4));
    CHECK(kgpu_get_free_masks(h, masks, N));

    /* a failing call: code + message, handle stays usable */
    if (kgpu_set_free_mask(h, N, 0) != KGPU_ERR_INVALID || strlen(kgpu_last_error(h)) == 0) {
        fprintf(stderr, "out-of-range index was not rejected\n");
        return 1;
    }
    printf("error path: %s\n", kgpu_last_error(h));
    printf("launches %lld upload_ms %.3f\n", (long long)kgpu_kernel_launches(h), kgpu_last_upload_ms(h));

    FILE *out = fopen(argv[6], "wb");
    if (!out) return 1;
    fwrite(keys1, 8, (size_t)P, out);
    fwrite(nk, 4, (size_t)NPAIR, out);
    fwrite(keys2, 8, (size_t)P, out);
    fwrite(keys3, 8, (size_t)P, out);
    fwrite(fit, 4, (size_t)NPAIR, out);
    fwrite(keys4, 8, (size_t)P, out);
    fwrite(masks, 4, (size_t)N, out);
    fclose(out);
    CHECK(kgpu_destroy(h));
    return 0;
}
