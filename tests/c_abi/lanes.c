/* A caller of libmprg.so that is not Python: plain C99 + pthreads over include/mprg.h, the way INTEGRATION.md
 * section 6 describes several builds in flight on one GPU.  `lanes` host threads each own a context
 * (mprg_create, mprg_set_workers(ctx, 1), mprg_set_wait_mode(ctx, 2)) and build the same small set of synthetic loci
 * `rounds` times from packed host rows (mprg_pack_rows -> mprg_build_packed); every PRG string must equal the one a
 * single context built first.  Exit codes: 0 = all equal, 77 = no CUDA device (mprg_create said MPRG_E_NO_DEVICE:
 * the library has no CPU fallback), 1 = anything else.
 *
 *     gcc -std=c99 -O2 -I include tests/c_abi/lanes.c -o lanes make_prg_b200/libmprg.so -lpthread
 *     ./lanes [lanes=4] [rounds=6]
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mprg.h"

#define N_LOCI 64

static uint8_t *g_packed;
static int64_t g_offsets[N_LOCI];
static int32_t g_rows[N_LOCI], g_cols[N_LOCI], g_flags[N_LOCI];
static char *g_want[N_LOCI];
static int64_t g_want_len[N_LOCI];
static int g_rounds = 6;

static uint32_t lcg(uint32_t *s) { return *s = *s * 1664525u + 1013904223u; }

/* a few haplotypes, SNP columns, one deletion: enough for matches, clusters and nested sites */
static void make_locus(int l, uint8_t *ascii, int rows, int cols) {
    uint32_t s = 12345u + 977u * (uint32_t)l;
    const char *b = "ACGT";
    uint8_t hap[4][512];
    for (int c = 0; c < cols; ++c) hap[0][c] = (uint8_t)b[(lcg(&s) >> 16) & 3];
    for (int h = 1; h < 4; ++h) {
        memcpy(hap[h], hap[h - 1], (size_t)cols);
        for (int k = 0; k < 2 + h; ++k) {
            const int c = (int)((lcg(&s) >> 8) % (uint32_t)cols);
            hap[h][c] = (uint8_t)b[(lcg(&s) >> 16) & 3];
        }
    }
    const int d0 = (int)((lcg(&s) >> 8) % (uint32_t)(cols > 8 ? cols - 8 : 1));
    for (int c = d0; c < d0 + 3 && c < cols; ++c) hap[3][c] = '-';
    for (int r = 0; r < rows; ++r) memcpy(ascii + (size_t)r * cols, hap[(lcg(&s) >> 20) & 3], (size_t)cols);
}

static int build_once(mprg_ctx *ctx, int check) {
    mprg_batch *batch = NULL;
    mprg_result *res = NULL;
    int rc = mprg_build_packed(ctx, g_packed, g_offsets, g_rows, g_cols, g_flags, N_LOCI, 5, 7, &batch, &res);
    if (rc != MPRG_OK) {
        fprintf(stderr, "mprg_build_packed: %d %s\n", rc, mprg_last_error(ctx));
        return 1;
    }
    int bad = 0;
    for (int l = 0; l < N_LOCI; ++l) {
        int64_t n = 0;
        const char *p = mprg_result_prg(res, l, &n);
        if (mprg_result_status(res, l) != MPRG_LOCUS_OK) bad++;
        else if (!check) {
            g_want[l] = (char *)malloc((size_t)n + 1);
            memcpy(g_want[l], p, (size_t)n);
            g_want_len[l] = n;
        } else if (n != g_want_len[l] || memcmp(p, g_want[l], (size_t)n) != 0) bad++;
    }
    mprg_result_free(res);
    mprg_batch_free(ctx, batch);
    return bad;
}

static void *lane(void *arg) {
    long bad = 0;
    mprg_ctx *ctx = NULL;
    (void)arg;
    if (mprg_create(0, &ctx) != MPRG_OK) return (void *)1L;
    if (mprg_set_workers(ctx, 1) != MPRG_OK || mprg_set_wait_mode(ctx, 2) != MPRG_OK) bad++;
    for (int r = 0; r < g_rounds && !bad; ++r) bad += build_once(ctx, 1);
    mprg_destroy(ctx);
    return (void *)bad;
}

int main(int argc, char **argv) {
    const int lanes = argc > 1 ? atoi(argv[1]) : 4;
    if (argc > 2) g_rounds = atoi(argv[2]);
    if (lanes < 1 || lanes > 32 || g_rounds < 1) return 1;
    mprg_ctx *ctx = NULL;
    const int rc = mprg_create(0, &ctx);
    if (rc == MPRG_E_NO_DEVICE) {
        printf("no CUDA device: mprg_create -> MPRG_E_NO_DEVICE (no CPU fallback)\n");
        return 77;
    }
    if (rc != MPRG_OK) return 1;
    int64_t total = 0;
    for (int l = 0; l < N_LOCI; ++l) {
        g_rows[l] = 6 + (l * 5) % 40;
        g_cols[l] = 40 + (l * 37) % 400;
        g_offsets[l] = total;
        total += (int64_t)g_rows[l] * 16 * ((g_cols[l] + 31) / 32);
    }
    g_packed = (uint8_t *)malloc((size_t)total);
    uint8_t *ascii = (uint8_t *)malloc(46 * 512);
    for (int l = 0; l < N_LOCI; ++l) {
        make_locus(l, ascii, g_rows[l], g_cols[l]);
        const int64_t cap = (int64_t)g_rows[l] * 16 * ((g_cols[l] + 31) / 32);
        if (mprg_pack_rows(ascii, g_rows[l], g_cols[l], g_packed + g_offsets[l], cap, &g_flags[l]) != MPRG_OK) return 1;
    }
    free(ascii);
    if (build_once(ctx, 0) != 0) return 1; /* the answers, from one context alone */
    mprg_destroy(ctx);
    pthread_t th[32];
    long bad = 0;
    for (int i = 0; i < lanes; ++i) pthread_create(&th[i], NULL, lane, NULL);
    for (int i = 0; i < lanes; ++i) {
        void *r = NULL;
        pthread_join(th[i], &r);
        bad += (long)r;
    }
    printf("%d lanes x %d builds x %d loci: %ld mismatches\n", lanes, g_rounds, N_LOCI, bad);
    return bad ? 1 : 0;
}
