/*
 * tests/harness/table_harness.c -- TEST INFRASTRUCTURE.
 *
 * Calls the 26 SINT32-data members of utils_arith_ntt_t (src/utils/arith/ntt.h:237-262) exactly as scheme code does:
 * through the function-pointer table that utils_arith_ntt(type) returns (arith.c:360-396), one polynomial per call,
 * host buffers, after init_reduce().  It does so on TWO libraries loaded side by side with dlopen(RTLD_LOCAL) --
 * libscgpu.so (the GPU drop-in) and oracle/_ref/libscref.so (the unmodified reference) -- on identical inputs and
 * compares every output word and return code.  With --threads T the whole sweep additionally runs from T pthreads
 * at once on one table (BLISS-B's worker threads call the members concurrently, bliss_b.c:74-175).
 *
 *   table_harness <libA.so> <libB.so> [--threads T] [--rounds R] [--time REPS]
 *
 * --time REPS: after the comparison, wall-clock per call of four members (BASELINE config C1: one polynomial per call
 * through the table, n = 512, q = 12289) on each library.
 *
 * Prints one line per (q, n, variant) and a summary; exit status 0 iff no word differs.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/scgpu_dropin.h"

typedef struct {
    void *h;
    const char *path;
    const utils_arith_ntt_t *(*table)(safecrypto_ntt_e);
    void (*init_reduce)(ntt_params_t *, size_t, SINT32);
} lib_t;

typedef struct {
    int q, n, has16;
    SINT16 *w16, *r16;      /* NULL when the modulus does not fit 16-bit tables */
    SINT32 *w32, *r32;
} pset_t;

#define MAXLEN 1280
#define NSLOTS 32
static const char *slot_name[NSLOTS] = {
    "modn_32", "muln_32", "sqrn_32", "pwr_32", "mul_32_sparse", "mul_32_sparse_16", "mul_32_pointwise",
    "mul_32_pointwise_16", "mul_32_scalar", "fft_32_32", "fft_32_32_large", "fft_32_16", "fft_32_16_large",
    "invert_32", "invert_32(zero coefficient)", "div_32", "flip_32", "center_32", "normalize_32",
    "normalize_32(3n)", "normalize_32(4n)", "normalize_32(5n)", "fwd_ntt_32_32", "inv_ntt_32_32(in place)",
    "fwd_ntt_32_32_large", "inv_ntt_32_32_large", "fwd_ntt_32_16", "inv_ntt_32_16(in place)",
    "fwd_ntt_32_16_large", "inv_ntt_32_16_large", "polymul composition", "div_32(zero denominator)"};

typedef struct { SINT32 v[NSLOTS][MAXLEN]; SINT32 rc[NSLOTS]; int used[NSLOTS]; } result_t;

static uint64_t rng_next(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static SINT32 rnd_range(uint64_t *s, SINT32 lo, SINT32 hi) { return lo + (SINT32)(rng_next(s) % (uint64_t)((int64_t)hi - lo + 1)); }

static void *xalloc(size_t bytes)
{
    void *p = NULL;
    if (posix_memalign(&p, 64, bytes ? bytes : 64)) { fprintf(stderr, "out of memory\n"); exit(2); }
    memset(p, 0, bytes);
    return p;
}

/* every member once, on inputs that depend only on (ps, seed) */
static void exercise(const lib_t *L, int variant, const pset_t *ps, uint64_t seed, result_t *R)
{
    const utils_arith_ntt_t *T = L->table((safecrypto_ntt_e)variant);
    ntt_params_t p;
    const int n = ps->n, q = ps->q;
    uint64_t s = seed;
    L->init_reduce(&p, (size_t)n, q);
    memset(R, 0, sizeof(*R));
    SINT32 *a = xalloc(sizeof(SINT32) * MAXLEN), *b = xalloc(sizeof(SINT32) * MAXLEN), *t = xalloc(sizeof(SINT32) * MAXLEN);
    SINT32 *u = xalloc(sizeof(SINT32) * MAXLEN);
    SINT16 *h = xalloc(sizeof(SINT16) * MAXLEN);
#define OUT(k) (R->used[k] = 1, R->v[k])
    /* scalars, extremes included */
    static const SINT32 edge[8] = {0, 1, -1, 2147483647, (-2147483647 - 1), 12289, -12289, 1 << 30};
    for (int i = 0; i < 64; i++) {
        SINT32 x = i < 8 ? edge[i] : (SINT32)rng_next(&s);
        SINT32 y = rnd_range(&s, -(q - 1), q - 1);
        SINT32 xs = rnd_range(&s, -4 * q, 4 * q);
        OUT(0)[i] = T->modn_32(x, &p);
        OUT(1)[i] = T->muln_32(xs, y, &p);
        OUT(2)[i] = T->sqrn_32(xs, &p);
        if (i < 16) OUT(3)[i] = T->pwr_32(rnd_range(&s, 1, q - 1), rnd_range(&s, 0, q - 1), &p);
    }
    /* sparse products: omega distinct positions */
    const UINT16 omega = 16;
    for (int i = 0; i < n; i++) { a[i] = rnd_range(&s, -200, 200); h[i] = (SINT16)a[i]; }
    for (int i = 0; i < omega; i++) u[i] = (SINT32)((rng_next(&s) % (uint64_t)(n / omega)) + (uint64_t)i * (n / omega));
    T->mul_32_sparse(OUT(4), (size_t)n, omega, a, u);
    T->mul_32_sparse_16(OUT(5), (size_t)n, omega, h, u);
    /* pointwise */
    for (int i = 0; i < n; i++) { a[i] = rnd_range(&s, -2 * q, 2 * q); b[i] = rnd_range(&s, 0, q - 1); h[i] = (SINT16)rnd_range(&s, -4000, 4000); }
    T->mul_32_pointwise(OUT(6), &p, a, b);
    T->mul_32_pointwise_16(OUT(7), &p, a, h);
    T->mul_32_scalar(OUT(8), &p, a, rnd_range(&s, -50, 50));
    /* bare fft passes on lazily reduced data */
    for (int i = 0; i < n; i++) a[i] = rnd_range(&s, -q, 2 * q);
    memcpy(OUT(9), a, sizeof(SINT32) * n);  T->fft_32_32(R->v[9], &p, ps->w32);
    memcpy(OUT(10), a, sizeof(SINT32) * n); T->fft_32_32_large(R->v[10], &p, ps->w32);
    if (ps->has16) {
        memcpy(OUT(11), a, sizeof(SINT32) * n); T->fft_32_16(R->v[11], &p, ps->w16);
        memcpy(OUT(12), a, sizeof(SINT32) * n); T->fft_32_16_large(R->v[12], &p, ps->w16);
    }
    /* invert / div with their return codes */
    for (int i = 0; i < n; i++) { a[i] = rnd_range(&s, 1, q - 1); b[i] = rnd_range(&s, 1, q - 1); }
    memcpy(OUT(13), a, sizeof(SINT32) * n); R->rc[13] = T->invert_32(R->v[13], &p, (size_t)n);
    memcpy(OUT(14), a, sizeof(SINT32) * n); R->v[14][37] = q; R->rc[14] = T->invert_32(R->v[14], &p, (size_t)n);
    memcpy(OUT(15), a, sizeof(SINT32) * n); R->rc[15] = T->div_32(R->v[15], b, &p, (size_t)n);
    memcpy(OUT(31), a, sizeof(SINT32) * n); b[n - 3] = 0; R->rc[31] = T->div_32(R->v[31], b, &p, (size_t)n);
    /* flip / center / normalize, normalize also on k*n words as module_lwe.c:742 passes */
    for (int i = 0; i < MAXLEN; i++) t[i] = rnd_range(&s, -(1 << 27), 1 << 27);
    memcpy(OUT(16), t, sizeof(SINT32) * n); T->flip_32(R->v[16], &p);
    memcpy(OUT(17), t, sizeof(SINT32) * n); T->center_32(R->v[17], (size_t)n, &p);
    memcpy(OUT(18), t, sizeof(SINT32) * n); T->normalize_32(R->v[18], (size_t)n, &p);
    for (int l = 3; l <= 5; l++) {
        memcpy(OUT(16 + l), t, sizeof(SINT32) * 256 * l);
        T->normalize_32(R->v[16 + l], (size_t)(256 * l), &p);
    }
    /* transforms: uniform or small signed coefficients */
    for (int i = 0; i < n; i++) {
        a[i] = (seed & 1) ? rnd_range(&s, -5, 5) : rnd_range(&s, 0, q - 1);
        b[i] = rnd_range(&s, 0, q - 1);
    }
    T->fwd_ntt_32_32(OUT(22), &p, a, ps->w32);
    memcpy(OUT(23), R->v[22], sizeof(SINT32) * n);
    T->inv_ntt_32_32(R->v[23], &p, R->v[23], ps->w32, ps->r32);
    T->fwd_ntt_32_32_large(OUT(24), &p, a, ps->w32);
    T->inv_ntt_32_32_large(OUT(25), &p, R->v[24], ps->w32, ps->r32);
    if (ps->has16) {
        T->fwd_ntt_32_16(OUT(26), &p, a, ps->w16);
        memcpy(OUT(27), R->v[26], sizeof(SINT32) * n);
        T->inv_ntt_32_16(R->v[27], &p, R->v[27], ps->w16, ps->r16);
        T->fwd_ntt_32_16_large(OUT(28), &p, a, ps->w16);
        T->inv_ntt_32_16_large(OUT(29), &p, R->v[28], ps->w16, ps->r16);
        /* the composition every scheme uses */
        T->fwd_ntt_32_16(t, &p, a, ps->w16);
        T->fwd_ntt_32_16(u, &p, b, ps->w16);
        T->mul_32_pointwise(t, &p, t, u);
        T->inv_ntt_32_16(OUT(30), &p, t, ps->w16, ps->r16);
    } else {
        T->fwd_ntt_32_32(t, &p, a, ps->w32);
        T->fwd_ntt_32_32(u, &p, b, ps->w32);
        T->mul_32_pointwise(t, &p, t, u);
        T->inv_ntt_32_32(OUT(30), &p, t, ps->w32, ps->r32);
    }
#undef OUT
    free(a); free(b); free(t); free(u); free(h);
}

static int compare(const result_t *A, const result_t *B, int q, int n, int variant, int verbose)
{
    int bad = 0;
    for (int k = 0; k < NSLOTS; k++) {
        if (A->used[k] != B->used[k]) { bad++; continue; }
        if (!A->used[k]) continue;
        if (memcmp(A->v[k], B->v[k], sizeof(A->v[k])) != 0 || A->rc[k] != B->rc[k]) {
            bad++;
            if (verbose) {
                int i = 0;
                while (i < MAXLEN && A->v[k][i] == B->v[k][i]) i++;
                fprintf(stderr, "MISMATCH q=%d n=%d variant=%d member %s: word %d  %d vs %d, rc %d vs %d\n", q, n, variant,
                        slot_name[k], i, i < MAXLEN ? A->v[k][i] : 0, i < MAXLEN ? B->v[k][i] : 0, A->rc[k], B->rc[k]);
            }
        }
    }
    return bad;
}

static lib_t load(const char *path)
{
    lib_t L;
    L.path = path;
    L.h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!L.h) { fprintf(stderr, "dlopen %s: %s\n", path, dlerror()); exit(2); }
    L.table = (const utils_arith_ntt_t *(*)(safecrypto_ntt_e))dlsym(L.h, "utils_arith_ntt");
    L.init_reduce = (void (*)(ntt_params_t *, size_t, SINT32))dlsym(L.h, "init_reduce");
    if (!L.table || !L.init_reduce) { fprintf(stderr, "%s lacks utils_arith_ntt / init_reduce\n", path); exit(2); }
    return L;
}

/* twiddles from roots_of_unity_s16/s32 of library B (the reference): what USE_RUNTIME_NTT_TABLES builds hand over */
static pset_t make_pset(const lib_t *B, int q, int n)
{
    pset_t ps;
    memset(&ps, 0, sizeof(ps));
    ps.q = q; ps.n = n; ps.has16 = q < 32768;
    SINT32 (*r32)(SINT32 *, SINT32 *, size_t, sc_ulimb_t, sc_ulimb_t, SINT32) = dlsym(B->h, "roots_of_unity_s32");
    SINT32 (*r16)(SINT16 *, SINT16 *, size_t, sc_ulimb_t, sc_ulimb_t, SINT32) = dlsym(B->h, "roots_of_unity_s16");
    if (!r32 || !r16) { fprintf(stderr, "roots_of_unity_* missing\n"); exit(2); }
    ps.w32 = xalloc(sizeof(SINT32) * n); ps.r32 = xalloc(sizeof(SINT32) * n);
    if (r32(ps.w32, ps.r32, (size_t)n, (sc_ulimb_t)q, 0, 0) != SC_FUNC_SUCCESS) { fprintf(stderr, "roots_of_unity_s32 failed\n"); exit(2); }
    if (ps.has16) {
        ps.w16 = xalloc(sizeof(SINT16) * n); ps.r16 = xalloc(sizeof(SINT16) * n);
        if (r16(ps.w16, ps.r16, (size_t)n, (sc_ulimb_t)q, 0, 0) != SC_FUNC_SUCCESS) { fprintf(stderr, "roots_of_unity_s16 failed\n"); exit(2); }
    }
    return ps;
}

typedef struct { const lib_t *A, *B; const pset_t *ps; int variant; uint64_t seed; int rounds; int bad; } job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    /* 64-byte aligned: the reference's AVX2 variant uses aligned vector accesses on caller buffers */
    result_t *ra = xalloc(sizeof(result_t)), *rb = xalloc(sizeof(result_t));
    for (int r = 0; r < j->rounds; r++) {
        exercise(j->A, j->variant, j->ps, j->seed + (uint64_t)r * 7919u, ra);
        exercise(j->B, j->variant, j->ps, j->seed + (uint64_t)r * 7919u, rb);
        j->bad += compare(ra, rb, j->ps->q, j->ps->n, j->variant, 1);
    }
    free(ra); free(rb);
    return NULL;
}

static double now_us(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

/* per-call latency of the members BLISS-B's sign core calls (bliss_b.c:1372-1384), n = 512, q = 12289 */
static void time_members(const lib_t *L, const pset_t *ps, int variant, int reps)
{
    const utils_arith_ntt_t *T = L->table((safecrypto_ntt_e)variant);
    ntt_params_t p;
    const int n = ps->n, q = ps->q;
    uint64_t s = 99;
    L->init_reduce(&p, (size_t)n, q);
    SINT32 *a = xalloc(sizeof(SINT32) * MAXLEN), *b = xalloc(sizeof(SINT32) * MAXLEN), *t = xalloc(sizeof(SINT32) * MAXLEN);
    SINT16 *h = xalloc(sizeof(SINT16) * MAXLEN);
    for (int i = 0; i < n; i++) { a[i] = rnd_range(&s, 0, q - 1); b[i] = rnd_range(&s, 0, q - 1); h[i] = (SINT16)rnd_range(&s, -2, 2); }
    double t0, us[5];
    for (int pass = 0; pass < 2; pass++) {           /* pass 0 warms up (plans, contexts, caches) */
        const int r = pass ? reps : 20;
        t0 = now_us(); for (int i = 0; i < r; i++) T->fwd_ntt_32_16(t, &p, a, ps->w16); us[0] = (now_us() - t0) / r;
        t0 = now_us(); for (int i = 0; i < r; i++) T->mul_32_pointwise_16(t, &p, t, h); us[1] = (now_us() - t0) / r;
        t0 = now_us(); for (int i = 0; i < r; i++) T->inv_ntt_32_16(t, &p, t, ps->w16, ps->r16); us[2] = (now_us() - t0) / r;
        t0 = now_us(); for (int i = 0; i < r; i++) T->normalize_32(t, (size_t)n, &p); us[3] = (now_us() - t0) / r;
        t0 = now_us(); for (int i = 0; i < r; i++) b[0] = T->muln_32(a[i % n], b[0] | 1, &p); us[4] = (now_us() - t0) / r;
    }
    printf("TIME %s variant=%d n=%d q=%d: fwd_ntt_32_16 %.2f us, mul_32_pointwise_16 %.2f us, inv_ntt_32_16 %.2f us, "
           "normalize_32 %.2f us, muln_32 %.2f us per call\n", L->path, variant, n, q, us[0], us[1], us[2], us[3], us[4]);
    free(a); free(b); free(t); free(h);
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s libA.so libB.so [--threads T] [--rounds R]\n", argv[0]); return 2; }
    int threads = 0, rounds = 2, time_reps = 0;
    for (int i = 3; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--time")) time_reps = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "--threads")) threads = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "--rounds")) rounds = atoi(argv[i + 1]);
    }
    lib_t A = load(argv[1]), B = load(argv[2]);
    static const int sets[4][2] = {{12289, 512}, {12289, 1024}, {7681, 256}, {8380417, 256}};
    /* safecrypto_ntt_e: 0 reference, 1 barrett, 2 fp, 3 avx, 4 solinas 7681, 5 solinas 8380417 */
    int total_bad = 0, total_members = 0;
    for (int si = 0; si < 4; si++) {
        pset_t ps = make_pset(&B, sets[si][0], sets[si][1]);
        for (int variant = 0; variant < 6; variant++) {
            if (variant == 4 && ps.q != 7681) continue;
            if (variant == 5 && ps.q != 8380417) continue;
            job_t j = {&A, &B, &ps, variant, 0x5CA1AB1Eull + (uint64_t)(si * 16 + variant), rounds, 0};
            worker(&j);
            int members = 0;
            for (int k = 0; k < NSLOTS; k++) members += (k < 11 || k > 12 || ps.has16) && (k < 26 || k > 29 || ps.has16);
            total_members += members * rounds;
            total_bad += j.bad;
            printf("q=%d n=%d variant=%d: %d member calls x %d rounds, %d mismatches\n", ps.q, ps.n, variant, members, rounds, j.bad);
        }
        if (threads > 1) {
            /* concurrent callers on one table, as BLISS-B's producer threads */
            pthread_t th[64];
            job_t jobs[64];
            if (threads > 64) threads = 64;
            for (int t = 0; t < threads; t++) {
                jobs[t] = (job_t){&A, &B, &ps, t % 4 < 3 ? t % 4 : 3, 0xC0FFEEull + (uint64_t)t * 104729u, rounds, 0};
                pthread_create(&th[t], NULL, worker, &jobs[t]);
            }
            int bad = 0;
            for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
            total_bad += bad;
            printf("q=%d n=%d: %d threads concurrently, %d mismatches\n", ps.q, ps.n, threads, bad);
        }
    }
    if (time_reps > 0) {
        pset_t ps = make_pset(&B, 12289, 512);
        for (int variant = 0; variant < 4; variant += 3) { time_members(&A, &ps, variant, time_reps); time_members(&B, &ps, variant, time_reps); }
    }
    printf("SUMMARY member_calls=%d mismatches=%d\n", total_members, total_bad);
    return total_bad ? 1 : 0;
}
