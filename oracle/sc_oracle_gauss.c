/*
 * oracle/sc_oracle_gauss.c -- TEST INFRASTRUCTURE ONLY (see sc_oracle.h).
 *
 * Restatement of the discrete Gaussian samplers of /root/reference/src/utils/sampling:
 *   CDF inversion, 32/64-bit          gaussian_cdf.c:536-774
 *   Knuth-Yao DDG walk, 32/64 rows    gaussian_knuth_yao.c:81-204,301-364
 *   Bernoulli (BLISS) rejection       gaussian_bernoulli.c:40-129,161-280
 *   vector wrappers + discard         sampling.c:68-228
 * Table construction uses the same C library calls on the same host (x87 long double expl,
 * float expf/powf/log2f), so tables are bit-identical to the reference's when built on the
 * same machine.
 */
#include "sc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* sc_math.c:447-452 over sc_log2 (floor log2) */
static size_t ceil_log2_sz(size_t x)
{
    size_t l = 0;
    while ((x >> (l + 1)) != 0) l++;
    if (x & (x - 1)) l++;
    return l;
}

/* sc_math.c:1066-1082 / 1084-1100: greedy binary expansion of a double in [0,1) */
static uint64_t bin_expansion(double x, int nbits)
{
    double val = 0, step = 0.5f;
    uint64_t res = 0;
    for (int i = 0; i < nbits; i++) {
        res <<= 1;
        if ((val + step) < x) { val += step; res |= 1; }
        step = step / 2;
    }
    return res;
}

/* ---- CDF ------------------------------------------------------------------------------- */

typedef struct { uint64_t *t64; uint32_t *t32; uint64_t *thi; int size; int precision; } cdf_t;

/* 128 / 192 / 256-bit tables are built by the reference's multi-precision float code (gaussian_cdf.c:192-318);
 * the port restates the SAMPLING over such a table, which the test harness injects here (one per precision). */
static uint64_t *g_high_tab[2];
static int g_high_size[2];
int orc_set_high_table(int precision, const uint64_t *words, int entries)
{
    /* 256-bit: gaussian_cdf_sample_256 stores its fourth draw in w[4] and reads w[3] uninitialised
     * (gaussian_cdf.c:519-522): no defined behaviour to restate */
    int k = precision == 128 ? 0 : precision == 192 ? 1 : -1;
    if (k < 0) return 1;
    if (!words) { free(g_high_tab[k]); g_high_tab[k] = NULL; return 0; }
    if (entries < 2) return 1;
    free(g_high_tab[k]);
    size_t bytes = (size_t)entries * (size_t)(precision / 8);
    g_high_tab[k] = malloc(bytes);
    memcpy(g_high_tab[k], words, bytes);
    g_high_size[k] = entries;
    return 0;
}

#define L_2_SQRTPI 1.128379167095512573896158903121545172L   /* SC_M_2_SQRTPIl */
#define L_SQRT1_2  0.707106781186547524400844362104849039L   /* SC_M_SQRT1_2l  */

static int cdf_build(cdf_t *c, int precision, int blinding, float tail, float sigma)
{
    int bits = (int)ceil_log2_sz((size_t)(tail * sigma));
    c->size = 1 << bits;
    c->precision = precision;
    c->t64 = NULL; c->t32 = NULL; c->thi = NULL;
    if (precision == 128 || precision == 192) {
        int k = precision == 128 ? 0 : 1;
        if (!g_high_tab[k]) return 1;
        size_t bytes = (size_t)g_high_size[k] * (size_t)(precision / 8);
        c->size = g_high_size[k];
        c->thi = malloc(bytes);
        memcpy(c->thi, g_high_tab[k], bytes);
        return 0;
    }
    if (precision == 64) {
        /* gaussian_cdf.c:555-610 */
        c->t64 = malloc(sizeof(uint64_t) * (size_t)c->size);
        if (blinding == ORC_BLINDING_SAMPLES) sigma *= L_SQRT1_2;
        long double d = L_2_SQRTPI * L_SQRT1_2 * 18446744073709551616.0L / sigma;
        long double e = -0.5L / (sigma * sigma);
        long double s = 0.5L * d;
        int i;
        c->t64[0] = 0;
        for (i = 1; i < c->size - 1; i++) {
            c->t64[i] = (uint64_t)s;
            if (c->t64[i] == 0) break;
            s += d * expl(e * ((long double)(i * i)));
        }
        for (; i < c->size; i++) c->t64[i] = 0xFFFFFFFFFFFFFFFFULL;
        return 0;
    }
    if (precision == 32) {
        /* gaussian_cdf.c:679-728: accumulators are FLOAT, constants are double */
        c->t32 = malloc(sizeof(uint32_t) * (size_t)c->size);
        if (blinding == ORC_BLINDING_SAMPLES) sigma *= M_SQRT1_2;
        float d = M_2_SQRTPI * M_SQRT1_2 * 4294967296.0 / sigma;
        float e = -0.5L / (sigma * sigma);
        float s = 0.5L * d;
        int i;
        c->t32[0] = 0;
        for (i = 1; i < c->size - 1; i++) {
            c->t32[i] = (uint32_t)s;
            if (c->t32[i] == 0) break;
            s += d * expl(e * ((float)(i * i)));
        }
        for (; i < c->size; i++) c->t32[i] = 0xFFFFFFFFu;
        return 0;
    }
    return 1;
}

static void cdf_free(cdf_t *c) { free(c->t64); free(c->t32); free(c->thi); }

/* gaussian_cdf.c:536-553 / 661-677: largest index a (by fixed halving steps) with l[a] < x */
static int32_t cdf_sample(const cdf_t *c, orc_prng_t *rng)
{
    uint32_t a = 0;
    if (c->precision > 64) {
        /* gaussian_cdf.c:480-532 (sample), :112-190 (compare_ge_prec, binary_search_*): x >= l[b] */
        int nw = c->precision / 64;
        uint64_t x[4];
        for (int i = 0; i < nw; i++) x[i] = orc_prng_64(rng);
        for (uint32_t st = (uint32_t)c->size >> 1; st > 0; st >>= 1) {
            uint32_t b = a + st;
            if (b >= (uint32_t)c->size) continue;
            const uint64_t *l = c->thi + (size_t)b * (size_t)nw;
            /* compare_ge_prec, gaussian_cdf.c:112-136, literally: retval = !x_lt_y | (equal & retval).  An equal
             * word gives !x_lt_y = 1, so the fold ends as "top word of x >= top word of l": the lower words
             * never decide (a quirk of the reference that parity has to keep). */
            unsigned ge = 1;
            for (int i = 0; i < nw; i++) {
                unsigned lt = x[i] < l[i], eq = x[i] == l[i];
                ge = (!lt) | (eq & ge);
            }
            if (ge) a = b;
        }
        return (x[0] & 1) ? (int32_t)a : -(int32_t)a;
    }
    if (c->precision == 64) {
        uint64_t x = orc_prng_64(rng);
        for (uint32_t st = (uint32_t)c->size >> 1; st > 0; st >>= 1) {
            uint32_t b = a + st;
            if (b < (uint32_t)c->size && c->t64[b] < x) a = b;
        }
        return (x & 1) ? (int32_t)a : -(int32_t)a;
    } else {
        uint32_t x = orc_prng_32(rng);
        for (uint32_t st = (uint32_t)c->size >> 1; st > 0; st >>= 1) {
            uint32_t b = a + st;
            if (b < (uint32_t)c->size && c->t32[b] < x) a = b;
        }
        return (x & 1) ? (int32_t)a : -(int32_t)a;
    }
}

int orc_cdf_table(int precision, int blinding, float tail, float sigma, void *out, size_t cap_entries)
{
    cdf_t c;
    if (cdf_build(&c, precision, blinding, tail, sigma)) return -1;
    if ((size_t)c.size <= cap_entries) {
        if (precision == 64) memcpy(out, c.t64, 8u * (size_t)c.size);
        else                 memcpy(out, c.t32, 4u * (size_t)c.size);
    }
    int size = c.size;
    cdf_free(&c);
    return size;
}

/* ---- Knuth-Yao ------------------------------------------------------------------------- */

typedef struct { int rows, cols, bound; uint8_t *pmat; } ky_t;

/* sc_math.c:1047-1062: 128 binary digits of a DOUBLE (the value itself carries 53) */
static unsigned __int128 bin_expansion_128(double x)
{
    double val = 0, step = 0.5f;
    unsigned __int128 res = 0;
    for (int i = 1; i < 129; i++) {
        res <<= 1;
        if ((val + step) < x) { val += step; res |= 1; }
        step = step / 2;
    }
    return res;
}

static void ky_build(ky_t *k, int bitwidth, int blinding, float tail, float sigma)
{
    /* gaussian_knuth_yao.c:126-189 with the 128-, 64- or 32-row table of :50-124; blinding scales sigma first (:144-146) */
    if (blinding == ORC_BLINDING_SAMPLES) sigma *= 0.7071067811865475244008443621L;
    k->bound = (int32_t)ceil(tail * sigma);
    k->rows = bitwidth;
    k->cols = k->bound + 1;
    k->pmat = malloc((size_t)k->rows * (size_t)k->cols);
    long double d = 0.7978845608028653558798L / sigma;
    long double e = -0.5L / (sigma * sigma);
    for (int col = 0; col < k->cols; col++) {
        long double pr = (col == 0) ? d : d * expl(e * ((long double)(col * col)));
        if (bitwidth == 128) {
            unsigned __int128 b128 = bin_expansion_128((double)pr);
            for (int row = 0; row < k->rows; row++)
                k->pmat[(size_t)row * (size_t)k->cols + (size_t)col] = (uint8_t)((b128 >> (127 - row)) & 1);
            continue;
        }
        uint64_t bitsv = bin_expansion((double)pr, bitwidth);
        for (int row = 0; row < k->rows; row++)
            k->pmat[(size_t)row * (size_t)k->cols + (size_t)col] = (uint8_t)((bitsv >> (bitwidth - 1 - row)) & 1);
    }
}

/* gaussian_knuth_yao.c:301-364.  After the first hit the reference `break`s out of the
 * column scan without re-aligning its table pointer, and keeps doubling the (now negative)
 * distance once per row: about 31 rows later the 32-bit distance wraps, turns positive and the
 * walk resumes from the drifted pointer, adding further columns to the sample.  That is what
 * the compiled reference returns, so it is restated literally, with the wrap-around made
 * explicit (unsigned arithmetic) instead of relying on signed overflow. */
static int32_t ky_sample(const ky_t *k, orc_prng_t *rng)
{
    for (;;) {
        int32_t dist = 0, sample = 0;
        const uint8_t *pm = k->pmat;
        uint32_t rnd = orc_prng_32(rng);
        for (int row = 0; row < k->rows; row++) {
            dist = (int32_t)(2u * (uint32_t)dist + (rnd & 1u));
            rnd >>= 1;
            if ((row & 0x1F) == 0x1F) rnd = orc_prng_32(rng);
            for (int col = 0; col < k->cols; col++) {
                dist = (int32_t)((uint32_t)dist - (uint32_t)*pm++);
                if (dist < 0) { sample += col; break; }
            }
        }
        rnd = orc_prng_32(rng);
        sample = sample % k->bound;
        if (sample == 0 && (rnd & 1)) continue;
        return (rnd & 2) ? sample : -sample;
    }
}

int orc_ky_table(int bitwidth, float tail, float sigma, uint8_t *pmat, size_t cap,
                 int32_t *rows, int32_t *cols, int32_t *bound)
{
    ky_t k;
    ky_build(&k, bitwidth & 0xFFF, (bitwidth >> 12) & 3, tail, sigma);        /* blinding rides in bits 12-13 */
    *rows = k.rows; *cols = k.cols; *bound = k.bound;
    size_t sz = (size_t)k.rows * (size_t)k.cols;
    if (sz <= cap) memcpy(pmat, k.pmat, sz);
    free(k.pmat);
    return (int)sz;
}

/* ---- Bernoulli ------------------------------------------------------------------------- */

typedef struct { int entries, maxval, maxlog; uint8_t tab[64][8]; } ber_t;

static void ber_build(ber_t *b, float tail, float sigma)
{
    /* gaussian_bernoulli.c:40-103 */
    float max_gauss_val = ceil(tail * sigma);
    b->maxval = (uint16_t)(int32_t)max_gauss_val;
    b->maxlog = (uint16_t)(int32_t)ceil(log2f(max_gauss_val));
    size_t max_val = ceil(log2f(tail * tail * sigma * sigma));
    b->entries = (int)max_val;
    for (size_t i = 0; i < max_val && i < 64; i++) {
        double temp = expf(-powf(2, i) / (2 * sigma * sigma));
        uint64_t bitsv = bin_expansion(temp, 64);
        for (int j = 0; j < 8; j++) b->tab[i][j] = (uint8_t)(bitsv >> (56 - 8 * j));
    }
}

/* gaussian_bernoulli.c:161-246 */
static uint32_t ber_candidate(const ber_t *b, orc_prng_t *rng)
{
    for (;;) {
        uint32_t val = orc_prng_var(rng, (size_t)b->maxlog);
        if (val >= (uint32_t)b->maxval) continue;
        uint32_t accept_mask = 0, x = val * val;
        int reject = 0;
        for (int j = 0; j < 8 && !reject; j++) {
            for (int i = b->entries; i--; ) {
                uint8_t r = (uint8_t)orc_prng_8(rng);
                uint8_t tv = b->tab[i][j];
                if (r < tv && ((accept_mask >> i) & 1) == 0) accept_mask |= (1u << i);
                if (r > tv && ((x >> i) & 1) == 1 && ((accept_mask >> i) & 1) == 0) { reject = 1; break; }
            }
        }
        if (!reject) return val;
    }
}

/* gaussian_bernoulli.c:248-280 */
static int32_t ber_sample(const ber_t *b, orc_prng_t *rng)
{
    for (;;) {
        int32_t val = (int32_t)ber_candidate(b, rng);
        uint32_t rnd = orc_prng_var(rng, 2);
        if (val == 0) { if (rnd < 2) continue; return 0; }
        return (rnd & 1) ? -val : val;
    }
}

int orc_ber_table(float tail, float sigma, uint8_t *tab, size_t cap,
                  int32_t *entries, int32_t *maxval, int32_t *maxlog)
{
    ber_t b;
    ber_build(&b, tail, sigma);
    *entries = b.entries; *maxval = b.maxval; *maxlog = b.maxlog;
    size_t sz = (size_t)b.entries * 8u;
    if (sz <= cap) memcpy(tab, b.tab, sz);
    return (int)sz;
}

/* ---- vector wrappers (sampling.c:68-228) --------------------------------------------------- */

typedef struct {
    int sampler;
    cdf_t cdf; ky_t ky; ber_t ber;
    orc_prng_t *rng;
    uint32_t thresh;
} smp_t;

static int32_t draw(smp_t *s)
{
    switch (s->sampler) {
    case ORC_SAMPLER_CDF:       return cdf_sample(&s->cdf, s->rng);
    case ORC_SAMPLER_KNUTH_YAO: return ky_sample(&s->ky, s->rng);
    default:                    return ber_sample(&s->ber, s->rng);
    }
}

/* sampling.c:85-105 */
static uint32_t discard_threshold(uint32_t discard)
{
    return discard == 2 ? 1u << 28 : discard == 4 ? 1u << 30 : discard == 6 ? 1u << 31 : 0;
}

static int discard_now(smp_t *s)
{
    if (s->thresh == 0) return 0;
    return orc_prng_32(s->rng) < s->thresh;
}

/* sampling.c:68-83 */
static size_t rand_range(orc_prng_t *rng, size_t x)
{
    size_t rem = 0xFFFFFFFFu % x;
    for (;;) {
        size_t y = orc_prng_32(rng);
        if (y >= (0xFFFFFFFFu - rem)) continue;
        return y % x;
    }
}

/* sampling.c:211-228 */
static void vec_normal(smp_t *s, int32_t *v, size_t n, int32_t centre)
{
    for (size_t i = 0; i < n; i++) {
        v[i] = draw(s) + centre;
        i -= (size_t)discard_now(s);
    }
}

/* sampling.c:127-145: inside-out Fisher-Yates; v[0] is drawn WITHOUT the centre */
static void vec_shuffle(smp_t *s, int32_t *v, size_t n, int32_t centre)
{
    v[0] = draw(s);
    for (size_t i = 1; i < n; i++) {
        size_t j = rand_range(s->rng, i);
        if (i != j) v[i] = v[j];
        v[j] = draw(s) + centre;
        i -= (size_t)discard_now(s);
    }
}

/* sampling.c:170-191 */
static void vec_blinding(smp_t *s, int32_t *v, size_t n, int32_t centre)
{
    vec_shuffle(s, v, n, centre);
    for (size_t i = 0; i < n; i++) v[i] -= draw(s);
    /* sampling.c:176,182-188: swap v[i] with v[i & (n - 1)] -- a no-op only when n is a power of two */
    uint32_t mask = (uint32_t)(n - 1);
    for (size_t i = 0; i < n; i++) {
        size_t j = i & mask;
        int32_t t = v[i]; v[i] = v[j]; v[j] = t;
    }
}

int orc_gauss_streams(int sampler, int precision, int blinding, int prng_type, float tail, float sigma,
                      uint32_t discard, const uint8_t *seeds, size_t seed_len, size_t nstreams,
                      size_t n, int32_t centre, int32_t *out, int threads, size_t calls_per_stream)
{
    smp_t proto;
    memset(&proto, 0, sizeof(proto));
    proto.sampler = sampler;
    proto.thresh = discard_threshold(discard);
    if (calls_per_stream == 0) calls_per_stream = 1;
    if (sampler == ORC_SAMPLER_CDF) {
        if (cdf_build(&proto.cdf, precision, blinding, tail, sigma)) return 1;
    } else if (sampler == ORC_SAMPLER_KNUTH_YAO) {
        if (precision != 32 && precision != 64 && precision != 128) return 1;
        ky_build(&proto.ky, precision, blinding, tail, sigma);
    } else if (sampler == ORC_SAMPLER_BERNOULLI) {
        ber_build(&proto.ber, tail, sigma);            /* blinding leaves these tables alone (gaussian_bernoulli.c:122-125) */
    } else {
        return 1;
    }
    int nt = threads > 0 ? threads : omp_get_max_threads();
    int fail = 0;
#pragma omp parallel for schedule(static) num_threads(nt) reduction(|:fail)
    for (size_t st = 0; st < nstreams; st++) {
        smp_t s = proto;
        s.rng = orc_prng_create(prng_type, seeds + st * seed_len, seed_len, 0);
        if (!s.rng) { fail |= 1; continue; }
        for (size_t c = 0; c < calls_per_stream; c++) {
            int32_t *v = out + (st * calls_per_stream + c) * n;
            /* ref_driver.c calls sample() directly for KY / Bernoulli (identical to sample_vector_32 when discard == 0);
             * here every sampler goes through the vector wrapper its blinding mode selects, as configure_sampler
             * wires them (sampling.c:395-413) in a build that has the sampler compiled in */
            if (blinding == ORC_SHUFFLE_SAMPLES) vec_shuffle(&s, v, n, centre);
            else if (blinding == ORC_BLINDING_SAMPLES)  vec_blinding(&s, v, n, centre);
            else                                         vec_normal(&s, v, n, centre);
        }
        orc_prng_destroy(s.rng);
    }
    if (sampler == ORC_SAMPLER_CDF) cdf_free(&proto.cdf);
    if (sampler == ORC_SAMPLER_KNUTH_YAO) free(proto.ky.pmat);
    return fail;
}
