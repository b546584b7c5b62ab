/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin batch driver compiled INTO oracle/_ref/libscref.so next to the unmodified reference
 * sources (see oracle/Makefile).  The reference API processes one polynomial / one sampler
 * stream per call; this file only loops those calls over a batch (OpenMP, thread-private
 * state) so that tests can compare whole batches and bench.py can time the reference's CPU
 * path.  All arithmetic is done by the reference's own functions, reached through its own
 * dispatch table utils_arith_ntt() (src/utils/arith/arith.c:360-396) and create_sampler()
 * (src/utils/sampling/sampling.c:425-469).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#include "safecrypto_types.h"
#include "safecrypto_private.h"
#include "utils/arith/arith.h"
#include "utils/arith/ntt.h"
#include "utils/sampling/sampling.h"
#include "utils/sampling/gaussian_cdf.h"
#include "utils/sampling/gaussian_knuth_yao.h"
#include "utils/sampling/gaussian_bernoulli.h"
#include "utils/sampling/gaussian_knuth_yao_fast.h"
#include "utils/crypto/prng.h"
#include "utils/arith/module_lwe.h"

/* op codes shared with oracle/sc_oracle.h and include/scgpu.h (SCGPU_OP_*) */
enum {
    OP_FWD = 0, OP_INV, OP_FWD_LARGE, OP_INV_LARGE, OP_FFT, OP_FFT_LARGE,
    OP_PW, OP_PW16, OP_NORMALIZE, OP_CENTER, OP_POLYMUL, OP_TRIPLE16,
    OP_MODN, OP_MULN, OP_SQRN, OP_FLIP, OP_INVERT, OP_DIV, OP_PWR, OP_SCALAR,
    OP_SPARSE32, OP_SPARSE16
};

size_t ref_sizeof_ntt_params(void) { return sizeof(ntt_params_t); }
size_t ref_sizeof_ntt_table(void)  { return sizeof(utils_arith_ntt_t); }

void ref_get_params(int n, int q, int32_t *m, int32_t *k, double *inv_q_dbl, float *inv_q_flt)
{
    ntt_params_t p;
    init_reduce(&p, (size_t)n, q);
    *m = p.u.ntt32.m; *k = p.u.ntt32.k; *inv_q_dbl = p.inv_q_dbl; *inv_q_flt = p.inv_q_flt;
}

static int max_threads(int threads)
{
    /* an explicit request wins over OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1) */
    return threads > 0 ? threads : omp_get_max_threads();
}

/* One reference call per batch row.  `a`, `b`, `out` are [count][n] SINT32 row-major
 * (b may be a single shared row when b_stride == 0; for *_16 operands b points at SINT16).
 * Returns the OR of the per-row return codes (invert/div) and writes them to rc[] if given. */
int ref_ntt_batch(int variant, int op, int n, int q, int tw_bits,
                  int32_t *out, const int32_t *a, const void *b, size_t b_stride,
                  const void *w, const void *r, size_t count, int threads, int32_t *rc,
                  int32_t scalar)
{
    const utils_arith_ntt_t *T = utils_arith_ntt((safecrypto_ntt_e)variant);
    int any = 0;
    ntt_params_t p;
    init_reduce(&p, (size_t)n, q);
    ntt_table = T;
    int nt = max_threads(threads);
#pragma omp parallel for schedule(static) num_threads(nt) reduction(|:any)
    for (size_t i = 0; i < count; i++) {
        int32_t *v = out + i * (size_t)n;
        const int32_t *t = a ? a + i * (size_t)n : NULL;
        const int32_t *u32 = b ? (const int32_t *)b + i * b_stride : NULL;
        const int16_t *u16 = b ? (const int16_t *)b + i * b_stride : NULL;
        int32_t ret = 0;
        switch (op) {
        case OP_FWD:
            if (tw_bits == 16) T->fwd_ntt_32_16(v, &p, t, (const SINT16 *)w);
            else               T->fwd_ntt_32_32(v, &p, t, (const SINT32 *)w);
            break;
        case OP_FWD_LARGE:
            if (tw_bits == 16) T->fwd_ntt_32_16_large(v, &p, t, (const SINT16 *)w);
            else               T->fwd_ntt_32_32_large(v, &p, t, (const SINT32 *)w);
            break;
        case OP_INV:
            if (tw_bits == 16) T->inv_ntt_32_16(v, &p, t, (const SINT16 *)w, (const SINT16 *)r);
            else               T->inv_ntt_32_32(v, &p, t, (const SINT32 *)w, (const SINT32 *)r);
            break;
        case OP_INV_LARGE:
            if (tw_bits == 16) T->inv_ntt_32_16_large(v, &p, t, (const SINT16 *)w, (const SINT16 *)r);
            else               T->inv_ntt_32_32_large(v, &p, t, (const SINT32 *)w, (const SINT32 *)r);
            break;
        case OP_FFT:
            memcpy(v, t, sizeof(int32_t) * (size_t)n);
            if (tw_bits == 16) T->fft_32_16(v, &p, (const SINT16 *)w);
            else               T->fft_32_32(v, &p, (const SINT32 *)w);
            break;
        case OP_FFT_LARGE:
            memcpy(v, t, sizeof(int32_t) * (size_t)n);
            if (tw_bits == 16) T->fft_32_16_large(v, &p, (const SINT16 *)w);
            else               T->fft_32_32_large(v, &p, (const SINT32 *)w);
            break;
        case OP_PW:        T->mul_32_pointwise(v, &p, t, u32); break;
        case OP_PW16:      T->mul_32_pointwise_16(v, &p, t, u16); break;
        case OP_NORMALIZE: memcpy(v, t, sizeof(int32_t) * (size_t)n); T->normalize_32(v, (size_t)n, &p); break;
        case OP_CENTER:    memcpy(v, t, sizeof(int32_t) * (size_t)n); T->center_32(v, (size_t)n, &p); break;
        case OP_FLIP:      memcpy(v, t, sizeof(int32_t) * (size_t)n); T->flip_32(v, &p); break;
        case OP_POLYMUL: {
            /* the composition every scheme uses: fwd(a), fwd(b), pointwise, inv */
            int32_t tmp[1024] __attribute__((aligned(32)));
            if (tw_bits == 16) {
                T->fwd_ntt_32_16(v, &p, t, (const SINT16 *)w);
                T->fwd_ntt_32_16(tmp, &p, u32, (const SINT16 *)w);
                T->mul_32_pointwise(v, &p, v, tmp);
                T->inv_ntt_32_16(v, &p, v, (const SINT16 *)w, (const SINT16 *)r);
            } else {
                T->fwd_ntt_32_32(v, &p, t, (const SINT32 *)w);
                T->fwd_ntt_32_32(tmp, &p, u32, (const SINT32 *)w);
                T->mul_32_pointwise(v, &p, v, tmp);
                T->inv_ntt_32_32(v, &p, v, (const SINT32 *)w, (const SINT32 *)r);
            }
        } break;
        case OP_TRIPLE16:
            /* BLISS-B sign/verify: bliss_b.c:1378-1384 / 1682-1684 */
            T->fwd_ntt_32_16(v, &p, t, (const SINT16 *)w);
            T->mul_32_pointwise_16(v, &p, v, u16);
            T->inv_ntt_32_16(v, &p, v, (const SINT16 *)w, (const SINT16 *)r);
            break;
        case OP_MODN: for (int j = 0; j < n; j++) v[j] = T->modn_32(t[j], &p); break;
        case OP_MULN: for (int j = 0; j < n; j++) v[j] = T->muln_32(t[j], u32[j], &p); break;
        case OP_SQRN: for (int j = 0; j < n; j++) v[j] = T->sqrn_32(t[j], &p); break;
        case OP_PWR:  for (int j = 0; j < n; j++) v[j] = T->pwr_32(t[j], u32[j], &p); break;
        case OP_INVERT:
            memcpy(v, t, sizeof(int32_t) * (size_t)n);
            ret = T->invert_32(v, &p, (size_t)n);
            break;
        case OP_DIV:
            memcpy(v, t, sizeof(int32_t) * (size_t)n);
            ret = T->div_32(v, u32, &p, (size_t)n);
            break;
        case OP_SCALAR:   T->mul_32_scalar(v, &p, t, scalar); break;
        case OP_SPARSE32: T->mul_32_sparse(v, (size_t)n, (UINT16)scalar, t, u32); break;
        case OP_SPARSE16: T->mul_32_sparse_16(v, (size_t)n, (UINT16)scalar, (const SINT16 *)(const void *)((const int16_t *)a + i * (size_t)n), u32); break;
        default: ret = -1;
        }
        if (rc) rc[i] = ret;
        any |= ret;
    }
    return any;
}

/* ---- PRNG -------------------------------------------------------------------------- */

static prng_ctx_t *make_prng(int prng_type, const uint8_t *seed, size_t seed_len, size_t seed_period)
{
    prng_ctx_t *ctx = prng_create(SC_ENTROPY_USER_PROVIDED, (safecrypto_prng_e)prng_type,
                                  SC_PRNG_THREADING_NONE, seed_period ? seed_period : 0x00100000);
    if (!ctx) return NULL;
    prng_set_entropy(ctx, seed, seed_len);
    prng_init(ctx, (const UINT8 *)"SAFEcrypto nonce", 16);
    return ctx;
}

/* `script` is a list of (kind,arg) pairs: kind 32 -> prng_32, 64 -> prng_64 (hi then lo word
 * written), 8 -> prng_8, 1 -> prng_bit, 0 -> prng_var(arg).  One output word per draw except
 * 64 (two). */
int ref_prng_script(int prng_type, const uint8_t *seed, size_t seed_len, size_t seed_period,
                    const int32_t *script, size_t ndraws, uint32_t *out)
{
    prng_ctx_t *ctx = make_prng(prng_type, seed, seed_len, seed_period);
    if (!ctx) return -1;
    size_t o = 0;
    for (size_t i = 0; i < ndraws; i++) {
        int kind = script[2 * i], arg = script[2 * i + 1];
        if (kind == 32) out[o++] = prng_32(ctx);
        else if (kind == 64) { UINT64 x = prng_64(ctx); out[o++] = (uint32_t)(x >> 32); out[o++] = (uint32_t)x; }
        else if (kind == 8) out[o++] = prng_8(ctx);
        else if (kind == 1) out[o++] = (uint32_t)prng_bit(ctx);
        else if (kind == 16) out[o++] = prng_16(ctx);
        else if (kind == 128) {
            UINT128 x = prng_128(ctx);
            out[o++] = (uint32_t)(x >> 96); out[o++] = (uint32_t)(x >> 64); out[o++] = (uint32_t)(x >> 32); out[o++] = (uint32_t)x;
        }
        else if (kind == 2) { FLOAT f = prng_float(ctx); memcpy(&out[o++], &f, 4); }
        else if (kind == 3) { DOUBLE d = prng_double(ctx); memcpy(&out[o], &d, 8); o += 2; }
        else if (kind == 4) {
            size_t nw = ((size_t)arg + 3) / 4;
            memset(out + o, 0, nw * 4);
            prng_mem(ctx, (UINT8 *)(out + o), arg);
            o += nw;
        }
        else if (kind == 5) {
            /* reset_chacha20 frees the generator (chacha20_csprng.c:58-67): only the DRBG has a defined reset */
            if (prng_type != SC_PRNG_AES_CTR_DRBG) { prng_destroy(ctx); return -2; }
            prng_reset(ctx);
        }
        else if (kind == 6) { out[o++] = (uint32_t)prng_get_csprng_bytes(ctx); out[o++] = (uint32_t)prng_get_out_bytes(ctx); }
        else out[o++] = prng_var(ctx, (size_t)arg);
    }
    prng_destroy(ctx);
    return (int)o;
}

/* ---- Gaussian samplers --------------------------------------------------------------- */

/* private struct layouts mirrored from gaussian_cdf.c:74-92, gaussian_knuth_yao.c:27-38,
 * gaussian_bernoulli.c:26-38 (test-side introspection of the tables only) */
SC_STRUCT_PACK_START
typedef struct { UINT64 *cdf; SINT32 cdf_size; SINT32 k; SINT32 use_kl; prng_ctx_t *prng; } SC_STRUCT_PACKED drv_cdf64_t;
typedef struct { UINT32 *cdf; SINT32 cdf_size; SINT32 k; SINT32 use_kl; prng_ctx_t *prng; } SC_STRUCT_PACKED drv_cdf32_t;
typedef struct { void *cdf_256; void *cdf_192; void *cdf_128; SINT32 cdf_size; SINT32 k; SINT32 use_kl; prng_ctx_t *prng; } SC_STRUCT_PACKED drv_cdfh_t;
typedef struct { SINT32 num_rows, num_cols; FLOAT tailcut; SINT32 bound; UINT8 *prelut; SINT32 *hamming; UINT8 *pmat; prng_ctx_t *prng; } SC_STRUCT_PACKED drv_ky_t;
typedef struct { UINT16 max_gauss_val, max_gauss_log; FLOAT sigma; UINT16 max_ber_entries, max_ber_bytes; SINT32 bits; UINT8 **ber_table; SINT32 reject_counter; prng_ctx_t *prng; } SC_STRUCT_PACKED drv_ber_t;
SC_STRUCT_PACK_END

int ref_cdf_table(int precision, int blinding, float tail, float sigma, void *out, size_t cap_entries)
{
    static const uint8_t z[64] = {0};
    prng_ctx_t *ctx = make_prng(SC_PRNG_CHACHA, z, 64, 0);
    int size = -1;
    if (precision == 64) {
        void *g = gaussian_cdf_create_64(ctx, tail, sigma, 0, (sample_blinding_e)blinding);
        drv_cdf64_t *c = (drv_cdf64_t *)g;
        size = c->cdf_size;
        if ((size_t)size <= cap_entries) memcpy(out, c->cdf, (size_t)size * 8);
        gaussian_cdf_destroy_64(&g);
    } else if (precision == 32) {
        void *g = gaussian_cdf_create_32(ctx, tail, sigma, 0, (sample_blinding_e)blinding);
        drv_cdf32_t *c = (drv_cdf32_t *)g;
        size = c->cdf_size;
        if ((size_t)size <= cap_entries) memcpy(out, c->cdf, (size_t)size * 4);
        gaussian_cdf_destroy_32(&g);
    } else if (precision == 128 || precision == 192 || precision == 256) {
        /* gauss_cdf_high_t (gaussian_cdf.c:60-72): entries of precision/64 limbs, limb 0 least significant */
        void *g = precision == 128 ? gaussian_cdf_create_128(ctx, tail, sigma, 0, (sample_blinding_e)blinding)
                : precision == 192 ? gaussian_cdf_create_192(ctx, tail, sigma, 0, (sample_blinding_e)blinding)
                                   : gaussian_cdf_create_256(ctx, tail, sigma, 0, (sample_blinding_e)blinding);
        drv_cdfh_t *c = (drv_cdfh_t *)g;
        const void *tab = precision == 128 ? c->cdf_128 : precision == 192 ? c->cdf_192 : c->cdf_256;
        size = c->cdf_size;
        if ((size_t)size <= cap_entries) memcpy(out, tab, (size_t)size * (size_t)(precision / 8));
        if (precision == 128) gaussian_cdf_destroy_128(&g);
        else if (precision == 192) gaussian_cdf_destroy_192(&g);
        else gaussian_cdf_destroy_256(&g);
    }
    prng_destroy(ctx);
    return size;
}

int ref_ky_table(int bitwidth, float tail, float sigma, uint8_t *pmat, size_t cap, int32_t *rows, int32_t *cols, int32_t *bound)
{
    static const uint8_t z[64] = {0};
    prng_ctx_t *ctx = make_prng(SC_PRNG_CHACHA, z, 64, 0);
    const sample_blinding_e bl = (sample_blinding_e)((bitwidth >> 12) & 3);     /* blinding rides in bits 12-13 */
    bitwidth &= 0xFFF;
    void *g = (bitwidth == 32) ? gaussian_knuth_yao_create_32(ctx, tail, sigma, 0, bl)
            : (bitwidth == 128) ? gaussian_knuth_yao_create_128(ctx, tail, sigma, 0, bl)
                                : gaussian_knuth_yao_create_64(ctx, tail, sigma, 0, bl);
    drv_ky_t *k = (drv_ky_t *)g;
    *rows = k->num_rows; *cols = k->num_cols; *bound = k->bound;
    size_t sz = (size_t)k->num_rows * (size_t)k->num_cols;
    if (sz <= cap) memcpy(pmat, k->pmat, sz);
    prng_destroy(ctx);
    return (int)sz;
}

int ref_ber_table(float tail, float sigma, uint8_t *tab, size_t cap, int32_t *entries, int32_t *maxval, int32_t *maxlog)
{
    static const uint8_t z[64] = {0};
    prng_ctx_t *ctx = make_prng(SC_PRNG_CHACHA, z, 64, 0);
    void *g = bernoulli_create_64(ctx, tail, sigma, 0, NORMAL_SAMPLES);
    drv_ber_t *b = (drv_ber_t *)g;
    *entries = b->max_ber_entries; *maxval = b->max_gauss_val; *maxlog = b->max_gauss_log;
    size_t sz = (size_t)b->max_ber_entries * b->max_ber_bytes;
    if (sz <= cap) memcpy(tab, b->ber_table[0], sz);
    bernoulli_destroy_64(&g);
    prng_destroy(ctx);
    return (int)sz;
}

/* gauss_knuth_yao_fast_t, gaussian_knuth_yao_fast.c:27-38.  The tables are constants of the reference's source; tests
 * read them out of the compiled library at run time and hand them to the GPU plan. */
SC_STRUCT_PACK_START
typedef struct { SINT32 num_rows, num_cols; UINT32 dist1_mask, dist2_mask; const UINT8 *pre_lut_1, *pre_lut_2, *pmat; prng_ctx_t *prng; } SC_STRUCT_PACKED drv_kyf_t;
SC_STRUCT_PACK_END

static void *kyfast_create(prng_ctx_t *ctx, int dimension)
{
    return dimension == 512 ? gaussian_knuth_yao_fast_512_create(ctx, 0.0f, 4.8591f, 0, NORMAL_SAMPLES)
                            : gaussian_knuth_yao_fast_256_create(ctx, 0.0f, 4.5120f, 0, NORMAL_SAMPLES);
}

/* dims: rows, cols, dist1_mask, dist2_mask, lut2 length */
int ref_kyfast_tables(int dimension, uint8_t *lut1, uint8_t *lut2, size_t lut2_cap, uint8_t *pmat, size_t pmat_cap, int32_t *dims)
{
    static const uint8_t z[64] = {0};
    prng_ctx_t *ctx = make_prng(SC_PRNG_CHACHA, z, 64, 0);
    drv_kyf_t *k = (drv_kyf_t *)kyfast_create(ctx, dimension);
    if (!k) { prng_destroy(ctx); return 1; }
    const size_t l2 = 32 * ((size_t)k->dist1_mask + 1), pm = (size_t)k->num_rows * (size_t)k->num_cols;
    dims[0] = k->num_rows; dims[1] = k->num_cols; dims[2] = (int32_t)k->dist1_mask; dims[3] = (int32_t)k->dist2_mask; dims[4] = (int32_t)l2;
    if (l2 > lut2_cap || pm > pmat_cap) { prng_destroy(ctx); return 2; }
    memcpy(lut1, k->pre_lut_1, 256);
    memcpy(lut2, k->pre_lut_2, l2);
    memcpy(pmat, k->pmat, pm);
    void *g = k;
    gaussian_knuth_yao_fast_destroy(&g);
    prng_destroy(ctx);
    return 0;
}

/* Optional caller-supplied 128 / 192-bit CDF table written over the one create_sampler() built: the reference's
 * own table construction needs GMP/MPFR to be meaningful (with USE_SAFECRYPTO_FLOAT_MP every entry comes out
 * as {2, 2, ..}), the SAMPLING over a table is what the GPU path is compared against. */
static const uint64_t *g_high_tab[2];
static int g_high_entries[2];
int ref_set_high_table(int precision, const uint64_t *words, int entries)
{
    int k = precision == 128 ? 0 : precision == 192 ? 1 : -1;
    if (k < 0) return 1;
    g_high_tab[k] = words;            /* caller keeps it alive; NULL clears */
    g_high_entries[k] = entries;
    return 0;
}

/* One independent PRNG stream + sampler per row; row i is seeded with seeds[i*seed_len ..].
 * sampler: 0 CDF (through create_sampler/get_vector_32), 1 Knuth-Yao, 5 Bernoulli (direct
 * create/sample calls as src/unit/unit_sampling.c does: they are not reachable through
 * create_sampler in a non-constrained build, sampling.h:23-34). */
int ref_gauss_streams(int sampler, int precision, int blinding, int prng_type, float tail, float sigma,
                      uint32_t discard, const uint8_t *seeds, size_t seed_len, size_t nstreams,
                      size_t n, int32_t centre, int32_t *out, int threads, size_t calls_per_stream)
{
    int fail = 0;
    int nt = max_threads(threads);
    if (calls_per_stream == 0) calls_per_stream = 1;
#pragma omp parallel for schedule(static) num_threads(nt) reduction(|:fail)
    for (size_t s = 0; s < nstreams; s++) {
        prng_ctx_t *ctx = make_prng(prng_type, seeds + s * seed_len, seed_len, 0);
        int32_t *v = out + s * n * calls_per_stream;
        if (!ctx) { fail |= 1; continue; }
        if (sampler == CDF_GAUSSIAN_SAMPLING) {
            utils_sampling_t *smp = create_sampler(CDF_GAUSSIAN_SAMPLING, (sample_precision_e)precision,
                (sample_blinding_e)blinding, (SINT32)n, SAMPLING_DISABLE_BOOTSTRAP, ctx, tail, sigma);
            if (!smp) { fail |= 1; prng_destroy(ctx); continue; }
            if (precision == 128 || precision == 192) {
                int k = precision == 128 ? 0 : 1;
                drv_cdfh_t *c = (drv_cdfh_t *)smp->gauss;
                if (g_high_tab[k]) {
                    if (g_high_entries[k] != c->cdf_size) { fail |= 1; destroy_sampler(&smp); prng_destroy(ctx); continue; }
                    memcpy(precision == 128 ? c->cdf_128 : c->cdf_192, g_high_tab[k], (size_t)c->cdf_size * (size_t)(precision / 8));
                }
            }
            set_discard(smp, discard);
            for (size_t c = 0; c < calls_per_stream; c++)
                get_vector_32(smp, v + c * n, n, (FLOAT)centre);
            destroy_sampler(&smp);
        } else if (sampler == KNUTH_YAO_GAUSSIAN_SAMPLING) {
            /* blinding only scales the table here (gaussian_knuth_yao.c:144-146); the loop is sample_vector_32's */
            if (blinding == SHUFFLE_SAMPLES) { fail |= 1; prng_destroy(ctx); continue; }
            void *g = (precision == 32) ? gaussian_knuth_yao_create_32(ctx, tail, sigma, 0, (sample_blinding_e)blinding)
                    : (precision == 128) ? gaussian_knuth_yao_create_128(ctx, tail, sigma, 0, (sample_blinding_e)blinding)
                                         : gaussian_knuth_yao_create_64(ctx, tail, sigma, 0, (sample_blinding_e)blinding);
            for (size_t j = 0; j < n * calls_per_stream; j++) v[j] = gaussian_knuth_yao_sample(g) + centre;
        } else if (sampler == KNUTH_YAO_FAST_GAUSSIAN_SAMPLING) {
            /* `precision` carries the dimension (256 / 512) that selects the reference's table set */
            void *g = kyfast_create(ctx, precision);
            if (!g) { fail |= 1; prng_destroy(ctx); continue; }
            for (size_t j = 0; j < n * calls_per_stream; j++) v[j] = gaussian_knuth_yao_fast_sample(g) + centre;
            gaussian_knuth_yao_fast_destroy(&g);
        } else if (sampler == BERNOULLI_GAUSSIAN_SAMPLING) {
            void *g = bernoulli_create_64(ctx, tail, sigma, 0, NORMAL_SAMPLES);
            for (size_t j = 0; j < n * calls_per_stream; j++) v[j] = bernoulli_sample_64(g) + centre;
            bernoulli_destroy_64(&g);
        } else {
            fail |= 1;
        }
        prng_destroy(ctx);
    }
    return fail;
}

/* ---- module product with a CSPRNG-sampled matrix (module_lwe.c:588-748) ------------------------------------------ */

/* One create_rand_product_{16,32}_csprng call per instance, exactly as kyber / dilithium make it: a CSPRNG created as
 * create_csprng() does (module_lwe.c:914-940: user-provided entropy = the instance's seed, 16 MiB reseed period).
 * y: [count][l][n], t: [count][k][n].  Also returns the matrix the same generator state would draw
 * (uniform_random_ring_q_csprng ring by ring) when A != NULL: [count][k l][n] in DRAW order. */
int ref_rand_product(int tw_bits, int variant, int n, int q, int q_bits, int k, int l, int transpose, int prng_type,
                     const uint8_t *seeds, size_t seed_len, const int32_t *y, int32_t *t, int32_t *A,
                     const void *w, const void *r, size_t count, int threads)
{
    static const UINT8 nonce[16] = "dilithiumcrystal";
    const utils_arith_ntt_t *T = utils_arith_ntt((safecrypto_ntt_e)variant);
    const utils_arith_poly_t *P = utils_arith_poly();
    int nt = max_threads(threads), fail = 0;
#pragma omp parallel for schedule(static) num_threads(nt) reduction(|:fail)
    for (size_t s = 0; s < count; s++) {
        ntt_params_t p;
        init_reduce(&p, (size_t)n, q);
        int32_t *yb = NULL, *c = NULL, *tmp = NULL, *tb = NULL;
        if (posix_memalign((void **)&yb, 64, sizeof(int32_t) * (size_t)l * n) || posix_memalign((void **)&c, 64, sizeof(int32_t) * (size_t)n) ||
            posix_memalign((void **)&tmp, 64, sizeof(int32_t) * (size_t)(l + k) * n) || posix_memalign((void **)&tb, 64, sizeof(int32_t) * (size_t)(k + l) * n)) { fail |= 1; continue; }
        memset(tb, 0, sizeof(int32_t) * (size_t)(k + l) * n);
        for (int pass = 0; pass < (A ? 2 : 1); pass++) {
            prng_ctx_t *ctx = prng_create(SC_ENTROPY_USER_PROVIDED, (safecrypto_prng_e)prng_type, SC_PRNG_THREADING_NONE, 0x01000000);
            if (!ctx) { fail |= 1; break; }
            prng_set_entropy(ctx, seeds + s * seed_len, seed_len);
            prng_init(ctx, nonce, 16);
            if (pass == 1) {
                for (int ring = 0; ring < k * l; ring++)
                    uniform_random_ring_q_csprng(ctx, A + (s * (size_t)(k * l) + ring) * n, (size_t)n, q, (UINT32)q_bits);
            } else {
                memcpy(yb, y + s * (size_t)l * n, sizeof(int32_t) * (size_t)l * n);
                if (tw_bits == 16)
                    create_rand_product_16_csprng(ctx, (UINT32)q, (UINT32)q_bits, tb, yb, (size_t)n, (size_t)k, (size_t)l, c, tmp,
                                                  RND_PRD_DISABLE_OVERWRITE, transpose, (const SINT16 *)w, (const SINT16 *)r, P, T, &p);
                else
                    create_rand_product_32_csprng(ctx, (UINT32)q, (UINT32)q_bits, tb, yb, (size_t)n, (size_t)k, (size_t)l, c, tmp,
                                                  RND_PRD_DISABLE_OVERWRITE, transpose, (const SINT32 *)w, (const SINT32 *)r, P, T, &p);
                memcpy(t + s * (size_t)k * n, tb, sizeof(int32_t) * (size_t)k * n);
            }
            prng_destroy(ctx);
        }
        free(yb); free(c); free(tmp); free(tb);
    }
    return fail;
}

/* ---- Micciancio-Walter bootstrap (mw_bootstrap.c, wired by create_sampler, sampling.c:449-457) ------------------- */

/* Per stream: create_sampler(CDF, 64-bit, NORMAL, SAMPLING_MW_BOOTSTRAP) and either n get_vector_32 samples at
 * (sigma, centre[0]) when per_sample == 0, or n get_bootstrap_sample(sigma_i, centre_i) calls with the arrays
 * sigmas / centres ([nstreams][n]) when per_sample != 0.  sigma2 is the create-time sigma squared in FLOAT. */
int ref_mw_streams(int prng_type, const uint8_t *seeds, size_t seed_len, size_t nstreams, size_t n, float tail, float sigma,
                   const float *sigmas, const float *centres, int per_sample, int32_t *out, int threads)
{
    int fail = 0, nt = max_threads(threads);
#pragma omp parallel for schedule(static) num_threads(nt) reduction(|:fail)
    for (size_t s = 0; s < nstreams; s++) {
        prng_ctx_t *ctx = make_prng(prng_type, seeds + s * seed_len, seed_len, 0);
        if (!ctx) { fail |= 1; continue; }
        utils_sampling_t *smp = create_sampler(CDF_GAUSSIAN_SAMPLING, SAMPLING_64BIT, NORMAL_SAMPLES, (SINT32)n,
                                               SAMPLING_MW_BOOTSTRAP, ctx, tail, sigma);
        if (!smp) { fail |= 1; prng_destroy(ctx); continue; }
        if (per_sample) {
            for (size_t i = 0; i < n; i++) out[s * n + i] = get_bootstrap_sample(smp, sigmas[s * n + i], centres[s * n + i]);
        } else {
            get_vector_32(smp, out + s * n, n, centres[0]);
        }
        destroy_sampler(&smp);
        prng_destroy(ctx);
    }
    return fail;
}

int ref_num_threads(void) { return omp_get_max_threads(); }
