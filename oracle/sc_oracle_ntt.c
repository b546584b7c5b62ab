/*
 * oracle/sc_oracle_ntt.c -- TEST INFRASTRUCTURE ONLY (see sc_oracle.h).
 *
 * CPU restatement of the 32-bit NTT surface of libsafecrypto.  Citations are to
 * /root/reference/src/utils/arith/ unless stated.  One translation unit covers all six
 * live reduction variants by switching on `variant` where the reference generates one
 * source file per variant from ntt_template.c.in (gen_ntt.sh).
 *
 * Signed-overflow note: the reference relies on two's-complement wrap-around in several
 * places (lazy butterflies, Barrett on large inputs).  Here every such operation is done on
 * unsigned types so the behaviour is defined and equal to what gcc -O2 emits for the
 * reference on x86-64.
 */
#include "sc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* ------------------------------------------------------------------------------------ */
/* wrap-around helpers                                                                    */

static inline int32_t add32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t sub32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t mul32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
static inline int64_t mul64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
static inline int64_t sub64(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }

/* sc_math.c:31-34, branch-free unsigned a < b */
static inline uint32_t ct_lt_u32(uint32_t a, uint32_t b)
{
    return ((((a ^ b) & ((a - b) ^ b)) ^ (a - b)) & 0x80000000u) >> 31;
}

void orc_init_reduce(orc_params_t *p, int n, int q)
{
    /* ntt.c:132-146 */
    p->n = n;
    p->q = q;
    p->k = 30;
    p->m = (1 << 30) / q;
    p->inv_q_dbl = 1.0 / (double)q;
    p->inv_q_flt = (float)(1.0 / (float)q);
}

/* ------------------------------------------------------------------------------------ */
/* scalar reductions                                                                      */

/* ntt.c:366-378 -- returns the low 32 bits of the 64-bit working value */
static inline int32_t barrett_reduce(int64_t a, const orc_params_t *p)
{
    int64_t t = mul64(a, p->m) >> p->k;
    int64_t c = sub64(a, mul64(t, p->q));
    c += (int64_t)((uint32_t)p->q * ((uint32_t)c >> 31));
    c -= (int64_t)(int32_t)((uint32_t)p->q * ct_lt_u32((uint32_t)p->q, (uint32_t)(c + 1)));
    return (int32_t)c;
}

/* ntt_template.c.in:707-730 (modn) / 795-820 (muln): three folds of 2^a = 2^b - 1 */
static inline int64_t solinas_fold(int64_t x, int hi_shift, int64_t lo_mask, int up_shift)
{
    for (int round = 0; round < 3; round++) {
        int64_t high = x >> hi_shift;
        int64_t low = x & lo_mask;
        x = low - high + (int64_t)((uint64_t)high << up_shift);
    }
    return x;
}

static inline int32_t solinas_fold32(int32_t x, int hi_shift, int32_t lo_mask, int up_shift)
{
    for (int round = 0; round < 3; round++) {
        int32_t high = x >> hi_shift;
        int32_t low = x & lo_mask;
        x = add32(sub32(low, high), (int32_t)((uint32_t)high << up_shift));
    }
    return x;
}

/* ntt_template.c.in:758-768 / 856-858: truncating double quotient (fp and avx variants) */
static inline int32_t fp_reduce(int64_t v, const orc_params_t *p)
{
    double quo = (double)v * p->inv_q_dbl;
    return (int32_t)sub64(v, mul64(p->q, (int64_t)quo));
}

int32_t orc_modn(int variant, int32_t x, const orc_params_t *p)
{
    switch (variant) {
    case ORC_NTT_REFERENCE:       return x % p->q;                               /* :701 */
    case ORC_NTT_BARRETT:         return barrett_reduce((int64_t)x, p);          /* :703-706 */
    case ORC_NTT_SOLINAS_7681:    return solinas_fold32(x, 13, 0x1FFF, 9);       /* :708-718 */
    case ORC_NTT_SOLINAS_8380417: return solinas_fold32(x, 23, 0x7FFFFF, 13);    /* :720-730 */
    default:                      return fp_reduce((int64_t)x, p);               /* :766-767 */
    }
}

static inline int32_t reduce_product(int variant, int64_t prod, const orc_params_t *p)
{
    switch (variant) {
    case ORC_NTT_REFERENCE:       return (int32_t)(prod % p->q);                          /* :788 */
    case ORC_NTT_BARRETT:         return barrett_reduce(prod, p);                         /* :790-794 */
    case ORC_NTT_SOLINAS_7681:    return (int32_t)solinas_fold(prod, 13, 0x1FFF, 9);      /* :796-807 */
    case ORC_NTT_SOLINAS_8380417: return (int32_t)solinas_fold(prod, 23, 0x7FFFFF, 13);   /* :809-820 */
    default:                      return fp_reduce(prod, p);                              /* :856-858 */
    }
}

int32_t orc_muln(int variant, int32_t x, int32_t y, const orc_params_t *p)
{
    return reduce_product(variant, (int64_t)x * (int64_t)y, p);
}

int32_t orc_sqrn(int variant, int32_t x, const orc_params_t *p)
{
    return reduce_product(variant, (int64_t)x * (int64_t)x, p);   /* :877-953 */
}

/* ntt_template.c.in:1689-1721 -- right-to-left square and multiply with constant-time select */
int32_t orc_pwr(int variant, int32_t x, int32_t e, const orc_params_t *p)
{
    int32_t y = (e & 1) ? x : 1;
    e >>= 1;
    while (e > 0) {
        x = orc_sqrn(variant, x, p);
        int32_t cand = orc_muln(variant, x, y, p);
        if (e & 1) y = cand;
        e >>= 1;
    }
    return y;
}

/* ------------------------------------------------------------------------------------ */
/* AVX2-variant lane arithmetic, restated per lane                                       */

#define MAGIC_DBL   6755399441055744.0          /* (double)0x0018000000000000 = 2^52 + 2^51 */
#define MAGIC_BITS  0x4338000000000000ULL       /* its IEEE-754 encoding                     */

/* ntt_template.c.in:37-41 */
static inline double lane_i64_to_dbl(int64_t x)
{
    uint64_t bits = (uint64_t)x + MAGIC_BITS;
    double d;
    memcpy(&d, &bits, 8);
    return d - MAGIC_DBL;
}

/* ntt_template.c.in:43-50 applied to the product td * (1/q).  The reference writes
 * _mm256_mul_pd followed (inside double_to_int64) by _mm256_add_pd of the magic constant;
 * gcc's default -ffp-contract=fast fuses the pair into one vfmadd on every FMA-capable
 * target (-march=native on any AVX2 host, x86-64-v3 here), i.e. a SINGLE rounding of
 * td * inv_q + magic.  That fused behaviour is what the compiled reference does and what is
 * restated (verified against oracle/_ref on products up to 2^62). */
static inline int64_t lane_quotient(double td, double inv_q)
{
    double y = fma(td, inv_q, MAGIC_DBL);
    uint64_t bits;
    memcpy(&bits, &y, 8);
    return (int64_t)(bits - MAGIC_BITS);
}

/* ntt_template.c.in:52-65 */
static inline double lane_i64_to_dbl_full(int64_t v)
{
    int64_t lo = v & 0xFFFFFFFFLL;
    int64_t hi = v >> 32;
    double hi_d = 4294967296.0 * lane_i64_to_dbl(hi);
    return hi_d + lane_i64_to_dbl(lo);
}

/* double-precision lane reduction: e.g. :984-999, :1110-1123, :1174-1185.
 * `_mm256_mul_epi32(res, b_q)` multiplies only the low signed 32 bits of the quotient. */
static inline int64_t lane_reduce_dbl(int64_t prod, int full_range, const orc_params_t *p)
{
    double td = full_range ? lane_i64_to_dbl_full(prod) : lane_i64_to_dbl(prod);
    int64_t quo = lane_quotient(td, p->inv_q_dbl);
    int64_t res = sub64(prod, mul64((int64_t)(int32_t)(uint32_t)quo, p->q));
    if (res < 0) res = (int64_t)((uint64_t)res + (uint64_t)(int64_t)p->q);
    return res;
}

/* single-precision lane reduction on the low 32 bits of the product: :1081-1094, :1372-1385 */
static inline int32_t lane_reduce_flt(int32_t p32, const orc_params_t *p)
{
    float qs_inv = (float)p->inv_q_dbl;
    float ts = (float)p32;
    float quo_f = ts * qs_inv;
    int32_t quo = (int32_t)lrintf(quo_f);           /* cvtps_epi32: round to nearest even */
    int32_t res = sub32(p32, mul32(quo, p->q));
    if (res < 0) res = add32(res, p->q);
    return res;
}

/* ------------------------------------------------------------------------------------ */
/* vector primitives                                                                      */

static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }

/* ntt.c:470-489 (LUT of (i, rev(i)) swap pairs) == full bit-reversal permutation */
static void bit_reverse_inplace(int32_t *v, int n)
{
    int bits = ilog2(n);
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < bits; b++) r |= ((i >> b) & 1) << (bits - 1 - b);
        if (i < r) { int32_t t = v[i]; v[i] = v[r]; v[r] = t; }
    }
}

/* :956-1015 */
static void pointwise32(int variant, int32_t *v, const orc_params_t *p, const int32_t *t, const int32_t *u)
{
    for (int i = 0; i < p->n; i++) {
        int64_t prod = (int64_t)t[i] * (int64_t)u[i];
        if (variant == ORC_NTT_AVX) v[i] = (int32_t)lane_reduce_dbl(prod, 1, p);
        else                        v[i] = reduce_product(variant, prod, p);
    }
}

/* :1018-1141 */
static void pointwise16(int variant, int32_t *v, const orc_params_t *p, const int32_t *t, const int16_t *u)
{
    for (int i = 0; i < p->n; i++) {
        int64_t prod = (int64_t)t[i] * (int64_t)u[i];
        if (variant == ORC_NTT_AVX) {
            if (p->q == 7681) v[i] = lane_reduce_flt((int32_t)prod, p);       /* :1074-1102 */
            else              v[i] = (int32_t)lane_reduce_dbl(prod, 0, p);    /* :1103-1131 */
        } else {
            v[i] = reduce_product(variant, prod, p);
        }
    }
}

static inline int32_t tw_at(const void *w, int tw_bits, int idx)
{
    return tw_bits == 16 ? (int32_t)((const int16_t *)w)[idx] : ((const int32_t *)w)[idx];
}

/*
 * Radix-2 decimation-in-time pass over bit-reversed input; stage `half` = 1,2,..,n/2 uses
 * twiddle w[j * (n/half)... ] exactly as the MK1 loops do (:1144-1244, :1246-1339,
 * :1341-1482, :1484-1539).  `large` = reduce after every add/sub.
 */
static void dit_fft(int variant, int32_t *v, const orc_params_t *p, const void *w, int tw_bits, int large)
{
    const int n = p->n;
    for (int half = 1, step = n; half < n; half <<= 1, step >>= 1) {
        const int span = half << 1;
        /* The AVX2 build vectorises the early stages of fft_32, large_fft_32 and fft_16
         * (never large_fft_16) and treats the j = 0 column like any other column there. */
        const int vec_stage = (variant == ORC_NTT_AVX) && (half < (n >> 3)) && !(large && tw_bits == 16);
        for (int j = 0; j < half; j++) {
            const int32_t y = tw_at(w, tw_bits, j * step);
            for (int k = j; k < n; k += span) {
                int32_t x;
                if (vec_stage) {
                    int64_t prod = (int64_t)v[k + half] * (int64_t)y;
                    if (tw_bits == 16 && p->q <= 12289) x = lane_reduce_flt((int32_t)prod, p);   /* :1367-1402 */
                    else if (tw_bits == 16)             x = (int32_t)lane_reduce_dbl(prod, 0, p); /* :1403-1436 */
                    else                                x = (int32_t)lane_reduce_dbl(prod, large, p); /* :1170-1185 / :1271-1286 */
                    int32_t lo = v[k];
                    v[k + half] = sub32(lo, x);
                    v[k] = add32(lo, x);
                    continue;                       /* no modn in the vector path even when large */
                }
                if (j == 0 && !(large && tw_bits == 32)) {
                    x = v[k + half];
                    if (tw_bits == 32 || large) x = orc_modn(variant, x, p);   /* :1208-1213, :1497-1498 */
                    /* fft_16: raw value, :1443-1447 */
                } else {
                    x = orc_muln(variant, v[k + half], y, p);
                }
                int32_t lo = v[k];
                int32_t d = sub32(lo, x), s = add32(lo, x);
                if (large) { d = orc_modn(variant, d, p); s = orc_modn(variant, s, p); }
                v[k + half] = d;
                v[k] = s;
            }
        }
    }
}

/* ntt.c:571-604 */
static void flip32(int32_t *v, const orc_params_t *p)
{
    const int n = p->n;
    for (int i = 1, j = n - 1; i <= ((n - 1) >> 1); i++, j--) {
        int32_t x = v[i]; v[i] = v[j]; v[j] = x;
    }
    v[0] = (int32_t)(0u - (uint32_t)v[0]);
    for (int i = 0; i < n; i++) {
        int32_t x = v[i];
        x = add32(x, (int32_t)((uint32_t)p->q * ((uint32_t)x >> 31)));
        x = sub32(x, (int32_t)((uint32_t)p->q * ct_lt_u32((uint32_t)p->q, (uint32_t)x + 1u)));
        v[i] = x;
    }
}

/* :1840-1901 */
static void normalize32(int variant, int32_t *v, size_t len, const orc_params_t *p)
{
    for (size_t i = 0; i < len; i++) {
        if (variant == ORC_NTT_AVX) { v[i] = lane_reduce_flt(v[i], p); continue; }   /* :1850-1869 */
        int32_t x = orc_modn(variant, v[i], p);
        x = add32(x, (int32_t)((uint32_t)p->q * ((uint32_t)x >> 31)));
        x = sub32(x, (int32_t)((uint32_t)p->q * ct_lt_u32((uint32_t)p->q, (uint32_t)x + 1u)));
        v[i] = x;
    }
}

/* :1777-1838 */
static void center32(int variant, int32_t *v, size_t len, const orc_params_t *p)
{
    const int32_t q = p->q;
    for (size_t i = 0; i < len; i++) {
        if (variant == ORC_NTT_AVX) {
            /* :1788-1810: residual of a round-to-nearest float quotient, then two masked
             * corrections against q>>1 and -(q>>1) */
            float qs_inv = (float)p->inv_q_dbl;
            int32_t quo = (int32_t)lrintf((float)v[i] * qs_inv);
            int32_t s = sub32(v[i], mul32(quo, q));
            if ((q >> 1) > s) s = sub32(s, q);
            if (-(q >> 1) > s) s = add32(s, q);
            v[i] = s;
            continue;
        }
        int32_t x = orc_modn(variant, v[i], p);
        const int32_t q2 = (q - 1) >> 1;
        while (x < -q2) x += q;
        while (x > q2) x -= q;
        v[i] = x;
    }
}

static void fwd_ntt(int variant, int32_t *v, const orc_params_t *p, const int32_t *t,
                    const void *w, int tw_bits, int large)
{
    /* :1541-1565, :1615-1639 */
    if (tw_bits == 16) pointwise16(variant, v, p, t, (const int16_t *)w);
    else               pointwise32(variant, v, p, t, (const int32_t *)w);
    bit_reverse_inplace(v, p->n);
    dit_fft(variant, v, p, w, tw_bits, large);
}

static void inv_ntt(int variant, int32_t *v, const orc_params_t *p, const int32_t *t,
                    const void *w, const void *r, int tw_bits, int large)
{
    /* :1567-1613, :1641-1687 */
    if (v != t) memmove(v, t, sizeof(int32_t) * (size_t)p->n);
    bit_reverse_inplace(v, p->n);
    dit_fft(variant, v, p, w, tw_bits, large);
    if (tw_bits == 16) pointwise16(variant, v, p, v, (const int16_t *)r);
    else               pointwise32(variant, v, p, v, (const int32_t *)r);
    flip32(v, p);
}

/* :1723-1748 / :1750-1775 */
static int32_t invert32(int variant, int32_t *v, const orc_params_t *p, size_t len, const int32_t *num_in, int32_t *num)
{
    for (size_t i = 0; i < len; i++) {
        int32_t x = orc_modn(variant, v[i], p);
        if (x == 0) return 1;                                   /* SC_FUNC_FAILURE */
        x = orc_pwr(variant, x, p->q - 2, p);
        if (num) num[i] = orc_muln(variant, num_in[i], x, p);
        else     v[i] = x;
    }
    return 0;
}

/* ntt.c:381-422 */
static void sparse_mul(int32_t *v, int n, int omega, const void *t, int t_bits, const int32_t *u)
{
    memset(v, 0, sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < omega; i++) {
        int pos = u[i];
        for (int j = 0; j < pos; j++)
            v[j] = add32(v[j], t_bits == 16 ? ((const int16_t *)t)[j + n - pos] : ((const int32_t *)t)[j + n - pos]);
        for (int j = pos; j < n; j++)
            v[j] = sub32(v[j], t_bits == 16 ? ((const int16_t *)t)[j - pos] : ((const int32_t *)t)[j - pos]);
    }
}

/* ntt.c:425-452: the HAVE_AVX2 build of ntt32_mult_scalar_generic (what an AVX2 host runs):
 * 64-bit product, +q when negative, truncated to 32 bits -- no modular reduction. */
static void scalar_mul(int32_t *v, const orc_params_t *p, const int32_t *t, int32_t c)
{
    for (int i = 0; i < p->n; i++) {
        int64_t res = (int64_t)t[i] * (int64_t)c;
        if (res < 0) res += p->q;
        v[i] = (int32_t)res;
    }
}

/* ------------------------------------------------------------------------------------ */
/* tables                                                                                 */

static int64_t powmod(int64_t b, int64_t e, int64_t q)
{
    __int128 r = 1, x = b % q;
    while (e > 0) { if (e & 1) r = (r * x) % q; x = (x * x) % q; e >>= 1; }
    return (int64_t)r;
}

/* roots_of_unity.c:67-88: smallest m >= 2 with m^n == -1 (mod q) */
static int64_t find_2nth_root(int64_t q, int n)
{
    for (int64_t m = 2; m < q - 1; m++)
        if (powmod(m, n, q) == q - 1) return m;
    return 0;
}

int64_t orc_find_primitive_root(int64_t q)
{
    /* roots_of_unity.c:40-64: smallest m whose powers m^1..m^(phi-1) hit 1 exactly ... the
     * reference's euler_phi() returns phi(q) only for odd q > 1; for prime q that is q-1 */
    for (int64_t m = 1; m < q - 1; m++) {
        int64_t hits = (m == 1), pw = m;
        for (int64_t l = 1; l < q - 1; l++) { pw = (int64_t)(((__int128)pw * m) % q); hits += (pw == 1); }
        if (hits == 1) return m;
    }
    return 0;
}

int orc_roots_of_unity(int64_t q, int n, int32_t *w, int32_t *r, int64_t *g_out)
{
    /* roots_of_unity.c:141-172: fwd[i] = g^i ; inv[0] = |x| with n*x + q*y = 1, which for
     * every table the reference ships is -(n^-1) mod q ; inv[i] = inv[i-1] * g */
    int64_t g = find_2nth_root(q, n);
    if (!g) return 1;
    int64_t ninv = powmod(n, q - 2, q);
    int64_t acc = 1, racc = (q - ninv) % q;
    for (int i = 0; i < n; i++) {
        w[i] = (int32_t)acc;
        r[i] = (int32_t)racc;
        acc = (int64_t)(((__int128)acc * g) % q);
        racc = (int64_t)(((__int128)racc * g) % q);
    }
    if (g_out) *g_out = g;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* batch driver (mirrors ref_ntt_batch in ref_driver.c)                                   */

int orc_num_threads(void) { return omp_get_max_threads(); }

int orc_ntt_batch(int variant, int op, int n, int q, int tw_bits,
                  int32_t *out, const int32_t *a, const void *b, size_t b_stride,
                  const void *w, const void *r, size_t count, int threads, int32_t *rc,
                  int32_t scalar)
{
    orc_params_t p;
    orc_init_reduce(&p, n, q);
    int nt = threads > 0 ? threads : omp_get_max_threads();
    int any = 0;
#pragma omp parallel for schedule(static) num_threads(nt) reduction(|:any)
    for (size_t i = 0; i < count; i++) {
        int32_t *v = out + i * (size_t)n;
        const int32_t *t = a ? a + i * (size_t)n : NULL;
        const int32_t *u32 = b ? (const int32_t *)b + i * b_stride : NULL;
        const int16_t *u16 = b ? (const int16_t *)b + i * b_stride : NULL;
        int32_t ret = 0;
        switch (op) {
        case ORC_OP_FWD:        fwd_ntt(variant, v, &p, t, w, tw_bits, 0); break;
        case ORC_OP_FWD_LARGE:  fwd_ntt(variant, v, &p, t, w, tw_bits, 1); break;
        case ORC_OP_INV:        inv_ntt(variant, v, &p, t, w, r, tw_bits, 0); break;
        case ORC_OP_INV_LARGE:  inv_ntt(variant, v, &p, t, w, r, tw_bits, 1); break;
        case ORC_OP_FFT:        memcpy(v, t, 4u * (size_t)n); dit_fft(variant, v, &p, w, tw_bits, 0); break;
        case ORC_OP_FFT_LARGE:  memcpy(v, t, 4u * (size_t)n); dit_fft(variant, v, &p, w, tw_bits, 1); break;
        case ORC_OP_PW:         pointwise32(variant, v, &p, t, u32); break;
        case ORC_OP_PW16:       pointwise16(variant, v, &p, t, u16); break;
        case ORC_OP_NORMALIZE:  memcpy(v, t, 4u * (size_t)n); normalize32(variant, v, (size_t)n, &p); break;
        case ORC_OP_CENTER:     memcpy(v, t, 4u * (size_t)n); center32(variant, v, (size_t)n, &p); break;
        case ORC_OP_FLIP:       memcpy(v, t, 4u * (size_t)n); flip32(v, &p); break;
        case ORC_OP_POLYMUL: {
            int32_t tmp[1024];
            fwd_ntt(variant, v, &p, t, w, tw_bits, 0);
            fwd_ntt(variant, tmp, &p, u32, w, tw_bits, 0);
            pointwise32(variant, v, &p, v, tmp);
            inv_ntt(variant, v, &p, v, w, r, tw_bits, 0);
        } break;
        case ORC_OP_TRIPLE16:
            fwd_ntt(variant, v, &p, t, w, 16, 0);
            pointwise16(variant, v, &p, v, u16);
            inv_ntt(variant, v, &p, v, w, r, 16, 0);
            break;
        case ORC_OP_MODN: for (int j = 0; j < n; j++) v[j] = orc_modn(variant, t[j], &p); break;
        case ORC_OP_MULN: for (int j = 0; j < n; j++) v[j] = orc_muln(variant, t[j], u32[j], &p); break;
        case ORC_OP_SQRN: for (int j = 0; j < n; j++) v[j] = orc_sqrn(variant, t[j], &p); break;
        case ORC_OP_PWR:  for (int j = 0; j < n; j++) v[j] = orc_pwr(variant, t[j], u32[j], &p); break;
        case ORC_OP_INVERT:
            memcpy(v, t, 4u * (size_t)n);
            ret = invert32(variant, v, &p, (size_t)n, NULL, NULL);
            break;
        case ORC_OP_DIV: {
            /* div_32(num, den): num[i] = muln(num[i], den[i]^(q-2)); a = num, b = den */
            int32_t den[1024];
            memcpy(v, t, 4u * (size_t)n);
            memcpy(den, u32, 4u * (size_t)n);
            ret = invert32(variant, den, &p, (size_t)n, v, v);
        } break;
        case ORC_OP_SCALAR:   scalar_mul(v, &p, t, scalar); break;
        case ORC_OP_SPARSE32: sparse_mul(v, n, scalar & 0xFFFF, t, 32, u32); break;
        case ORC_OP_SPARSE16: sparse_mul(v, n, scalar & 0xFFFF, (const int16_t *)(const void *)a + i * (size_t)n, 16, u32); break;
        default: ret = -1;
        }
        if (rc) rc[i] = ret;
        any |= ret;
    }
    return any;
}
