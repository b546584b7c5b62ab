/*
 * oracle/sc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the libsafecrypto (0.2.0-79) lattice hot path: the `*_32` members
 * of utils_arith_ntt_t (src/utils/arith/ntt.h:237-262), the Gaussian samplers behind
 * create_sampler()/get_vector_32() (src/utils/sampling/sampling.h:88-110) and the part of
 * the PRNG (src/utils/crypto/prng.c, chacha20_csprng.c, ctr_drbg.c) whose word stream they
 * consume.  It is the checker for the CUDA path; nothing under libsafecrypto_b200/ may link,
 * import or execute it.  Parity status: PINNED -- tests/test_oracle_vs_ref.py compares every
 * function here against oracle/_ref/libscref.so (the reference's own sources compiled by
 * oracle/Makefile) and tests/golden/ holds vectors generated from that library.
 */
#ifndef SC_ORACLE_H
#define SC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reduction variants: values of safecrypto_ntt_e, src/utils/arith/ntt.h:106-123 */
enum {
    ORC_NTT_REFERENCE = 0,
    ORC_NTT_BARRETT = 1,
    ORC_NTT_FLOATING_POINT = 2,
    ORC_NTT_AVX = 3,
    ORC_NTT_SOLINAS_7681 = 4,
    ORC_NTT_SOLINAS_8380417 = 5
};

/* batch op codes (same numbering as oracle/ref_driver.c and include/scgpu.h) */
enum {
    ORC_OP_FWD = 0, ORC_OP_INV, ORC_OP_FWD_LARGE, ORC_OP_INV_LARGE, ORC_OP_FFT, ORC_OP_FFT_LARGE,
    ORC_OP_PW, ORC_OP_PW16, ORC_OP_NORMALIZE, ORC_OP_CENTER, ORC_OP_POLYMUL, ORC_OP_TRIPLE16,
    ORC_OP_MODN, ORC_OP_MULN, ORC_OP_SQRN, ORC_OP_FLIP, ORC_OP_INVERT, ORC_OP_DIV, ORC_OP_PWR,
    ORC_OP_SCALAR, ORC_OP_SPARSE32, ORC_OP_SPARSE16
};

/* the fields of ntt_params_t (ntt.h:91-103) the 32-bit path reads */
typedef struct {
    int32_t n;
    int32_t q;
    int32_t m;          /* Barrett multiplier floor(2^k / q)      ntt.c:142-146 */
    int32_t k;          /* Barrett shift, always 30                              */
    double  inv_q_dbl;  /* 1.0 / (double) q                        ntt.c:138     */
    float   inv_q_flt;  /* 1.0 / (float) q                         ntt.c:139     */
} orc_params_t;

void orc_init_reduce(orc_params_t *p, int n, int q);

/* scalar primitives */
int32_t orc_modn(int variant, int32_t x, const orc_params_t *p);
int32_t orc_muln(int variant, int32_t x, int32_t y, const orc_params_t *p);
int32_t orc_sqrn(int variant, int32_t x, const orc_params_t *p);
int32_t orc_pwr(int variant, int32_t x, int32_t e, const orc_params_t *p);

/* twiddle tables: w[i] = g^i, r[i] = -(n^-1) g^i  (roots_of_unity.c:141-207) */
int orc_roots_of_unity(int64_t q, int n, int32_t *w, int32_t *r, int64_t *g_out);
int64_t orc_find_primitive_root(int64_t q);

/* Same calling convention as ref_ntt_batch() in oracle/ref_driver.c. */
int orc_ntt_batch(int variant, int op, int n, int q, int tw_bits,
                  int32_t *out, const int32_t *a, const void *b, size_t b_stride,
                  const void *w, const void *r, size_t count, int threads, int32_t *rc,
                  int32_t scalar);

/* PRNG (prng_types.h / safecrypto_types.h:237-254 numbering) */
enum { ORC_PRNG_AES_CTR_DRBG = 0, ORC_PRNG_CHACHA = 2 };

typedef struct orc_prng orc_prng_t;
orc_prng_t *orc_prng_create(int prng_type, const uint8_t *seed, size_t seed_len, size_t seed_period);
void     orc_prng_destroy(orc_prng_t *ctx);
uint32_t orc_prng_32(orc_prng_t *ctx);
uint64_t orc_prng_64(orc_prng_t *ctx);
uint32_t orc_prng_var(orc_prng_t *ctx, size_t n);
uint32_t orc_prng_8(orc_prng_t *ctx);
int32_t  orc_prng_bit(orc_prng_t *ctx);
float    orc_prng_float(orc_prng_t *ctx);
double   orc_prng_double(orc_prng_t *ctx);
int32_t  orc_prng_mem(orc_prng_t *ctx, uint8_t *mem, int32_t length);
int      orc_prng_reset(orc_prng_t *ctx);                 /* AES-CTR-DRBG only, see sc_oracle_prng.c */
uint64_t orc_prng_csprng_bytes(orc_prng_t *ctx);
uint64_t orc_prng_out_bytes(orc_prng_t *ctx);
/* script = (kind, arg) pairs: 32 prng_32, 64 prng_64 (hi, lo), 8 prng_8, 1 prng_bit, 16 prng_16, 0 prng_var(arg),
 * 128 prng_128 (4 words, most significant first), 2 prng_float (bits), 3 prng_double (bits, low word first),
 * 4 prng_mem(arg bytes; ceil(arg/4) words, zero padded), 5 prng_reset (no output), 6 statistics (csprng bytes,
 * out bytes; low 32 bits each) */
int orc_prng_script(int prng_type, const uint8_t *seed, size_t seed_len, size_t seed_period,
                    const int32_t *script, size_t ndraws, uint32_t *out);
void orc_aes256_encrypt_block(const uint8_t key[32], const uint8_t in[16], uint8_t out[16]);

/* Gaussian samplers (random_sampling_e numbering, safecrypto_private.h:174-181) */
enum { ORC_SAMPLER_CDF = 0, ORC_SAMPLER_KNUTH_YAO = 1, ORC_SAMPLER_BERNOULLI = 5 };
enum { ORC_NORMAL_SAMPLES = 0, ORC_BLINDING_SAMPLES = 1, ORC_SHUFFLE_SAMPLES = 2 };

int orc_cdf_table(int precision, int blinding, float tail, float sigma, void *out, size_t cap_entries);
/* inject a reference-built 128 / 192 / 256-bit CDF table (entries x precision/64 words) for orc_gauss_streams */
int orc_set_high_table(int precision, const uint64_t *words, int entries);
int orc_ky_table(int bitwidth, float tail, float sigma, uint8_t *pmat, size_t cap,
                 int32_t *rows, int32_t *cols, int32_t *bound);
int orc_ber_table(float tail, float sigma, uint8_t *tab, size_t cap,
                  int32_t *entries, int32_t *maxval, int32_t *maxlog);
int orc_gauss_streams(int sampler, int precision, int blinding, int prng_type, float tail, float sigma,
                      uint32_t discard, const uint8_t *seeds, size_t seed_len, size_t nstreams,
                      size_t n, int32_t centre, int32_t *out, int threads, size_t calls_per_stream);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
