/*
 * scgpu_dropin.h -- the reference's own types and symbols for the hot path, re-declared so that
 * libscgpu.so can stand in for src/utils/arith (NTT table) and src/utils/sampling (+ the part
 * of src/utils/crypto/prng.c the samplers draw from).  Layouts are binary compatible with
 * libsafecrypto 0.2.0-79 built for x86-64 (HAVE_64BIT, packed structs); tests/test_abi.py
 * checks the sizes against the compiled reference.
 *
 * When building INSIDE the reference tree, include the reference's headers instead and link
 * libscgpu.so in place of the objects listed in INTEGRATION.md; the symbols are the same.
 */
#ifndef SCGPU_DROPIN_H
#define SCGPU_DROPIN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef SCGPU_NO_SC_TYPES            /* include/safecrypto_types.h */
typedef int8_t SINT8;   typedef uint8_t UINT8;
typedef int16_t SINT16; typedef uint16_t UINT16;
typedef int32_t SINT32; typedef uint32_t UINT32;
typedef int64_t SINT64; typedef uint64_t UINT64;
typedef float FLOAT;    typedef double DOUBLE;
typedef int64_t sc_slimb_t;          /* src/utils/arith/limb.h, 64-bit limbs */
typedef uint64_t sc_ulimb_t;
#define SC_FUNC_SUCCESS 0
#define SC_FUNC_FAILURE 1
#endif

/* ---- src/utils/arith/ntt.h:49-124 ------------------------------------------------------- */
#pragma pack(push, 1)
typedef struct ntt16_params_t { SINT16 q; UINT16 q_inv; SINT16 m; SINT16 k; } ntt16_params_t;
typedef struct ntt32_params_t { SINT32 q; UINT32 q_inv; SINT32 m; SINT32 k; } ntt32_params_t;
typedef struct ntt64_params_t { SINT64 q; UINT64 q_inv; SINT64 m; SINT64 k; } ntt64_params_t;
typedef struct nttlimb_params_t { sc_slimb_t q; sc_ulimb_t q_inv; sc_slimb_t m; sc_slimb_t k; } nttlimb_params_t;
typedef struct ntt_params_t {
    DOUBLE q_dbl;
    DOUBLE inv_q_dbl;
    FLOAT  inv_q_flt;
    size_t n;
    union ntt_u {
        nttlimb_params_t nttlimb;
        ntt64_params_t   ntt64;
        ntt32_params_t   ntt32;
        ntt16_params_t   ntt16;
    } u;
} ntt_params_t;
#pragma pack(pop)

typedef enum safecrypto_ntt {
    SC_NTT_REFERENCE = 0, SC_NTT_BARRETT, SC_NTT_FLOATING_POINT, SC_NTT_AVX,
    SC_NTT_SOLINAS_7681, SC_NTT_SOLINAS_8380417, SC_NTT_SOLINAS_16813057, SC_NTT_SOLINAS_134348801,
    SC_NTT_REFERENCE_REV, SC_NTT_BARRETT_REV, SC_NTT_FLOATING_POINT_REV, SC_NTT_AVX_REV,
    SC_NTT_SOLINAS_7681_REV, SC_NTT_SOLINAS_8380417_REV, SC_NTT_SOLINAS_16813057_REV,
    SC_NTT_SOLINAS_134348801_REV
} safecrypto_ntt_e;

/* ---- utils_arith_ntt_t, ntt.h:217-297: 18 SINT16-data, 26 SINT32-data, 33 limb-data members.
 * Only the 26 `*_32` members are live in the reference's schemes and GPU backed
 * here; the others are populated with a stub that reports the call and aborts. */
typedef const ntt_params_t *scP;
#pragma pack(push, 1)
typedef struct _utils_arith_ntt {
    /* SINT16 data */
    SINT16 (*modn_16)(SINT16, scP);
    SINT16 (*muln_16)(SINT16, SINT16, scP);
    SINT16 (*sqrn_16)(SINT16, scP);
    void   (*mul_16_sparse)(SINT16 *, size_t, UINT16, const SINT16 *, const SINT16 *);
    void   (*mul_16_pointwise)(SINT16 *, scP, const SINT16 *, const SINT16 *);
    void   (*mul_16_scalar)(SINT16 *, scP, const SINT16 *, SINT16);
    void   (*fft_16)(SINT16 *, scP, const SINT16 *);
    void   (*large_fft_16)(SINT16 *, scP, const SINT16 *);
    SINT32 (*pwr_16)(SINT16, SINT16, scP);
    SINT32 (*invert_16)(SINT16 *, scP, size_t);
    SINT32 (*div_16)(SINT16 *, const SINT16 *, scP, size_t);
    void   (*flip_16)(SINT16 *, scP);
    void   (*center_16)(SINT16 *, size_t, scP);
    void   (*normalize_16)(SINT16 *, size_t, scP);
    void   (*fwd_ntt_16)(SINT16 *, scP, const SINT16 *, const SINT16 *);
    void   (*inv_ntt_16)(SINT16 *, scP, const SINT16 *, const SINT16 *, const SINT16 *);
    void   (*fwd_ntt_16_large)(SINT16 *, scP, const SINT16 *, const SINT16 *);
    void   (*inv_ntt_16_large)(SINT16 *, scP, const SINT16 *, const SINT16 *, const SINT16 *);
    /* SINT32 data -- the hot path */
    SINT32 (*modn_32)(SINT32, scP);
    SINT32 (*muln_32)(SINT32, SINT32, scP);
    SINT32 (*sqrn_32)(SINT32, scP);
    void   (*mul_32_sparse)(SINT32 *, size_t, UINT16, const SINT32 *, const SINT32 *);
    void   (*mul_32_sparse_16)(SINT32 *, size_t, UINT16, const SINT16 *, const SINT32 *);
    void   (*mul_32_pointwise)(SINT32 *, scP, const SINT32 *, const SINT32 *);
    void   (*mul_32_pointwise_16)(SINT32 *, scP, const SINT32 *, const SINT16 *);
    void   (*mul_32_scalar)(SINT32 *, scP, const SINT32 *, SINT32);
    void   (*fft_32_32)(SINT32 *, scP, const SINT32 *);
    void   (*fft_32_32_large)(SINT32 *, scP, const SINT32 *);
    void   (*fft_32_16)(SINT32 *, scP, const SINT16 *);
    void   (*fft_32_16_large)(SINT32 *, scP, const SINT16 *);
    SINT32 (*pwr_32)(SINT32, SINT32, scP);
    SINT32 (*invert_32)(SINT32 *, scP, size_t);
    SINT32 (*div_32)(SINT32 *, const SINT32 *, scP, size_t);
    void   (*flip_32)(SINT32 *, scP);
    void   (*center_32)(SINT32 *, size_t, scP);
    void   (*normalize_32)(SINT32 *, size_t, scP);
    void   (*fwd_ntt_32_32)(SINT32 *, scP, const SINT32 *, const SINT32 *);
    void   (*inv_ntt_32_32)(SINT32 *, scP, const SINT32 *, const SINT32 *, const SINT32 *);
    void   (*fwd_ntt_32_32_large)(SINT32 *, scP, const SINT32 *, const SINT32 *);
    void   (*inv_ntt_32_32_large)(SINT32 *, scP, const SINT32 *, const SINT32 *, const SINT32 *);
    void   (*fwd_ntt_32_16)(SINT32 *, scP, const SINT32 *, const SINT16 *);
    void   (*inv_ntt_32_16)(SINT32 *, scP, const SINT32 *, const SINT16 *, const SINT16 *);
    void   (*fwd_ntt_32_16_large)(SINT32 *, scP, const SINT32 *, const SINT16 *);
    void   (*inv_ntt_32_16_large)(SINT32 *, scP, const SINT32 *, const SINT16 *, const SINT16 *);
    /* limb data */
    sc_slimb_t (*modn_limb)(sc_slimb_t, scP);
    sc_slimb_t (*muln_limb)(sc_slimb_t, sc_slimb_t, scP);
    sc_slimb_t (*sqrn_limb)(sc_slimb_t, scP);
    void   (*mul_limb_sparse)(sc_slimb_t *, size_t, UINT16, const SINT32 *, const sc_slimb_t *);
    void   (*mul_limb_sparse_16)(sc_slimb_t *, size_t, UINT16, const SINT16 *, const sc_slimb_t *);
    void   (*mul_limb_pointwise)(sc_slimb_t *, scP, const sc_slimb_t *, const sc_slimb_t *);
    void   (*mul_limb_pointwise_32)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT32 *);
    void   (*mul_limb_pointwise_16)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT16 *);
    void   (*mul_limb_scalar)(sc_slimb_t *, scP, const sc_slimb_t *, sc_slimb_t);
    void   (*fft_limb)(sc_slimb_t *, scP, const sc_slimb_t *);
    void   (*fft_limb_large)(sc_slimb_t *, scP, const sc_slimb_t *);
    void   (*fft_limb_32)(sc_slimb_t *, scP, const SINT32 *);
    void   (*fft_limb_32_large)(sc_slimb_t *, scP, const SINT32 *);
    void   (*fft_limb_16)(sc_slimb_t *, scP, const SINT16 *);
    void   (*fft_limb_16_large)(sc_slimb_t *, scP, const SINT16 *);
    sc_slimb_t (*pwr_limb)(sc_slimb_t, sc_slimb_t, scP);
    SINT32 (*invert_limb)(sc_slimb_t *, scP, size_t);
    SINT32 (*div_limb)(sc_slimb_t *, const sc_slimb_t *, scP, size_t);
    void   (*flip_limb)(sc_slimb_t *, scP);
    void   (*center_limb)(sc_slimb_t *, size_t, scP);
    void   (*normalize_limb)(sc_slimb_t *, size_t, scP);
    void   (*fwd_ntt_limb)(sc_slimb_t *, scP, const sc_slimb_t *, const sc_slimb_t *);
    void   (*inv_ntt_limb)(sc_slimb_t *, scP, const sc_slimb_t *, const sc_slimb_t *, const sc_slimb_t *);
    void   (*fwd_ntt_limb_large)(sc_slimb_t *, scP, const sc_slimb_t *, const sc_slimb_t *);
    void   (*inv_ntt_limb_large)(sc_slimb_t *, scP, const sc_slimb_t *, const sc_slimb_t *, const sc_slimb_t *);
    void   (*fwd_ntt_limb_32)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT32 *);
    void   (*inv_ntt_limb_32)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT32 *, const SINT32 *);
    void   (*fwd_ntt_limb_32_large)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT32 *);
    void   (*inv_ntt_limb_32_large)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT32 *, const SINT32 *);
    void   (*fwd_ntt_limb_16)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT16 *);
    void   (*inv_ntt_limb_16)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT16 *, const SINT16 *);
    void   (*fwd_ntt_limb_16_large)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT16 *);
    void   (*inv_ntt_limb_16_large)(sc_slimb_t *, scP, const sc_slimb_t *, const SINT16 *, const SINT16 *);
} utils_arith_ntt_t;
#pragma pack(pop)

/* src/utils/arith/arith.h:134, ntt.h:300,332-333 */
extern const utils_arith_ntt_t *ntt_table;
const utils_arith_ntt_t *utils_arith_ntt(safecrypto_ntt_e type);
void init_reduce(ntt_params_t *p, size_t n, SINT32 q);
void barrett_init(ntt_params_t *p);
/* roots_of_unity.h: run-time twiddle generation (USE_RUNTIME_NTT_TABLES, bliss_b.c:365-374) */
SINT32 roots_of_unity_s32(SINT32 *fwd, SINT32 *inv, size_t n, sc_ulimb_t p, sc_ulimb_t prim, SINT32 ternary);
SINT32 roots_of_unity_s16(SINT16 *fwd, SINT16 *inv, size_t n, sc_ulimb_t p, sc_ulimb_t prim, SINT32 ternary);

/* ---- PRNG front end: src/utils/crypto/prng.h:38-100, prng_types.h:61-74 ------------------- */
typedef enum safecrypto_entropy {
    SC_ENTROPY_RANDOM = 0, SC_ENTROPY_DEV_RANDOM, SC_ENTROPY_DEV_URANDOM, SC_ENTROPY_DEV_HWRNG,
    SC_ENTROPY_CALLBACK, SC_ENTROPY_USER_PROVIDED
} safecrypto_entropy_e;
typedef enum safecrypto_prng_threading { SC_PRNG_THREADING_NONE = 0 } safecrypto_prng_threading_e;
typedef enum safecrypto_prng {           /* include/safecrypto_types.h:237-254 */
    SC_PRNG_AES_CTR_DRBG = 0, SC_PRNG_AES_CTR, SC_PRNG_CHACHA, SC_PRNG_SALSA, SC_PRNG_ISAAC, SC_PRNG_KISS,
    SC_PRNG_HASH_DRBG_SHA2_256, SC_PRNG_HASH_DRBG_SHA2_512, SC_PRNG_HASH_DRBG_SHA3_256,
    SC_PRNG_HASH_DRBG_SHA3_512, SC_PRNG_HASH_DRBG_BLAKE2_256, SC_PRNG_HASH_DRBG_BLAKE2_512,
    SC_PRNG_HASH_DRBG_WHIRLPOOL_512, SC_PRNG_FILE, SC_PRNG_HIGH_ENTROPY, SC_PRNG_MAX
} safecrypto_prng_e;
typedef void (*prng_entropy_callback)(size_t, UINT8 *);
typedef struct prng_ctx_t prng_ctx_t;    /* opaque here: libscgpu's own context, see INTEGRATION.md */

prng_ctx_t *prng_create(safecrypto_entropy_e entropy, safecrypto_prng_e type,
                        safecrypto_prng_threading_e mt, size_t seed_period);
SINT32 prng_set_entropy(prng_ctx_t *ctx, const UINT8 *entropy, size_t len);
SINT32 prng_set_entropy_callback(prng_entropy_callback cb);
SINT32 prng_init(prng_ctx_t *ctx, const UINT8 *nonce, size_t len_nonce);
safecrypto_prng_e prng_get_type(prng_ctx_t *ctx);
SINT32 prng_destroy(prng_ctx_t *ctx);
void prng_reset(prng_ctx_t *ctx);                              /* prng.c:861-932 */
UINT64 prng_get_csprng_bytes(prng_ctx_t *ctx);
UINT64 prng_get_out_bytes(prng_ctx_t *ctx);
SINT32 prng_bit(prng_ctx_t *ctx);
UINT64 prng_64(prng_ctx_t *ctx);
UINT32 prng_32(prng_ctx_t *ctx);
UINT16 prng_16(prng_ctx_t *ctx);
UINT8  prng_8(prng_ctx_t *ctx);
UINT32 prng_var(prng_ctx_t *ctx, size_t n);
#ifdef __SIZEOF_INT128__
unsigned __int128 prng_128(prng_ctx_t *ctx);                   /* prng.c:950-960 (HAVE_128BIT, x86-64) */
#endif
FLOAT  prng_float(prng_ctx_t *ctx);                            /* prng.c:1005-1008 */
DOUBLE prng_double(prng_ctx_t *ctx);                           /* prng.c:1010-1015 */
SINT32 prng_mem(prng_ctx_t *ctx, UINT8 *mem, SINT32 length);   /* prng.c:1050-1105 */

/* ---- samplers: src/utils/sampling/sampling.h:37-110, safecrypto_private.h:154-181 ---------- */
typedef enum sample_precision {
    SAMPLING_32BIT = 32, SAMPLING_64BIT = 64, SAMPLING_128BIT = 128, SAMPLING_192BIT = 192, SAMPLING_256BIT = 256
} sample_precision_e;
typedef enum sample_bootstrap { SAMPLING_DISABLE_BOOTSTRAP = 0, SAMPLING_MW_BOOTSTRAP } sample_bootstrap_e;
typedef enum sample_blinding { NORMAL_SAMPLES = 0, BLINDING_SAMPLES, SHUFFLE_SAMPLES } sample_blinding_e;
typedef enum random_sampling_e {
    CDF_GAUSSIAN_SAMPLING = 0, KNUTH_YAO_GAUSSIAN_SAMPLING, BAC_GAUSSIAN_SAMPLING, HUFFMAN_GAUSSIAN_SAMPLING,
    ZIGGURAT_GAUSSIAN_SAMPLING, BERNOULLI_GAUSSIAN_SAMPLING, KNUTH_YAO_FAST_GAUSSIAN_SAMPLING, SAMPLING_MAX
} random_sampling_e;
#define SCA_PATTERN_SAMPLE_DISCARD_LO 0x00000002
#define SCA_PATTERN_SAMPLE_DISCARD_MD 0x00000004
#define SCA_PATTERN_SAMPLE_DISCARD_HI 0x00000006

typedef struct _utils_sampling utils_sampling_t;
#pragma pack(push, 1)
struct _utils_sampling {                 /* sampling.h:68-86, same member order */
    void *(*create)(prng_ctx_t *, FLOAT, FLOAT, size_t, sample_blinding_e);
    SINT32 (*destroy)(void **);
    prng_ctx_t *(*get_prng)(void *);
    SINT32 (*sample)(void *);
    SINT32 (*vector_16)(const utils_sampling_t *, SINT16 *, size_t, SINT32);
    SINT32 (*vector_32)(const utils_sampling_t *, SINT32 *, size_t, SINT32);
    sample_precision_e precision;
    SINT32 dimension;
    sample_bootstrap_e bootstrapped;
    FLOAT tail;
    FLOAT sigma;
    FLOAT sigma2;
    void *gauss;
    prng_ctx_t *prng_ctx;
    UINT32 discard;
    void *bootstrap;
};
#pragma pack(pop)

utils_sampling_t *create_sampler(random_sampling_e type, sample_precision_e precision,
                                 sample_blinding_e blinding, SINT32 dimension,
                                 sample_bootstrap_e bootstrapped, prng_ctx_t *prng_ctx,
                                 FLOAT tail, FLOAT sigma);
SINT32 destroy_sampler(utils_sampling_t **sampler);
SINT32 set_discard(utils_sampling_t *sampler, UINT32 discard);
SINT32 get_sample(utils_sampling_t *sampler);
SINT32 get_bootstrap_sample(utils_sampling_t *sampler, FLOAT sigma, FLOAT centre);
SINT32 get_vector_16(utils_sampling_t *sampler, SINT16 *v, size_t n, FLOAT centre);
SINT32 get_vector_32(utils_sampling_t *sampler, SINT32 *v, size_t n, FLOAT centre);

#ifdef __cplusplus
}
#endif
#endif /* SCGPU_DROPIN_H */
