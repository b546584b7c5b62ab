/*
 * oracle/sc_oracle_prng.c -- TEST INFRASTRUCTURE ONLY (see sc_oracle.h).
 *
 * Restatement of the word stream the reference samplers consume: the 4096-word bit pool of
 * src/utils/crypto/prng.c:95-132,963-1048 on top of either the ChaCha20-CSPRNG framing
 * (chacha20_csprng.c:21-106 over chacha/chacha20.c:88-210) or the AES-256 CTR-DRBG framing
 * (ctr_drbg.c:37-199, prng_get_func.c:159-193).  Entropy is the user-provided ring buffer
 * of prng_get_func.c:108-119.  AES-256 itself is FIPS-197, written from the standard.
 */
#include "sc_oracle.h"

#include <stdlib.h>
#include <string.h>

#define POOL_WORDS      4096        /* RANDOM_POOL_SIZE, prng.h:31         */
#define DRBG_BUF_BYTES  1024        /* CSPRNG_BUFFER_SIZE, prng_types.h:52 */
#define DRBG_MIN_RESEED 0x00001000u /* ctr_drbg.h:31                       */
#define DRBG_MAX_RESEED 0x80000000u /* ctr_drbg.h:27                       */

/* ---- AES-256 (FIPS-197) ---------------------------------------------------------------- */

static uint8_t SBOX[256];
static int sbox_ready;

static uint8_t gf_mul(uint8_t a, uint8_t b)
{
    uint8_t r = 0;
    while (b) { if (b & 1) r ^= a; a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1B : 0)); b >>= 1; }
    return r;
}

static void sbox_init(void)
{
    if (sbox_ready) return;
    /* multiplicative inverse followed by the affine map */
    for (int x = 0; x < 256; x++) {
        uint8_t inv = 0;
        if (x) for (int y = 1; y < 256; y++) if (gf_mul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        uint8_t s = inv, r = inv;
        for (int i = 0; i < 4; i++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        SBOX[x] = (uint8_t)(s ^ 0x63);
    }
    sbox_ready = 1;
}

typedef struct { uint8_t rk[15][16]; } aes256_ks_t;

static void aes256_expand(aes256_ks_t *ks, const uint8_t key[32])
{
    sbox_init();
    uint8_t w[240];
    memcpy(w, key, 32);
    uint8_t rcon = 1;
    for (int i = 32; i < 240; i += 4) {
        uint8_t t[4] = { w[i - 4], w[i - 3], w[i - 2], w[i - 1] };
        if (i % 32 == 0) {
            uint8_t t0 = t[0];
            t[0] = (uint8_t)(SBOX[t[1]] ^ rcon); t[1] = SBOX[t[2]]; t[2] = SBOX[t[3]]; t[3] = SBOX[t0];
            rcon = gf_mul(rcon, 2);
        } else if (i % 32 == 16) {
            for (int j = 0; j < 4; j++) t[j] = SBOX[t[j]];
        }
        for (int j = 0; j < 4; j++) w[i + j] = (uint8_t)(w[i - 32 + j] ^ t[j]);
    }
    memcpy(ks->rk, w, 240);
}

static void aes256_encrypt(const aes256_ks_t *ks, const uint8_t in[16], uint8_t out[16])
{
    uint8_t s[16], t[16];
    for (int i = 0; i < 16; i++) s[i] = (uint8_t)(in[i] ^ ks->rk[0][i]);
    for (int round = 1; round <= 14; round++) {
        /* SubBytes + ShiftRows (state is column-major: byte index = 4*col + row) */
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 4; r++)
                t[4 * c + r] = SBOX[s[4 * ((c + r) & 3) + r]];
        if (round < 14) {
            for (int c = 0; c < 4; c++) {
                uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                s[4 * c + 0] = (uint8_t)(gf_mul(a0, 2) ^ gf_mul(a1, 3) ^ a2 ^ a3);
                s[4 * c + 1] = (uint8_t)(a0 ^ gf_mul(a1, 2) ^ gf_mul(a2, 3) ^ a3);
                s[4 * c + 2] = (uint8_t)(a0 ^ a1 ^ gf_mul(a2, 2) ^ gf_mul(a3, 3));
                s[4 * c + 3] = (uint8_t)(gf_mul(a0, 3) ^ a1 ^ a2 ^ gf_mul(a3, 2));
            }
        } else {
            memcpy(s, t, 16);
        }
        for (int i = 0; i < 16; i++) s[i] ^= ks->rk[round][i];
    }
    memcpy(out, s, 16);
}

void orc_aes256_encrypt_block(const uint8_t key[32], const uint8_t in[16], uint8_t out[16])
{
    aes256_ks_t ks;
    aes256_expand(&ks, key);
    aes256_encrypt(&ks, in, out);
}

/* ---- ChaCha20 block (chacha/chacha20.c:88-210, 20 rounds, 64-bit block counter) -------- */

static inline uint32_t rotl32(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
#define QR(a, b, c, d) \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

static void chacha20_block(const uint32_t in[16], uint32_t out[16])
{
    uint32_t x[16];
    memcpy(x, in, 64);
    for (int i = 0; i < 10; i++) {
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

static inline uint32_t le32(const uint8_t *p)
{
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* ---- context --------------------------------------------------------------------------- */

struct orc_prng {
    int type;
    /* user entropy ring buffer (prng_get_func.c:108-119) */
    const uint8_t *ent; size_t ent_len, ent_idx;
    uint8_t *ent_copy;
    size_t seed_period;
    /* bit pool (prng.c:95-132) */
    uint32_t pool[POOL_WORDS];
    int32_t bits, rd_idx;
    uint32_t var_buf; size_t var_bits;
    /* ChaCha20-CSPRNG (chacha20_csprng.h:31-43) */
    uint32_t cc_in[16];
    uint8_t cc_data[16];
    size_t cc_count;
    uint32_t cc_reseed_ctr;
    /* CTR-DRBG (ctr_drbg.h:40-50) + transfer buffer (prng_types.h:111-116) */
    aes256_ks_t ks;
    uint8_t key[32];
    uint32_t counter, drbg_reseed_ctr, drbg_period;
    uint8_t buf[DRBG_BUF_BYTES];
    int rng_cnt;
    /* statistics (prng.c: stats_csprng_bytes / stats_out_bytes) */
    uint64_t csprng_bytes, out_bytes;
};

static void entropy(orc_prng_t *c, size_t n, uint8_t *dst)
{
    for (size_t i = 0; i < n; i++) {
        dst[i] = c->ent[c->ent_idx++];
        if (c->ent_idx == c->ent_len) c->ent_idx = 0;
    }
}

/* chacha20_csprng.c:21-29 */
static void chacha_reseed(orc_prng_t *c)
{
    uint8_t seed[40];
    c->cc_reseed_ctr = 0;
    entropy(c, 40, seed);
    static const uint32_t sigma[4] = { 0x61707865, 0x3320646e, 0x79622d32, 0x6b206574 };
    for (int i = 0; i < 4; i++) c->cc_in[i] = sigma[i];
    for (int i = 0; i < 8; i++) c->cc_in[4 + i] = le32(seed + 4 * i);
    c->cc_in[12] = 0; c->cc_in[13] = 0;          /* state->ctr is never written: stays 0 */
    c->cc_in[14] = le32(seed + 32);
    c->cc_in[15] = le32(seed + 36);
    memset(c->cc_data, 0, 16);
}

/* chacha20_csprng.c:72-84: one block per refill, only its first 16 bytes are used and they
 * are XORed onto the previous 16 output bytes (encrypt-in-place of `data`) */
static uint32_t chacha_next32(orc_prng_t *c)
{
    c->cc_count += 4;
    if (c->cc_count == 16) {
        uint32_t ks[16];
        c->cc_count = 0;
        chacha20_block(c->cc_in, ks);
        for (int i = 0; i < 4; i++) {
            uint32_t v = le32(c->cc_data + 4 * i) ^ ks[i];
            c->cc_data[4 * i] = (uint8_t)v; c->cc_data[4 * i + 1] = (uint8_t)(v >> 8);
            c->cc_data[4 * i + 2] = (uint8_t)(v >> 16); c->cc_data[4 * i + 3] = (uint8_t)(v >> 24);
        }
        if (++c->cc_in[12] == 0) c->cc_in[13]++;
    }
    const uint8_t *d = c->cc_data + c->cc_count;
    return ((uint32_t)d[0] << 24) | ((uint32_t)d[1] << 16) | ((uint32_t)d[2] << 8) | (uint32_t)d[3];
}

/* chacha20_csprng.c:99-106 */
static uint64_t chacha_random64(orc_prng_t *c)
{
    c->csprng_bytes += 8;                                /* get_random_64_chacha, prng_get_func.c:242-261 */
    c->cc_reseed_ctr += 8;
    if ((uint32_t)c->seed_period <= c->cc_reseed_ctr) chacha_reseed(c);
    uint64_t hi = chacha_next32(c);
    return (hi << 32) | chacha_next32(c);
}

/* ctr_drbg.c:100-147 */
static void drbg_reseed(orc_prng_t *c)
{
    uint8_t bytes[48], blk[16], ctrb[4];
    c->drbg_reseed_ctr = 0;
    for (int block = 3; block > 0; ) {
        c->counter++;
        for (int i = 0; i < 16; i++) blk[i] = (uint8_t)(c->counter >> (8 * (i & 3)));
        block--;
        aes256_encrypt(&c->ks, blk, bytes + 16 * block);
    }
    entropy(c, 4, ctrb);
    entropy(c, 32, c->key);
    for (int i = 0; i < 32; i++) c->key[i] ^= bytes[12 + i];
    c->counter ^= le32(ctrb);
    aes256_expand(&c->ks, c->key);
}

/* ctr_drbg.c:164-199 */
static void drbg_fill(orc_prng_t *c)
{
    uint8_t blk[16];
    for (int off = 0; off < DRBG_BUF_BYTES; off += 16) {
        uint32_t ctr = c->counter++;
        for (int i = 0; i < 16; i++) blk[i] = (uint8_t)(ctr >> (8 * (i & 3)));
        aes256_encrypt(&c->ks, blk, c->buf + off);
    }
    if (++c->drbg_reseed_ctr >= c->drbg_period) drbg_reseed(c);
}

/* prng_get_func.c:174-193 */
static uint64_t drbg_random64(orc_prng_t *c)
{
    if (c->rng_cnt == DRBG_BUF_BYTES / 8) { c->rng_cnt = 0; drbg_fill(c); }
    c->csprng_bytes += 8;
    const uint8_t *b = c->buf + 8 * c->rng_cnt++;
    return (uint64_t)le32(b) | ((uint64_t)le32(b + 4) << 32);
}

orc_prng_t *orc_prng_create(int prng_type, const uint8_t *seed, size_t seed_len, size_t seed_period)
{
    if (prng_type != ORC_PRNG_CHACHA && prng_type != ORC_PRNG_AES_CTR_DRBG) return NULL;
    orc_prng_t *c = calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->type = prng_type;
    c->ent_copy = malloc(seed_len);
    memcpy(c->ent_copy, seed, seed_len);
    c->ent = c->ent_copy; c->ent_len = seed_len;
    c->seed_period = seed_period ? seed_period : 0x00100000;
    c->rng_cnt = DRBG_BUF_BYTES / 8;                      /* prng.c:630 */
    if (prng_type == ORC_PRNG_CHACHA) {
        chacha_reseed(c);                                 /* create_chacha20, chacha20_csprng.c:31-47 */
    } else {
        /* ctr_drbg_create, ctr_drbg.c:37-70: key = 0 (zeroed allocation), counter = 0 */
        size_t blocks = c->seed_period >> 4;
        if (blocks > DRBG_MAX_RESEED) blocks = DRBG_MAX_RESEED;
        else if (blocks < DRBG_MIN_RESEED) blocks = DRBG_MIN_RESEED;
        c->drbg_period = (uint32_t)blocks;
        aes256_expand(&c->ks, c->key);
        drbg_reseed(c);
    }
    return c;
}

void orc_prng_destroy(orc_prng_t *c)
{
    if (!c) return;
    free(c->ent_copy);
    free(c);
}

uint32_t orc_prng_var(orc_prng_t *c, size_t n);

/* prng.c:95-132: refill the whole pool with 2048 64-bit draws, high word first */
static void pool_refill(orc_prng_t *c)
{
    if (c->bits != 0) return;
    for (int i = 0; i < POOL_WORDS; i += 2) {
        uint64_t d = (c->type == ORC_PRNG_CHACHA) ? chacha_random64(c) : drbg_random64(c);
        c->pool[i] = (uint32_t)(d >> 32);
        c->pool[i + 1] = (uint32_t)d;
    }
    c->rd_idx = 0;
    c->bits = 32 * POOL_WORDS;
}

uint32_t orc_prng_32(orc_prng_t *c)
{
    pool_refill(c);
    uint32_t v = c->pool[c->rd_idx];
    c->bits -= 32;
    if (++c->rd_idx >= POOL_WORDS) c->rd_idx = 0;
    c->out_bytes += 4;
    return v;
}

uint64_t orc_prng_64(orc_prng_t *c)
{
    uint64_t hi = orc_prng_32(c);           /* prng.c:963-978 */
    return (hi << 32) | orc_prng_32(c);
}

/* prng.c:1005-1015 */
float orc_prng_float(orc_prng_t *c) { return ((float)orc_prng_32(c)) / UINT32_MAX; }
double orc_prng_double(orc_prng_t *c)
{
    uint32_t a = orc_prng_var(c, 27);
    uint32_t b = orc_prng_var(c, 26);
    return (a * 67108864.0 + b) * 1.11022302462516e-16;
}

/* prng.c:1050-1105: whole 64-byte blocks of eight generator draws each, taken from the generator
 * directly -- the words already sitting in the bit pool are NOT consumed and stay ahead of them */
int32_t orc_prng_mem(orc_prng_t *c, uint8_t *mem, int32_t length)
{
    int32_t num_blocks = (length + 63) >> 6;
    uint8_t *p = mem;
    while (num_blocks--) {
        uint64_t d[8];
        for (int i = 0; i < 8; i++) d[i] = (c->type == ORC_PRNG_CHACHA) ? chacha_random64(c) : drbg_random64(c);
        memcpy(p, d, (size_t)(length >= 64 ? 64 : length));      /* little-endian host, as the reference's union */
        length -= 64;
        p += 64;
    }
    return 0;
}

/* prng.c:861-932 with ctr_drbg_reset (ctr_drbg.c:84-101).  The 1 KiB transfer buffer position rng_cnt is NOT reset,
 * so the draws that follow first drain what the old key left in it.  For ChaCha20 the reference calls
 * reset_chacha20, which frees the generator (chacha20_csprng.c:58-67): anything after that is a use after free,
 * there is no behaviour to restate and this function refuses. */
int orc_prng_reset(orc_prng_t *c)
{
    if (c->type != ORC_PRNG_AES_CTR_DRBG) return 1;
    c->bits = 0; c->rd_idx = 0; c->var_bits = 0;
    c->csprng_bytes = 0; c->out_bytes = 0;
    c->counter = 0;
    memset(c->key, 0, 32);
    aes256_expand(&c->ks, c->key);
    drbg_reseed(c);
    return 0;
}
uint64_t orc_prng_csprng_bytes(orc_prng_t *c) { return c->csprng_bytes; }
uint64_t orc_prng_out_bytes(orc_prng_t *c) { return c->out_bytes; }

/* prng.c:1017-1048: LSB-first bit buffer refilled from prng_32 */
uint32_t orc_prng_var(orc_prng_t *c, size_t n)
{
    uint32_t mask;
    if (n >= 32) { n = 32; mask = 0xFFFFFFFFu; } else mask = (1u << n) - 1u;
    uint32_t ret = c->var_buf;
    if (c->var_bits < n) {
        size_t need = n - c->var_bits;
        /* need == 32 (n = 32 on an empty buffer) shifts by the type width, undefined in C.  As compiled with the
         * reference's flags on a BMI2 host (shlx / bzhi, oracle/Makefile) the shift leaves the value alone and the
         * mask keeps all 32 bits: the result is stale_var_buf | fresh word.  var_buf is zero whenever var_bits is,
         * except right after prng_reset, so this is only observable there. */
        ret = need >= 32 ? ret : ret << need;
        c->var_buf = orc_prng_32(c);
        ret |= c->var_buf & (need >= 32 ? 0xFFFFFFFFu : ((1u << need) - 1u));
        c->var_buf = need >= 32 ? c->var_buf : c->var_buf >> need;
        c->var_bits = 32 - need;
    } else {
        c->var_buf >>= n;
        c->var_bits -= n;
    }
    return ret & mask;
}

uint32_t orc_prng_8(orc_prng_t *c) { return orc_prng_var(c, 8) & 0xFFu; }
int32_t orc_prng_bit(orc_prng_t *c) { return (int32_t)orc_prng_var(c, 1); }

int orc_prng_script(int prng_type, const uint8_t *seed, size_t seed_len, size_t seed_period,
                    const int32_t *script, size_t ndraws, uint32_t *out)
{
    orc_prng_t *c = orc_prng_create(prng_type, seed, seed_len, seed_period);
    if (!c) return -1;
    size_t o = 0;
    for (size_t i = 0; i < ndraws; i++) {
        int kind = script[2 * i], arg = script[2 * i + 1];
        if (kind == 32) out[o++] = orc_prng_32(c);
        else if (kind == 64) { uint64_t x = orc_prng_64(c); out[o++] = (uint32_t)(x >> 32); out[o++] = (uint32_t)x; }
        else if (kind == 8) out[o++] = orc_prng_8(c);
        else if (kind == 1) out[o++] = (uint32_t)orc_prng_bit(c);
        else if (kind == 16) out[o++] = orc_prng_var(c, 16);
        else if (kind == 128) {                                          /* prng_128: prng_64 << 64 | prng_64 */
            for (int h = 0; h < 2; h++) { uint64_t x = orc_prng_64(c); out[o++] = (uint32_t)(x >> 32); out[o++] = (uint32_t)x; }
        }
        else if (kind == 2) { float f = orc_prng_float(c); memcpy(&out[o++], &f, 4); }
        else if (kind == 3) { double d = orc_prng_double(c); memcpy(&out[o], &d, 8); o += 2; }
        else if (kind == 4) {                                            /* prng_mem(arg bytes), zero padded to words */
            size_t nw = ((size_t)arg + 3) / 4;
            memset(out + o, 0, nw * 4);
            orc_prng_mem(c, (uint8_t *)(out + o), arg);
            o += nw;
        }
        else if (kind == 5) { if (orc_prng_reset(c)) { orc_prng_destroy(c); return -2; } }
        else if (kind == 6) { out[o++] = (uint32_t)orc_prng_csprng_bytes(c); out[o++] = (uint32_t)orc_prng_out_bytes(c); }
        else out[o++] = orc_prng_var(c, (size_t)arg);
    }
    orc_prng_destroy(c);
    return (int)o;
}
