// csprng.cuh -- device-side restatement of the word stream libsafecrypto's samplers consume.
//
//   prng_32 / prng_64 / prng_var / prng_8 / prng_bit      src/utils/crypto/prng.c:963-1048
//   ChaCha20-CSPRNG framing                               src/utils/crypto/chacha20_csprng.c:21-106
//   ChaCha20 block, 20 rounds, 64-bit block counter       src/utils/crypto/chacha/chacha20.c:88-210
//   AES-256 CTR-DRBG framing                              src/utils/crypto/ctr_drbg.c:37-199,
//                                                         prng_get_func.c:159-193
//   user-provided entropy ring buffer                     prng_get_func.c:108-119
//
// The 4096-word pool of prng.c only buffers: the sequence of 32-bit words handed out is the
// generator's 64-bit draws split high word first, which is what PrngState::next32() yields.
// AES is FIPS-197 with T-tables built in shared memory from the S-box.
#pragma once
#include <cstdint>

namespace scgpu {

enum { PRNG_AES = 0, PRNG_CHACHA20 = 2 };

// Resumable per-stream generator state (also the layout kept in device memory for the drop-in
// prng_ctx_t, see host_sampling.cu).
struct PrngState {
    uint32_t type;
    uint32_t seed_len;
    uint32_t ent_idx;          // ring-buffer read position in the seed bytes
    uint32_t seed_period;      // ChaCha: bytes between reseeds; DRBG: 1 KiB updates between reseeds
    // ChaCha20-CSPRNG
    uint32_t cc_key[8];
    uint32_t cc_iv[2];
    uint32_t cc_ctr_lo, cc_ctr_hi;
    uint32_t cc_data[4];       // running XOR of the first 16 bytes of every block since (re)seed
    uint32_t cc_count;         // data_count / 4, 0..3
    uint32_t cc_reseed_ctr;
    // CTR-DRBG
    uint32_t drbg_counter;
    uint32_t drbg_blocks;      // blocks produced since the last 1 KiB boundary bookkeeping (0..63)
    uint32_t drbg_reseed_ctr;
    uint32_t drbg_rk[60];      // AES-256 round keys, big-endian words
    uint32_t drbg_buf[4];      // current 16-byte block as 4 output words (hi0, lo0, hi1, lo1)
    uint32_t drbg_pos;         // next word of drbg_buf (4 = empty)
    // prng.c front end
    uint32_t have_lo;          // second half of a 64-bit draw pending
    uint32_t lo_word;
    uint32_t var_buf, var_bits;
    uint32_t error;            // 2: the fresh-entropy ring ran dry inside a launch (never re-use entropy)
    // pooled mode = the drop-in prng_ctx_t: prng_32 & co. are served from the 4096-word bit pool that prng.c:95-132
    // refills with 2048 generator draws at a time, prng_mem draws from the generator BEHIND that pool, and the DRBG
    // keeps its 1 KiB transfer buffer (prng_get_func.c:174-193) so that prng_reset's stale reads are reproduced
    uint32_t pooled;
    uint32_t pool_rd, pool_fill;
    uint32_t rng_cnt;          // next u64 of the DRBG transfer buffer, 128 = empty
    uint32_t ent_fresh;        // 1: the ring holds fresh (callback / OS) entropy and must never wrap onto used bytes
    uint32_t ent_avail;        // fresh bytes left ahead of ent_idx
    uint32_t instantiated;     // 0: init() has not run yet (the host creates the header, the device the generator)
    uint64_t words_out;        // 32-bit words handed out so far (stats_out_bytes / 4)
    uint64_t draws64;          // generator draws so far (stats_csprng_bytes / 8)
};
constexpr uint32_t kPoolWords = 4096;      // RANDOM_POOL_SIZE, prng.h:31
constexpr uint32_t kDrbgBufWords = 256;    // CSPRNG_BUFFER_SIZE / 4, prng_types.h:52

__device__ __constant__ uint8_t kAesSbox[256] = {
    0x63,0x7c,0x77,0x7b,0xf2,0x6b,0x6f,0xc5,0x30,0x01,0x67,0x2b,0xfe,0xd7,0xab,0x76,0xca,0x82,0xc9,0x7d,0xfa,0x59,0x47,0xf0,
    0xad,0xd4,0xa2,0xaf,0x9c,0xa4,0x72,0xc0,0xb7,0xfd,0x93,0x26,0x36,0x3f,0xf7,0xcc,0x34,0xa5,0xe5,0xf1,0x71,0xd8,0x31,0x15,
    0x04,0xc7,0x23,0xc3,0x18,0x96,0x05,0x9a,0x07,0x12,0x80,0xe2,0xeb,0x27,0xb2,0x75,0x09,0x83,0x2c,0x1a,0x1b,0x6e,0x5a,0xa0,
    0x52,0x3b,0xd6,0xb3,0x29,0xe3,0x2f,0x84,0x53,0xd1,0x00,0xed,0x20,0xfc,0xb1,0x5b,0x6a,0xcb,0xbe,0x39,0x4a,0x4c,0x58,0xcf,
    0xd0,0xef,0xaa,0xfb,0x43,0x4d,0x33,0x85,0x45,0xf9,0x02,0x7f,0x50,0x3c,0x9f,0xa8,0x51,0xa3,0x40,0x8f,0x92,0x9d,0x38,0xf5,
    0xbc,0xb6,0xda,0x21,0x10,0xff,0xf3,0xd2,0xcd,0x0c,0x13,0xec,0x5f,0x97,0x44,0x17,0xc4,0xa7,0x7e,0x3d,0x64,0x5d,0x19,0x73,
    0x60,0x81,0x4f,0xdc,0x22,0x2a,0x90,0x88,0x46,0xee,0xb8,0x14,0xde,0x5e,0x0b,0xdb,0xe0,0x32,0x3a,0x0a,0x49,0x06,0x24,0x5c,
    0xc2,0xd3,0xac,0x62,0x91,0x95,0xe4,0x79,0xe7,0xc8,0x37,0x6d,0x8d,0xd5,0x4e,0xa9,0x6c,0x56,0xf4,0xea,0x65,0x7a,0xae,0x08,
    0xba,0x78,0x25,0x2e,0x1c,0xa6,0xb4,0xc6,0xe8,0xdd,0x74,0x1f,0x4b,0xbd,0x8b,0x8a,0x70,0x3e,0xb5,0x66,0x48,0x03,0xf6,0x0e,
    0x61,0x35,0x57,0xb9,0x86,0xc1,0x1d,0x9e,0xe1,0xf8,0x98,0x11,0x69,0xd9,0x8e,0x94,0x9b,0x1e,0x87,0xe9,0xce,0x55,0x28,0xdf,
    0x8c,0xa1,0x89,0x0d,0xbf,0xe6,0x42,0x68,0x41,0x99,0x2d,0x0f,0xb0,0x54,0xbb,0x16};

// Shared-memory AES tables: te0[x] = (2s, s, s, 3s) big-endian word, plus the S-box.  te1..te3 are
// byte rotations of te0 (PRMT), so one 1 KiB table serves all four.
struct AesTables { uint32_t te0[256]; uint8_t sbox[256]; };

__device__ __forceinline__ void aes_tables_init(AesTables &t)
{
    for (int x = threadIdx.x; x < 256; x += blockDim.x) {
        uint32_t s = kAesSbox[x];
        uint32_t s2 = ((s << 1) ^ ((s & 0x80) ? 0x1B : 0)) & 0xFF;
        uint32_t s3 = s2 ^ s;
        t.te0[x] = (s2 << 24) | (s << 16) | (s << 8) | s3;
        t.sbox[x] = (uint8_t)s;
    }
}

__device__ __forceinline__ uint32_t rotr8(uint32_t v, int bytes) { return __funnelshift_r(v, v, 8 * bytes); }

__device__ __forceinline__ uint32_t aes_subword(const AesTables &t, uint32_t w)
{
    return ((uint32_t)t.sbox[w >> 24] << 24) | ((uint32_t)t.sbox[(w >> 16) & 0xFF] << 16) |
           ((uint32_t)t.sbox[(w >> 8) & 0xFF] << 8) | (uint32_t)t.sbox[w & 0xFF];
}

// key: 8 big-endian words -> 60 round-key words
__device__ __forceinline__ void aes256_expand(const AesTables &t, const uint32_t key[8], uint32_t *rk)
{
    for (int i = 0; i < 8; i++) rk[i] = key[i];
    uint32_t rcon = 0x01000000u;
    for (int i = 8; i < 60; i++) {
        uint32_t tmp = rk[i - 1];
        if ((i & 7) == 0) {
            tmp = aes_subword(t, (tmp << 8) | (tmp >> 24)) ^ rcon;
            rcon = ((rcon << 1) ^ ((rcon & 0x80000000u) ? 0x1B000000u : 0)) & 0xFF000000u;
        } else if ((i & 7) == 4) {
            tmp = aes_subword(t, tmp);
        }
        rk[i] = rk[i - 8] ^ tmp;
    }
}

// state words are big-endian columns
__device__ __forceinline__ void aes256_encrypt(const AesTables &t, const uint32_t *rk, uint32_t s0, uint32_t s1,
                                               uint32_t s2, uint32_t s3, uint32_t out[4])
{
    s0 ^= rk[0]; s1 ^= rk[1]; s2 ^= rk[2]; s3 ^= rk[3];
#pragma unroll 1
    for (int r = 1; r < 14; r++) {
        uint32_t t0 = t.te0[s0 >> 24] ^ rotr8(t.te0[(s1 >> 16) & 0xFF], 1) ^ rotr8(t.te0[(s2 >> 8) & 0xFF], 2) ^ rotr8(t.te0[s3 & 0xFF], 3) ^ rk[4 * r];
        uint32_t t1 = t.te0[s1 >> 24] ^ rotr8(t.te0[(s2 >> 16) & 0xFF], 1) ^ rotr8(t.te0[(s3 >> 8) & 0xFF], 2) ^ rotr8(t.te0[s0 & 0xFF], 3) ^ rk[4 * r + 1];
        uint32_t t2 = t.te0[s2 >> 24] ^ rotr8(t.te0[(s3 >> 16) & 0xFF], 1) ^ rotr8(t.te0[(s0 >> 8) & 0xFF], 2) ^ rotr8(t.te0[s1 & 0xFF], 3) ^ rk[4 * r + 2];
        uint32_t t3 = t.te0[s3 >> 24] ^ rotr8(t.te0[(s0 >> 16) & 0xFF], 1) ^ rotr8(t.te0[(s1 >> 8) & 0xFF], 2) ^ rotr8(t.te0[s2 & 0xFF], 3) ^ rk[4 * r + 3];
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    const uint8_t *sb = t.sbox;
    out[0] = (((uint32_t)sb[s0 >> 24] << 24) | ((uint32_t)sb[(s1 >> 16) & 0xFF] << 16) | ((uint32_t)sb[(s2 >> 8) & 0xFF] << 8) | sb[s3 & 0xFF]) ^ rk[56];
    out[1] = (((uint32_t)sb[s1 >> 24] << 24) | ((uint32_t)sb[(s2 >> 16) & 0xFF] << 16) | ((uint32_t)sb[(s3 >> 8) & 0xFF] << 8) | sb[s0 & 0xFF]) ^ rk[57];
    out[2] = (((uint32_t)sb[s2 >> 24] << 24) | ((uint32_t)sb[(s3 >> 16) & 0xFF] << 16) | ((uint32_t)sb[(s0 >> 8) & 0xFF] << 8) | sb[s1 & 0xFF]) ^ rk[58];
    out[3] = (((uint32_t)sb[s3 >> 24] << 24) | ((uint32_t)sb[(s0 >> 16) & 0xFF] << 16) | ((uint32_t)sb[(s1 >> 8) & 0xFF] << 8) | sb[s2 & 0xFF]) ^ rk[59];
}

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// Bank-replicated T-tables for the throughput kernel: te_k[x * 32 + lane] -- every lane reads its own bank, so
// the 16 data-dependent lookups of a round are conflict-free (the 1 KiB table costs ~3.5 wavefronts per LDS
// with 32 random indices; profiles/gauss_r01).  The S-box of the last round is byte 2 of the same entry.
// Four tables te0..te3 (te_k = te0 rotated right by k bytes), 32 KiB each, so that a round needs no rotations:
// per column 4 x (shift + LOP3) index computations, 4 LDS and 2 three-input XORs.
__device__ __forceinline__ void aes_rep_init(uint32_t *te0r)
{
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
        uint32_t s = kAesSbox[i >> 5];
        uint32_t s2 = ((s << 1) ^ ((s & 0x80) ? 0x1B : 0)) & 0xFF;
        const uint32_t t0 = (s2 << 24) | (s << 16) | (s << 8) | (s2 ^ s);
        te0r[i] = t0;
        te0r[i + 8192] = __funnelshift_r(t0, t0, 8);
        te0r[i + 16384] = __funnelshift_r(t0, t0, 16);
        te0r[i + 24576] = __funnelshift_r(t0, t0, 24);
    }
}

// rk: 60 round-key words; te0r = the replicated table, l4 = 4 * lane.  A lookup's byte offset is
// ((s >> (8k - 7)) & 0x7F80) | l4: one shift + one 3-input LOP3, the table base rides in the LDS address.
__device__ __forceinline__ uint32_t aes_rep_at(const uint32_t *te0r, uint32_t off)
{
    return *reinterpret_cast<const uint32_t *>(reinterpret_cast<const unsigned char *>(te0r) + off);
}
__device__ __forceinline__ void aes256_encrypt_rep(const uint32_t *te0r, uint32_t l4, const uint32_t *rk, uint32_t s0, uint32_t s1,
                                                   uint32_t s2, uint32_t s3, uint32_t out[4])
{
#define SCGPU_T(x) aes_rep_at(te0r, (x))
#define SCGPU_T1(x) aes_rep_at(te0r, (x) + 32768u)
#define SCGPU_T2(x) aes_rep_at(te0r, (x) + 65536u)
#define SCGPU_T3(x) aes_rep_at(te0r, (x) + 98304u)
#define SCGPU_B3(v) ((((v) >> 17) & 0x7F80u) | l4)
#define SCGPU_B2(v) ((((v) >> 9) & 0x7F80u) | l4)
#define SCGPU_B1(v) ((((v) >> 1) & 0x7F80u) | l4)
#define SCGPU_B0(v) ((((v) << 7) & 0x7F80u) | l4)
    s0 ^= __ldg(rk + 0); s1 ^= __ldg(rk + 1); s2 ^= __ldg(rk + 2); s3 ^= __ldg(rk + 3);
#pragma unroll 2
    for (int r = 1; r < 14; r++) {
        const uint4 k = __ldg(reinterpret_cast<const uint4 *>(rk) + r);
        uint32_t t0 = SCGPU_T(SCGPU_B3(s0)) ^ SCGPU_T1(SCGPU_B2(s1)) ^ SCGPU_T2(SCGPU_B1(s2)) ^ SCGPU_T3(SCGPU_B0(s3)) ^ k.x;
        uint32_t t1 = SCGPU_T(SCGPU_B3(s1)) ^ SCGPU_T1(SCGPU_B2(s2)) ^ SCGPU_T2(SCGPU_B1(s3)) ^ SCGPU_T3(SCGPU_B0(s0)) ^ k.y;
        uint32_t t2 = SCGPU_T(SCGPU_B3(s2)) ^ SCGPU_T1(SCGPU_B2(s3)) ^ SCGPU_T2(SCGPU_B1(s0)) ^ SCGPU_T3(SCGPU_B0(s1)) ^ k.z;
        uint32_t t3 = SCGPU_T(SCGPU_B3(s3)) ^ SCGPU_T1(SCGPU_B2(s0)) ^ SCGPU_T2(SCGPU_B1(s1)) ^ SCGPU_T3(SCGPU_B0(s2)) ^ k.w;
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    // last round: S-box = byte 2 of the table entry, placed into byte 3 / 2 / 1 / 0
#define SCGPU_S(x) (SCGPU_T(x) & 0x00FF0000u)
    const uint4 k = __ldg(reinterpret_cast<const uint4 *>(rk) + 14);
    out[0] = ((SCGPU_S(SCGPU_B3(s0)) << 8) | SCGPU_S(SCGPU_B2(s1)) | (SCGPU_S(SCGPU_B1(s2)) >> 8) | (SCGPU_S(SCGPU_B0(s3)) >> 16)) ^ k.x;
    out[1] = ((SCGPU_S(SCGPU_B3(s1)) << 8) | SCGPU_S(SCGPU_B2(s2)) | (SCGPU_S(SCGPU_B1(s3)) >> 8) | (SCGPU_S(SCGPU_B0(s0)) >> 16)) ^ k.y;
    out[2] = ((SCGPU_S(SCGPU_B3(s2)) << 8) | SCGPU_S(SCGPU_B2(s3)) | (SCGPU_S(SCGPU_B1(s0)) >> 8) | (SCGPU_S(SCGPU_B0(s1)) >> 16)) ^ k.z;
    out[3] = ((SCGPU_S(SCGPU_B3(s3)) << 8) | SCGPU_S(SCGPU_B2(s0)) | (SCGPU_S(SCGPU_B1(s1)) >> 8) | (SCGPU_S(SCGPU_B0(s2)) >> 16)) ^ k.w;
#undef SCGPU_S
#undef SCGPU_T
#undef SCGPU_T1
#undef SCGPU_T2
#undef SCGPU_T3
#undef SCGPU_B3
#undef SCGPU_B2
#undef SCGPU_B1
#undef SCGPU_B0
}

// CTR-DRBG block input: the 32-bit counter replicated four times in native (little-endian) byte
// order (ctr_drbg.c:180-186); as big-endian AES columns that is bswap(counter) four times.
// Output words in prng order for the two 64-bit values of the block: (hi0, lo0, hi1, lo1) where
// value k is the little-endian u64 at byte 8k of the ciphertext.
__device__ __forceinline__ void drbg_block_words(const AesTables &t, const uint32_t *rk, uint32_t counter, uint32_t w[4])
{
    uint32_t c = bswap32(counter), o[4];
    aes256_encrypt(t, rk, c, c, c, c, o);
    // ciphertext bytes 4i..4i+3 are the big-endian bytes of o[i]; little-endian reads swap them
    w[0] = bswap32(o[1]); w[1] = bswap32(o[0]); w[2] = bswap32(o[3]); w[3] = bswap32(o[2]);
}

// ---- ChaCha20 ----------------------------------------------------------------------------------------
// One of the four rotations of a quarter-round is issued on the fma pipe as  hi32(x * 2^r) + x * 2^r  (IMAD.HI +
// IMAD): the ALU pipe (LOP3, SHF) is this kernel's bottleneck -- 8 of a quarter-round's 12 operations -- while
// the fma-heavy pipe only carries the 4 additions.
__device__ __forceinline__ uint32_t rotl_fma(uint32_t x, uint32_t pow2)
{
    uint32_t hi, r;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(hi) : "r"(x), "r"(pow2));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(pow2), "r"(hi));
    return r;
}
#define SCGPU_QR(a, b, c, d)                                                                  \
    a += b; d ^= a; d = rotl_fma(d, 1u << 16); c += d; b ^= c; b = __funnelshift_l(b, b, 12); \
    a += b; d ^= a; d = __funnelshift_l(d, d, 8);  c += d; b ^= c; b = __funnelshift_l(b, b, 7);

// first four keystream words of one block
__device__ __forceinline__ void chacha20_first16(const uint32_t key[8], uint32_t ctr_lo, uint32_t ctr_hi,
                                                 uint32_t iv0, uint32_t iv1, uint32_t ks[4])
{
    uint32_t x0 = 0x61707865, x1 = 0x3320646e, x2 = 0x79622d32, x3 = 0x6b206574;
    uint32_t x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3], x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
    uint32_t x12 = ctr_lo, x13 = ctr_hi, x14 = iv0, x15 = iv1;
#pragma unroll 1
    for (int i = 0; i < 10; i++) {
        SCGPU_QR(x0, x4, x8, x12) SCGPU_QR(x1, x5, x9, x13) SCGPU_QR(x2, x6, x10, x14) SCGPU_QR(x3, x7, x11, x15)
        SCGPU_QR(x0, x5, x10, x15) SCGPU_QR(x1, x6, x11, x12) SCGPU_QR(x2, x7, x8, x13) SCGPU_QR(x3, x4, x9, x14)
    }
    ks[0] = x0 + 0x61707865; ks[1] = x1 + 0x3320646e; ks[2] = x2 + 0x79622d32; ks[3] = x3 + 0x6b206574;
}

// ctr_drbg_create + the first ctr_drbg_reseed (ctr_drbg.c:37-70, 100-147) of a FRESH stream, inlined: zero key, zero
// counter, three counter blocks encrypted, key = entropy[4..36) ^ those bytes, counter = 3 ^ entropy[0..4).  What
// PrngStream::init computes for PRNG_AES, without the resumable state (the throughput kernels set up one generator
// per stream and the out-of-line version kept the whole PrngState in local memory).
__device__ __forceinline__ void drbg_instantiate(const AesTables &t, const uint8_t *seed, uint32_t seed_len, uint32_t *rk, uint32_t &counter)
{
    uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    aes256_expand(t, zero, rk);
    uint32_t bytes_be[12];
    uint32_t ctr = 0;
    for (int block = 3; block > 0;) {
        ctr++;
        const uint32_t c = bswap32(ctr);
        uint32_t o[4];
        block--;
        aes256_encrypt(t, rk, c, c, c, c, o);
        for (int i = 0; i < 4; i++) bytes_be[4 * block + i] = o[i];
    }
    uint32_t e = 0;
    auto le32 = [&]() {
        uint32_t v = 0;
        for (int b = 0; b < 4; b++) { v |= (uint32_t)seed[e] << (8 * b); if (++e == seed_len) e = 0; }
        return v;
    };
    const uint32_t ctr_le = le32();
    uint32_t key[8];
    for (int i = 0; i < 8; i++) key[i] = bswap32(le32()) ^ bytes_be[3 + i];
    counter = ctr ^ ctr_le;
    aes256_expand(t, key, rk);
}

// ---- the sequential generator -------------------------------------------------------------------------
struct PrngStream {
    PrngState s;
    const uint8_t *seed;
    const AesTables *aes;
    uint32_t *pool = nullptr;      // pooled mode: kPoolWords words of device memory
    uint32_t *buf1k = nullptr;     // pooled mode, DRBG: kDrbgBufWords words in prng output order (hi, lo per u64)

    __device__ __forceinline__ uint8_t ent_byte()
    {
        if (s.ent_fresh) {
            if (s.ent_avail == 0) { s.error = 2; return 0; }
            s.ent_avail--;
        }
        uint8_t b = seed[s.ent_idx++];
        if (s.ent_idx == s.seed_len) s.ent_idx = 0;
        return b;
    }
    __device__ __forceinline__ uint32_t ent_le32()
    {
        uint32_t v = ent_byte();
        v |= (uint32_t)ent_byte() << 8; v |= (uint32_t)ent_byte() << 16; v |= (uint32_t)ent_byte() << 24;
        return v;
    }
    // chacha20_csprng.c:21-29
    __device__ __noinline__ void chacha_reseed()
    {
        s.cc_reseed_ctr = 0;
        for (int i = 0; i < 8; i++) s.cc_key[i] = ent_le32();
        s.cc_iv[0] = ent_le32(); s.cc_iv[1] = ent_le32();
        s.cc_ctr_lo = 0; s.cc_ctr_hi = 0;
        s.cc_data[0] = s.cc_data[1] = s.cc_data[2] = s.cc_data[3] = 0;
    }
    // ctr_drbg.c:100-147
    __device__ __noinline__ void drbg_reseed()
    {
        uint32_t bytes_be[12];                         // three ciphertext blocks as big-endian words
        s.drbg_reseed_ctr = 0;
        for (int block = 3; block > 0;) {
            s.drbg_counter++;
            uint32_t c = bswap32(s.drbg_counter), o[4];
            block--;
            aes256_encrypt(*aes, s.drbg_rk, c, c, c, c, o);
            for (int i = 0; i < 4; i++) bytes_be[4 * block + i] = o[i];
        }
        uint32_t ctr_le = ent_le32();
        uint32_t key[8];
        for (int i = 0; i < 8; i++) {
            // key bytes 4i..4i+3 = entropy ^ bytes[12 + 4i ..]; as a big-endian word
            uint32_t e = bswap32(ent_le32());
            key[i] = e ^ bytes_be[3 + i];
        }
        s.drbg_counter ^= ctr_le;
        aes256_expand(*aes, key, s.drbg_rk);
    }
    // `pooled`, `ent_fresh` and `ent_avail` are set by the caller beforehand
    __device__ __noinline__ void init(uint32_t type, uint32_t seed_len, uint32_t seed_period)
    {
        s.type = type; s.seed_len = seed_len; s.ent_idx = 0;
        s.cc_count = 0; s.have_lo = 0; s.lo_word = 0; s.var_buf = 0; s.var_bits = 0; s.error = 0; s.words_out = 0;
        s.drbg_pos = 4; s.drbg_blocks = 0; s.instantiated = 1;
        s.pool_rd = 0; s.pool_fill = 0; s.rng_cnt = kDrbgBufWords / 2; s.draws64 = 0;
        if (type == PRNG_CHACHA20) {
            s.seed_period = seed_period;
            chacha_reseed();
        } else {
            // ctr_drbg_create, ctr_drbg.c:37-70: zero key, zero counter, period in 16-byte units clamped
            uint64_t blocks = (uint64_t)seed_period >> 4;
            if (blocks > 0x80000000ull) blocks = 0x80000000ull;
            else if (blocks < 0x1000ull) blocks = 0x1000ull;
            s.seed_period = (uint32_t)blocks;
            uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            s.drbg_counter = 0;
            aes256_expand(*aes, zero, s.drbg_rk);
            drbg_reseed();
        }
    }
    // chacha20_csprng.c:72-84
    __device__ __forceinline__ uint32_t chacha_next32()
    {
        s.cc_count++;
        if (s.cc_count == 4) {
            uint32_t ks[4];
            s.cc_count = 0;
            chacha20_first16(s.cc_key, s.cc_ctr_lo, s.cc_ctr_hi, s.cc_iv[0], s.cc_iv[1], ks);
            s.cc_data[0] ^= ks[0]; s.cc_data[1] ^= ks[1]; s.cc_data[2] ^= ks[2]; s.cc_data[3] ^= ks[3];
            if (++s.cc_ctr_lo == 0) s.cc_ctr_hi++;
        }
        return bswap32(s.cc_data[s.cc_count]);
    }
    // ctr_drbg_reset, ctr_drbg.c:84-101, as prng_reset drives it (prng.c:861-932): the transfer buffer position
    // rng_cnt is left alone, so later draws first drain the blocks the OLD key left in the buffer
    // ChaCha20: the reference's reset_chacha20 frees the generator (chacha20_csprng.c:58-67); what is done here is
    // the swapped destroy_chacha20 body (:49-56): data_count = 0 and a reseed.
    __device__ __noinline__ void reset_pooled()
    {
        s.pool_rd = 0; s.pool_fill = 0; s.var_bits = 0; s.words_out = 0; s.draws64 = 0;
        if (s.type == PRNG_CHACHA20) { s.cc_count = 0; chacha_reseed(); return; }
        uint32_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        s.drbg_counter = 0;
        aes256_expand(*aes, zero, s.drbg_rk);
        drbg_reseed();
    }
    // one 64-bit generator draw, split high word first (prng.c:108-127)
    // not inlined: every sampler of k_stream_seq draws through next32() in several places, and an inlined copy of the
    // ChaCha20 / AES block per call site made the kernel take a quarter of an hour to assemble
    __device__ __noinline__ void draw64(uint32_t &hi, uint32_t &lo)
    {
        s.draws64++;
        if (s.type == PRNG_CHACHA20) {
            s.cc_reseed_ctr += 8;                                   // chacha20_csprng.c:99-106
            if (s.seed_period <= s.cc_reseed_ctr) chacha_reseed();
            hi = chacha_next32();
            lo = chacha_next32();
        } else if (s.pooled) {
            // get_random_64_aes + ctr_drbg_update literally: 64 blocks per refill, reseed bookkeeping right after
            if (s.rng_cnt == kDrbgBufWords / 2) {
                s.rng_cnt = 0;
                for (uint32_t blk = 0; blk < kDrbgBufWords / 4; blk++) {
                    uint32_t w[4];
                    drbg_block_words(*aes, s.drbg_rk, s.drbg_counter++, w);
                    buf1k[4 * blk] = w[0]; buf1k[4 * blk + 1] = w[1]; buf1k[4 * blk + 2] = w[2]; buf1k[4 * blk + 3] = w[3];
                }
                if (++s.drbg_reseed_ctr >= s.seed_period) drbg_reseed();
            }
            hi = buf1k[2 * s.rng_cnt]; lo = buf1k[2 * s.rng_cnt + 1];
            s.rng_cnt++;
        } else {
            if (s.drbg_pos >= 4) {
                if (s.drbg_blocks == 64) {                          // a 1 KiB update completed (ctr_drbg.c:190-196)
                    s.drbg_blocks = 0;
                    if (++s.drbg_reseed_ctr >= s.seed_period) drbg_reseed();
                }
                drbg_block_words(*aes, s.drbg_rk, s.drbg_counter++, s.drbg_buf);
                s.drbg_blocks++;
                s.drbg_pos = 0;
            }
            hi = s.drbg_buf[s.drbg_pos]; lo = s.drbg_buf[s.drbg_pos + 1];
            s.drbg_pos += 2;
        }
    }
    // update_pool, prng.c:95-132: the whole pool at once, 2048 draws, high word first
    __device__ __noinline__ void pool_refill()
    {
        for (uint32_t i = 0; i < kPoolWords; i += 2) {
            uint32_t hi, lo;
            draw64(hi, lo);
            pool[i] = hi; pool[i + 1] = lo;
        }
        s.pool_rd = 0; s.pool_fill = kPoolWords;
    }
    __device__ __forceinline__ uint32_t next32()
    {
        s.words_out++;
        if (s.pooled) {
            if (s.pool_rd >= s.pool_fill) pool_refill();
            return pool[s.pool_rd++];
        }
        if (s.have_lo) { s.have_lo = 0; return s.lo_word; }
        uint32_t hi, lo;
        draw64(hi, lo);
        s.lo_word = lo; s.have_lo = 1;
        return hi;
    }
    __device__ __forceinline__ uint64_t next64()
    {
        uint64_t hi = next32();
        return (hi << 32) | next32();
    }
    // prng.c:1017-1048
    __device__ __forceinline__ uint32_t var(uint32_t n)
    {
        uint32_t mask = n >= 32 ? 0xFFFFFFFFu : (1u << n) - 1u;
        if (n > 32) n = 32;
        uint32_t ret = s.var_buf;
        if (s.var_bits < n) {
            uint32_t need = n - s.var_bits;
            // need == 32 shifts by the type width (undefined in C): the reference as compiled for a BMI2 host (shlx / bzhi)
            // returns stale var_buf | fresh word; var_buf is zero whenever var_bits is, except right after prng_reset
            ret = need >= 32 ? ret : ret << need;
            s.var_buf = next32();
            ret |= s.var_buf & (need >= 32 ? 0xFFFFFFFFu : ((1u << need) - 1u));
            s.var_buf = need >= 32 ? s.var_buf : s.var_buf >> need;
            s.var_bits = 32 - need;
        } else {
            s.var_buf >>= n;
            s.var_bits -= n;
        }
        return ret & mask;
    }
};

}  // namespace scgpu
