// rand_product.cu -- create_rand_product_{16,32}_csprng (src/utils/arith/module_lwe.c:588-748) for a batch of
// instances, with the matrix sampled ON THE DEVICE.
//
// The reference draws the k x l matrix of a Kyber / Dilithium instance from a CSPRNG seeded with the instance's
// 32-byte rho (create_csprng, module_lwe.c:914-940): ring after ring, uniform_random_ring_q_csprng (:519-535) takes
// 512 bytes of prng_mem per 256 coefficients, reads them as UINT16, masks to q_bits and subtracts q once.  Round 1
// took the matrix as an input: k l n words per instance had to exist in HBM (and cross PCIe when the caller held
// seeds).  Here a generation kernel expands the seeds of a chunk of instances into a scratch matrix that never
// leaves the L2 (a chunk is sized to 48 MB of matrix against 126 MB of L2), and the fused mat-vec kernel of
// warp32.cuh consumes it: HBM sees 4 n (l + k) bytes + the seed per instance instead of 4 n (k l + l + k).
//
// k_gen_rings: one warp per instance.
//   ChaCha20-CSPRNG  word g of a fresh stream is 0 for g < 3, else word (g-3)%4 of D[(g-3)/4] byte-swapped, D[b] =
//                    XOR_{c<=b} of the first 16 keystream bytes of block c (chacha20_csprng.c:72-84).  Lane L
//                    encrypts blocks [L C, (L+1) C), a warp XOR-scan turns the per-lane sums into the running XOR,
//                    a second pass emits.  prng_mem stores each 64-bit draw (hi << 32 | lo) little-endian, so
//                    coefficients 4d, 4d+1 come from the draw's LOW word (the later one), 4d+2, 4d+3 from its high.
//   AES-CTR-DRBG     ciphertext block b = AES-256_K(counter + b replicated), bytes in order; lane L encrypts blocks
//                    L, L + 32, ...; lane 0 instantiates the DRBG (ctr_drbg.c:37-147) and shares the round keys.
// No reseed can fall inside an instance: create_csprng uses a 16 MiB period, an instance draws k l n 2 bytes.
#include "scgpu_internal.h"
#include <cstdlib>
#include "csprng.cuh"
#include "../../include/scgpu.h"

namespace scgpu {

namespace {

struct GenArgs {
    const uint8_t *seeds;
    uint32_t seed_len, seed_period;
    size_t count;
    int32_t *out;               // [count][rings][n]
    int n, rings;               // rings = k l
    int k, l, transpose;        // destination of ring r: transposed draws are j-major (module_lwe.c:701-722)
    int32_t q;
    uint32_t mask;
    int cache_blocks;           // ChaCha: keystream blocks per lane kept in shared memory between the passes
};

__device__ __forceinline__ int dest_ring(const GenArgs &a, int r)
{
    if (!a.transpose) return r;                      // drawn i-major: A[i][j] = ring i l + j
    const int j = r / a.k, i = r % a.k;              // drawn j-major: ring j k + i multiplies y_j into t_i
    return i * a.l + j;
}

__device__ __forceinline__ int32_t ring_coeff(uint32_t v16, const GenArgs &a)
{
    int32_t x = (int32_t)(v16 & a.mask);
    return x - (x >= a.q ? a.q : 0);
}

// word g of the stream -> coefficients c0, c0 + 1 (c0 even)
__device__ __forceinline__ void emit_word(const GenArgs &a, int32_t *inst_out, size_t g, uint32_t word)
{
    const uint32_t c0 = 4u * (uint32_t)(g >> 1) + ((g & 1) ? 0u : 2u);
    const uint32_t total = (uint32_t)a.rings << 8;                  // n = 256 (the only defined ring length)
    if (c0 >= total) return;
    const int ring = (int)(c0 >> 8), pos = (int)(c0 & 255u);
    int32_t *dst = inst_out + (size_t)dest_ring(a, ring) * a.n + pos;
    *reinterpret_cast<int2 *>(dst) = make_int2(ring_coeff(word & 0xFFFFu, a), ring_coeff(word >> 16, a));
}

template <int PRNG>
__global__ void __launch_bounds__(128) k_gen_rings(GenArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ AesTables aes;
    __shared__ uint32_t s_rk[4][64];                 // per warp: 60 round-key words + counter
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (PRNG == PRNG_AES) aes_tables_init(aes);
    __syncthreads();
    const size_t inst = blockIdx.x * (size_t)(blockDim.x >> 5) + warp;
    if (inst >= a.count) return;
    const uint8_t *seed = a.seeds + inst * a.seed_len;
    int32_t *inst_out = a.out + inst * (size_t)a.rings * a.n;
    const size_t words = (size_t)a.rings * a.n / 2;                  // 32-bit words of generator output

    if (PRNG == PRNG_CHACHA20) {
        uint32_t key[8], iv[2];
        {
            uint32_t e = 0;
            auto le32 = [&]() {
                uint32_t v = 0;
                for (int b = 0; b < 4; b++) { v |= (uint32_t)seed[e] << (8 * b); if (++e == a.seed_len) e = 0; }
                return v;
            };
            for (int i = 0; i < 8; i++) key[i] = le32();
            iv[0] = le32(); iv[1] = le32();
        }
        const size_t nblocks = words > 3 ? (words - 3 + 3) / 4 : 0;
        const size_t C = (nblocks + 31) / 32;
        const size_t b0 = (size_t)lane * C, b1 = (b0 + C < nblocks) ? b0 + C : nblocks;
        uint4 *cache = reinterpret_cast<uint4 *>(smem_raw) + (size_t)warp * a.cache_blocks * 32 + lane;
        const bool cached = C <= (size_t)a.cache_blocks;
        uint32_t acc[4] = {0, 0, 0, 0};
        for (size_t b = b0; b < b1; b++) {
            uint32_t ks[4];
            chacha20_first16(key, (uint32_t)b, (uint32_t)(b >> 32), iv[0], iv[1], ks);
            acc[0] ^= ks[0]; acc[1] ^= ks[1]; acc[2] ^= ks[2]; acc[3] ^= ks[3];
            if (cached) cache[(b - b0) * 32] = make_uint4(ks[0], ks[1], ks[2], ks[3]);
        }
        uint32_t run[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t v = acc[i];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, off);
                if (lane >= off) v ^= o;
            }
            run[i] = v ^ acc[i];                                     // exclusive scan
        }
        if (lane == 0) for (size_t g = 0; g < 3 && g < words; g++) emit_word(a, inst_out, g, 0u);
        for (size_t b = b0; b < b1; b++) {
            uint32_t ks[4];
            if (cached) {
                const uint4 v = cache[(b - b0) * 32];
                ks[0] = v.x; ks[1] = v.y; ks[2] = v.z; ks[3] = v.w;
            } else {
                chacha20_first16(key, (uint32_t)b, (uint32_t)(b >> 32), iv[0], iv[1], ks);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                run[i] ^= ks[i];
                const size_t g = 3 + 4 * b + i;
                if (g < words) emit_word(a, inst_out, g, bswap32(run[i]));
            }
        }
    } else {
        if (lane == 0) {
            uint32_t rk0[60], counter;
            drbg_instantiate(aes, seed, a.seed_len, rk0, counter);
            for (int i = 0; i < 60; i++) s_rk[warp][i] = rk0[i];
            s_rk[warp][60] = counter;
        }
        __syncwarp();
        const uint32_t *rk = s_rk[warp];
        const uint32_t c0 = rk[60];
        const size_t nblocks = (words + 3) / 4;
        for (size_t b = lane; b < nblocks; b += 32) {
            const uint32_t c = bswap32(c0 + (uint32_t)b);
            uint32_t o[4];
            aes256_encrypt(aes, rk, c, c, c, c, o);
            // ciphertext bytes 4i .. 4i+3 are the big-endian bytes of o[i]; prng_mem copies them in order, so the
            // little-endian UINT16 pair of word i is bswap(o[i]); coefficient index = byte offset / 2
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t cidx = 8u * (uint32_t)b + 2u * (uint32_t)i;
                const uint32_t total = (uint32_t)a.rings << 8;
                if (cidx >= total) continue;
                const uint32_t word = bswap32(o[i]);
                const int ring = (int)(cidx >> 8), pos = (int)(cidx & 255u);
                int32_t *dst = inst_out + (size_t)dest_ring(a, ring) * a.n + pos;
                *reinterpret_cast<int2 *>(dst) = make_int2(ring_coeff(word & 0xFFFFu, a), ring_coeff(word >> 16, a));
            }
        }
    }
}

// ---- AES-CTR-DRBG, throughput form (the library's default generator, so create_csprng's default too) ---------------
// One warp per instance left 31 lanes idle while lane 0 instantiated the DRBG (key schedule + entropy mix, about as much
// work as a lane's share of the blocks) and read a 1 KiB T-table with bank conflicts.  Now: k_gen_setup_aes instantiates
// one DRBG per THREAD into a scratch of round keys, and k_gen_rings_aes encrypts one counter-addressed block per thread
// over the bank-replicated tables of the CDF kernel (csprng.cuh: aes_rep_init, 128 KiB, one 1024-thread CTA per SM).
__global__ void __launch_bounds__(128) k_gen_setup_aes(const uint8_t *seeds, uint32_t seed_len, size_t count, uint32_t *keys)
{
    __shared__ AesTables aes;
    aes_tables_init(aes);
    __syncthreads();
    const size_t inst = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (inst >= count) return;
    uint32_t rk[60], counter;
    drbg_instantiate(aes, seeds + inst * seed_len, seed_len, rk, counter);
    uint32_t *k = keys + inst * 64;
    for (int i = 0; i < 60; i++) k[i] = rk[i];
    k[60] = counter;
}

constexpr int kGenAesCta = 1024;
constexpr size_t kGenAesTabBytes = 4 * 32768;

__global__ void __launch_bounds__(kGenAesCta) k_gen_rings_aes(GenArgs a, const uint32_t *__restrict__ keys)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *te0r = reinterpret_cast<uint32_t *>(smem_raw);
    aes_rep_init(te0r);
    __syncthreads();
    const uint32_t l4 = (threadIdx.x & 31) * 4;
    const uint32_t total_c = (uint32_t)a.rings << 8;                 // coefficients per instance (n = 256)
    const size_t nblocks = ((size_t)total_c + 7) / 8;                // 8 coefficients per 16-byte block
    const size_t total = a.count * nblocks;
    for (size_t item = blockIdx.x * (size_t)blockDim.x + threadIdx.x; item < total; item += (size_t)gridDim.x * blockDim.x) {
        const size_t inst = item / nblocks, b = item % nblocks;
        const uint32_t *k = keys + inst * 64;
        const uint32_t c = bswap32(__ldg(k + 60) + (uint32_t)b);
        uint32_t o[4];
        aes256_encrypt_rep(te0r, l4, k, c, c, c, c, o);
        int32_t *inst_out = a.out + inst * (size_t)a.rings * a.n;
        // ciphertext bytes in order are the little-endian UINT16 pairs of bswap(o[i]) (see k_gen_rings); the eight
        // coefficients of a block lie in one ring (256 is a multiple of 8): two 128-bit stores
        const uint32_t cidx = 8u * (uint32_t)b;
        const int ring = (int)(cidx >> 8), pos = (int)(cidx & 255u);
        int32_t *dst = inst_out + (size_t)dest_ring(a, ring) * a.n + pos;
        int32_t v[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t word = bswap32(o[i]);
            v[2 * i] = ring_coeff(word & 0xFFFFu, a);
            v[2 * i + 1] = ring_coeff(word >> 16, a);
        }
        if (cidx + 8 <= total_c) {
            *reinterpret_cast<int4 *>(dst) = make_int4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<int4 *>(dst + 4) = make_int4(v[4], v[5], v[6], v[7]);
        } else {
            for (uint32_t i = 0; cidx + i < total_c; i++) dst[i] = v[i];
        }
    }
}

}  // namespace

int launch_gen_rings(int prng_type, const uint8_t *seeds, size_t seed_len, size_t count, int32_t *out, int n, int k, int l,
                     int transpose, int32_t q, uint32_t q_bits, cudaStream_t st)
{
    if (count == 0) return SCGPU_OK;
    if (n != 256) { set_error("gen_rings: n = %d (256 only)", n); return SCGPU_ERR_UNSUPPORTED; }
    GenArgs a;
    a.seeds = seeds; a.seed_len = (uint32_t)seed_len; a.seed_period = 0x01000000u;      // create_csprng, module_lwe.c:921
    a.count = count; a.out = out; a.n = n; a.rings = k * l; a.k = k; a.l = l; a.transpose = transpose;
    a.q = q; a.mask = q_bits >= 32 ? 0xFFFFFFFFu : (1u << q_bits) - 1u;
    const size_t words = (size_t)k * l * n / 2;
    const size_t blocks = (words + 3) / 4;
    a.cache_blocks = (int)((blocks + 31) / 32);
    if (a.cache_blocks > 24) a.cache_blocks = 0;          // beyond 48 KB of cache per CTA the second pass recomputes
    const unsigned grid = (unsigned)((count + 3) / 4);
    if (prng_type == PRNG_CHACHA20) {
        const size_t smem = (size_t)4 * a.cache_blocks * 32 * 16;
        SCGPU_CUDA_CHECK(cudaFuncSetAttribute(k_gen_rings<PRNG_CHACHA20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        k_gen_rings<PRNG_CHACHA20><<<grid, 128, smem, st>>>(a);
    } else if (getenv("SCGPU_GEN_AES_WARP") && atoi(getenv("SCGPU_GEN_AES_WARP")) != 0) {
        k_gen_rings<PRNG_AES><<<grid, 128, 0, st>>>(a);              // one warp per instance (round-2 first version; A/B)
    } else {
        uint32_t *keys = nullptr;
        SCGPU_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&keys), count * 64 * sizeof(uint32_t), st));
        k_gen_setup_aes<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(seeds, (uint32_t)seed_len, count, keys);
        count_launch();
        const size_t items = count * (((size_t)k * l * n + 7) / 8);
        int sms = 148;
        { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
        size_t g2 = (items + kGenAesCta - 1) / kGenAesCta;
        if (g2 > (size_t)sms) g2 = (size_t)sms;
        SCGPU_CUDA_CHECK(cudaFuncSetAttribute(k_gen_rings_aes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGenAesTabBytes));
        k_gen_rings_aes<<<(unsigned)g2, kGenAesCta, kGenAesTabBytes, st>>>(a, keys);
        const cudaError_t fe = cudaFreeAsync(keys, st);
        if (fe != cudaSuccess) { set_error("cudaFreeAsync failed: %s", cudaGetErrorString(fe)); return SCGPU_ERR_CUDA; }
    }
    count_launch();
    SCGPU_CUDA_CHECK(cudaGetLastError());
    return SCGPU_OK;
}

}  // namespace scgpu
