// gauss_plan.h -- device-side view of one sampler's tables (built on the host by host_sampling.cu).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cuda_runtime.h>

namespace scgpu {

constexpr int kGuideBits = 10;

struct PrngState;

struct GaussTablesDev {
    int sampler;            // random_sampling_e
    int precision;          // 32 / 64, or 128 / 192 / 256 with a caller-built table (cdfh)
    int blinding;           // sample_blinding_e
    // CDF (gaussian_cdf.c:555-610, 679-728)
    const uint64_t *cdf64;
    const uint32_t *cdf32;
    uint32_t cdf_size;
    // optional guide for the throughput kernels (32 / 64-bit, sorted tables only): entry b = lo | hi << 16, the
    // search results of the smallest and the largest x whose top kGuideBits bits are b
    const uint32_t *cdf_guide;
    // high-precision CDF (gaussian_cdf.c:112-532): cdf_size entries of precision/64 words, word 0 least
    // significant (the reference's u128_t / u192_t / u256_t arrays on a 64-bit-limb build)
    const uint64_t *cdfh;
    // Knuth-Yao (gaussian_knuth_yao.c:81-189): sorted flat indices (row * cols + col) of the one-bits of
    // the row-major probability matrix
    int ky_rows, ky_cols, ky_bound;
    uint32_t ky_nones;
    const uint32_t *ky_flat;
    // the same matrix as a bitmap (bit p of word p / 32 = entry p) with the number of one-bits before each word:
    // rank(p) = ky_rank[p / 32] + popc(low bits), so a whole row of the walk costs two table reads instead of a scan
    const uint32_t *ky_bits, *ky_rank;
    // Knuth-Yao "fast" (gaussian_knuth_yao_fast.c:303-368): two byte look-up tables and a small byte-per-bit matrix,
    // all handed over by the caller (they are source-embedded constants of the reference)
    const uint8_t *kf_lut1, *kf_lut2, *kf_pmat;
    int kf_rows, kf_cols;
    uint32_t kf_d1mask, kf_d2mask;
    // Bernoulli (gaussian_bernoulli.c:40-103): entries x 8 bytes, most significant byte first
    const uint8_t *ber_tab;
    int ber_entries, ber_maxval, ber_maxlog;
};

// Micciancio-Walter bootstrap (mw_bootstrap.c) over a base sampler of sigma 16: constants of one launch
struct MwParams {
    int32_t z[3][2];        // combiner weights of the three levels (mw_bootstrap_create, :112-160)
    int32_t k;              // rounding steps (29)
    double scale;           // sqrt((sigma^2 - rr_sigma2) / wide_sigma2), computed on the host in long double
    float centre;           // used when centres == nullptr
    const float *centres;   // optional per-sample centres, [nstreams][n]
    int32_t clamp, lim_lo, lim_hi;      // get_vector_32's integer limits (sampling.c:560-573)
};

int launch_gauss_seq(const GaussTablesDev &g, int prng_type, const uint8_t *seeds, size_t seed_len,
                     uint32_t seed_period, PrngState *states, size_t nstreams, size_t n, size_t calls,
                     int32_t centre, uint32_t discard, int32_t *out, int mode, cudaStream_t st,
                     uint32_t *pool_mem = nullptr, const MwParams *mw = nullptr);
int set_fixed_probe_search(int on);
int launch_gauss_fast(const GaussTablesDev &g, int prng_type, const uint8_t *seeds, size_t seed_len,
                      uint32_t seed_period, size_t nstreams, size_t per_stream, int32_t centre, int32_t *out,
                      uint32_t *key_scratch, int sm_count, cudaStream_t st);

}  // namespace scgpu
