// gauss.cu -- discrete Gaussian sampling kernels (src/utils/sampling of the reference).
//
//   k_stream_seq   one thread per PRNG stream, any sampler / vector mode / discard setting: the exact
//                  sequential semantics of sample_vector_32, shuffle_sample_vector_32,
//                  blinding_sample_vector_32 (sampling.c:68-228) over gaussian_cdf_sample_32/64
//                  (gaussian_cdf.c:536-553,639-659,661-677,760-774), gaussian_knuth_yao_sample
//                  (gaussian_knuth_yao.c:301-364) and bernoulli_sample_64 (gaussian_bernoulli.c:161-280).
//   k_cdf_aes      CDF + AES-CTR-DRBG, NORMAL_SAMPLES, no discard: sample j of a stream depends only on
//                  DRBG block j/2 (64-bit) or j/4 (32-bit), so one thread produces one block's samples;
//                  CDF table and AES tables in shared memory, coalesced 64/128-bit stores.
//   k_cdf_chacha   CDF + ChaCha20-CSPRNG, NORMAL_SAMPLES, no discard: one warp per stream; each lane
//                  encrypts a contiguous run of blocks, a warp XOR-scan rebuilds the running XOR the
//                  reference's encrypt-in-place framing creates, samples go out through shared memory.
#include "scgpu_internal.h"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include "csprng.cuh"
#include "gauss_plan.h"
#include "../../include/scgpu.h"

namespace scgpu {

namespace {

// ---- samplers on a sequential stream ----------------------------------------------------------------

// gaussian_cdf.c:536-553: fixed log2(size) probe steps, largest a with cdf[a] < x
template <typename T>
__device__ __forceinline__ uint32_t cdf_search(const T *cdf, uint32_t size, T x)
{
    uint32_t a = 0;
    if ((size & (size - 1)) == 0) {
        // power-of-two table (every table the reference builds): a + st < size always holds
        for (uint32_t st = size >> 1; st > 0; st >>= 1) {
            const uint32_t b = a + st;
            a = (cdf[b] < x) ? b : a;
        }
        return a;
    }
    for (uint32_t st = size >> 1; st > 0; st >>= 1) {
        uint32_t b = a + st;
        a = (b < size && cdf[b] < x) ? b : a;
    }
    return a;
}

// gaussian_cdf.c:112-190, 480-532: x = precision/64 successive prng_64 draws, word 0 first; fixed halving
// steps keep the largest a with "x >= l[a]" as compare_ge_prec evaluates it: retval = !x_lt_y | (equal & retval)
// folded from word 0 up.  An equal word yields !x_lt_y = 1, so the fold reduces to the TOP words' >= and the
// lower words never decide; reproduced literally.  Sign from bit 0 of word 0.
__device__ __forceinline__ int32_t sample_cdf_high(const GaussTablesDev &g, PrngStream &rng)
{
    const int nw = g.precision >> 6;
    uint64_t x[4];
    for (int i = 0; i < nw; i++) x[i] = rng.next64();
    uint32_t a = 0;
    for (uint32_t st = g.cdf_size >> 1; st > 0; st >>= 1) {
        const uint32_t b = a + st;
        if (b >= g.cdf_size) continue;
        const uint64_t *l = g.cdfh + (size_t)b * nw;
        unsigned ge = 1;
        for (int i = 0; i < nw; i++) {
            const unsigned lt = x[i] < l[i], eq = x[i] == l[i];
            ge = (!lt) | (eq & ge);
        }
        if (ge) a = b;
    }
    return (x[0] & 1) ? (int32_t)a : -(int32_t)a;
}

// The same result through a guide table: the fixed-step search above returns the largest a with cdf[a] < x
// (0 if none) whenever the table is sorted (checked on the host before a guide is built); guide[b] brackets that
// index for every x whose top kGuideBits bits are b, and a bisection of the bracket finishes the job.  For
// sigma = 215 the bracket is a single entry for 2 out of 3 draws; a warp needs ~3 probes instead of 12.
template <typename T>
__device__ __forceinline__ uint32_t cdf_search_guided(const T *cdf, const uint32_t *guide, T x)
{
    const uint32_t g = guide[(uint32_t)(x >> (sizeof(T) * 8 - kGuideBits))];
    uint32_t lo = g & 0xFFFFu, hi = g >> 16;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (cdf[mid] < x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ int32_t sample_cdf(const GaussTablesDev &g, PrngStream &rng)
{
    if (g.precision > 64) return sample_cdf_high(g, rng);
    if (g.precision == 64) {
        uint64_t x = rng.next64();
        uint32_t a = cdf_search<uint64_t>(g.cdf64, g.cdf_size, x);
        return (x & 1) ? (int32_t)a : -(int32_t)a;
    }
    uint32_t x = rng.next32();
    uint32_t a = cdf_search<uint32_t>(g.cdf32, g.cdf_size, x);
    return (x & 1) ? (int32_t)a : -(int32_t)a;
}

// Knuth-Yao walk (gaussian_knuth_yao.c:301-364).  The reference walks a byte-per-bit probability matrix
// with a running pointer: each row doubles the 32-bit distance, adds a random bit, then subtracts matrix
// entries until the distance goes negative (adding the number of entries scanned to the sample) or a whole
// row's worth of entries is consumed.  It does NOT re-align the pointer after a hit, and the distance keeps
// doubling and eventually wraps around, so later rows keep contributing.  With the sorted flat positions of
// the one-bits and a bitmap with prefix counts this is evaluated per row in O(1): "ones in [ptr, ptr+len)" is a
// difference of two ranks, "k-th one after ptr" one read of the flat table.
__device__ __forceinline__ uint32_t ky_rank(const GaussTablesDev &g, uint32_t pos)         // one-bits before position pos
{
    const uint32_t wd = pos >> 5;
    return __ldg(g.ky_rank + wd) + (uint32_t)__popc(__ldg(g.ky_bits + wd) & ((1u << (pos & 31)) - 1u));
}

__device__ __forceinline__ int32_t sample_ky(const GaussTablesDev &g, PrngStream &rng)
{
    const uint32_t cols = (uint32_t)g.ky_cols;
    for (;;) {
        uint32_t dist = 0;              // the reference's SINT32, with explicit wrap-around
        int32_t sample = 0;
        uint32_t ptr = 0;
        uint32_t rnd = rng.next32();
        for (int row = 0; row < g.ky_rows; row++) {
            dist = 2u * dist + (rnd & 1u);
            rnd >>= 1;
            if ((row & 0x1F) == 0x1F) rnd = rng.next32();
            uint32_t len = cols, col0 = 0;
            uint32_t i0 = ky_rank(g, ptr);
            if ((int32_t)dist < 0) {
                // first entry of the scan
                const uint32_t bit = (__ldg(g.ky_bits + (ptr >> 5)) >> (ptr & 31)) & 1u;
                dist -= bit;
                ptr += 1; i0 += bit;
                if ((int32_t)dist < 0) continue;        // hit at column 0: sample += 0
                len = cols - 1; col0 = 1;               // wrapped to INT_MAX: keep scanning this row
            }
            const uint32_t cnt = ky_rank(g, ptr + len) - i0;      // ones in the window
            if (dist < cnt) {
                const uint32_t f = __ldg(g.ky_flat + i0 + dist);  // the (dist+1)-th one takes the distance to -1
                sample += (int32_t)(col0 + (f - ptr));
                ptr = f + 1;
                dist = 0xFFFFFFFFu;
            } else {
                dist -= cnt;
                ptr += len;
            }
        }
        rnd = rng.next32();
        sample = sample % g.ky_bound;
        if (sample == 0 && (rnd & 1)) continue;
        return (rnd & 2) ? sample : -sample;
    }
}

// One candidate of the Bernoulli sampler (gaussian_bernoulli.c:161-246 for the candidate, :248-280 for zero / sign):
// false = rejected, draw again.  Kept as a single attempt so that the vector loop of k_stream_seq can let every lane
// move on to its next sample as soon as its own candidate is accepted (a warp that loops per sample until all 32
// lanes have accepted runs for the slowest lane of every sample: ~4x the mean number of candidates at sigma = 215).
__device__ __forceinline__ bool ber_try(const GaussTablesDev &g, PrngStream &rng, int32_t &out)
{
    const uint32_t val = rng.var(g.ber_maxlog);
    if (val >= (uint32_t)g.ber_maxval) return false;
    uint32_t accept_mask = 0, x = val * val;
    for (int j = 0; j < 8; j++) {
        for (int i = g.ber_entries; i--;) {
            uint32_t r = rng.var(8) & 0xFF;
            uint32_t tv = g.ber_tab[i * 8 + j];
            if (r < tv && ((accept_mask >> i) & 1) == 0) accept_mask |= (1u << i);
            if (r > tv && ((x >> i) & 1) == 1 && ((accept_mask >> i) & 1) == 0) return false;
        }
    }
    const uint32_t rnd = rng.var(2);
    if (val == 0) { if (rnd < 2) return false; out = 0; return true; }
    out = (rnd & 1) ? -(int32_t)val : (int32_t)val;
    return true;
}

__device__ __forceinline__ int32_t sample_ber(const GaussTablesDev &g, PrngStream &rng)
{
    int32_t s;
    while (!ber_try(g, rng, s)) { }
    return s;
}

// gaussian_knuth_yao_fast_sample (gaussian_knuth_yao_fast.c:303-368): a byte indexes the first table; a miss carries a
// distance into the second table (5 more random bits), a second miss walks the probability matrix from column 13 on,
// bottom row first, one random bit per column.  prng_8 / prng_bit are prng_var(8) / prng_var(1).
__device__ __forceinline__ int32_t sample_ky_fast(const GaussTablesDev &g, PrngStream &rng)
{
    int32_t sample = g.kf_lut1[rng.var(8) & 0xFF];
    if ((sample & 16) == 0) {
        sample &= 0xF;
        return rng.var(1) ? -sample : sample;
    }
    int32_t distance = sample & (int32_t)g.kf_d1mask;
    sample = g.kf_lut2[(int32_t)(rng.var(8) & 0x1F) + 32 * distance];
    if ((sample & 0x20) == 0) {
        sample &= 0x1F;
        return rng.var(1) ? -sample : sample;
    }
    distance = sample & (int32_t)g.kf_d2mask;
    for (int col = 13; col < g.kf_cols; col++) {
        distance = distance * 2 + (int32_t)rng.var(1);
        for (int row = g.kf_rows - 1; row >= 0; row--) {
            distance -= g.kf_pmat[row * g.kf_cols + col];
            if (distance < 0) return rng.var(1) ? -row : row;
        }
    }
    return 0;
}

// mw_bootstrap_sample (mw_bootstrap.c:243-259) over the CDF base sampler of sigma 16.  One sample is a fixed function
// of 8 base samples (the combiner network, left operand first, :85-99), one prng_64 and 29 more base samples:
//  * c = centre + x * scale, ci = floor(c);
//  * mw_flip_and_round (:213-241) scales the fractional part by 1UL << precision with precision = 64, i.e. by 1 on
//    x86-64 (the shift count wraps), so the scaled centre is 0: the coin flips consume random bits up to the first
//    one-bit and always end in mw_round(0);
//  * mw_round (:193-211): 29 steps of  sample = base_centre[center & 1] + base sample  (FLOAT + int, truncated),
//    minus one for an odd negative centre, center = center / 2 + sample.
__device__ __forceinline__ int32_t sample_mw(const GaussTablesDev &g, const MwParams &m, PrngStream &rng, float centre)
{
    int32_t b[8];
#pragma unroll 1
    for (int i = 0; i < 8; i++) b[i] = sample_cdf(g, rng);
    int32_t l0[4], l1[2];
    for (int i = 0; i < 4; i++) l0[i] = m.z[0][0] * b[2 * i] + m.z[0][1] * b[2 * i + 1];
    for (int i = 0; i < 2; i++) l1[i] = m.z[1][0] * l0[2 * i] + m.z[1][1] * l0[2 * i + 1];
    const int32_t x = m.z[2][0] * l1[0] + m.z[2][1] * l1[1];
    const double c = __dadd_rn((double)centre, __dmul_rn((double)x, m.scale));
    const double ci = floor(c);
    while (rng.next64() == 0) { }                       // the flips stop at the first one-bit
    long long center = 0;
#pragma unroll 1
    for (int it = 0; it < m.k; it++) {
        const int32_t smp = sample_cdf(g, rng);
        int32_t sample = (int32_t)__fadd_rn((center & 1) ? 0.5f : 0.0f, (float)smp);      // truncation toward zero
        if ((center & 1) && center < 0) sample--;
        center /= 2;
        center += sample;
    }
    return (int32_t)ci + (int32_t)center;
}

__device__ __forceinline__ int32_t draw(const GaussTablesDev &g, PrngStream &rng)
{
    if (g.sampler == SCGPU_SAMPLER_CDF) return sample_cdf(g, rng);
    if (g.sampler == SCGPU_SAMPLER_KNUTH_YAO) return sample_ky(g, rng);
    if (g.sampler == SCGPU_SAMPLER_KNUTH_YAO_FAST) return sample_ky_fast(g, rng);
    return sample_ber(g, rng);
}

// sampling.c:68-83
__device__ __forceinline__ uint32_t rand_range(PrngStream &rng, uint32_t x)
{
    uint32_t rem = 0xFFFFFFFFu % x;
    for (;;) {
        uint32_t y = rng.next32();
        if (y >= (0xFFFFFFFFu - rem)) continue;
        return y % x;
    }
}

struct SeqArgs {
    GaussTablesDev g;
    const uint8_t *seeds;
    PrngState *states;          // optional resumable states (drop-in prng_ctx_t); NULL = fresh streams from seeds
    uint32_t seed_len, seed_period;
    uint32_t prng_type;
    size_t nstreams, n, calls;
    int32_t centre;
    uint32_t thresh;
    int32_t *out;
    int mode;                   // 0 sampler vector calls, 1 raw words, 2 single get_sample, 3 instantiate only,
                                // 4 prng_mem (n = 64-byte blocks), 5 prng_reset, 6 refill the bit pool
    uint32_t *pool_mem;         // pooled states (drop-in prng_ctx_t): kPoolWords + kDrbgBufWords words per stream
    MwParams mw;                // mode 7: Micciancio-Walter bootstrap samples
};

__global__ void __launch_bounds__(128) k_stream_seq(SeqArgs a)
{
    __shared__ AesTables aes;
    aes_tables_init(aes);
    __syncthreads();
    const size_t sidx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (sidx >= a.nstreams) return;
    PrngStream rng;
    rng.aes = &aes;
    rng.seed = a.seeds + sidx * a.seed_len;
    if (a.pool_mem) {
        rng.pool = a.pool_mem + sidx * (size_t)(kPoolWords + kDrbgBufWords);
        rng.buf1k = rng.pool + kPoolWords;
    }
    if (a.states) rng.s = a.states[sidx];
    else { rng.s.pooled = 0; rng.s.ent_fresh = 0; rng.s.ent_avail = 0; }
    if (!a.states || !rng.s.instantiated) rng.init(a.prng_type, a.seed_len, a.seed_period);

    int32_t *v = a.out + sidx * a.n * a.calls;
    if (a.mode == 1) {
        for (size_t i = 0; i < a.n * a.calls; i++) v[i] = (int32_t)rng.next32();
    } else if (a.mode == 2) {
        v[0] = draw(a.g, rng);
    } else if (a.mode == 3) {
        // instantiate only: the state is written back below
    } else if (a.mode == 4) {
        // prng_mem, prng.c:1050-1105: eight generator draws per 64-byte block, each stored as a little-endian u64;
        // the bit pool is bypassed
        for (size_t i = 0; i < a.n * 8; i++) {
            uint32_t hi, lo;
            rng.draw64(hi, lo);
            v[2 * i] = (int32_t)lo; v[2 * i + 1] = (int32_t)hi;
        }
    } else if (a.mode == 5) {
        rng.reset_pooled();
    } else if (a.mode == 6) {
        rng.pool_refill();
    } else if (a.mode == 7) {
        // get_vector_32's bootstrap branch (sampling.c:560-573) / get_bootstrap_sample (:519-538)
        for (size_t i = 0; i < a.n * a.calls; i++) {
            const float centre = a.mw.centres ? a.mw.centres[sidx * a.n * a.calls + i] : a.mw.centre;
            int32_t r = sample_mw(a.g, a.mw, rng, centre);
            if (a.mw.clamp) r = r < a.mw.lim_lo ? a.mw.lim_lo : (r > a.mw.lim_hi ? a.mw.lim_hi : r);
            v[i] = r;
        }
    } else {
        for (size_t call = 0; call < a.calls; call++, v += a.n) {
            const size_t n = a.n;
            if (a.g.blinding == SCGPU_NORMAL_SAMPLES) {
                // sampling.c:211-228 (every sampler is driven the same way: sample() + centre)
                if (a.g.sampler == SCGPU_SAMPLER_BERNOULLI) {
                    // one candidate per trip; the lanes of a warp sit at different samples i
                    for (size_t i = 0; i < n;) {
                        int32_t smp;
                        if (!ber_try(a.g, rng, smp)) continue;
                        v[i] = smp + a.centre;
                        bool discard = a.thresh && rng.next32() < a.thresh;
                        if (!discard) i++;
                    }
                } else
                for (size_t i = 0; i < n;) {
                    v[i] = draw(a.g, rng) + a.centre;
                    bool discard = a.thresh && rng.next32() < a.thresh;
                    if (!discard) i++;
                }
            } else {
                // sampling.c:127-145: inside-out shuffle; v[0] is drawn without the centre
                v[0] = draw(a.g, rng);
                for (size_t i = 1; i < n;) {
                    uint32_t j = rand_range(rng, (uint32_t)i);
                    if (j != i) v[i] = v[j];
                    v[j] = draw(a.g, rng) + a.centre;
                    bool discard = a.thresh && rng.next32() < a.thresh;
                    if (!discard) i++;
                }
                if (a.g.blinding == SCGPU_BLINDING_SAMPLES) {    // sampling.c:170-191
                    for (size_t i = 0; i < n; i++) v[i] -= draw(a.g, rng);
                    // :176,182-188: swap v[i] with v[i & (n - 1)], a no-op only when n is a power of two
                    const uint32_t mask = (uint32_t)(n - 1);
                    for (size_t i = 0; i < n; i++) {
                        const size_t j = i & mask;
                        const int32_t t = v[i]; v[i] = v[j]; v[j] = t;
                    }
                }
            }
        }
    }
    if (a.states) a.states[sidx] = rng.s;
}

// ---- Bernoulli sampler: one lane per stream, ONE random draw per trip ---------------------------------------------
// bernoulli_sample_64 (gaussian_bernoulli.c:161-280) draws a 12-bit candidate, then compares up to 8 x 23 random
// bytes with its table and rejects at the first byte that decides against the candidate.  An accepted candidate costs
// 184 byte draws, a rejected one a handful, and ~15 candidates are rejected per accepted one (sigma = 215): with one
// sample() call per lane per trip a warp always waits for its one lane that is inside an accepted candidate
// (3.4e7 samples/s in round 1).  Here every lane advances its own state machine by one draw per trip -- candidate,
// byte (j, i), zero / sign, discard -- so all lanes make progress all the time, and the generator is stepped for the
// whole warp at once: each lane keeps a 16-word FIFO in shared memory, and when any lane runs dry every lane with room
// draws its next generator block, so the ChaCha20 / AES blocks of the 32 streams are computed side by side instead of
// one lane at a time (with a 4-word FIFO only 17 of 32 lanes took part in an average refill, profiles/ber_lanes_r2_ncu.json).
// Fresh streams, NORMAL_SAMPLES (any discard setting); everything else stays on k_stream_seq.
constexpr int kBerFifo = 16;            // words per lane; refills add one generator block (4 words) at a time
__global__ void __launch_bounds__(128) k_ber_lanes(SeqArgs a)
{
    __shared__ AesTables aes;
    __shared__ uint32_t fifo[kBerFifo][128];            // lane-interleaved: conflict-free
    aes_tables_init(aes);
    __syncthreads();
    const size_t sidx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t total = a.n * a.calls;
    bool done = sidx >= a.nstreams || total == 0;
    PrngStream rng;
    rng.aes = &aes;
    rng.seed = a.seeds + (done ? 0 : sidx) * a.seed_len;
    rng.s.pooled = 0; rng.s.ent_fresh = 0; rng.s.ent_avail = 0;
    if (!done) rng.init(a.prng_type, a.seed_len, a.seed_period);
    int32_t *v = a.out + (done ? 0 : sidx) * total;
    uint32_t (*q)[128] = fifo;
    const int tid = threadIdx.x;
    int head = 0, cnt = 0;                              // ring buffer state of this lane
    uint32_t var_buf = 0, var_bits = 0;                 // prng_var's bit buffer (prng.c:1017-1048)
    auto pop = [&]() { const uint32_t x = q[head][tid]; head = (head + 1) & (kBerFifo - 1); cnt--; return x; };
    auto var = [&](uint32_t n) {                         // n < 32 here
        uint32_t ret = var_buf;
        if (var_bits < n) {
            const uint32_t need = n - var_bits;
            ret <<= need;
            var_buf = pop();
            ret |= var_buf & ((1u << need) - 1u);
            var_buf >>= need;
            var_bits = 32 - need;
        } else {
            var_buf >>= n;
            var_bits -= n;
        }
        return ret & ((1u << n) - 1u);
    };
    const int entries = a.g.ber_entries;
    int phase = 0, i = 0, j = 0;
    uint32_t val = 0, x = 0, accept_mask = 0;
    size_t idx = 0;
    while (__any_sync(0xFFFFFFFFu, !done)) {
        // Refill event: some lane is dry.  Every lane with room for a whole generator block (two 64-bit draws) takes
        // one, so the 32 ChaCha20 / AES blocks run side by side; the deep FIFO averages out the differences in the
        // lanes' consumption so that (almost) every lane takes part in (almost) every event.
        if (__any_sync(0xFFFFFFFFu, !done && cnt == 0)) {
            if (!done && cnt <= kBerFifo - 4) {
#pragma unroll 1
                for (int d = 0; d < 2; d++) {
                    uint32_t hi, lo;
                    rng.draw64(hi, lo);
                    q[(head + cnt) & (kBerFifo - 1)][tid] = hi;
                    q[(head + cnt + 1) & (kBerFifo - 1)][tid] = lo;
                    cnt += 2;
                }
            }
        }
        if (done) continue;
        if (phase == 0) {                               // candidate (gaussian_bernoulli.c:167-175)
            val = var((uint32_t)a.g.ber_maxlog);
            if (val < (uint32_t)a.g.ber_maxval) { x = val * val; accept_mask = 0; j = 0; i = entries - 1; phase = 1; }
        } else if (phase == 1) {                        // one table byte (:177-243)
            const uint32_t r = var(8);
            const uint32_t tv = __ldg(a.g.ber_tab + i * 8 + j);
            const uint32_t undecided = ((accept_mask >> i) & 1u) ^ 1u;
            if (r < tv && undecided) accept_mask |= 1u << i;
            if (r > tv && ((x >> i) & 1u) && undecided) phase = 0;                 // rejected: next candidate
            else if (i-- == 0) { i = entries - 1; if (++j == 8) phase = 2; }
        } else if (phase == 2) {                        // zero with probability 1/2, sign (:248-280)
            const uint32_t rnd = var(2);
            if (val == 0 && rnd < 2) phase = 0;
            else {
                const int32_t smp = val == 0 ? 0 : ((rnd & 1) ? -(int32_t)val : (int32_t)val);
                v[idx] = smp + a.centre;
                if (a.thresh) phase = 3;
                else { phase = 0; done = ++idx >= total; }
            }
        } else {                                        // discard_sample (sampling.c:95-105): a whole prng_32 word
            if (!(pop() < a.thresh)) done = ++idx >= total;
            phase = 0;
        }
    }
}

// ---- fast path 1: CDF over the AES-CTR-DRBG ---------------------------------------------------------------
constexpr int kKsCache = 8;             // ChaCha20 blocks per lane kept between the two passes of k_cdf_chacha

struct FastArgs {
    GaussTablesDev g;
    const uint8_t *seeds;
    uint32_t seed_len, seed_period;
    size_t nstreams, per_stream;        // per_stream = calls * n samples
    int32_t centre;
    int32_t *out;
    uint32_t *keys;                      // [nstreams][64]: 60 round-key words + initial counter (16-byte aligned rows)
    unsigned long long *ctr;             // work counter of k_cdf_chacha (streams beyond the first grid-full), or nullptr
};

// one thread per stream: DRBG instantiation (zero-key block encryptions, entropy mix, key schedule)
__global__ void __launch_bounds__(128) k_drbg_setup(FastArgs a)
{
    __shared__ AesTables aes;
    aes_tables_init(aes);
    __syncthreads();
    const size_t sidx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (sidx >= a.nstreams) return;
    uint32_t rk[60], counter;
    drbg_instantiate(aes, a.seeds + sidx * a.seed_len, a.seed_len, rk, counter);
    uint32_t *k = a.keys + sidx * 64;
    for (int i = 0; i < 60; i++) k[i] = rk[i];
    k[60] = counter;
}

constexpr int kAesCta = 1024;            // one CTA per SM: 128 KiB of replicated AES tables are shared by 32 warps
constexpr size_t kAesTabBytes = 4 * 32768;

template <int PREC>
__global__ void __launch_bounds__(kAesCta) k_cdf_aes(FastArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *te0r = reinterpret_cast<uint32_t *>(smem_raw);                 // 4 tables x 256 x 32 words, one copy per bank
    uint64_t *cdf64 = reinterpret_cast<uint64_t *>(smem_raw + kAesTabBytes);
    uint32_t *cdf32 = reinterpret_cast<uint32_t *>(smem_raw + kAesTabBytes);
    aes_rep_init(te0r);
    if (PREC == 64) for (uint32_t i = threadIdx.x; i < a.g.cdf_size; i += blockDim.x) cdf64[i] = a.g.cdf64[i];
    else            for (uint32_t i = threadIdx.x; i < a.g.cdf_size; i += blockDim.x) cdf32[i] = a.g.cdf32[i];
    uint32_t *guide = reinterpret_cast<uint32_t *>(smem_raw + kAesTabBytes + a.g.cdf_size * (PREC == 64 ? 8 : 4));
    const bool guided = a.g.cdf_guide != nullptr;
    if (guided) for (uint32_t i = threadIdx.x; i < (1u << kGuideBits); i += blockDim.x) guide[i] = a.g.cdf_guide[i];
    __syncthreads();
    const uint32_t l4 = (threadIdx.x & 31) * 4;
    constexpr int SPB = PREC == 64 ? 2 : 4;                   // samples per 16-byte DRBG block
    const size_t blocks_per_stream = (a.per_stream + SPB - 1) / SPB;
    const size_t total = a.nstreams * blocks_per_stream;
#pragma unroll 2
    for (size_t item = blockIdx.x * (size_t)blockDim.x + threadIdx.x; item < total; item += (size_t)gridDim.x * blockDim.x) {
        const size_t sidx = item / blocks_per_stream, blk = item % blocks_per_stream;
        const uint32_t *k = a.keys + sidx * 64;
        uint32_t w[4];
        {
            // drbg_block_words() with the replicated table
            const uint32_t cnt = bswap32(__ldg(k + 60) + (uint32_t)blk);
            uint32_t o[4];
            aes256_encrypt_rep(te0r, l4, k, cnt, cnt, cnt, cnt, o);
            w[0] = bswap32(o[1]); w[1] = bswap32(o[0]); w[2] = bswap32(o[3]); w[3] = bswap32(o[2]);
        }
        int32_t res[4];
        if (PREC == 64) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                uint64_t x = ((uint64_t)w[2 * i] << 32) | w[2 * i + 1];
                uint32_t s = guided ? cdf_search_guided<uint64_t>(cdf64, guide, x) : cdf_search<uint64_t>(cdf64, a.g.cdf_size, x);
                res[i] = ((x & 1) ? (int32_t)s : -(int32_t)s) + a.centre;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint32_t s = guided ? cdf_search_guided<uint32_t>(cdf32, guide, w[i]) : cdf_search<uint32_t>(cdf32, a.g.cdf_size, w[i]);
                res[i] = ((w[i] & 1) ? (int32_t)s : -(int32_t)s) + a.centre;
            }
        }
        int32_t *o = a.out + sidx * a.per_stream + blk * SPB;
        const size_t left = a.per_stream - blk * SPB;
        if (left >= SPB && ((reinterpret_cast<uintptr_t>(o) & (SPB * 4 - 1)) == 0)) {
            if (PREC == 64) *reinterpret_cast<int2 *>(o) = make_int2(res[0], res[1]);
            else            *reinterpret_cast<int4 *>(o) = make_int4(res[0], res[1], res[2], res[3]);
        } else {
            for (size_t i = 0; i < SPB && i < left; i++) o[i] = res[i];
        }
    }
}

// ---- fast path 2: CDF over the ChaCha20-CSPRNG ----------------------------------------------------------------
// Word i of a fresh stream: 0 for i < 3, else word (i-3)%4 of D[(i-3)/4], D[j] = XOR_{k<=j} first16(block k),
// byte-swapped.  One warp per stream; lane l owns blocks [l*C, (l+1)*C).
template <int PREC>
__global__ void __launch_bounds__(256) k_cdf_chacha(FastArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *cdf64 = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *cdf32 = reinterpret_cast<uint32_t *>(smem_raw);
    if (PREC == 64) for (uint32_t i = threadIdx.x; i < a.g.cdf_size; i += blockDim.x) cdf64[i] = a.g.cdf64[i];
    else            for (uint32_t i = threadIdx.x; i < a.g.cdf_size; i += blockDim.x) cdf32[i] = a.g.cdf32[i];
    uint32_t *guide = reinterpret_cast<uint32_t *>(smem_raw + a.g.cdf_size * (PREC == 64 ? 8 : 4) + (size_t)8 * kKsCache * 32 * 16);
    const bool guided = a.g.cdf_guide != nullptr;
    if (guided) for (uint32_t i = threadIdx.x; i < (1u << kGuideBits); i += blockDim.x) guide[i] = a.g.cdf_guide[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    // keystream cache: the first pass keeps each block's 16 bytes so the second pass does not run ChaCha20
    // again (kKsCache blocks per lane, lane-interleaved uint4 -> conflict-free 128-bit accesses)
    uint4 *kscache = reinterpret_cast<uint4 *>(smem_raw + a.g.cdf_size * (PREC == 64 ? 8 : 4)) + (threadIdx.x >> 5) * (kKsCache * 32) + lane;
    constexpr int WPS = PREC == 64 ? 2 : 1;                                   // words per sample
    const size_t words = a.per_stream * WPS;
    const size_t nblocks = words > 3 ? (words - 3 + 3) / 4 : 0;                // blocks D[0..nblocks)
    const size_t C = (nblocks + 31) / 32;
    // streams beyond the first grid-full are claimed from a global counter: the warps of a scheduler do not advance
    // at the same rate, and with a static stride the favoured ones retire early (24 resident warps, 20 alive on
    // average in profiles/gauss_chacha_r03b_ncu.json)
    for (size_t sidx = warp; sidx < a.nstreams;) {
        unsigned long long claim = 0;
        if (a.ctr != nullptr && lane == 0) claim = atomicAdd(a.ctr, 1ull) + nwarps;
        const uint8_t *seed = a.seeds + sidx * a.seed_len;
        // key / iv: first 40 entropy bytes (ring buffer)
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
        // pass 1: XOR of this lane's blocks
        uint32_t acc[4] = {0, 0, 0, 0};
        const size_t b0 = (size_t)lane * C, b1 = (b0 + C < nblocks) ? b0 + C : nblocks;
        const bool cached = C <= (size_t)kKsCache;
        for (size_t b = b0; b < b1; b++) {
            uint32_t ks[4];
            chacha20_first16(key, (uint32_t)b, (uint32_t)(b >> 32), iv[0], iv[1], ks);
            acc[0] ^= ks[0]; acc[1] ^= ks[1]; acc[2] ^= ks[2]; acc[3] ^= ks[3];
            if (cached) kscache[(b - b0) * 32] = make_uint4(ks[0], ks[1], ks[2], ks[3]);
        }
        // exclusive XOR scan over lanes
        uint32_t pre[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t v = acc[i];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, off);
                if (lane >= off) v ^= o;
            }
            pre[i] = v ^ acc[i];
        }
        // pass 2: regenerate the lane's blocks, emit its words -> samples
        int32_t *orow = a.out + sidx * a.per_stream;
        if (lane == 0) {
            // words 0..2 are zero: samples built only from them
            if (PREC == 64) { if (a.per_stream > 0) orow[0] = a.centre + 0; }     // x = 0 -> a = 0, sign -> -0
            else for (size_t i = 0; i < 3 && i < a.per_stream; i++) orow[i] = a.centre;
        }
        uint32_t run[4] = {pre[0], pre[1], pre[2], pre[3]};
        // the 64-bit sample that straddles two blocks needs the last word of the previous block:
        // word index of D[j] word k is 3 + 4j + k; sample s (64-bit) uses words 2s, 2s+1.
        uint32_t carry = 0;                                                        // word 4*b0+2 (= D[b0-1] word 3)
        if (b0 > 0 && b0 <= nblocks) carry = bswap32(pre[3]);
        for (size_t b = b0; b < b1; b++) {
            uint32_t ks[4];
            if (cached) {
                const uint4 v = kscache[(b - b0) * 32];
                ks[0] = v.x; ks[1] = v.y; ks[2] = v.z; ks[3] = v.w;
            } else {
                chacha20_first16(key, (uint32_t)b, (uint32_t)(b >> 32), iv[0], iv[1], ks);
            }
            run[0] ^= ks[0]; run[1] ^= ks[1]; run[2] ^= ks[2]; run[3] ^= ks[3];
            const uint32_t w0 = bswap32(run[0]), w1 = bswap32(run[1]), w2 = bswap32(run[2]), w3 = bswap32(run[3]);
            const size_t wi = 3 + 4 * b;                                           // index of w0
            if (PREC == 64) {
                // words wi-1 (carry), wi .. wi+3: samples (wi-1)/2 = 2b+1 and 2b+2
                const size_t s0 = 2 * b + 1;
                uint64_t x0 = ((uint64_t)carry << 32) | w0;
                uint64_t x1 = ((uint64_t)w1 << 32) | w2;
                if (s0 < a.per_stream) {
                    uint32_t s = guided ? cdf_search_guided<uint64_t>(cdf64, guide, x0) : cdf_search<uint64_t>(cdf64, a.g.cdf_size, x0);
                    orow[s0] = ((x0 & 1) ? (int32_t)s : -(int32_t)s) + a.centre;
                }
                if (s0 + 1 < a.per_stream) {
                    uint32_t s = guided ? cdf_search_guided<uint64_t>(cdf64, guide, x1) : cdf_search<uint64_t>(cdf64, a.g.cdf_size, x1);
                    orow[s0 + 1] = ((x1 & 1) ? (int32_t)s : -(int32_t)s) + a.centre;
                }
                carry = w3;
            } else {
                const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (wi + k < a.per_stream) {
                        uint32_t s = guided ? cdf_search_guided<uint32_t>(cdf32, guide, w[k]) : cdf_search<uint32_t>(cdf32, a.g.cdf_size, w[k]);
                        orow[wi + k] = ((w[k] & 1) ? (int32_t)s : -(int32_t)s) + a.centre;
                    }
                }
            }
        }
        if (a.ctr != nullptr) {
            const uint32_t lo = __shfl_sync(0xFFFFFFFFu, (uint32_t)claim, 0), hi = __shfl_sync(0xFFFFFFFFu, (uint32_t)(claim >> 32), 0);
            sidx = (size_t)(((unsigned long long)hi << 32) | lo);
        } else {
            sidx += nwarps;
        }
    }
}

unsigned cap_grid(size_t want, int sms, int per_sm)
{
    size_t cap = (size_t)sms * per_sm;
    if (want > cap) want = cap;
    if (want == 0) want = 1;
    return (unsigned)want;
}

}  // namespace

int launch_gauss_seq(const GaussTablesDev &g, int prng_type, const uint8_t *seeds, size_t seed_len,
                     uint32_t seed_period, PrngState *states, size_t nstreams, size_t n, size_t calls,
                     int32_t centre, uint32_t discard, int32_t *out, int mode, cudaStream_t st, uint32_t *pool_mem,
                     const MwParams *mw)
{
    if (nstreams == 0 || (n * calls == 0 && mode != 2 && mode != 3 && mode != 5 && mode != 6)) return SCGPU_OK;
    SeqArgs a;
    a.g = g; a.seeds = seeds; a.states = states; a.seed_len = (uint32_t)seed_len; a.seed_period = seed_period;
    a.prng_type = (uint32_t)prng_type; a.nstreams = nstreams; a.n = n; a.calls = calls; a.centre = centre;
    a.thresh = discard == 2 ? 1u << 28 : discard == 4 ? 1u << 30 : discard == 6 ? 1u << 31 : 0;   // sampling.c:85-92
    a.out = out; a.mode = mode; a.pool_mem = pool_mem;
    memset(&a.mw, 0, sizeof(a.mw));
    if (mw) a.mw = *mw;
    const unsigned grid = (unsigned)((nstreams + 127) / 128);
    if (mode == 0 && states == nullptr && g.sampler == SCGPU_SAMPLER_BERNOULLI && g.blinding == SCGPU_NORMAL_SAMPLES && g.ber_maxlog < 32)
        k_ber_lanes<<<grid, 128, 0, st>>>(a);
    else
        k_stream_seq<<<grid, 128, 0, st>>>(a);
    count_launch();
    SCGPU_CUDA_CHECK(cudaGetLastError());
    return SCGPU_OK;
}

// 1 (the default): the throughput kernels run the reference's fixed probe sequence (log2(size) probes for every
// draw, gaussian_cdf.c:536-553) -- the number and order of table lookups does not depend on the draw.  0
// (scgpu_set_fixed_probe_search(0) or SCGPU_GUIDED_SEARCH=1): guide-bracketed bisection, ~3 probes per draw but a
// data-dependent trip count.  Same samples either way.
static std::atomic<int> g_fixed_probe_search{-1};
static int fixed_probe_search()
{
    int m = g_fixed_probe_search.load(std::memory_order_relaxed);
    if (m < 0) {
        const char *e = getenv("SCGPU_GUIDED_SEARCH");
        m = (e && atoi(e) != 0) ? 0 : 1;
        const char *f = getenv("SCGPU_FIXED_PROBE_SEARCH");
        if (f) m = atoi(f) != 0 ? 1 : 0;
        int expect = -1;
        if (!g_fixed_probe_search.compare_exchange_strong(expect, m)) m = expect;
    }
    return m;
}
int set_fixed_probe_search(int on)
{
    const int old = fixed_probe_search();
    g_fixed_probe_search.store(on ? 1 : 0);
    return old;
}

int launch_gauss_fast(const GaussTablesDev &g, int prng_type, const uint8_t *seeds, size_t seed_len,
                      uint32_t seed_period, size_t nstreams, size_t per_stream, int32_t centre, int32_t *out,
                      uint32_t *key_scratch, int sm_count, cudaStream_t st)
{
    if (nstreams == 0 || per_stream == 0) return SCGPU_OK;
    FastArgs a;
    a.g = g;
    if (fixed_probe_search()) a.g.cdf_guide = nullptr;
    a.seeds = seeds; a.seed_len = (uint32_t)seed_len; a.seed_period = seed_period;
    a.nstreams = nstreams; a.per_stream = per_stream; a.centre = centre; a.out = out; a.keys = key_scratch;
    a.ctr = nullptr;
    const size_t table_bytes = (size_t)g.cdf_size * (g.precision == 64 ? 8 : 4);
    if (prng_type == PRNG_AES) {
        k_drbg_setup<<<(unsigned)((nstreams + 127) / 128), 128, 0, st>>>(a);
        count_launch();
        const size_t smem = kAesTabBytes + table_bytes + (sizeof(uint32_t) << kGuideBits);
        const int spb = g.precision == 64 ? 2 : 4;
        const size_t items = nstreams * ((per_stream + spb - 1) / spb);
        if (smem > 220 * 1024) { set_error("CDF table of %zu bytes does not fit beside the AES tables", table_bytes); return SCGPU_ERR_UNSUPPORTED; }
        const unsigned grid = cap_grid((items + kAesCta - 1) / kAesCta, sm_count, 1);
        if (g.precision == 64) {
            SCGPU_CUDA_CHECK(cudaFuncSetAttribute(k_cdf_aes<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_cdf_aes<64><<<grid, kAesCta, smem, st>>>(a);
        } else {
            SCGPU_CUDA_CHECK(cudaFuncSetAttribute(k_cdf_aes<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_cdf_aes<32><<<grid, kAesCta, smem, st>>>(a);
        }
    } else {
        const unsigned grid = cap_grid((nstreams + 7) / 8, sm_count, 3);
        if (nstreams > (size_t)grid * 8) { const int e = next_work_counter(st, &a.ctr); if (e != SCGPU_OK) return e; }
        const size_t cc_smem = table_bytes + (size_t)8 * kKsCache * 32 * 16 + (sizeof(uint32_t) << kGuideBits);   // table, 8 warps of cache, guide
        if (g.precision == 64) {
            SCGPU_CUDA_CHECK(cudaFuncSetAttribute(k_cdf_chacha<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cc_smem));
            k_cdf_chacha<64><<<grid, 256, cc_smem, st>>>(a);
        } else {
            SCGPU_CUDA_CHECK(cudaFuncSetAttribute(k_cdf_chacha<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cc_smem));
            k_cdf_chacha<32><<<grid, 256, cc_smem, st>>>(a);
        }
    }
    count_launch();
    SCGPU_CUDA_CHECK(cudaGetLastError());
    return SCGPU_OK;
}

}  // namespace scgpu
