// host_sampling.cu -- host side of the sampler path: table construction, the batch C-ABI and the
// reference's drop-in create_sampler()/get_vector_32()/prng_*() surface.
//
// Tables are built on the HOST with the reference's own formulas and C library calls (x87 long double
// expl for the CDF and Knuth-Yao tables, float expf/powf/log2f for Bernoulli), exactly as
// gaussian_cdf_create_64 (gaussian_cdf.c:555-610), gaussian_cdf_create_32 (:679-728),
// create_knuth_yao_table_32/64 (gaussian_knuth_yao.c:81-124) and gen_ber_table_64
// (gaussian_bernoulli.c:61-103) do, then uploaded; they are never recomputed on the device
// (SURVEY.md 7 "hard parts").  All random words and all samples are produced by kernels in gauss.cu.
#include "scgpu_internal.h"
#include "csprng.cuh"
#include "gauss_plan.h"
#include "../../include/scgpu.h"
#include "../../include/scgpu_dropin.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

using namespace scgpu;

struct scgpu_gauss_plan {
    GaussTablesDev t;
    int device, sm_count;
    void *d_cdf = nullptr;
    uint32_t *d_flat = nullptr, *d_kybits = nullptr, *d_kyrank = nullptr;
    uint8_t *d_ber = nullptr, *d_kf = nullptr;
    uint32_t *d_guide = nullptr;
    std::mutex mu;
    // Micciancio-Walter network constants (mw_bootstrap_create, mw_bootstrap.c:112-175), long double as in the reference
    bool mw = false;
    int32_t mw_z[3][2] = {};
    int32_t mw_k = 0;
    float mw_tail = 0;
    long double mw_inv_wide_sigma2 = 0, mw_rr_sigma2 = 0;
    cudaStream_t hstreams[3] = {nullptr, nullptr, nullptr};     // pipeline of the *_host entry point
};

namespace {

// sc_math.c:447-452
size_t ceil_log2_sz(size_t x)
{
    size_t l = 0;
    while ((x >> (l + 1)) != 0) l++;
    if (x & (x - 1)) l++;
    return l;
}

// sc_math.c:1066-1100
uint64_t bin_expansion(double x, int nbits)
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

// the same loop with up to 128 result bits (get_binary_expansion_fraction_32/64/128, sc_math.c:1047-1100)
unsigned __int128 bin_expansion128(double x, int nbits)
{
    double val = 0, step = 0.5f;
    unsigned __int128 res = 0;
    for (int i = 0; i < nbits; i++) {
        res <<= 1;
        if ((val + step) < x) { val += step; res |= 1; }
        step = step / 2;
    }
    return res;
}

#define L_2_SQRTPI 1.128379167095512573896158903121545172L
#define L_SQRT1_2  0.707106781186547524400844362104849039L

std::vector<uint64_t> build_cdf64(int blinding, float tail, float sigma)
{
    int bits = (int)ceil_log2_sz((size_t)(tail * sigma));
    int size = 1 << bits;
    std::vector<uint64_t> cdf((size_t)size);
    if (blinding == SCGPU_BLINDING_SAMPLES) sigma *= L_SQRT1_2;
    long double d = L_2_SQRTPI * L_SQRT1_2 * 18446744073709551616.0L / sigma;
    long double e = -0.5L / (sigma * sigma);
    long double s = 0.5L * d;
    int i;
    cdf[0] = 0;
    for (i = 1; i < size - 1; i++) {
        cdf[i] = (uint64_t)s;
        if (cdf[i] == 0) break;
        s += d * expl(e * ((long double)(i * i)));
    }
    for (; i < size; i++) cdf[i] = 0xFFFFFFFFFFFFFFFFULL;
    return cdf;
}

std::vector<uint32_t> build_cdf32(int blinding, float tail, float sigma)
{
    int bits = (int)ceil_log2_sz((size_t)(tail * sigma));
    int size = 1 << bits;
    std::vector<uint32_t> cdf((size_t)size);
    if (blinding == SCGPU_BLINDING_SAMPLES) sigma *= M_SQRT1_2;
    float d = M_2_SQRTPI * M_SQRT1_2 * 4294967296.0 / sigma;
    float e = -0.5L / (sigma * sigma);
    float s = 0.5L * d;
    int i;
    cdf[0] = 0;
    for (i = 1; i < size - 1; i++) {
        cdf[i] = (uint32_t)s;
        if (cdf[i] == 0) break;
        s += d * expl(e * ((float)(i * i)));
    }
    for (; i < size; i++) cdf[i] = 0xFFFFFFFFu;
    return cdf;
}

template <typename T>
int upload(T **dst, const std::vector<T> &src)
{
    SCGPU_CUDA_CHECK(cudaMalloc(dst, sizeof(T) * (src.size() ? src.size() : 1)));
    SCGPU_CUDA_CHECK(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return SCGPU_OK;
}

template <typename T>
int ensure_cap(T **buf, size_t *cap, size_t want)
{
    if (*cap >= want) return SCGPU_OK;
    cudaFree(*buf);
    *buf = nullptr; *cap = 0;
    SCGPU_CUDA_CHECK(cudaMalloc(buf, want * sizeof(T)));
    *cap = want;
    return SCGPU_OK;
}

const uint32_t kDefaultSeedPeriod = 0x00100000;     // safecrypto.c:379

// Can the position-addressable kernels serve this request?  They assume no reseed inside a stream.
// Guide of the throughput kernels (gauss.cu: cdf_search_guided).  Built only for a sorted table: then the
// reference's fixed-step search (gaussian_cdf.c:536-553) IS the predecessor search the guided bisection performs.
template <typename T>
std::vector<uint32_t> build_guide(const std::vector<T> &cdf)
{
    std::vector<uint32_t> g;
    const size_t size = cdf.size();
    if (size < 2 || size > 65536) return g;
    for (size_t i = 1; i < size; i++) if (cdf[i] < cdf[i - 1]) return g;
    auto search = [&](T x) {
        uint32_t a = 0;
        for (uint32_t st = (uint32_t)size >> 1; st > 0; st >>= 1) {
            uint32_t b = a + st;
            if (b < size && cdf[b] < x) a = b;
        }
        return a;
    };
    const int shift = (int)sizeof(T) * 8 - kGuideBits;
    g.resize((size_t)1 << kGuideBits);
    for (uint32_t b = 0; b < g.size(); b++) {
        const T xmin = (T)((T)b << shift);
        const T xmax = (T)(xmin | (T)(((T)1 << shift) - 1));
        g[b] = search(xmin) | (search(xmax) << 16);
    }
    return g;
}

bool fast_path_ok(const GaussTablesDev &t, int prng_type, size_t per_stream, uint32_t discard)
{
    if (t.sampler != SCGPU_SAMPLER_CDF || t.blinding != SCGPU_NORMAL_SAMPLES || discard != 0) return false;
    if (t.precision > 64) return false;                  // high-precision tables: sequential kernel
    // the throughput kernels keep the CDF table in shared memory beside 128 KiB of AES tables / 32 KiB of
    // keystream cache (gauss.cu); larger tables go through the sequential kernel
    const size_t table_bytes = (size_t)t.cdf_size * (t.precision == 64 ? 8 : 4);
    if (table_bytes + (prng_type == PRNG_AES ? 4 * 32768 : 32768) + 4096 > 216 * 1024) return false;
    const size_t words = per_stream * (t.precision == 64 ? 2 : 1);
    if (prng_type == PRNG_CHACHA20) return words <= 2 * ((size_t)kDefaultSeedPeriod / 8 - 1);   // first reseed epoch
    const size_t blocks = (words + 3) / 4;
    return blocks <= (size_t)(kDefaultSeedPeriod >> 4) * 64;                                      // first DRBG epoch
}

}  // namespace

extern "C" int scgpu_gauss_plan_create(scgpu_gauss_plan_t **out, int sampler, int precision, int blinding,
                                       float tail, float sigma, int device)
{
    if (!out) { set_error("gauss_plan_create: null argument"); return SCGPU_ERR_ARG; }
    if (!(sigma > 0) || !(tail > 0)) { set_error("gauss_plan_create: tail/sigma must be positive"); return SCGPU_ERR_ARG; }
    if (blinding < 0 || blinding > 2) { set_error("gauss_plan_create: blinding %d", blinding); return SCGPU_ERR_ARG; }
    int ndev = 0;
    SCGPU_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("gauss_plan_create: device %d of %d", device, ndev); return SCGPU_ERR_ARG; }
    SCGPU_CUDA_CHECK(cudaSetDevice(device));
    if (sampler == SCGPU_SAMPLER_CDF && (precision == 128 || precision == 192)) {
        // gaussian_cdf_create_128 / _192 (gaussian_cdf.c:385-397): the table of gauss_cdf_create_high_precision is built
        // on the host (cdf_hp.cu), the sampling runs over it as over a caller-built table
        if (!(tail * sigma >= 2.0f)) { set_error("gauss_plan_create: tail * sigma too small"); return SCGPU_ERR_ARG; }
        const std::vector<uint64_t> t = build_cdf_high(precision, blinding, tail, sigma);
        return scgpu_gauss_plan_create_table(out, precision, blinding, t.data(), t.size() / (size_t)(precision / 64), device);
    }
    scgpu_gauss_plan *p = new scgpu_gauss_plan();
    memset(&p->t, 0, sizeof(p->t));
    p->t.sampler = sampler; p->t.precision = precision; p->t.blinding = blinding;
    p->device = device;
    cudaDeviceProp prop;
    SCGPU_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    p->sm_count = prop.multiProcessorCount;
    { const int e = init_work_counters(); if (e != SCGPU_OK) return e; }
    int rc = SCGPU_OK;
    if (sampler == SCGPU_SAMPLER_CDF && precision == 64) {
        std::vector<uint64_t> cdf = build_cdf64(blinding, tail, sigma);
        uint64_t *d = nullptr;
        rc = upload(&d, cdf);
        p->d_cdf = d; p->t.cdf64 = d; p->t.cdf_size = (uint32_t)cdf.size();
        const std::vector<uint32_t> guide = build_guide(cdf);
        if (rc == SCGPU_OK && !guide.empty()) { rc = upload(&p->d_guide, guide); p->t.cdf_guide = p->d_guide; }
    } else if (sampler == SCGPU_SAMPLER_CDF && precision == 32) {
        std::vector<uint32_t> cdf = build_cdf32(blinding, tail, sigma);
        uint32_t *d = nullptr;
        rc = upload(&d, cdf);
        p->d_cdf = d; p->t.cdf32 = d; p->t.cdf_size = (uint32_t)cdf.size();
        const std::vector<uint32_t> guide = build_guide(cdf);
        if (rc == SCGPU_OK && !guide.empty()) { rc = upload(&p->d_guide, guide); p->t.cdf_guide = p->d_guide; }
    } else if (sampler == SCGPU_SAMPLER_KNUTH_YAO && (precision == 32 || precision == 64 || precision == 128)) {
        // gaussian_knuth_yao.c:126-189 (create), :50-124 (the 128 / 64 / 32-row tables); blinding scales sigma (:144-146).
        // The byte-per-bit matrix is kept as (a) the sorted flat positions of its one-bits, (b) a bitmap with
        // per-word prefix counts.
        const int rows = precision;
        if (blinding == SCGPU_BLINDING_SAMPLES) sigma *= 0.7071067811865475244008443621L;
        const int bound = (int32_t)ceil(tail * sigma);
        const int cols = bound + 1;
        long double d = 0.7978845608028653558798L / sigma;
        long double e = -0.5L / (sigma * sigma);
        std::vector<unsigned __int128> colbits((size_t)cols);
        for (int col = 0; col < cols; col++) {
            long double pr = (col == 0) ? d : d * expl(e * ((long double)(col * col)));
            colbits[col] = bin_expansion128((double)pr, rows);
        }
        const size_t total = (size_t)rows * (size_t)cols;
        std::vector<uint32_t> flat, bits(total / 32 + 2, 0), rank(total / 32 + 2, 0);
        for (int row = 0; row < rows; row++)
            for (int col = 0; col < cols; col++)
                if ((colbits[col] >> (rows - 1 - row)) & 1) {
                    const size_t pos = (size_t)row * cols + col;
                    flat.push_back((uint32_t)pos);
                    bits[pos >> 5] |= 1u << (pos & 31);
                }
        for (size_t wd = 1; wd < rank.size(); wd++) rank[wd] = rank[wd - 1] + (uint32_t)__builtin_popcount(bits[wd - 1]);
        flat.push_back(0xFFFFFFFFu);                    // sentinel: reads one past the last one-bit stay in bounds
        rc = upload(&p->d_flat, flat);
        if (rc == SCGPU_OK) rc = upload(&p->d_kybits, bits);
        if (rc == SCGPU_OK) rc = upload(&p->d_kyrank, rank);
        p->t.ky_rows = rows; p->t.ky_cols = cols; p->t.ky_bound = bound;
        p->t.ky_nones = (uint32_t)flat.size() - 1; p->t.ky_flat = p->d_flat;
        p->t.ky_bits = p->d_kybits; p->t.ky_rank = p->d_kyrank;
    } else if (sampler == SCGPU_SAMPLER_BERNOULLI && precision == 64) {
        // gaussian_bernoulli.c:40-103
        float max_gauss_val = ceil(tail * sigma);
        p->t.ber_maxval = (uint16_t)(int32_t)max_gauss_val;
        p->t.ber_maxlog = (uint16_t)(int32_t)ceil(log2f(max_gauss_val));
        size_t max_val = ceil(log2f(tail * tail * sigma * sigma));
        if (max_val > 32) { delete p; set_error("Bernoulli table needs %zu entries (accept mask is 32 bits)", max_val); return SCGPU_ERR_UNSUPPORTED; }
        std::vector<uint8_t> tab(max_val * 8);
        for (size_t i = 0; i < max_val; i++) {
            double temp = expf(-powf(2, i) / (2 * sigma * sigma));
            uint64_t bitsv = bin_expansion(temp, 64);
            for (int j = 0; j < 8; j++) tab[i * 8 + j] = (uint8_t)(bitsv >> (56 - 8 * j));
        }
        rc = upload(&p->d_ber, tab);
        p->t.ber_entries = (int)max_val; p->t.ber_tab = p->d_ber;
    } else {
        delete p;
        set_error("gauss_plan_create: sampler %d at %d-bit precision (blinding %d) is not on the GPU path", sampler, precision, blinding);
        return SCGPU_ERR_UNSUPPORTED;
    }
    if (rc != SCGPU_OK) { scgpu_gauss_plan_destroy(p); return rc; }
    *out = p;
    return SCGPU_OK;
}

// CDF sampler over a table the CALLER built (gauss_cdf_create_high_precision, gaussian_cdf.c:192-318, needs
// the reference's multi-precision float library; the table is host set-up, the sampling is the hot path).
extern "C" int scgpu_gauss_plan_create_table(scgpu_gauss_plan_t **out, int precision, int blinding,
                                             const uint64_t *table, size_t entries, int device)
{
    if (!out || !table) { set_error("gauss_plan_create_table: null argument"); return SCGPU_ERR_ARG; }
    if (precision == 256) {
        set_error("gauss_plan_create_table: the reference's 256-bit sampler reads an uninitialised word (gaussian_cdf.c:519-522): there is no behaviour to reproduce");
        return SCGPU_ERR_UNSUPPORTED;
    }
    if (precision != 128 && precision != 192) { set_error("gauss_plan_create_table: precision %d (128 or 192)", precision); return SCGPU_ERR_ARG; }
    if (blinding < 0 || blinding > 2) { set_error("gauss_plan_create_table: blinding %d", blinding); return SCGPU_ERR_ARG; }
    if (entries < 2 || entries > (1u << 24)) { set_error("gauss_plan_create_table: %zu entries", entries); return SCGPU_ERR_ARG; }
    int ndev = 0;
    SCGPU_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("gauss_plan_create_table: device %d of %d", device, ndev); return SCGPU_ERR_ARG; }
    SCGPU_CUDA_CHECK(cudaSetDevice(device));
    scgpu_gauss_plan *p = new scgpu_gauss_plan();
    memset(&p->t, 0, sizeof(p->t));
    p->t.sampler = SCGPU_SAMPLER_CDF; p->t.precision = precision; p->t.blinding = blinding;
    p->device = device;
    cudaDeviceProp prop;
    SCGPU_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    p->sm_count = prop.multiProcessorCount;
    { const int e = init_work_counters(); if (e != SCGPU_OK) return e; }
    std::vector<uint64_t> words(table, table + entries * (size_t)(precision / 64));
    uint64_t *d = nullptr;
    const int rc = upload(&d, words);
    p->d_cdf = d; p->t.cdfh = d; p->t.cdf_size = (uint32_t)entries;
    if (rc != SCGPU_OK) { scgpu_gauss_plan_destroy(p); return rc; }
    *out = p;
    return SCGPU_OK;
}

extern "C" int scgpu_gauss_plan_create_ky_fast(scgpu_gauss_plan_t **out, const uint8_t *lut1, const uint8_t *lut2, size_t lut2_len,
                                               const uint8_t *pmat, int rows, int cols, uint32_t dist1_mask, uint32_t dist2_mask,
                                               int blinding, int device)
{
    if (!out || !lut1 || !lut2 || !pmat) { set_error("gauss_plan_create_ky_fast: null argument"); return SCGPU_ERR_ARG; }
    if (rows < 1 || cols < 14 || rows > 4096 || cols > 4096) { set_error("gauss_plan_create_ky_fast: %d x %d matrix", rows, cols); return SCGPU_ERR_ARG; }
    if (lut2_len < 32 * ((size_t)dist1_mask + 1)) { set_error("gauss_plan_create_ky_fast: lut2 holds %zu bytes, the first-level distances index %zu", lut2_len, 32 * ((size_t)dist1_mask + 1)); return SCGPU_ERR_ARG; }
    if (blinding == SCGPU_BLINDING_SAMPLES) { set_error("gauss_plan_create_ky_fast: configure_sampler refuses blinding for this sampler (sampling.c:372-374)"); return SCGPU_ERR_UNSUPPORTED; }
    if (blinding < 0 || blinding > 2) { set_error("gauss_plan_create_ky_fast: blinding %d", blinding); return SCGPU_ERR_ARG; }
    int ndev = 0;
    SCGPU_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("gauss_plan_create_ky_fast: device %d of %d", device, ndev); return SCGPU_ERR_ARG; }
    SCGPU_CUDA_CHECK(cudaSetDevice(device));
    scgpu_gauss_plan *p = new scgpu_gauss_plan();
    memset(&p->t, 0, sizeof(p->t));
    p->t.sampler = SCGPU_SAMPLER_KNUTH_YAO_FAST; p->t.precision = 64; p->t.blinding = blinding;
    p->device = device;
    cudaDeviceProp prop;
    SCGPU_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    p->sm_count = prop.multiProcessorCount;
    { const int e = init_work_counters(); if (e != SCGPU_OK) { delete p; return e; } }
    std::vector<uint8_t> all(256 + lut2_len + (size_t)rows * cols);
    memcpy(all.data(), lut1, 256);
    memcpy(all.data() + 256, lut2, lut2_len);
    memcpy(all.data() + 256 + lut2_len, pmat, (size_t)rows * cols);
    const int rc = upload(&p->d_kf, all);
    if (rc != SCGPU_OK) { scgpu_gauss_plan_destroy(p); return rc; }
    p->t.kf_lut1 = p->d_kf; p->t.kf_lut2 = p->d_kf + 256; p->t.kf_pmat = p->d_kf + 256 + lut2_len;
    p->t.kf_rows = rows; p->t.kf_cols = cols; p->t.kf_d1mask = dist1_mask; p->t.kf_d2mask = dist2_mask;
    *out = p;
    return SCGPU_OK;
}

// mw_bootstrap_create(sampler, base, 16.0f, 4, 1, 64, 35, 2.5f) -- the only configuration create_sampler uses
static void mw_constants(scgpu_gauss_plan *p)
{
    const float base_sigma = 16.0f, eta = 2.5f;
    const size_t max_slevels = 4, log_base = 1, precision = 64, max_flips = 35;
    const double inv_two_eta_2 = 1.0 / (2.0 * eta * eta);
    long double wide = (long double)base_sigma * (long double)base_sigma;
    const long double base_sigma2 = wide;
    for (size_t i = 0; i < max_slevels - 1; i++) {
        const int32_t z1 = (int32_t)floor(sqrt((double)(wide * inv_two_eta_2)));
        const int32_t z2 = z1 - 1 > 1 ? z1 - 1 : 1;
        p->mw_z[i][0] = z1; p->mw_z[i][1] = z2;
        wide = (z1 * z1 + z2 * z2) * wide;
    }
    p->mw_inv_wide_sigma2 = 1 / wide;
    p->mw_k = (int32_t)ceil((double)(precision - max_flips) / log_base);
    long double rr = 1, t = 1.0 / (1UL << (2 * log_base)), s = 1.0;
    for (size_t i = (size_t)p->mw_k - 1; i--;) { s *= t; rr += s; }
    p->mw_rr_sigma2 = rr * base_sigma2;
    p->mw = true;
}

// the per-launch constants for (sigma^2 as the FLOAT product the reference forms, centre); false: sigma below the floor
static bool mw_params(const scgpu_gauss_plan *p, float sigma2, float centre, MwParams *m)
{
    memset(m, 0, sizeof(*m));
    memcpy(m->z, p->mw_z, sizeof(m->z));
    m->k = p->mw_k;
    const long double v = ((long double)(double)sigma2 - p->mw_rr_sigma2) * p->mw_inv_wide_sigma2;
    if (!(v >= 0)) return false;
    m->scale = sqrt((double)v);
    m->centre = centre;
    return true;
}

extern "C" int scgpu_gauss_plan_create_mw(scgpu_gauss_plan_t **out, int precision, int blinding, float tail, int device)
{
    if (precision != 32 && precision != 64) { set_error("gauss_plan_create_mw: base sampler precision %d (32 or 64)", precision); return SCGPU_ERR_UNSUPPORTED; }
    const int e = scgpu_gauss_plan_create(out, SCGPU_SAMPLER_CDF, precision, blinding, tail, 16.0f, device);
    if (e != SCGPU_OK) return e;
    mw_constants(*out);
    (*out)->mw_tail = tail;
    return SCGPU_OK;
}

extern "C" int scgpu_gauss_mw_streams(const scgpu_gauss_plan_t *plan, int prng_type, const uint8_t *seeds, size_t seed_len,
                                      size_t nstreams, size_t n, float sigma, float centre, const float *centres, int32_t *out,
                                      void *stream)
{
    if (!plan || !seeds || !out || seed_len == 0) { set_error("gauss_mw_streams: null/empty argument"); return SCGPU_ERR_ARG; }
    if (!plan->mw) { set_error("gauss_mw_streams: the plan was not created by scgpu_gauss_plan_create_mw"); return SCGPU_ERR_ARG; }
    if (prng_type != PRNG_AES && prng_type != PRNG_CHACHA20) { set_error("PRNG type %d is not on the GPU path", prng_type); return SCGPU_ERR_UNSUPPORTED; }
    MwParams m;
    if (!mw_params(plan, sigma * sigma, centre, &m)) { set_error("gauss_mw_streams: sigma %g is below the combiner network's noise floor", (double)sigma); return SCGPU_ERR_ARG; }
    m.centres = centres;
    const float limit = sigma * plan->mw_tail;
    m.clamp = centres ? 0 : 1;                        // integer limits exist per call; per-sample centres are clamped by the caller
    m.lim_lo = (int32_t)(-limit + centre); m.lim_hi = (int32_t)(limit + centre);
    SCGPU_CUDA_CHECK(cudaSetDevice(plan->device));
    return launch_gauss_seq(plan->t, prng_type, seeds, seed_len, kDefaultSeedPeriod, nullptr, nstreams, n, 1, 0, 0, out, 7,
                            static_cast<cudaStream_t>(stream), nullptr, &m);
}

extern "C" int scgpu_set_fixed_probe_search(int on) { return set_fixed_probe_search(on); }

extern "C" void scgpu_gauss_plan_destroy(scgpu_gauss_plan_t *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    cudaFree(p->d_cdf); cudaFree(p->d_flat); cudaFree(p->d_kybits); cudaFree(p->d_kyrank); cudaFree(p->d_ber); cudaFree(p->d_kf); cudaFree(p->d_guide);
    for (int i = 0; i < 3; i++) if (p->hstreams[i]) { cudaStreamSynchronize(p->hstreams[i]); cudaStreamDestroy(p->hstreams[i]); }
    delete p;
}

static int gauss_dispatch(scgpu_gauss_plan *p, int prng_type, const uint8_t *d_seeds, size_t seed_len, size_t nstreams,
                          size_t n, size_t calls, int32_t centre, uint32_t discard, int32_t *d_out, cudaStream_t st)
{
    if (prng_type != PRNG_AES && prng_type != PRNG_CHACHA20) { set_error("PRNG type %d is not on the GPU path (0 = AES-CTR-DRBG, 2 = ChaCha20)", prng_type); return SCGPU_ERR_UNSUPPORTED; }
    if (seed_len == 0 || !d_seeds || !d_out) { set_error("gauss_streams: null/empty argument"); return SCGPU_ERR_ARG; }
    if (discard != 0 && discard != 2 && discard != 4 && discard != 6) { set_error("gauss_streams: discard %u", discard); return SCGPU_ERR_ARG; }
    if (fast_path_ok(p->t, prng_type, n * calls, discard)) {
        // DRBG round keys of this call: stream-ordered scratch, so calls on different streams (or a second call while
        // the first is still running) never share it and nothing here synchronises the device
        uint32_t *keys = nullptr;
        if (prng_type == PRNG_AES) SCGPU_CUDA_CHECK(cudaMallocAsync(&keys, nstreams * 64 * sizeof(uint32_t), st));
        const int e = launch_gauss_fast(p->t, prng_type, d_seeds, seed_len, kDefaultSeedPeriod, nstreams, n * calls, centre,
                                        d_out, keys, p->sm_count, st);
        if (keys) SCGPU_CUDA_CHECK(cudaFreeAsync(keys, st));
        return e;
    }
    return launch_gauss_seq(p->t, prng_type, d_seeds, seed_len, kDefaultSeedPeriod, nullptr, nstreams, n, calls, centre,
                            discard, d_out, 0, st);
}

extern "C" int scgpu_gauss_streams(const scgpu_gauss_plan_t *plan, int prng_type, const uint8_t *seeds,
                                   size_t seed_len, size_t nstreams, size_t n, size_t calls, int32_t centre,
                                   uint32_t discard, int32_t *out, void *stream)
{
    if (!plan) { set_error("gauss_streams: null plan"); return SCGPU_ERR_ARG; }
    scgpu_gauss_plan *p = const_cast<scgpu_gauss_plan *>(plan);
    SCGPU_CUDA_CHECK(cudaSetDevice(p->device));
    return gauss_dispatch(p, prng_type, seeds, seed_len, nstreams, n, calls, centre, discard, out, static_cast<cudaStream_t>(stream));
}

extern "C" int scgpu_gauss_streams_host(const scgpu_gauss_plan_t *plan, int prng_type, const uint8_t *seeds,
                                        size_t seed_len, size_t nstreams, size_t n, size_t calls, int32_t centre,
                                        uint32_t discard, int32_t *out)
{
    if (!plan || !seeds || !out) { set_error("gauss_streams_host: null argument"); return SCGPU_ERR_ARG; }
    if (nstreams == 0 || n * calls == 0) return SCGPU_OK;
    scgpu_gauss_plan *p = const_cast<scgpu_gauss_plan *>(plan);
    std::lock_guard<std::mutex> lock(p->mu);
    SCGPU_CUDA_CHECK(cudaSetDevice(p->device));
    // chunks of streams through three streams: the D2H of one chunk (4 bytes per sample, the long pole) overlaps the
    // kernels of the next ones
    constexpr int kS = 3;
    for (int i = 0; i < kS; i++)
        if (!p->hstreams[i]) SCGPU_CUDA_CHECK(cudaStreamCreateWithFlags(&p->hstreams[i], cudaStreamNonBlocking));
    const size_t row = n * calls * sizeof(int32_t);
    size_t rows = (32u << 20) / row;
    if (rows < 1) rows = 1;
    if (rows > nstreams) rows = nstreams;
    int status = SCGPU_OK;
    void *ds[kS] = {}, *dout[kS] = {};
    auto cuda_ok = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && status == SCGPU_OK) { set_error("%s failed: %s", what, cudaGetErrorString(e)); status = SCGPU_ERR_CUDA; }
        return e == cudaSuccess;
    };
    for (int i = 0; i < kS && status == SCGPU_OK; i++) {
        cuda_ok(cudaMallocAsync(&ds[i], rows * seed_len, p->hstreams[i]), "cudaMallocAsync");
        cuda_ok(cudaMallocAsync(&dout[i], rows * row, p->hstreams[i]), "cudaMallocAsync");
    }
    for (size_t off = 0, ci = 0; off < nstreams && status == SCGPU_OK; off += rows, ci++) {
        const int s = (int)(ci % kS);
        cudaStream_t st = p->hstreams[s];
        const size_t cnt = nstreams - off < rows ? nstreams - off : rows;
        if (!cuda_ok(cudaMemcpyAsync(ds[s], seeds + off * seed_len, cnt * seed_len, cudaMemcpyHostToDevice, st), "H2D copy")) break;
        const int e = gauss_dispatch(p, prng_type, static_cast<const uint8_t *>(ds[s]), seed_len, cnt, n, calls, centre, discard,
                                     static_cast<int32_t *>(dout[s]), st);
        if (e != SCGPU_OK) { status = e; break; }
        cuda_ok(cudaMemcpyAsync(reinterpret_cast<char *>(out) + off * row, dout[s], cnt * row, cudaMemcpyDeviceToHost, st), "D2H copy");
    }
    for (int i = 0; i < kS; i++) {
        if (ds[i]) cudaFreeAsync(ds[i], p->hstreams[i]);
        if (dout[i]) cudaFreeAsync(dout[i], p->hstreams[i]);
        cuda_ok(cudaStreamSynchronize(p->hstreams[i]), "cudaStreamSynchronize");
    }
    return status;
}

extern "C" int scgpu_prng_words(int prng_type, const uint8_t *seeds, size_t seed_len, size_t seed_period,
                                size_t nstreams, size_t nwords, uint32_t *out, void *stream)
{
    if (prng_type != PRNG_AES && prng_type != PRNG_CHACHA20) { set_error("PRNG type %d is not on the GPU path", prng_type); return SCGPU_ERR_UNSUPPORTED; }
    if (!seeds || !out || seed_len == 0) { set_error("prng_words: null/empty argument"); return SCGPU_ERR_ARG; }
    GaussTablesDev none;
    memset(&none, 0, sizeof(none));
    return launch_gauss_seq(none, prng_type, seeds, seed_len, seed_period ? (uint32_t)seed_period : kDefaultSeedPeriod,
                            nullptr, nstreams, nwords, 1, 0, 0, reinterpret_cast<int32_t *>(out), 1, static_cast<cudaStream_t>(stream));
}

// =======================================================================================================
// Drop-in PRNG front end and sampler objects
// =======================================================================================================
//
// prng_ctx_t here is libscgpu's own context (the reference's is opaque to its callers, prng.h:38-100).  The
// generator state, the 4096-word bit pool of prng.c:95-132 and the DRBG's 1 KiB transfer buffer live on the DEVICE
// (PrngState in pooled mode, csprng.cuh); every random word -- handed to host callers through prng_32()/prng_var()/
// prng_mem() or consumed by a sampler kernel -- is produced by the kernels in gauss.cu from that state.  The host
// keeps a mirror of the state header and of the pool: it serves prng_32 & co. from the mirrored pool (exactly the
// reference's pool, refilled 4096 words at a time) and pushes the header back before the next launch, so host and
// device draws interleave in the reference's order, prng_mem draws behind the pool as the reference's does, and
// prng_reset leaves the DRBG's stale buffer in place as the reference's does.
//
// Entropy: SC_ENTROPY_USER_PROVIDED is the caller's ring buffer, wrapping as prng_get_func.c:108-119 does.  Callback
// and OS entropy is FRESH for every reseed, as in chacha20_csprng.c:21-29 / ctr_drbg.c:128-147: the device reads a
// ring of unused seeds that the host tops up before every launch with as many seeds as the launch could possibly
// consume; the device never wraps onto used bytes (PrngState::ent_avail) and a launch that runs the ring dry aborts.

struct prng_ctx_t {
    safecrypto_prng_e type;
    safecrypto_entropy_e entropy;
    size_t seed_period;
    std::vector<uint8_t> ring;          // host mirror of the device entropy ring
    bool ring_dirty = false;
    const UINT8 *user_entropy = nullptr; size_t user_len = 0;
    bool inited = false;
    int device = 0;
    cudaStream_t st = nullptr;
    uint8_t *d_ring = nullptr; size_t d_ring_cap = 0;
    PrngState *d_state = nullptr;
    PrngState hs;                       // host mirror, authoritative between launches
    uint32_t *d_poolmem = nullptr;      // kPoolWords of pool + kDrbgBufWords of DRBG transfer buffer
    uint32_t pool[kPoolWords];
    int32_t *d_out = nullptr; size_t out_cap = 0;
    std::mutex mu;
};

static prng_entropy_callback g_entropy_cb = nullptr;

namespace {

[[noreturn]] void prng_fatal(const char *what)
{
    fprintf(stderr, "libscgpu: %s: %s\n", what, scgpu_last_error());
    abort();
}

#define PRNG_CUDA(expr) do { if ((expr) != cudaSuccess) { set_error("%s", #expr); prng_fatal("CUDA call failed"); } } while (0)

GaussTablesDev no_tables()
{
    GaussTablesDev t;
    memset(&t, 0, sizeof(t));
    return t;
}

size_t per_seed_bytes(const prng_ctx_t *c) { return c->type == SC_PRNG_CHACHA ? 40 : 36; }

// one seed's worth of fresh entropy, requested the way the generator's reseed requests it
// (chacha20_csprng.c:25: 40 bytes at once; ctr_drbg.c:128-129: 4 bytes, then 32)
bool fresh_seed(const prng_ctx_t *c, uint8_t *dst)
{
    if (c->entropy == SC_ENTROPY_CALLBACK) {
        if (!g_entropy_cb) return false;
        if (c->type == SC_PRNG_AES_CTR_DRBG) { g_entropy_cb(4, dst); g_entropy_cb(32, dst + 4); }
        else g_entropy_cb(40, dst);
        return true;
    }
    // SC_ENTROPY_RANDOM / DEV_RANDOM / DEV_URANDOM: the kernel's CSPRNG (the reference's random()-based source is
    // seeded from the clock, prng.c:145-160; there is nothing to reproduce bit for bit)
    FILE *fp = fopen("/dev/urandom", "rb");
    const size_t want = per_seed_bytes(c);
    const bool ok = fp && fread(dst, 1, want, fp) == want;
    if (fp) fclose(fp);
    return ok;
}

// upper bound on the reseeds one launch that hands out `words` 32-bit words can trigger
size_t reseeds_for(const prng_ctx_t *c, size_t words)
{
    const size_t bytes = words * 4 + 4 * kPoolWords + 4 * kDrbgBufWords;     // + one pool refill, one DRBG update
    if (c->type == SC_PRNG_CHACHA) return bytes / (c->seed_period ? c->seed_period : 1) + 2;
    size_t period = c->seed_period >> 4;                                     // ctr_drbg.c:46-53, in 1 KiB updates
    if (period < 0x1000) period = 0x1000;
    return bytes / 1024 / period + 2;
}

// fresh-entropy mode: make sure `reseeds` unused seeds lie ahead of the device's read position
void top_up_entropy(prng_ctx_t *c, size_t reseeds)
{
    if (!c->hs.ent_fresh) return;
    const size_t ps = per_seed_bytes(c);
    size_t L = c->ring.size();
    if (reseeds * ps > L) {
        // grow: keep the unused seeds in order, the rest is filled below
        std::vector<uint8_t> nr(reseeds * ps, 0);
        for (size_t i = 0; i < c->hs.ent_avail; i++) nr[i] = c->ring[(c->hs.ent_idx + i) % L];
        memset(c->ring.data(), 0, c->ring.size());
        c->ring.swap(nr);
        c->hs.ent_idx = 0;
        L = c->ring.size();
        c->hs.seed_len = (uint32_t)L;
    }
    if (c->hs.ent_avail == L) return;
    // used region: [ent_idx + ent_avail, ent_idx + L) mod L, a whole number of seeds starting on a seed boundary
    for (size_t off = c->hs.ent_avail; off < L; off += ps) {
        uint8_t seed[40];
        if (!fresh_seed(c, seed)) { set_error("entropy source failed"); prng_fatal("prng reseed entropy"); }
        for (size_t i = 0; i < ps; i++) c->ring[(c->hs.ent_idx + off + i) % L] = seed[i];
        memset(seed, 0, sizeof(seed));
    }
    c->hs.ent_avail = (uint32_t)L;
    c->ring_dirty = true;
}

// One kernel on the context's state: push the header (and the ring if the host refreshed it), run, pull the header
// and -- when the kernel refilled it -- the pool.  `words` bounds the 32-bit words the launch may hand out.
void run_state_kernel(prng_ctx_t *c, const GaussTablesDev &t, size_t n, size_t calls, int32_t centre, uint32_t discard,
                      int mode, int32_t *d_out, size_t words, const MwParams *mw = nullptr)
{
    PRNG_CUDA(cudaSetDevice(c->device));
    top_up_entropy(c, reseeds_for(c, words));
    if (c->ring_dirty || c->d_ring_cap < c->ring.size()) {
        if (c->d_ring_cap < c->ring.size()) {
            if (c->d_ring) { PRNG_CUDA(cudaMemsetAsync(c->d_ring, 0, c->d_ring_cap, c->st)); PRNG_CUDA(cudaStreamSynchronize(c->st)); cudaFree(c->d_ring); }
            PRNG_CUDA(cudaMalloc(&c->d_ring, c->ring.size()));
            c->d_ring_cap = c->ring.size();
        }
        PRNG_CUDA(cudaMemcpyAsync(c->d_ring, c->ring.data(), c->ring.size(), cudaMemcpyHostToDevice, c->st));
        c->ring_dirty = false;
    }
    const uint64_t draws_before = c->hs.draws64;
    PRNG_CUDA(cudaMemcpyAsync(c->d_state, &c->hs, sizeof(PrngState), cudaMemcpyHostToDevice, c->st));
    if (launch_gauss_seq(t, c->type, c->d_ring, c->ring.size(), (uint32_t)c->seed_period, c->d_state, 1, n, calls, centre,
                         discard, d_out, mode, c->st, c->d_poolmem, mw) != SCGPU_OK)
        prng_fatal("prng kernel");
    PRNG_CUDA(cudaMemcpyAsync(&c->hs, c->d_state, sizeof(PrngState), cudaMemcpyDeviceToHost, c->st));
    PRNG_CUDA(cudaStreamSynchronize(c->st));
    if (c->hs.error) {
        set_error("the device ran out of fresh entropy inside one launch (%zu words requested); entropy is never re-used", words);
        prng_fatal("prng entropy");
    }
    if (mode != 4 && c->hs.draws64 != draws_before && c->hs.pool_fill) {
        PRNG_CUDA(cudaMemcpyAsync(c->pool, c->d_poolmem, sizeof(c->pool), cudaMemcpyDeviceToHost, c->st));
        PRNG_CUDA(cudaStreamSynchronize(c->st));
    }
}

int32_t *ensure_out(prng_ctx_t *c, size_t words)
{
    if (c->out_cap < words) {
        cudaFree(c->d_out);
        c->d_out = nullptr;
        PRNG_CUDA(cudaMalloc(&c->d_out, sizeof(int32_t) * words));
        c->out_cap = words;
    }
    return c->d_out;
}

uint32_t host_next32(prng_ctx_t *c)
{
    if (c->hs.pool_rd >= c->hs.pool_fill) run_state_kernel(c, no_tables(), 0, 0, 0, 0, 6, nullptr, kPoolWords);
    c->hs.words_out++;
    return c->pool[c->hs.pool_rd++];
}

// prng.c:1017-1048 on the mirrored bit buffer
uint32_t host_var(prng_ctx_t *c, size_t n)
{
    UINT32 mask = n >= 32 ? 0xFFFFFFFFu : (1u << n) - 1u;
    if (n > 32) n = 32;
    UINT32 ret = c->hs.var_buf;
    if (c->hs.var_bits < n) {
        size_t need = n - c->hs.var_bits;
        ret = need >= 32 ? ret : ret << need;
        c->hs.var_buf = host_next32(c);
        ret |= c->hs.var_buf & (need >= 32 ? 0xFFFFFFFFu : ((1u << need) - 1u));     // need == 32: see csprng.cuh var()
        c->hs.var_buf = need >= 32 ? c->hs.var_buf : c->hs.var_buf >> need;
        c->hs.var_bits = (uint32_t)(32 - need);
    } else {
        c->hs.var_buf >>= n;
        c->hs.var_bits -= (uint32_t)n;
    }
    return ret & mask;
}

}  // namespace

extern "C" {

prng_ctx_t *prng_create(safecrypto_entropy_e entropy, safecrypto_prng_e type, safecrypto_prng_threading_e mt,
                        size_t seed_period)
{
    (void)mt;
    // prng.c:564-626: unknown entropy sources, unknown types and a zero period are rejected; only the two generators
    // whose byte stream is reproduced on the device are available here
    if (entropy != SC_ENTROPY_RANDOM && entropy != SC_ENTROPY_DEV_RANDOM && entropy != SC_ENTROPY_DEV_URANDOM &&
        entropy != SC_ENTROPY_DEV_HWRNG && entropy != SC_ENTROPY_CALLBACK && entropy != SC_ENTROPY_USER_PROVIDED) return NULL;
    if (type != SC_PRNG_AES_CTR_DRBG && type != SC_PRNG_CHACHA) return NULL;
    if (seed_period == 0) return NULL;
    prng_ctx_t *c = new prng_ctx_t();
    c->type = type; c->entropy = entropy; c->seed_period = seed_period;
    memset(&c->hs, 0, sizeof(c->hs));
    const char *env = getenv("SCGPU_DEVICE");
    c->device = env ? atoi(env) : 0;
    return c;
}

SINT32 prng_set_entropy(prng_ctx_t *ctx, const UINT8 *entropy, size_t len)
{
    if (!ctx) return SC_FUNC_FAILURE;
    ctx->user_entropy = entropy; ctx->user_len = len;
    return SC_FUNC_SUCCESS;
}

SINT32 prng_set_entropy_callback(prng_entropy_callback cb)
{
    g_entropy_cb = cb;
    return SC_FUNC_SUCCESS;
}

SINT32 prng_init(prng_ctx_t *ctx, const UINT8 *nonce, size_t len_nonce)
{
    (void)nonce; (void)len_nonce;       // unused by these two generators (prng.c:219-262)
    if (!ctx || ctx->inited) return SC_FUNC_FAILURE;
    memset(&ctx->hs, 0, sizeof(ctx->hs));
    ctx->hs.pooled = 1;
    switch (ctx->entropy) {
    case SC_ENTROPY_USER_PROVIDED:
        if (!ctx->user_entropy || ctx->user_len == 0) return SC_FUNC_FAILURE;
        ctx->ring.assign(ctx->user_entropy, ctx->user_entropy + ctx->user_len);
        ctx->ring_dirty = true;
        break;
    case SC_ENTROPY_CALLBACK:
        if (!g_entropy_cb) return SC_FUNC_FAILURE;
        ctx->hs.ent_fresh = 1;
        break;
    case SC_ENTROPY_RANDOM: case SC_ENTROPY_DEV_RANDOM: case SC_ENTROPY_DEV_URANDOM: case SC_ENTROPY_DEV_HWRNG:
        ctx->hs.ent_fresh = 1;
        break;
    default:
        return SC_FUNC_FAILURE;
    }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return SC_FUNC_FAILURE;
    PRNG_CUDA(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
    PRNG_CUDA(cudaMalloc(&ctx->d_state, sizeof(PrngState)));
    PRNG_CUDA(cudaMalloc(&ctx->d_poolmem, sizeof(uint32_t) * (kPoolWords + kDrbgBufWords)));
    PRNG_CUDA(cudaMemsetAsync(ctx->d_poolmem, 0, sizeof(uint32_t) * (kPoolWords + kDrbgBufWords), ctx->st));
    ctx->inited = true;
    // instantiate the generator on the device (hs.instantiated == 0); in fresh mode the ring is first filled by
    // top_up_entropy
    run_state_kernel(ctx, no_tables(), 0, 0, 0, 0, 3, nullptr, 0);
    return SC_FUNC_SUCCESS;
}

safecrypto_prng_e prng_get_type(prng_ctx_t *ctx) { return ctx ? ctx->type : SC_PRNG_MAX; }

SINT32 prng_destroy(prng_ctx_t *ctx)
{
    if (!ctx) return SC_FUNC_FAILURE;
    if (ctx->inited) {
        cudaSetDevice(ctx->device);
        // zeroise what held key material before releasing it
        if (ctx->d_ring) cudaMemsetAsync(ctx->d_ring, 0, ctx->d_ring_cap, ctx->st);
        cudaMemsetAsync(ctx->d_state, 0, sizeof(PrngState), ctx->st);
        cudaMemsetAsync(ctx->d_poolmem, 0, sizeof(uint32_t) * (kPoolWords + kDrbgBufWords), ctx->st);
        if (ctx->d_out) cudaMemsetAsync(ctx->d_out, 0, sizeof(int32_t) * ctx->out_cap, ctx->st);
        cudaStreamSynchronize(ctx->st);
        cudaFree(ctx->d_ring); cudaFree(ctx->d_state); cudaFree(ctx->d_poolmem); cudaFree(ctx->d_out);
        cudaStreamDestroy(ctx->st);
    }
    if (!ctx->ring.empty()) memset(ctx->ring.data(), 0, ctx->ring.size());
    memset(&ctx->hs, 0, sizeof(ctx->hs));
    memset(ctx->pool, 0, sizeof(ctx->pool));
    delete ctx;
    return SC_FUNC_SUCCESS;
}

// prng.c:861-932.  CTR-DRBG: ctr_drbg_reset (zero key and counter, reseed); the transfer buffer keeps its position,
// so the next draws drain what the old key left there, as in the reference.  ChaCha20: the reference calls
// reset_chacha20, which FREES the generator (chacha20_csprng.c:58-67, the bodies of reset/destroy are swapped) and
// every later draw is a use after free; here the front end is reset and the generator reseeded, which is what the
// swapped function (destroy_chacha20, :49-56) does.
void prng_reset(prng_ctx_t *ctx)
{
    if (!ctx || !ctx->inited) return;
    std::lock_guard<std::mutex> lock(ctx->mu);
    run_state_kernel(ctx, no_tables(), 0, 0, 0, 0, 5, nullptr, 0);
}

UINT64 prng_get_csprng_bytes(prng_ctx_t *ctx) { return ctx->hs.draws64 * 8; }
UINT64 prng_get_out_bytes(prng_ctx_t *ctx) { return ctx->hs.words_out * 4; }

UINT32 prng_32(prng_ctx_t *ctx)
{
    std::lock_guard<std::mutex> lock(ctx->mu);
    return host_next32(ctx);
}

UINT64 prng_64(prng_ctx_t *ctx)
{
    std::lock_guard<std::mutex> lock(ctx->mu);
    UINT64 hi = host_next32(ctx);
    return (hi << 32) | host_next32(ctx);
}

#ifdef __SIZEOF_INT128__
// prng.c:950-960
unsigned __int128 prng_128(prng_ctx_t *ctx)
{
    unsigned __int128 v = prng_64(ctx);
    v <<= 64;
    v |= prng_64(ctx);
    return v;
}
#endif

UINT32 prng_var(prng_ctx_t *ctx, size_t n)
{
    std::lock_guard<std::mutex> lock(ctx->mu);
    return host_var(ctx, n);
}

SINT32 prng_bit(prng_ctx_t *ctx) { return (SINT32)prng_var(ctx, 1); }
UINT16 prng_16(prng_ctx_t *ctx) { return (UINT16)prng_var(ctx, 16); }
UINT8 prng_8(prng_ctx_t *ctx) { return (UINT8)prng_var(ctx, 8); }

// prng.c:1005-1015
FLOAT prng_float(prng_ctx_t *ctx) { return ((FLOAT)prng_32(ctx)) / UINT32_MAX; }
DOUBLE prng_double(prng_ctx_t *ctx)
{
    std::lock_guard<std::mutex> lock(ctx->mu);
    UINT32 a = host_var(ctx, 27);
    UINT32 b = host_var(ctx, 26);
    return (a * 67108864.0 + b) * 1.11022302462516e-16;
}

// prng.c:1050-1105: ceil(length / 64) blocks of eight generator draws, copied out as little-endian u64; the draws
// come from the generator itself, behind whatever the bit pool already holds
SINT32 prng_mem(prng_ctx_t *ctx, UINT8 *mem, SINT32 length)
{
    if (!ctx || !ctx->inited) return SC_FUNC_FAILURE;
    if (length <= 0) return SC_FUNC_SUCCESS;
    std::lock_guard<std::mutex> lock(ctx->mu);
    const size_t blocks = ((size_t)length + 63) >> 6;
    int32_t *d = ensure_out(ctx, blocks * 16);
    run_state_kernel(ctx, no_tables(), blocks, 1, 0, 0, 4, d, blocks * 16);
    PRNG_CUDA(cudaMemcpyAsync(mem, d, (size_t)length, cudaMemcpyDeviceToHost, ctx->st));
    PRNG_CUDA(cudaStreamSynchronize(ctx->st));
    return SC_FUNC_SUCCESS;
}

}  // extern "C"

// ---- samplers ----------------------------------------------------------------------------------------------

namespace {

struct GaussObj {            // what utils_sampling_t::gauss points at
    scgpu_gauss_plan *plan;
    prng_ctx_t *prng;
};

// one launch of mw_bootstrap_sample calls on the context's device state (unclamped)
void run_mw_on_ctx(GaussObj *o, int32_t *host_out, size_t n, float sigma2, float centre)
{
    prng_ctx_t *c = o->prng;
    MwParams m;
    if (!mw_params(o->plan, sigma2, centre, &m)) { set_error("bootstrap sampler: sigma^2 = %g is below the combiner network's noise floor", (double)sigma2); prng_fatal("get_bootstrap_sample"); }
    std::lock_guard<std::mutex> lock(c->mu);
    int32_t *d = ensure_out(c, n);
    // 8 + 29 base samples and one flip word per sample, 64-bit draws
    run_state_kernel(c, o->plan->t, n, 1, 0, 0, 7, d, n * 80 + 65536, &m);
    PRNG_CUDA(cudaMemcpyAsync(host_out, d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->st));
    PRNG_CUDA(cudaStreamSynchronize(c->st));
}

void *gauss_create_stub(prng_ctx_t *, FLOAT, FLOAT, size_t, sample_blinding_e) { return nullptr; }

SINT32 gauss_destroy(void **g)
{
    if (!g || !*g) return SC_FUNC_FAILURE;
    GaussObj *o = static_cast<GaussObj *>(*g);
    scgpu_gauss_plan_destroy(o->plan);
    delete o;
    *g = nullptr;
    return SC_FUNC_SUCCESS;
}

prng_ctx_t *gauss_get_prng(void *g) { return g ? static_cast<GaussObj *>(g)->prng : nullptr; }

// run `n` samples (vector mode) or one bare sample (mode 2) of the sampler on the context's device state
void run_on_ctx(GaussObj *o, int32_t *host_out, size_t n, int32_t centre, uint32_t discard, int mode)
{
    prng_ctx_t *c = o->prng;
    std::lock_guard<std::mutex> lock(c->mu);
    int32_t *d = ensure_out(c, n);
    // words one launch may draw: the CDF samplers are bounded by the mode; Knuth-Yao and Bernoulli restart
    // data-dependently, so they get a generous bound (a launch that still runs the fresh-entropy ring dry aborts)
    const GaussTablesDev &t = o->plan->t;
    size_t per = (size_t)(t.precision > 32 ? t.precision / 32 : 1);
    if (t.sampler == SCGPU_SAMPLER_KNUTH_YAO) per = 64;
    if (t.sampler == SCGPU_SAMPLER_BERNOULLI) per = 4096;
    if (t.blinding != SCGPU_NORMAL_SAMPLES) per = 2 * per + 8;
    if (discard) per *= 4;
    run_state_kernel(c, t, n, 1, centre, discard, mode, d, n * per + 65536);
    PRNG_CUDA(cudaMemcpyAsync(host_out, d, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->st));
    PRNG_CUDA(cudaStreamSynchronize(c->st));
}

SINT32 gauss_sample(void *g)
{
    int32_t v = 0;
    run_on_ctx(static_cast<GaussObj *>(g), &v, 1, 0, 0, 2);
    return v;
}

SINT32 vector_32(const utils_sampling_t *s, SINT32 *v, size_t n, SINT32 centre)
{
    run_on_ctx(static_cast<GaussObj *>(s->gauss), v, n, centre, s->discard, 0);
    return SC_FUNC_SUCCESS;
}

SINT32 vector_16(const utils_sampling_t *s, SINT16 *v, size_t n, SINT32 centre)
{
    std::vector<int32_t> tmp(n);
    run_on_ctx(static_cast<GaussObj *>(s->gauss), tmp.data(), n, centre, s->discard, 0);
    for (size_t i = 0; i < n; i++) v[i] = (SINT16)tmp[i];        // the reference stores each sample into SINT16
    return SC_FUNC_SUCCESS;
}

}  // namespace

extern "C" {

// sampling.c:425-469.  Unsupported combinations return NULL like the reference's configure_sampler.
utils_sampling_t *create_sampler(random_sampling_e type, sample_precision_e precision, sample_blinding_e blinding,
                                 SINT32 dimension, sample_bootstrap_e bootstrapped, prng_ctx_t *prng_ctx,
                                 FLOAT tail, FLOAT sigma)
{
    if (!prng_ctx || !prng_ctx->inited) return NULL;
    if (bootstrapped == SAMPLING_MW_BOOTSTRAP) {
        // sampling.c:449-457: a base sampler of sigma 16 of the configured type, combined by mw_bootstrap_create(.., 16.0f,
        // 4, 1, 64, 35, 2.5f); the CDF base samplers are on the GPU path
        if (type != CDF_GAUSSIAN_SAMPLING || (precision != SAMPLING_32BIT && precision != SAMPLING_64BIT)) return NULL;
        scgpu_gauss_plan *mplan = nullptr;
        if (scgpu_gauss_plan_create_mw(&mplan, (int)precision, (int)blinding, tail, prng_ctx->device) != SCGPU_OK) return NULL;
        utils_sampling_t *s = static_cast<utils_sampling_t *>(calloc(1, sizeof(utils_sampling_t)));
        if (!s) { scgpu_gauss_plan_destroy(mplan); return NULL; }
        GaussObj *o = new GaussObj{mplan, prng_ctx};
        s->create = gauss_create_stub; s->destroy = gauss_destroy; s->get_prng = gauss_get_prng; s->sample = gauss_sample;
        s->vector_16 = vector_16; s->vector_32 = vector_32;
        s->precision = precision; s->dimension = dimension; s->bootstrapped = bootstrapped;
        s->tail = tail; s->sigma = sigma; s->sigma2 = sigma * sigma;
        s->gauss = o; s->prng_ctx = prng_ctx; s->discard = 0;
        s->bootstrap = o;                         // mw_bootstrap_destroy's counterpart is gauss_destroy
        return s;
    }
    if (bootstrapped != SAMPLING_DISABLE_BOOTSTRAP) return NULL;
    // configure_sampler refuses blinding for the Knuth-Yao samplers (sampling.c:341-343, 372-374)
    if ((type == KNUTH_YAO_GAUSSIAN_SAMPLING || type == KNUTH_YAO_FAST_GAUSSIAN_SAMPLING) && blinding == BLINDING_SAMPLES) return NULL;
    scgpu_gauss_plan *plan = nullptr;
    if (scgpu_gauss_plan_create(&plan, (int)type, (int)precision, (int)blinding, tail, sigma, prng_ctx->device) != SCGPU_OK)
        return NULL;
    utils_sampling_t *s = static_cast<utils_sampling_t *>(calloc(1, sizeof(utils_sampling_t)));
    if (!s) { scgpu_gauss_plan_destroy(plan); return NULL; }
    GaussObj *o = new GaussObj{plan, prng_ctx};
    s->create = gauss_create_stub;
    s->destroy = gauss_destroy;
    s->get_prng = gauss_get_prng;
    s->sample = gauss_sample;
    s->vector_16 = vector_16;
    s->vector_32 = vector_32;
    s->precision = precision;
    s->dimension = dimension;
    s->bootstrapped = bootstrapped;
    s->tail = tail;
    s->sigma = sigma;
    s->gauss = o;
    s->prng_ctx = prng_ctx;
    s->discard = 0;
    s->bootstrap = NULL;
    return s;
}

SINT32 destroy_sampler(utils_sampling_t **sampler)
{
    if (!sampler || !*sampler) return SC_FUNC_FAILURE;
    utils_sampling_t *s = *sampler;
    if (s->destroy(&s->gauss) == SC_FUNC_FAILURE) return SC_FUNC_FAILURE;
    free(s);
    *sampler = NULL;
    return SC_FUNC_SUCCESS;
}

SINT32 set_discard(utils_sampling_t *sampler, UINT32 discard) { sampler->discard = discard; return SC_FUNC_SUCCESS; }
SINT32 get_sample(utils_sampling_t *sampler) { return sampler->sample(sampler->gauss); }
// sampling.c:519-538
SINT32 get_bootstrap_sample(utils_sampling_t *sampler, FLOAT sigma, FLOAT centre)
{
    if (sampler->bootstrapped != SAMPLING_MW_BOOTSTRAP) return 0;
    SINT32 sample = 0;
    run_mw_on_ctx(static_cast<GaussObj *>(sampler->gauss), &sample, 1, sigma * sigma, centre);
    const FLOAT limit = sigma * sampler->tail;
    if (sample < (-limit + centre)) sample = (-limit + centre);
    if (sample > (limit + centre)) sample = (limit + centre);
    return sample;
}
// sampling.c:540-580: the bootstrap branch clamps to integer limits, the plain branch truncates the centre
static void mw_vector(utils_sampling_t *sampler, SINT32 *v, size_t n, FLOAT centre)
{
    run_mw_on_ctx(static_cast<GaussObj *>(sampler->gauss), v, n, sampler->sigma2, centre);
    const FLOAT limit = sampler->sigma * sampler->tail;
    const SINT32 limits[2] = {(SINT32)(-limit + centre), (SINT32)(limit + centre)};
    for (size_t i = 0; i < n; i++) v[i] = v[i] < limits[0] ? limits[0] : (v[i] > limits[1] ? limits[1] : v[i]);
}
SINT32 get_vector_16(utils_sampling_t *sampler, SINT16 *v, size_t n, FLOAT centre)
{
    if (sampler->bootstrapped == SAMPLING_MW_BOOTSTRAP) {
        std::vector<SINT32> tmp(n);
        mw_vector(sampler, tmp.data(), n, centre);
        for (size_t i = 0; i < n; i++) v[i] = (SINT16)tmp[i];
        return SC_FUNC_SUCCESS;
    }
    return sampler->vector_16(sampler, v, n, (SINT32)centre);
}
SINT32 get_vector_32(utils_sampling_t *sampler, SINT32 *v, size_t n, FLOAT centre)
{
    if (sampler->bootstrapped == SAMPLING_MW_BOOTSTRAP) { mw_vector(sampler, v, n, centre); return SC_FUNC_SUCCESS; }
    return sampler->vector_32(sampler, v, n, (SINT32)centre);
}

}  // extern "C"
