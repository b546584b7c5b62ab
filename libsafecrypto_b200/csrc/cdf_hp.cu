// cdf_hp.cu -- host set-up: the 128 / 192 / 256-bit CDF tables of gauss_cdf_create_high_precision
// (src/utils/sampling/gaussian_cdf.c:192-318, sized by gaussian_cdf_create_high, :328-383).
//
// The reference runs this function on its sc_mpf layer.  Built with MPFR, every sc_mpf call is the mpfr_* function of
// the same name with MPFR_RNDZ at `precision` bits (src/utils/arith/sc_mpf.c:38, :42-54): each operation returns the
// exact result TRUNCATED to `precision` significant bits.  That is a complete specification, and it is what this file
// computes, operation for operation, in plain multi-word integer arithmetic (no MPFR, no GMP):
//
//   pi = trunc(pi); two_sqrt_2pi = trunc(2 / trunc(sqrt(2 pi))); sigma' = sigma (or trunc(sigma trunc(sqrt(1/2))) with
//   blinding); d = trunc(trunc(2^precision / sigma') two_sqrt_2pi); e = -trunc(1/2 / trunc(sigma'^2)); s_1 = d / 2;
//   table[i] = floor(s_i);  s_(i+1) = trunc(s_i + trunc(d trunc(exp(trunc(e i^2)))));  table[0] = 0, table[last] = ~0.
//
// exp is evaluated with 96 guard bits (argument reduction by ln 2, alternating Taylor series in fixed point) and then
// truncated, pi / ln 2 are 384-bit constants: the truncation of the true value unless it lies within 2^-80 ulp of a
// representable number.  The reference build WITHOUT MPFR -- the only one possible in this repository's container --
// has empty bodies for sc_mpf_exp / sc_mpf_get_pi and produces a degenerate table, so the compiled reference cannot
// pin this function; tests/test_cdf_high.py compares it bit for bit with tables computed independently by
// tests/golden/make_cdf_high.py over mpmath's correctly rounded arithmetic, and with the (pinned) 64-bit table.
#include "scgpu_internal.h"
#include "../../include/scgpu.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace scgpu {
namespace {

// ---- unsigned multi-word integers, little-endian 32-bit limbs (sizes here: a few hundred bits) ---------------------
struct Big {
    std::vector<uint32_t> w;
    Big() {}
    explicit Big(uint64_t v) { if (v) { w.push_back((uint32_t)v); if (v >> 32) w.push_back((uint32_t)(v >> 32)); } }
    void trim() { while (!w.empty() && w.back() == 0) w.pop_back(); }
    bool zero() const { return w.empty(); }
    int bits() const
    {
        if (w.empty()) return 0;
        return 32 * (int)(w.size() - 1) + (32 - __builtin_clz(w.back()));
    }
    static Big from_hex(const char *h)
    {
        Big r;
        const size_t n = strlen(h);
        for (size_t i = 0; i < n; i++) {
            const char c = h[n - 1 - i];
            const uint32_t v = (uint32_t)(c <= '9' ? c - '0' : (c | 0x20) - 'a' + 10);
            if (i / 8 >= r.w.size()) r.w.push_back(0);
            r.w[i / 8] |= v << (4 * (i % 8));
        }
        r.trim();
        return r;
    }
};

int cmp(const Big &a, const Big &b)
{
    if (a.w.size() != b.w.size()) return a.w.size() < b.w.size() ? -1 : 1;
    for (size_t i = a.w.size(); i-- > 0;)
        if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
}
Big add(const Big &a, const Big &b)
{
    Big r;
    uint64_t c = 0;
    for (size_t i = 0; i < std::max(a.w.size(), b.w.size()) || c; i++) {
        c += (i < a.w.size() ? a.w[i] : 0ull) + (i < b.w.size() ? b.w[i] : 0ull);
        r.w.push_back((uint32_t)c);
        c >>= 32;
    }
    r.trim();
    return r;
}
Big sub(const Big &a, const Big &b)                       // a >= b
{
    Big r;
    int64_t c = 0;
    for (size_t i = 0; i < a.w.size(); i++) {
        c += (int64_t)a.w[i] - (i < b.w.size() ? (int64_t)b.w[i] : 0);
        r.w.push_back((uint32_t)c);
        c >>= 32;                                          // arithmetic shift: borrow is -1
    }
    r.trim();
    return r;
}
Big shl(const Big &a, int s)
{
    if (a.zero() || s == 0) return a;
    Big r;
    const int ws = s / 32, bs = s % 32;
    r.w.assign(a.w.size() + ws + 1, 0);
    for (size_t i = 0; i < a.w.size(); i++) {
        const uint64_t v = (uint64_t)a.w[i] << bs;
        r.w[i + ws] |= (uint32_t)v;
        r.w[i + ws + 1] |= (uint32_t)(v >> 32);
    }
    r.trim();
    return r;
}
Big shr(const Big &a, int s)                               // floor(a / 2^s)
{
    const int ws = s / 32, bs = s % 32;
    if ((size_t)ws >= a.w.size()) return Big();
    Big r;
    r.w.assign(a.w.size() - ws, 0);
    for (size_t i = 0; i < r.w.size(); i++) {
        uint64_t v = a.w[i + ws];
        if (i + ws + 1 < a.w.size()) v |= (uint64_t)a.w[i + ws + 1] << 32;
        r.w[i] = (uint32_t)(v >> bs);
    }
    r.trim();
    return r;
}
Big mul(const Big &a, const Big &b)
{
    if (a.zero() || b.zero()) return Big();
    Big r;
    r.w.assign(a.w.size() + b.w.size(), 0);
    for (size_t i = 0; i < a.w.size(); i++) {
        uint64_t c = 0;
        for (size_t j = 0; j < b.w.size(); j++) {
            c += (uint64_t)a.w[i] * b.w[j] + r.w[i + j];
            r.w[i + j] = (uint32_t)c;
            c >>= 32;
        }
        r.w[i + b.w.size()] += (uint32_t)c;
    }
    r.trim();
    return r;
}
Big div_small(const Big &a, uint32_t d)                    // floor(a / d)
{
    Big r;
    r.w.assign(a.w.size(), 0);
    uint64_t rem = 0;
    for (size_t i = a.w.size(); i-- > 0;) {
        rem = (rem << 32) | a.w[i];
        r.w[i] = (uint32_t)(rem / d);
        rem %= d;
    }
    r.trim();
    return r;
}
Big div(const Big &a, const Big &b)                        // floor(a / b), shift-and-subtract
{
    Big q, rem;
    const int n = a.bits();
    q.w.assign((size_t)(n + 31) / 32, 0);
    for (int i = n - 1; i >= 0; i--) {
        rem = shl(rem, 1);
        if ((a.w[(size_t)i / 32] >> (i % 32)) & 1u) { if (rem.w.empty()) rem.w.push_back(1); else rem.w[0] |= 1u; }
        if (cmp(rem, b) >= 0) { rem = sub(rem, b); q.w[(size_t)i / 32] |= 1u << (i % 32); }
    }
    q.trim();
    return q;
}
Big isqrt(const Big &a)                                    // floor(sqrt(a)), bit by bit
{
    Big r;
    for (int i = (a.bits() + 1) / 2; i >= 0; i--) {
        Big t = add(r, shl(Big(1), i));
        if (cmp(mul(t, t), a) <= 0) r = t;
    }
    return r;
}

// ---- positive floating-point numbers with a P-bit mantissa, every operation truncating (MPFR_RNDZ) ----------------
struct Hp { Big m; long e; };                              // m 2^e, m has exactly P bits (or is zero)

Hp norm(Big m, long e, int P)
{
    const int b = m.bits();
    if (b == 0) return Hp{Big(), 0};
    if (b > P) return Hp{shr(m, b - P), e + (b - P)};
    return Hp{shl(m, P - b), e - (P - b)};
}
Hp from_double(double v, int P)                            // exact (53 <= P)
{
    int ex = 0;
    const double fr = frexp(v, &ex);
    return norm(Big((uint64_t)ldexp(fr, 53)), (long)ex - 53, P);
}
Hp hp_mul(const Hp &a, const Hp &b, int P) { return norm(mul(a.m, b.m), a.e + b.e, P); }
Hp hp_div(const Hp &a, const Hp &b, int P) { return norm(div(shl(a.m, P + 2), b.m), a.e - b.e - (P + 2), P); }
Hp hp_sqrt(const Hp &a, int P)
{
    Big m = a.m;
    long e = a.e;
    if (e & 1) { m = shl(m, 1); e -= 1; }
    const int K = P + 2 + (P & 1);                         // even: sqrt(m 2^K) has at least P bits
    return norm(isqrt(shl(m, K)), (e - K) / 2, P);
}
Hp hp_add(const Hp &a, const Hp &b, int P)
{
    if (a.m.zero()) return b;
    if (b.m.zero()) return a;
    const Hp &hi = a.e >= b.e ? a : b, &lo = a.e >= b.e ? b : a;
    const long diff = hi.e - lo.e;
    if (diff > P + 2) return hi;                           // the small term is below the last bit: truncation drops it
    return norm(add(shl(hi.m, (int)diff), lo.m), lo.e, P);
}
Big to_int(const Hp &a)                                    // floor
{
    if (a.m.zero()) return Big();
    return a.e >= 0 ? shl(a.m, (int)a.e) : shr(a.m, (int)-a.e);
}

const char kPiHex[] = "c90fdaa22168c234c4c6628b80dc1cd129024e088a67cc74020bbea63b139b22514a08798e3404ddef9519b3cd3a431b";   // floor(pi 2^382)
const char kLn2Hex[] = "b17217f7d1cf79abc9e3b39803f2f6af40f343267298b62d8a0d175b8baafa2be7b876206debac98559552fb4afa1b10";  // floor(ln2 2^384)

// trunc(exp(-x)) for x > 0
Hp hp_exp_neg(const Hp &x, int P)
{
    const int F = P + 96;
    const long sh = x.e + F;
    const Big xf = sh >= 0 ? shl(x.m, (int)sh) : shr(x.m, (int)-sh);           // floor(x 2^F)
    const Big ln2f = shr(Big::from_hex(kLn2Hex), 384 - F);
    Big r = xf;
    long k = 0;
    {                                                                           // k = floor(x / ln 2): a few hundred at most
        const Big q = div(xf, ln2f);
        k = q.zero() ? 0 : (long)q.w[0];
        r = sub(xf, mul(q, ln2f));
    }
    Big pos = shl(Big(1), F), neg, term = pos;
    for (uint32_t n = 1; !term.zero(); n++) {
        term = div_small(shr(mul(term, r), F), n);
        if (n & 1) neg = add(neg, term); else pos = add(pos, term);
    }
    return norm(sub(pos, neg), -(long)F - k, P);
}

size_t ceil_log2(size_t x)
{
    size_t l = 0;
    while ((x >> (l + 1)) != 0) l++;
    if (x & (x - 1)) l++;
    return l;
}

}  // namespace

// table: entries x (precision / 64) words, word 0 least significant (the reference's cdf_128 / cdf_192 / cdf_256 array)
std::vector<uint64_t> build_cdf_high(int precision, int blinding, float tail, float sigma)
{
    const int P = precision, nw = P / 64;
    const size_t size = (size_t)1 << ceil_log2((size_t)(tail * sigma));        // gaussian_cdf.c:340, :375
    std::vector<uint64_t> out(size * (size_t)nw, 0);
    const Hp pi = norm(Big::from_hex(kPiHex), -382, P);
    const Hp two = from_double(2.0, P), half = from_double(0.5, P);
    const Hp two_sqrt_2pi = hp_div(two, hp_sqrt(Hp{pi.m, pi.e + 1}, P), P);
    const Hp sqrt_1_2 = hp_sqrt(half, P);
    Hp s128 = from_double((double)sigma, P);
    if (blinding == SCGPU_BLINDING_SAMPLES) s128 = hp_mul(s128, sqrt_1_2, P);
    const Hp d = hp_mul(hp_div(norm(Big(1), P, P), s128, P), two_sqrt_2pi, P);
    const Hp e = hp_div(half, hp_mul(s128, s128, P), P);                       // magnitude of the (negative) exponent scale
    Hp s = Hp{d.m, d.e - 1};
    for (size_t i = 1; i + 1 < size; i++) {
        const Big ip = to_int(s);
        for (int j = 0; j < nw; j++) {
            uint64_t v = 0;
            if ((size_t)(2 * j) < ip.w.size()) v |= ip.w[(size_t)(2 * j)];
            if ((size_t)(2 * j + 1) < ip.w.size()) v |= (uint64_t)ip.w[(size_t)(2 * j + 1)] << 32;
            out[i * (size_t)nw + (size_t)j] = v;
        }
        const Hp x = hp_mul(e, norm(Big((uint64_t)i * (uint64_t)i), 0, P), P);  // trunc(e i^2): an exact integer times e
        s = hp_add(s, hp_mul(d, hp_exp_neg(x, P), P), P);
    }
    for (int j = 0; j < nw; j++) out[(size - 1) * (size_t)nw + (size_t)j] = ~0ull;
    return out;
}

}  // namespace scgpu

// Host-only (no device needed): writes the table into `table` (capacity in entries) and its size into *entries.
extern "C" int scgpu_gauss_cdf_table_high(uint64_t *table, size_t capacity, size_t *entries, int precision, int blinding,
                                          float tail, float sigma)
{
    using namespace scgpu;
    if (!entries) { set_error("gauss_cdf_table_high: null argument"); return SCGPU_ERR_ARG; }
    if (precision != 128 && precision != 192 && precision != 256) { set_error("gauss_cdf_table_high: precision %d (128, 192 or 256)", precision); return SCGPU_ERR_ARG; }
    if (!(sigma > 0) || !(tail > 0) || !(tail * sigma >= 2.0f) || tail * sigma > 16777216.0f) { set_error("gauss_cdf_table_high: tail / sigma out of range"); return SCGPU_ERR_ARG; }
    if (blinding < 0 || blinding > 2) { set_error("gauss_cdf_table_high: blinding %d", blinding); return SCGPU_ERR_ARG; }
    size_t size = 1;
    while (size < (size_t)(tail * sigma)) size <<= 1;
    *entries = size;
    if (!table) return SCGPU_OK;                             // size query
    if (capacity < size) { set_error("gauss_cdf_table_high: %zu entries needed, %zu given", size, capacity); return SCGPU_ERR_ARG; }
    const std::vector<uint64_t> t = build_cdf_high(precision, blinding, tail, sigma);
    memcpy(table, t.data(), t.size() * sizeof(uint64_t));
    return SCGPU_OK;
}
