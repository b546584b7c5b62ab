// reduce.cuh -- device-side modular reduction policies.
//
// Two families:
//  * Exact<V>: reproduces, bit for bit, what reduction variant V of libsafecrypto returns
//    (ntt_template.c.in:698-953 and the AVX2 lane code :8-67,970-1131,1164-1203,1361-1438),
//    including the lazily reduced / non-canonical representatives.  Used by every kernel whose
//    output the reference defines per variant (fwd_ntt, pointwise, modn, ...).
//  * Mont: signed Montgomery arithmetic (R = 2^32) for the fused kernels whose output is the
//    canonical residue, where any internally consistent arithmetic gives the reference's bits.
//
// Everything is INT-pipe work (IMAD / IMAD.HI / IADD3 / LOP3 / SHF); the fp and avx variants
// additionally use the FP64 / FP32 pipes exactly where the reference's C does.
#pragma once
#include <cstdint>

namespace scgpu {

enum Variant : int { V_REFERENCE = 0, V_BARRETT = 1, V_FP = 2, V_AVX = 3, V_SOL7681 = 4, V_SOL8380417 = 5 };

// Reduction constants, read from the caller's ntt_params_t (never recomputed) plus the
// derived reciprocals the exact signed remainder needs.
struct RedConst {
    int32_t q;
    int32_t m;           // Barrett multiplier (p->u.ntt32.m)
    int32_t k;           // Barrett shift      (p->u.ntt32.k)
    float qs_inv;        // (FLOAT) p->inv_q_dbl, the AVX2 lanes' single-precision reciprocal
    double inv_q_dbl;    // p->inv_q_dbl
    uint64_t recip64;    // floor((2^64 - 1) / q)
    uint32_t recip32;    // floor((2^32 - 1) / q)
    uint32_t pad;
};

__device__ __forceinline__ uint32_t ct_lt_u32(uint32_t a, uint32_t b)
{
    // sc_math.c:31-34
    return ((((a ^ b) & ((a - b) ^ b)) ^ (a - b)) & 0x80000000u) >> 31;
}

// x += q if negative, x -= q if x >= q, in the reference's branch-free form (ntt.c:594-595).
// ct_lt_u32(q, u + 1) is the unsigned comparison q < u + 1 (including the wrap of u + 1 at u = 2^32 - 1); written as
// a comparison it compiles to ISETP + select instead of the eight bit operations of the reference's branch-free C
// (a GPU select does not branch either).
__device__ __forceinline__ int32_t cond_fix(int32_t x, int32_t q)
{
    uint32_t u = (uint32_t)x;
    u += (uint32_t)q & (uint32_t)(x >> 31);
    u -= ((uint32_t)q < u + 1u) ? (uint32_t)q : 0u;
    return (int32_t)u;
}

// ---- exact signed remainder (C `%`: sign of the dividend) ------------------------------------
__device__ __forceinline__ int32_t rem_s32(int32_t x, const RedConst &c)
{
    uint32_t ax = x < 0 ? 0u - (uint32_t)x : (uint32_t)x;
    uint32_t t = __umulhi(ax, c.recip32);          // floor(ax/q) or one less
    uint32_t r = ax - t * (uint32_t)c.q;           // in [0, 2q)
    r -= (r >= (uint32_t)c.q) ? (uint32_t)c.q : 0u;
    return x < 0 ? -(int32_t)r : (int32_t)r;
}

__device__ __forceinline__ int32_t rem_s64(int64_t p, const RedConst &c)
{
    uint64_t ap = p < 0 ? 0ull - (uint64_t)p : (uint64_t)p;
    uint64_t t = __umul64hi(ap, c.recip64);        // floor(ap/q) or one less
    uint32_t r = (uint32_t)ap - (uint32_t)t * (uint32_t)c.q;   // low 32 bits suffice: r in [0, 2q)
    r -= (r >= (uint32_t)c.q) ? (uint32_t)c.q : 0u;
    return p < 0 ? -(int32_t)r : (int32_t)r;
}

// ---- Barrett, ntt.c:366-378, on the low 32 bits of the 64-bit working value -------------------
__device__ __forceinline__ int32_t barrett_s64(int64_t a, const RedConst &c)
{
    int64_t am = (int64_t)((uint64_t)a * (uint64_t)(int64_t)c.m);
    uint32_t t = (uint32_t)(am >> c.k);
    uint32_t v = (uint32_t)a - t * (uint32_t)c.q;
    return cond_fix((int32_t)v, c.q);
}

// ---- truncating double quotient, ntt_template.c.in:766-767 / 856-858 ---------------------------
__device__ __forceinline__ int32_t fp_s64(int64_t v, const RedConst &c)
{
    double quo = __dmul_rn(__ll2double_rn(v), c.inv_q_dbl);
    long long qi = __double2ll_rz(quo);
    return (int32_t)((uint32_t)v - (uint32_t)c.q * (uint32_t)qi);
}
__device__ __forceinline__ int32_t fp_s32(int32_t v, const RedConst &c)
{
    double quo = __dmul_rn(__int2double_rn(v), c.inv_q_dbl);
    long long qi = __double2ll_rz(quo);
    return (int32_t)((uint32_t)v - (uint32_t)c.q * (uint32_t)qi);
}

// ---- Solinas folds, ntt_template.c.in:707-730 / 795-820 -----------------------------------------
template <int HI, int UP>
__device__ __forceinline__ int32_t solinas_s32(int32_t x)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        int32_t high = x >> HI;
        int32_t low = x & ((1 << HI) - 1);
        x = (int32_t)((uint32_t)low - (uint32_t)high + ((uint32_t)high << UP));
    }
    return x;
}
template <int HI, int UP>
__device__ __forceinline__ int32_t solinas_s64(int64_t x)
{
#pragma unroll
    for (int i = 0; i < 3; i++) {
        int64_t high = x >> HI;
        int64_t low = x & ((1ll << HI) - 1);
        x = (int64_t)((uint64_t)low - (uint64_t)high + ((uint64_t)high << UP));
    }
    return (int32_t)x;
}

// ---- AVX2 lanes ----------------------------------------------------------------------------------
constexpr double kMagicDbl = 6755399441055744.0;              // 2^52 + 2^51
constexpr unsigned long long kMagicBits = 0x4338000000000000ull;

__device__ __forceinline__ double lane_i64_to_dbl(int64_t x)                    // :37-41
{
    return __dsub_rn(__longlong_as_double((long long)((uint64_t)x + kMagicBits)), kMagicDbl);
}
__device__ __forceinline__ double lane_i64_to_dbl_full(int64_t v)               // :52-65
{
    double lo = lane_i64_to_dbl(v & 0xFFFFFFFFll);
    double hi = lane_i64_to_dbl(v >> 32);
    return __fma_rn(4294967296.0, hi, lo);
}
// Double-precision lane: quotient = fused(td * (1/q) + magic) as gcc contracts it on FMA hosts
// (see oracle/sc_oracle_ntt.c lane_quotient), product with q taken on the quotient's low 32 bits.
__device__ __forceinline__ int32_t lane_dbl(int64_t prod, bool full_range, const RedConst &c)
{
    double td = full_range ? lane_i64_to_dbl_full(prod) : lane_i64_to_dbl(prod);
    double y = __fma_rn(td, c.inv_q_dbl, kMagicDbl);
    int64_t quo = (int64_t)((uint64_t)__double_as_longlong(y) - kMagicBits);
    int64_t res = (int64_t)((uint64_t)prod - (uint64_t)((int64_t)(int32_t)(uint32_t)quo * (int64_t)c.q));
    if (res < 0) res += c.q;
    return (int32_t)res;
}
// Single-precision lane on the low 32 bits of the product (:1081-1094, :1372-1385)
__device__ __forceinline__ int32_t lane_flt(int32_t p32, const RedConst &c)
{
    float quo_f = __fmul_rn(__int2float_rn(p32), c.qs_inv);
    int32_t quo = __float2int_rn(quo_f);
    int32_t res = (int32_t)((uint32_t)p32 - (uint32_t)quo * (uint32_t)c.q);
    if (res < 0) res = (int32_t)((uint32_t)res + (uint32_t)c.q);
    return res;
}

// The same lane for 512 < q: cvtps_epi32 rounds to nearest even, and |quo_f| <= 2^31 / q < 2^22, so adding 1.5 * 2^23
// performs exactly that rounding in the mantissa (one FADD + one IADD instead of a conversion on the XU pipe).
__device__ __forceinline__ int32_t lane_flt_magic(int32_t p32, const RedConst &c)
{
    const float quo_f = __fmul_rn(__int2float_rn(p32), c.qs_inv);
    const int32_t quo = __float_as_int(__fadd_rn(quo_f, 12582912.0f)) - 0x4B400000;
    int32_t res = (int32_t)((uint32_t)p32 - (uint32_t)quo * (uint32_t)c.q);
    if (res < 0) res = (int32_t)((uint32_t)res + (uint32_t)c.q);
    return res;
}

// ---- the policy ------------------------------------------------------------------------------------
template <int V>
struct Exact {
    // modn_32
    static __device__ __forceinline__ int32_t modn(int32_t x, const RedConst &c)
    {
        if (V == V_REFERENCE) return rem_s32(x, c);
        if (V == V_BARRETT) return barrett_s64((int64_t)x, c);
        if (V == V_SOL7681) return solinas_s32<13, 9>(x);
        if (V == V_SOL8380417) return solinas_s32<23, 13>(x);
        return fp_s32(x, c);                                    // fp and avx scalar code
    }
    // reduction of a 64-bit product (muln_32 / sqrn_32 / scalar pointwise)
    static __device__ __forceinline__ int32_t redprod(int64_t p, const RedConst &c)
    {
        if (V == V_REFERENCE) return rem_s64(p, c);
        if (V == V_BARRETT) return barrett_s64(p, c);
        if (V == V_SOL7681) return solinas_s64<13, 9>(p);
        if (V == V_SOL8380417) return solinas_s64<23, 13>(p);
        return fp_s64(p, c);
    }
    static __device__ __forceinline__ int32_t muln(int32_t x, int32_t y, const RedConst &c)
    {
        return redprod((int64_t)x * (int64_t)y, c);
    }
    // mul_32_pointwise element (:956-1015)
    static __device__ __forceinline__ int32_t pw32(int32_t x, int32_t y, const RedConst &c)
    {
        if (V == V_AVX) return lane_dbl((int64_t)x * (int64_t)y, true, c);
        return muln(x, y, c);
    }
    // mul_32_pointwise_16 element (:1018-1141); y is a sign-extended SINT16
    static __device__ __forceinline__ int32_t pw16(int32_t x, int32_t y, const RedConst &c)
    {
        if (V == V_AVX) {
            int64_t p = (int64_t)x * (int64_t)y;
            return c.q == 7681 ? lane_flt((int32_t)p, c) : lane_dbl(p, false, c);
        }
        return muln(x, y, c);
    }
    // normalize_32 element (:1840-1901)
    static __device__ __forceinline__ int32_t normalize(int32_t x, const RedConst &c)
    {
        if (V == V_AVX) return lane_flt(x, c);
        return cond_fix(modn(x, c), c.q);
    }
    // center_32 element (:1777-1838)
    static __device__ __forceinline__ int32_t center(int32_t x, const RedConst &c)
    {
        const int32_t q = c.q;
        if (V == V_AVX) {
            int32_t quo = __float2int_rn(__fmul_rn(__int2float_rn(x), c.qs_inv));
            int32_t s = (int32_t)((uint32_t)x - (uint32_t)quo * (uint32_t)q);
            if ((q >> 1) > s) s = (int32_t)((uint32_t)s - (uint32_t)q);
            if (-(q >> 1) > s) s = (int32_t)((uint32_t)s + (uint32_t)q);
            return s;
        }
        int32_t v = modn(x, c);
        const int32_t q2 = (q - 1) >> 1;
        // modn leaves |v| within a few q of zero for every variant; the reference's while loops
        while (v < -q2) v += q;
        while (v > q2) v -= q;
        return v;
    }
    // pwr_32 (:1689-1721)
    static __device__ int32_t pwr(int32_t x, int32_t e, const RedConst &c)
    {
        int32_t y = (e & 1) ? x : 1;
        e >>= 1;
        while (e > 0) {
            x = muln(x, x, c);
            int32_t cand = muln(x, y, c);
            y = (e & 1) ? cand : y;
            e >>= 1;
        }
        return y;
    }
};

// ---- signed Montgomery, R = 2^32 --------------------------------------------------------------------
// mont(x, w, wq) = x * w * 2^-32 mod q in (-q, q), for |x * w| < q * 2^31, with wq = w * q^-1 mod 2^32
// precomputed per twiddle: two IMAD.HI + one IMAD, the subtraction folds into the butterfly's IADD3.
struct MontTw { int32_t w; int32_t wq; };

__device__ __forceinline__ int32_t mont_hi(int32_t x, int32_t w) { return __mulhi(x, w); }
__device__ __forceinline__ int32_t mont_lo(int32_t x, int32_t wq, int32_t q) { return __mulhi(x * wq, q); }
__device__ __forceinline__ int32_t mont_mul(int32_t x, MontTw t, int32_t q)
{
    return mont_hi(x, t.w) - mont_lo(x, t.wq, q);
}
// data x data: x * y * 2^-32 mod q, |x * y| < q * 2^31
__device__ __forceinline__ int32_t mont_mul2(int32_t x, int32_t y, int32_t qinv, int32_t q)
{
    int32_t lo = x * y;
    return __mulhi(x, y) - __mulhi(lo * qinv, q);
}

}  // namespace scgpu
