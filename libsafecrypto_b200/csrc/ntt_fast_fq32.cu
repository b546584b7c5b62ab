// ntt_fast_fq32.cu -- fused negacyclic product, float-quotient arithmetic (fq_arith.cuh), WARP-LOCAL schedule:
// 32 coefficients per thread, n/32 threads per polynomial (8 / 16 / 32 lanes of ONE warp for n = 256 / 512 /
// 1024), so a whole product needs one shared-memory exchange per transform and only __syncwarp():
//
//   pass 0   stages 0..4   thread tau holds elements tau + (n/32) m, m = 0..31.  The 31 twiddles of these
//                          stages are the same for every thread: they live in the kernel's constant bank.
//   pass 1   stages 5..    thread tau holds the 32 contiguous elements 32 tau .. 32 tau + 31, i.e. 32 / SUB
//                          independent sub-chunks of SUB = n/32 elements; both operands of a sub-chunk are
//                          transformed together (shared twiddles), multiplied pointwise and taken back through
//                          the inverse stages before the next sub-chunk is touched.
//
// Against the 8-coefficient schedule of ntt_fast_fq.cu (per product, n = 512): 2 exchanges instead of 6
// (96 instead of 192 shared-memory wavefronts), no named barriers, address arithmetic amortised over 4x more
// butterflies, 16 independent butterflies per stage per thread.  The tile is padded (4 words per 32 elements)
// so that both the strided 32-bit accesses of pass 0 and the 128-bit accesses of pass 1 are conflict-free and
// every address is  thread base + compile-time offset.  Pass-1 twiddles are stored thread-major per stage
// (fq32_slot) so that a warp's 128-bit twiddle loads are contiguous.
#include "scgpu_internal.h"
#include "fq_host.h"
#include "../../include/scgpu.h"

#include <cstdlib>
#include <vector>

namespace scgpu {

namespace {

using fq::Tw;
using fq::kBias;
typedef uint32_t u32;

constexpr int kThreads32 = 128;
#ifndef FQ32_MINB
#define FQ32_MINB 5
#endif

template <int LOGN>
struct Cfg32 {
    static constexpr int N = 1 << LOGN;
    static constexpr int T = N / 32;             // threads per polynomial
    static constexpr int PW = 32 / T;            // polynomials per warp
    static constexpr int SUB = N / 32 > 16 ? 16 : N / 32;   // elements of one pass-1 sub-chunk
    static constexpr int NSUB = 32 / SUB;
    // first stage of the sub-chunk loop.  n = 1024: stage 5 couples all 32 elements of a thread; it runs as a
    // separate step over the whole chunk (one operand at a time) so that the sub-chunk loop holds 2 x 16
    // coefficients instead of 2 x 32 (168 -> 96 registers, 3 -> 5 CTAs per SM)
    static constexpr int S1 = LOGN == 10 ? 6 : 5;
    static constexpr int TS = N + N / 8 + (T & 31);   // tile stride in words (bank offset T between polynomials)
    static constexpr int POLYS = (kThreads32 / 32) * PW;
};

struct Fq32Const {
    const int32_t *pf;                           // pass-1 forward table, thread-major (fq32_slot): w | wq | k | c,
    const int32_t *pi;                           // pass-1 inverse table                      n words each
    Tw f0[31], i0[31];                           // entries 1..31: stages 0..4
    Tw ninv, one;
    int32_t q, nq, x0, pwk, kf, ki;
    float invq;
    uint32_t M;
    int r0;                                      // reduce every coefficient at the entry of inverse pass 0
};

__device__ __forceinline__ int32_t bred(int32_t p, const Fq32Const &c)
{
    const int32_t qe = (int32_t)(((int64_t)p * (int64_t)c.M + 0x80000000ll) >> 32);
    return qe * c.nq + p;
}
__device__ __forceinline__ bool out_of_range(int32_t v, const Fq32Const &c)
{
    return ((u32)v + (u32)c.x0) > (u32)(2 * c.x0);
}

__device__ __forceinline__ void ct(u32 &lo, u32 &hi, const Tw &z, int32_t nq)
{
    const u32 t = (u32)fq::mul((int32_t)hi, z, nq);
    hi = lo - t;
    lo = lo + t;
}
__device__ __forceinline__ void gs(u32 &lo, u32 &hi, const Tw &z, int32_t nq)
{
    const u32 d = lo - hi + (u32)kBias;
    lo = lo + hi - (u32)kBias;
    hi = (u32)fq::mul((int32_t)d, z, nq);
}

// pass-0 entry from the constant bank: w and wq are used as constant operands, k and c come as one 64-bit load
__device__ __forceinline__ Tw cb_entry(const Tw &e)
{
    const int2 kc = *reinterpret_cast<const int2 *>(&e.k);
    Tw t;
    t.w = e.w; t.wq = e.wq; t.k = kc.x; t.c = __int_as_float(kc.y);
    return t;
}

// padded tile position of element e
__host__ __device__ constexpr int pos32(int e) { return e + 4 * (e >> 5); }

// ---- pass 0: stages 0..4 on x[m] = element tau + T m ---------------------------------------------------
__device__ __forceinline__ void fwd_pass0(u32 (&x)[32], const Fq32Const &c)
{
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int half = 16 >> s;
#pragma unroll
        for (int m = 0; m < 32; m++)
            if ((m & half) == 0) ct(x[m], x[m + half], cb_entry(c.f0[(1 << s) - 1 + (m >> (5 - s))]), c.nq);
    }
}

// stages 4..1, then stage 0 with n^-1 folded into both branches; returns canonical residues
__device__ __forceinline__ void inv_pass0(u32 (&x)[32], const Fq32Const &c)
{
#pragma unroll
    for (int s = 4; s >= 1; s--) {
        const int half = 16 >> s;
#pragma unroll
        for (int m = 0; m < 32; m++)
            if ((m & half) == 0) gs(x[m], x[m + half], cb_entry(c.i0[(1 << s) - 1 + (m >> (5 - s))]), c.nq);
    }
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const u32 s = x[m] + x[m + 16] - (u32)kBias;
        const u32 d = x[m] - x[m + 16] + (u32)kBias;
        const u32 ys = (u32)fq::mul((int32_t)s, c.ninv, c.nq);
        const u32 yd = (u32)fq::mul((int32_t)d, c.i0[0], c.nq);
        x[m] = min(ys, ys + (u32)c.q);
        x[m + 16] = min(yd, yd + (u32)c.q);
    }
}

// ---- pass 1 -------------------------------------------------------------------------------------------
// CNT consecutive entries r0 .. r0 + CNT - 1 of thread tau for stage S (CNT in 1, 2, 4)
template <int LOGN, int S, int CNT>
__device__ __forceinline__ void load_entries(Tw (&tw)[CNT], const int32_t *tab, int tau, int r0)
{
    using C = Cfg32<LOGN>;
    constexpr int LEN = C::N >> (S + 1);
    constexpr int G = 16 / LEN;
    constexpr int V = G < 4 ? G : 4;
    constexpr int N = C::N;
    const int32_t *p = tab + (1 << S) + ((r0 / V) * C::T + tau) * V + (r0 % V);
    if (CNT == 4) {
        const int4 w = __ldg(reinterpret_cast<const int4 *>(p));
        const int4 f = __ldg(reinterpret_cast<const int4 *>(p + N));
        const int4 k = __ldg(reinterpret_cast<const int4 *>(p + 2 * N));
        const int4 e = __ldg(reinterpret_cast<const int4 *>(p + 3 * N));
        tw[0] = Tw{w.x, __int_as_float(f.x), k.x, __int_as_float(e.x)};
        tw[1 % CNT] = Tw{w.y, __int_as_float(f.y), k.y, __int_as_float(e.y)};
        tw[2 % CNT] = Tw{w.z, __int_as_float(f.z), k.z, __int_as_float(e.z)};
        tw[3 % CNT] = Tw{w.w, __int_as_float(f.w), k.w, __int_as_float(e.w)};
    } else if (CNT == 2) {
        const int2 w = __ldg(reinterpret_cast<const int2 *>(p));
        const int2 f = __ldg(reinterpret_cast<const int2 *>(p + N));
        const int2 k = __ldg(reinterpret_cast<const int2 *>(p + 2 * N));
        const int2 e = __ldg(reinterpret_cast<const int2 *>(p + 3 * N));
        tw[0] = Tw{w.x, __int_as_float(f.x), k.x, __int_as_float(e.x)};
        tw[1 % CNT] = Tw{w.y, __int_as_float(f.y), k.y, __int_as_float(e.y)};
    } else {
        tw[0] = Tw{__ldg(p), __int_as_float(__ldg(p + N)), __ldg(p + 2 * N), __int_as_float(__ldg(p + 3 * N))};
    }
}

// one radix-2 stage S (forward: Cooley-Tukey, inverse: Gentleman-Sande) on sub-chunk h of NOPS operands
template <int LOGN, int S, int NOPS, bool INV, int CH>
__device__ __forceinline__ void stage1(u32 (&xa)[CH], u32 (&xb)[CH], const Fq32Const &c, int tau, int h)
{
    using C = Cfg32<LOGN>;
    constexpr int LEN = C::N >> (S + 1);
    constexpr int CNT = CH / (2 * LEN);              // twiddles of this (sub-)chunk in this stage
    constexpr int GRP = CNT < 4 ? CNT : 4;
    const int32_t *tab = INV ? c.pi : c.pf;
#pragma unroll
    for (int g0 = 0; g0 < CNT; g0 += GRP) {
        Tw tw[GRP];
        load_entries<LOGN, S, GRP>(tw, tab, tau, h * CNT + g0);
#pragma unroll
        for (int g = 0; g < GRP; g++) {
#pragma unroll
            for (int j = 0; j < LEN; j++) {
                const int i = (g0 + g) * 2 * LEN + j;
                if (INV) {
                    gs(xa[i], xa[i + LEN], tw[g], c.nq);
                } else {
                    ct(xa[i], xa[i + LEN], tw[g], c.nq);
                    if (NOPS == 2) ct(xb[i], xb[i + LEN], tw[g], c.nq);
                }
            }
        }
    }
}

template <int LOGN, int S, int NOPS>
__device__ __forceinline__ void fwd_stages1(u32 (&xa)[Cfg32<LOGN>::SUB], u32 (&xb)[Cfg32<LOGN>::SUB],
                                            const Fq32Const &c, int tau, int h)
{
    if constexpr (S < LOGN) {
        stage1<LOGN, S, NOPS, false, Cfg32<LOGN>::SUB>(xa, xb, c, tau, h);
        fwd_stages1<LOGN, S + 1, NOPS>(xa, xb, c, tau, h);
    }
}
template <int LOGN, int S>
__device__ __forceinline__ void inv_stages1(u32 (&x)[Cfg32<LOGN>::SUB], const Fq32Const &c, int tau, int h)
{
    if constexpr (S >= Cfg32<LOGN>::S1) {
        stage1<LOGN, S, 1, true, Cfg32<LOGN>::SUB>(x, x, c, tau, h);
        inv_stages1<LOGN, S - 1>(x, c, tau, h);
    }
}

template <int SUB>
__device__ __forceinline__ void load_sub(const int32_t *p, u32 (&x)[SUB])
{
#pragma unroll
    for (int k = 0; k < SUB; k += 4) {
        const int4 v = *reinterpret_cast<const int4 *>(p + k);
        x[k] = (u32)v.x; x[k + 1] = (u32)v.y; x[k + 2] = (u32)v.z; x[k + 3] = (u32)v.w;
    }
}
template <int SUB>
__device__ __forceinline__ void store_sub(int32_t *p, const u32 (&x)[SUB])
{
#pragma unroll
    for (int k = 0; k < SUB; k += 4)
        *reinterpret_cast<int4 *>(p + k) = make_int4((int32_t)x[k], (int32_t)x[k + 1], (int32_t)x[k + 2], (int32_t)x[k + 3]);
}

// n = 1024 only: stage 5 (forward) / its inverse on the thread's whole 32-element chunk, in place in the tile
template <int LOGN, bool INV>
__device__ __forceinline__ void chunk_stage5(int32_t *p, const Fq32Const &c, int tau)
{
    if constexpr (Cfg32<LOGN>::S1 > 5) {
        u32 x[32];
        load_sub<32>(p, x);
        stage1<LOGN, 5, 1, INV, 32>(x, x, c, tau, 0);
        store_sub<32>(p, x);
    }
}

template <int LOGN>
__device__ __forceinline__ void load_operand(u32 (&x)[32], const int32_t *row, int tau, const Fq32Const &c)
{
    constexpr int T = Cfg32<LOGN>::T;
    int32_t v[32];
    bool wide = false;
#pragma unroll
    for (int m = 0; m < 32; m++) {
        v[m] = __ldg(row + tau + m * T);
        wide |= out_of_range(v[m], c);
    }
    if (__any_sync(0xFFFFFFFFu, wide)) {
#pragma unroll
        for (int m = 0; m < 32; m++) v[m] = bred(v[m], c);
    }
#pragma unroll
    for (int m = 0; m < 32; m++) x[m] = (u32)v[m] + (u32)kBias;
}

// same, from a raw row that a bulk copy (TMA) has staged at the start of the tile region
template <int LOGN>
__device__ __forceinline__ void load_operand_staged(u32 (&x)[32], const int32_t *raw, int tau, const Fq32Const &c)
{
    constexpr int T = Cfg32<LOGN>::T;
    int32_t v[32];
    bool wide = false;
#pragma unroll
    for (int m = 0; m < 32; m++) {
        v[m] = raw[tau + m * T];
        wide |= out_of_range(v[m], c);
    }
    if (__any_sync(0xFFFFFFFFu, wide)) {
#pragma unroll
        for (int m = 0; m < 32; m++) v[m] = bred(v[m], c);
    }
#pragma unroll
    for (int m = 0; m < 32; m++) x[m] = (u32)v[m] + (u32)kBias;
}

// ---- TMA bulk copies (cp.async.bulk) + mbarrier: the next product's operand rows are fetched into tile
// regions the current product no longer needs, by one lane per warp, while the warp computes ----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int LOGN>
__device__ __forceinline__ void store_pass0(int32_t *tile, const u32 (&x)[32], int tau)
{
    constexpr int T = Cfg32<LOGN>::T;
#pragma unroll
    for (int m = 0; m < 32; m++) tile[tau + pos32(T * m)] = (int32_t)x[m];
}
template <int LOGN>
__device__ __forceinline__ void load_pass0(const int32_t *tile, u32 (&x)[32], int tau)
{
    constexpr int T = Cfg32<LOGN>::T;
#pragma unroll
    for (int m = 0; m < 32; m++) x[m] = (u32)tile[tau + pos32(T * m)];
}

enum { FQ_POLYMUL = 0, FQ_KEY16 = 1, FQ_KEY32 = 2 };

// reference NTT-domain index of the thread's pass-1 element e (position 32 tau + e of the bit-reversed order):
// brev(32 tau + e) = brev5(e) << (LOGN - 5) | brev_{LOGN-5}(tau); for fixed e the lanes of a polynomial read
// n/32 consecutive coefficients
template <int LOGN>
__device__ __forceinline__ int ntt_index(int tau, int e)
{
    return (int)((__brev((unsigned)e) >> 27) << (LOGN - 5)) | (int)(__brev((unsigned)tau) >> (32 - (LOGN - 5)));
}

// TMA = true: operand rows arrive by bulk copy (16-byte aligned rows); false: plain LDG (any alignment)
template <int LOGN, int MODE, bool TMA>
__global__ void __launch_bounds__(kThreads32, FQ32_MINB)
k_polymul_fq32(int32_t *__restrict__ out, const int32_t *__restrict__ a, const void *__restrict__ bsrc,
               size_t b_stride, size_t count, const __grid_constant__ Fq32Const c)
{
    using C = Cfg32<LOGN>;
    constexpr int N = C::N, T = C::T, SUB = C::SUB;
    __shared__ __align__(16) int32_t tiles[2][C::POLYS][C::TS];      // slot stride TS = T (mod 32) banks
    const int lane = threadIdx.x & 31;
    const int tau = lane % T;
    const int slot = (threadIdx.x / 32) * C::PW + lane / T;       // polynomial slot inside the CTA
    int32_t *ta = tiles[0][slot];
    int32_t *tb = tiles[1][slot];
    // bulk-copy pipeline state: bars[warp][0 / 1] complete when the a / b rows of the warp's next product have
    // landed; `other` = byte offset between a polynomial's two tile regions, whose roles swap every iteration
    __shared__ __align__(8) uint64_t bars[kThreads32 / 32][2];
    const int warp = threadIdx.x / 32;
    uint32_t parity = 0;
    constexpr uint32_t ROW_BYTES = (uint32_t)N * 4u;
    auto fetch = [&](int op, size_t nbase, int32_t *region0) {
        // lane 0: rows of the warp's PW polynomials of the product group starting at nbase -> region0[p]
        mbar_expect_tx(&bars[warp][op], ROW_BYTES * C::PW);
#pragma unroll
        for (int p = 0; p < C::PW; p++) {
            size_t row = nbase + (size_t)warp * C::PW + p;
            if (row >= count) row = 0;
            const int32_t *src = op == 0 ? a + row * N : static_cast<const int32_t *>(bsrc) + row * b_stride;
            bulk_g2s(region0 + p * C::TS, src, ROW_BYTES, &bars[warp][op]);
        }
    };
    const size_t first = (size_t)blockIdx.x * C::POLYS;
    if (TMA) {
        if (lane == 0) {
            mbar_init(&bars[warp][0], 1);
            mbar_init(&bars[warp][1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0 && first < count) {
            fetch(0, first, tiles[0][warp * C::PW]);
            if (MODE == FQ_POLYMUL) fetch(1, first, tiles[1][warp * C::PW]);
        }
    }

    for (size_t base = first; base < count; base += (size_t)gridDim.x * C::POLYS) {
        const size_t poly = base + slot;
        const bool live = poly < count;
        const size_t prow = live ? poly : 0;
        // rolled loops (operand, sub-chunk): the fully unrolled body was 60 KB of SASS and spent 2 of every
        // 7 stall cycles waiting for instructions (profiles/polymul_r02c_*); the L1.5 I-cache holds 32 KB
#pragma unroll 1
        for (int op = 0; op < (MODE == FQ_POLYMUL ? 2 : 1); op++) {
            u32 x[32];
            int32_t *tile = op == 0 ? ta : tb;
            if (TMA) {
                mbar_wait(&bars[warp][op], parity);
                load_operand_staged<LOGN>(x, tile, tau, c);
                __syncwarp();                     // the padded result overwrites the raw row in place
            } else {
                const int32_t *row = op == 0 ? a + prow * N : static_cast<const int32_t *>(bsrc) + prow * b_stride;
                load_operand<LOGN>(x, row, tau, c);
            }
            fwd_pass0(x, c);
            store_pass0<LOGN>(tile, x, tau);
        }
        __syncwarp();
        if constexpr (C::S1 > 5) {
#pragma unroll 1
            for (int op = 0; op < (MODE == FQ_POLYMUL ? 2 : 1); op++) chunk_stage5<LOGN, false>((op == 0 ? ta : tb) + 36 * tau, c, tau);
        }
#pragma unroll 1
        for (int h = 0; h < C::NSUB; h++) {
            u32 xa[SUB], xb[SUB];
            int32_t *pa = ta + 36 * tau + SUB * h;
            load_sub<SUB>(pa, xa);
            if (MODE == FQ_POLYMUL) {
                load_sub<SUB>(tb + 36 * tau + SUB * h, xb);
                fwd_stages1<LOGN, C::S1, 2>(xa, xb, c, tau, h);
#pragma unroll
                for (int i = 0; i < SUB; i++)
                    xa[i] = (u32)fq::mul_var((int32_t)(xa[i] - (u32)kBias), (int32_t)(xb[i] - (u32)kBias), c.invq, c.pwk, c.nq) + (u32)kBias;
            } else {
                fwd_stages1<LOGN, C::S1, 1>(xa, xb, c, tau, h);
                int32_t kv[SUB];
                bool wide = false;
#pragma unroll
                for (int i = 0; i < SUB; i++) {
                    const int j = ntt_index<LOGN>(tau, SUB * h + i);
                    if (MODE == FQ_KEY16) kv[i] = (int32_t)__ldg(static_cast<const int16_t *>(bsrc) + prow * b_stride + j);
                    else {
                        kv[i] = __ldg(static_cast<const int32_t *>(bsrc) + prow * b_stride + j);
                        wide |= out_of_range(kv[i], c);
                    }
                }
                if (MODE == FQ_KEY32 && __any_sync(0xFFFFFFFFu, wide)) {
#pragma unroll
                    for (int i = 0; i < SUB; i++) kv[i] = bred(kv[i], c);
                }
#pragma unroll
                for (int i = 0; i < SUB; i++)
                    xa[i] = (u32)fq::mul_var((int32_t)(xa[i] - (u32)kBias), kv[i], c.invq, c.pwk, c.nq) + (u32)kBias;
            }
            inv_stages1<LOGN, LOGN - 1>(xa, c, tau, h);
            store_sub<SUB>(pa, xa);
        }
        chunk_stage5<LOGN, true>(ta + 36 * tau, c, tau);
        const size_t nbase = base + (size_t)gridDim.x * C::POLYS;
        if (TMA) fence_proxy_async();
        __syncwarp();
        // tb is free from here on: the next product's a rows go there
        if (TMA && lane == 0 && nbase < count) fetch(0, nbase, tb - (lane / T) * C::TS);
        {
            u32 x[32];
            load_pass0<LOGN>(ta, x, tau);
            if (TMA && MODE == FQ_POLYMUL) {
                fence_proxy_async();
                __syncwarp();
                // every coefficient is in registers: ta takes the next product's b rows
                if (lane == 0 && nbase < count) fetch(1, nbase, ta - (lane / T) * C::TS);
            }
            if (c.r0) {
#pragma unroll
                for (int m = 0; m < 32; m++) x[m] = (u32)fq::mul((int32_t)x[m], c.one, c.nq);
            }
            inv_pass0(x, c);
            if (live) {
                int32_t *orow = out + poly * N;
#pragma unroll
                for (int m = 0; m < 32; m++) orow[tau + m * T] = (int32_t)x[m];
            }
        }
        if (TMA) {
            int32_t *t = ta; ta = tb; tb = t;         // the regions swap roles
            parity ^= 1u;
        } else {
            __syncwarp();
        }
    }
}


// ---- module-LWE matrix-vector product  t_i = INTT(sum_j A_ij o NTT(s_j))  (module_lwe.c:588-748) -------------
// Same warp-local schedule.  The l transformed vectors stay in shared memory (padded pass-1 layout, unbiased),
// each output row accumulates its l pointwise products in registers.  HBM traffic is dominated by A
// (k l rows per instance against l + k for s and t), so with TMA = true every row moves by bulk copy:
//   * A_ij rows (contiguous n words) land in a one-row staging buffer per instance; as soon as a thread group
//     has pulled its 32 coefficients into registers the next row (or the next instance's first row) is in
//     flight, i.e. a row's DRAM latency is covered by one whole pointwise step;
//   * the next instance's s rows are fetched into the stash tiles during the last inverse transform.
// The reference's NTT-domain order is our bit-reversed one: element e of thread tau is coefficient
// brev5(e) * (n/32) + brev(tau) of the row (ntt_index), so for a fixed e the lanes of an instance read n/32
// consecutive words: conflict-free from the staging row (instances 8 banks apart), one full sector from HBM
// in the LDG variant.
template <int LOGN, bool TMA>
__global__ void __launch_bounds__(kThreads32)
k_matvec_fq32(int32_t *__restrict__ out, const int32_t *__restrict__ A, const int32_t *__restrict__ s,
              int k, int l, size_t count, const __grid_constant__ Fq32Const c)
{
    using C = Cfg32<LOGN>;
    constexpr int N = C::N, T = C::T, SUB = C::SUB, NSUB = C::NSUB;
    constexpr int AROW = N + T;                                      // staging row stride: T banks between instances
    constexpr uint32_t ROW_BYTES = (uint32_t)N * 4u;
    extern __shared__ __align__(16) int32_t dyn_tiles[];             // [l + 1][POLYS][TS], then [POLYS][AROW]
    __shared__ __align__(8) uint64_t bars[kThreads32 / 32][2];       // [warp][0: s rows, 1: A row]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x / 32;
    const int tau = lane % T;
    const int slot = warp * C::PW + lane / T;
    int32_t *xt = dyn_tiles + slot * C::TS;                          // exchange tile of the inverse transform
    int32_t *astage = dyn_tiles + (size_t)(l + 1) * C::POLYS * C::TS + slot * AROW;
    const int taurev = (int)(__brev((unsigned)tau) >> (32 - (LOGN - 5)));
    uint32_t par_s = 0, par_a = 0;

    // lane 0: rows of the warp's PW instances (clamped to instance 0 beyond the batch)
    auto fetch_s = [&](size_t nbase) {
        mbar_expect_tx(&bars[warp][0], ROW_BYTES * C::PW * (uint32_t)l);
        for (int p = 0; p < C::PW; p++) {
            size_t row = nbase + (size_t)warp * C::PW + p;
            if (row >= count) row = 0;
            for (int j = 0; j < l; j++)
                bulk_g2s(dyn_tiles + ((size_t)(j + 1) * C::POLYS + warp * C::PW + p) * C::TS, s + (row * l + j) * N, ROW_BYTES, &bars[warp][0]);
        }
    };
    auto fetch_a = [&](size_t nbase, int step) {
        mbar_expect_tx(&bars[warp][1], ROW_BYTES * C::PW);
        for (int p = 0; p < C::PW; p++) {
            size_t row = nbase + (size_t)warp * C::PW + p;
            if (row >= count) row = 0;
            bulk_g2s(dyn_tiles + (size_t)(l + 1) * C::POLYS * C::TS + (warp * C::PW + p) * AROW,
                     A + (row * k * l + step) * N, ROW_BYTES, &bars[warp][1]);
        }
    };
    const size_t first = (size_t)blockIdx.x * C::POLYS;
    if (TMA) {
        if (lane == 0) {
            mbar_init(&bars[warp][0], 1);
            mbar_init(&bars[warp][1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0 && first < count) { fetch_s(first); fetch_a(first, 0); }
    }

    for (size_t base = first; base < count; base += (size_t)gridDim.x * C::POLYS) {
        const size_t inst = base + slot;
        const bool live = inst < count;
        const size_t irow = live ? inst : 0;
        const size_t nbase = base + (size_t)gridDim.x * C::POLYS;
        if (TMA) { mbar_wait(&bars[warp][0], par_s); par_s ^= 1u; }
#pragma unroll 1
        for (int j = 0; j < l; j++) {
            u32 x[32];
            int32_t *tile = dyn_tiles + ((size_t)(j + 1) * C::POLYS + slot) * C::TS;
            if (TMA) {
                load_operand_staged<LOGN>(x, tile, tau, c);
                __syncwarp();
            } else {
                load_operand<LOGN>(x, s + (irow * l + j) * N, tau, c);
            }
            fwd_pass0(x, c);
            store_pass0<LOGN>(tile, x, tau);
        }
        __syncwarp();
#pragma unroll 1
        for (int j = 0; j < l; j++) {
#pragma unroll 1
            for (int h = 0; h < NSUB; h++) {
                u32 xa[SUB], xb[SUB];
                int32_t *p = dyn_tiles + ((size_t)(j + 1) * C::POLYS + slot) * C::TS + 36 * tau + SUB * h;
                load_sub<SUB>(p, xa);
                fwd_stages1<LOGN, C::S1, 1>(xa, xb, c, tau, h);
#pragma unroll
                for (int i = 0; i < SUB; i++) xa[i] -= (u32)kBias;
                store_sub<SUB>(p, xa);                               // only this thread reads it again
            }
        }
#pragma unroll 1
        for (int i = 0; i < k; i++) {
            u32 acc[NSUB][SUB];
#pragma unroll
            for (int h = 0; h < NSUB; h++)
#pragma unroll
                for (int e = 0; e < SUB; e++) acc[h][e] = (u32)kBias;
#pragma unroll 1
            for (int j = 0; j < l; j++) {
                int32_t av[32];
                bool wide = false;
                if (TMA) {
                    mbar_wait(&bars[warp][1], par_a); par_a ^= 1u;
#pragma unroll
                    for (int e = 0; e < 32; e++) {
                        av[e] = astage[taurev + (int)((__brev((unsigned)e) >> 27) << (LOGN - 5))];
                        wide |= out_of_range(av[e], c);
                    }
                    // the staging row is in registers: put the next row (this instance's next step, or step 0
                    // of the warp's next instances) in flight before the arithmetic
                    fence_proxy_async();
                    __syncwarp();
                    const int step = i * l + j + 1;
                    if (lane == 0) {
                        if (step < k * l) fetch_a(base, step);
                        else if (nbase < count) fetch_a(nbase, 0);
                    }
                } else {
                    const int32_t *arow = A + ((irow * k + i) * l + j) * N + taurev;
#pragma unroll
                    for (int e = 0; e < 32; e++) {
                        av[e] = __ldg(arow + (int)((__brev((unsigned)e) >> 27) << (LOGN - 5)));
                        wide |= out_of_range(av[e], c);
                    }
                }
                // A is canonical in the reference (sampled in [0, q)); anything else is reduced first
                if (__any_sync(0xFFFFFFFFu, wide)) {
#pragma unroll
                    for (int e = 0; e < 32; e++) av[e] = bred(av[e], c);
                }
                const int32_t *sp = dyn_tiles + ((size_t)(j + 1) * C::POLYS + slot) * C::TS + 36 * tau;
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const int4 sv = *reinterpret_cast<const int4 *>(sp + e);
                    acc[e / SUB][e % SUB]           += (u32)fq::mul_var(av[e], sv.x, c.invq, c.pwk, c.nq);
                    acc[(e + 1) / SUB][(e + 1) % SUB] += (u32)fq::mul_var(av[e + 1], sv.y, c.invq, c.pwk, c.nq);
                    acc[(e + 2) / SUB][(e + 2) % SUB] += (u32)fq::mul_var(av[e + 2], sv.z, c.invq, c.pwk, c.nq);
                    acc[(e + 3) / SUB][(e + 3) % SUB] += (u32)fq::mul_var(av[e + 3], sv.w, c.invq, c.pwk, c.nq);
                }
            }
            if (TMA && i == k - 1) {
                // the stash is dead: the next instances' s rows travel during the last inverse transform
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && nbase < count) fetch_s(nbase);
            }
#pragma unroll
            for (int h = 0; h < NSUB; h++) {
                inv_stages1<LOGN, LOGN - 1>(acc[h], c, tau, h);
                store_sub<SUB>(xt + 36 * tau + SUB * h, acc[h]);
            }
            __syncwarp();
            {
                u32 x[32];
                load_pass0<LOGN>(xt, x, tau);
                if (c.r0) {
#pragma unroll
                    for (int m = 0; m < 32; m++) x[m] = (u32)fq::mul((int32_t)x[m], c.one, c.nq);
                }
                inv_pass0(x, c);
                if (live) {
                    int32_t *orow = out + (inst * k + i) * N;
#pragma unroll
                    for (int m = 0; m < 32; m++) orow[tau + m * T] = (int32_t)x[m];
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace

// position of entry r (0 .. 16/len - 1) of thread tau in stage s of the thread-major pass-1 table
static int fq32_slot(int logn, int s, int tau, int r)
{
    const int n = 1 << logn, T = n / 32, len = n >> (s + 1), G = 16 / len, V = G < 4 ? G : 4;
    return (1 << s) + ((r / V) * T + tau) * V + (r % V);
}

int build_fq32_tables(NttPlanDev &p, const int32_t *w_host)
{
    p.fq32_ok = 0; p.fq32_tab = nullptr;
    if (p.logn < 8 || p.logn > 10) return SCGPU_OK;
    int r0 = 0; int32_t x0 = 0;
    if (!fq::analyse32(p.logn, p.rc.q, 1, &r0, &x0)) return SCGPU_OK;
    int r0_mv = 0; int32_t x0_mv = 0;
    p.fq32_mv_ok = fq::analyse32(p.logn, p.rc.q, 8, &r0_mv, &x0_mv) ? 1 : 0;     // up to 8 accumulated products
    p.fq32_r0_mv = r0_mv;
    std::vector<Tw> zf, zi;
    Tw ninv, one;
    if (!fq::build_tables(p.logn, p.rc.q, w_host, zf, zi, ninv, one)) return SCGPU_OK;
    const int n = p.n, T = n / 32;
    // forward [w | wq | k | c] then inverse [w | wq | k | c], n words each, entries of stages >= 5 thread-major
    std::vector<int32_t> pack(8 * n, 0);
    for (int s = 5; s < p.logn; s++) {
        const int len = n >> (s + 1), G = 16 / len;
        for (int tau = 0; tau < T; tau++)
            for (int r = 0; r < G; r++) {
                const int nat = (1 << s) + tau * G + r, at = fq32_slot(p.logn, s, tau, r);
                const Tw *src[2] = {&zf[nat], &zi[nat]};
                for (int d = 0; d < 2; d++) {
                    int32_t *dst = pack.data() + 4 * n * d + at;
                    dst[0] = src[d]->w;
                    memcpy(dst + n, &src[d]->wq, 4);
                    dst[2 * n] = src[d]->k;
                    memcpy(dst + 3 * n, &src[d]->c, 4);
                }
            }
    }
    SCGPU_CUDA_CHECK(cudaMalloc(&p.fq32_tab, sizeof(int32_t) * 8 * n));
    SCGPU_CUDA_CHECK(cudaMemcpy(p.fq32_tab, pack.data(), sizeof(int32_t) * 8 * n, cudaMemcpyHostToDevice));
    memcpy(p.fq32_pass0, &zf[1], sizeof(Tw) * 31);
    memcpy(p.fq32_pass0 + sizeof(Tw) * 31, &zi[1], sizeof(Tw) * 31);
    memcpy(p.fq_ninv, &ninv, sizeof(Tw));
    memcpy(p.fq_one, &one, sizeof(Tw));
    p.fq32_r0 = r0;
    p.fq32_x0 = x0;
    p.fq32_ok = 1;
    return SCGPU_OK;
}

void free_fq32_tables(NttPlanDev &p)
{
    if (p.fq32_tab) cudaFree(p.fq32_tab);
    p.fq32_tab = nullptr;
    p.fq32_ok = 0;
}

static Fq32Const fq32_const(const NttPlanDev &p, int r0)
{
    Fq32Const c;
    const int n = p.n;
    c.pf = static_cast<const int32_t *>(p.fq32_tab);
    c.pi = c.pf + 4 * n;
    memcpy(c.f0, p.fq32_pass0, sizeof(Tw) * 31);
    memcpy(c.i0, p.fq32_pass0 + sizeof(Tw) * 31, sizeof(Tw) * 31);
    memcpy(&c.ninv, p.fq_ninv, sizeof(Tw));
    memcpy(&c.one, p.fq_one, sizeof(Tw));
    c.q = p.rc.q; c.nq = -p.rc.q; c.x0 = p.fq32_x0;
    c.pwk = (int32_t)((uint32_t)kBias * (uint32_t)p.rc.q);
    c.kf = c.pwk;
    c.ki = (int32_t)((uint32_t)c.pwk + (uint32_t)kBias);
    c.invq = (float)(1.0 / (double)p.rc.q);
    c.M = (uint32_t)((1ull << 32) / (uint64_t)p.rc.q);
    c.r0 = r0;
    return c;
}

// n = 256 (Kyber); returns SCGPU_ERR_UNSUPPORTED when this schedule does not apply (the caller falls back)
int launch_matvec_fq32(const NttPlanDev &p, int32_t *out, const int32_t *A, const int32_t *s, int k, int l,
                       size_t count, cudaStream_t st)
{
    if (!p.fq32_ok || !p.fq32_mv_ok || p.logn != 8 || l > 8) return SCGPU_ERR_UNSUPPORTED;
    const Fq32Const c = fq32_const(p, p.fq32_r0_mv);
    using C = Cfg32<8>;
    const char *no_tma = getenv("SCGPU_NO_TMA");
    const bool tma = ((uintptr_t)A % 16) == 0 && ((uintptr_t)s % 16) == 0 && !(no_tma && atoi(no_tma) != 0);
    const size_t smem = ((size_t)(l + 1) * C::POLYS * C::TS + (tma ? (size_t)C::POLYS * (C::N + C::T) : 0)) * sizeof(int32_t);
    static bool attr_set = false;
    if (!attr_set) {
        SCGPU_CUDA_CHECK(cudaFuncSetAttribute(k_matvec_fq32<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SCGPU_CUDA_CHECK(cudaFuncSetAttribute(k_matvec_fq32<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    const int sms = p.sm_count > 0 ? p.sm_count : 148;
    int per_sm = (int)((227 * 1024) / (smem + 1024 + 64));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    const size_t groups = (count + C::POLYS - 1) / C::POLYS;
    size_t grid = (size_t)sms * per_sm;
    if (grid > groups) grid = groups;
    if (tma) k_matvec_fq32<8, true><<<(unsigned)grid, kThreads32, smem, st>>>(out, A, s, k, l, count, c);
    else     k_matvec_fq32<8, false><<<(unsigned)grid, kThreads32, smem, st>>>(out, A, s, k, l, count, c);
    count_launch();
    SCGPU_CUDA_CHECK(cudaGetLastError());
    return SCGPU_OK;
}

int launch_polymul_fq32(const NttPlanDev &p, int mode, int32_t *out, const int32_t *a, const void *b,
                        size_t b_stride, size_t count, cudaStream_t st)
{
    const Fq32Const c = fq32_const(p, p.fq32_r0);
    const int sms = p.sm_count > 0 ? p.sm_count : 148;
    // bulk copies need 16-byte aligned rows; the polymul's second operand is only staged in FQ_POLYMUL mode
    const char *no_tma = getenv("SCGPU_NO_TMA");
    bool tma = ((uintptr_t)a % 16) == 0 && !(no_tma && atoi(no_tma) != 0);
    if (mode == FQ_POLYMUL) tma = tma && ((uintptr_t)b % 16) == 0 && (b_stride % 4) == 0;
#define FQ32_LAUNCH(L)                                                                                     \
    {                                                                                                      \
        const size_t groups = (count + Cfg32<L>::POLYS - 1) / Cfg32<L>::POLYS;                             \
        size_t grid = (size_t)sms * FQ32_MINB;                                                     \
        if (grid > groups) grid = groups;                                                                  \
        if (tma) {                                                                                         \
            if (mode == FQ_POLYMUL)    k_polymul_fq32<L, FQ_POLYMUL, true><<<(unsigned)grid, kThreads32, 0, st>>>(out, a, b, b_stride, count, c); \
            else if (mode == FQ_KEY16) k_polymul_fq32<L, FQ_KEY16, true><<<(unsigned)grid, kThreads32, 0, st>>>(out, a, b, b_stride, count, c);   \
            else                       k_polymul_fq32<L, FQ_KEY32, true><<<(unsigned)grid, kThreads32, 0, st>>>(out, a, b, b_stride, count, c);   \
        } else {                                                                                           \
            if (mode == FQ_POLYMUL)    k_polymul_fq32<L, FQ_POLYMUL, false><<<(unsigned)grid, kThreads32, 0, st>>>(out, a, b, b_stride, count, c); \
            else if (mode == FQ_KEY16) k_polymul_fq32<L, FQ_KEY16, false><<<(unsigned)grid, kThreads32, 0, st>>>(out, a, b, b_stride, count, c);   \
            else                       k_polymul_fq32<L, FQ_KEY32, false><<<(unsigned)grid, kThreads32, 0, st>>>(out, a, b, b_stride, count, c);   \
        }                                                                                                  \
    }
    switch (p.logn) {
    case 8:  FQ32_LAUNCH(8); break;
    case 9:  FQ32_LAUNCH(9); break;
    case 10: FQ32_LAUNCH(10); break;
    default: set_error("unsupported n=%d", p.n); return SCGPU_ERR_UNSUPPORTED;
    }
#undef FQ32_LAUNCH
    count_launch();
    SCGPU_CUDA_CHECK(cudaGetLastError());
    return SCGPU_OK;
}

}  // namespace scgpu
