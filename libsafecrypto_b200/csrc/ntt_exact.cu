// ntt_exact.cu -- variant-exact batched kernels for the `*_32` members of utils_arith_ntt_t.
//
// These reproduce the reference's dataflow graph (pre-twist by w, bit reversal, radix-2 DIT
// stages with twiddle w[j * n/half], lazy sums, post-twist by r, index flip) with the
// variant's own reduction, so outputs carry the same lazily reduced representative as
// ntt_template.c.in produces on the CPU.  Integer butterflies are independent within a stage,
// so evaluating a stage across n/2 threads gives the same bits as the reference's loop order.
//
// Layout: one polynomial per CTA at a time, resident in shared memory across all log2(n)
// stages; n/2 threads, one butterfly per thread per stage; w and r staged once per CTA in
// shared memory; 128-bit coalesced global loads/stores; persistent grid-stride over the batch.
#include "scgpu_internal.h"
#include "../../include/scgpu.h"

namespace scgpu {

namespace {

template <int LOGN>
__device__ __forceinline__ int brev(int i) { return (int)(__brev((unsigned)i) >> (32 - LOGN)); }

// One DIT pass over v[] (bit-reversed order in, natural order out).  ntt_template.c.in:
// fft_32 :1144-1244, large_fft_32 :1246-1339, fft_16 :1341-1482, large_fft_16 :1484-1539.
// TW16 / LARGE are compile-time and the stage loop is unrolled: the generic runtime-flag version spent a third of
// its issue slots on branches, predicate and uniform-datapath bookkeeping (profiles/exact_fwd_r03_ncu_mix.txt).
template <int V, int LOGN, bool TW16, bool LARGE>
__device__ __forceinline__ void dit_fft_t(int32_t *v, const int32_t *w, const RedConst &c)
{
    constexpr int N = 1 << LOGN;
    const int t = threadIdx.x;
#pragma unroll
    for (int s = 0; s < LOGN; s++) {
        const int half = 1 << s;
        const int j = t & (half - 1);
        const int k = j + ((t >> s) << (s + 1));
        int32_t lo = v[k], hi = v[k + half], x;
        // AVX2 build: stages with half < n/8 of fft_32, large_fft_32 and fft_16 run the vector
        // lanes on every column (j = 0 included) and never re-reduce the sums.
        const bool vec = (V == V_AVX) && (half < (N >> 3)) && !(LARGE && TW16);
        bool reduce_out = LARGE;
        if (vec) {
            const int32_t y = w[j << (LOGN - s)];
            int64_t prod = (int64_t)hi * (int64_t)y;
            if (TW16) x = (c.q <= 12289) ? lane_flt((int32_t)prod, c) : lane_dbl(prod, false, c);
            else      x = lane_dbl(prod, LARGE, c);
            reduce_out = false;
        } else if (LARGE && !TW16) {
            x = Exact<V>::muln(hi, w[j << (LOGN - s)], c);                 // every column multiplied
        } else {
            // column j = 0 is not multiplied: reduced only (fft_32, large_fft_16) or passed through (fft_16)
            const int32_t x0 = (!TW16 || LARGE) ? Exact<V>::modn(hi, c) : hi;
            if (half == 1) x = x0;                                           // stage 0: j = 0 for every thread
            else {
                const int32_t xm = Exact<V>::muln(hi, w[j << (LOGN - s)], c);
                x = (j == 0) ? x0 : xm;
            }
        }
        int32_t d = (int32_t)((uint32_t)lo - (uint32_t)x);
        int32_t a = (int32_t)((uint32_t)lo + (uint32_t)x);
        if (reduce_out) { d = Exact<V>::modn(d, c); a = Exact<V>::modn(a, c); }
        v[k + half] = d;
        v[k] = a;
        __syncthreads();
    }
}

template <int V, int LOGN, bool LARGE>
__device__ __forceinline__ void dit_fft(int32_t *v, const int32_t *w, const RedConst &c, bool tw16)
{
    if (tw16) dit_fft_t<V, LOGN, true, LARGE>(v, w, c);
    else      dit_fft_t<V, LOGN, false, LARGE>(v, w, c);
}

// global row -> shared, optional pre-twist, optional bit reversal.  Each thread moves two
// coefficients per 64-bit access pair (n/2 threads): coalesced 8-byte loads.
template <int V, int LOGN>
__device__ __forceinline__ void load_row(int32_t *v, const int32_t *src, const int32_t *tw, bool twist,
                                         bool tw16, bool bitrev, const RedConst &c)
{
    const int t = threadIdx.x;
    int2 x = reinterpret_cast<const int2 *>(src)[t];
    int i0 = 2 * t, i1 = 2 * t + 1;
    if (twist) {
        x.x = tw16 ? Exact<V>::pw16(x.x, tw[i0], c) : Exact<V>::pw32(x.x, tw[i0], c);
        x.y = tw16 ? Exact<V>::pw16(x.y, tw[i1], c) : Exact<V>::pw32(x.y, tw[i1], c);
    }
    if (bitrev) { i0 = brev<LOGN>(i0); i1 = brev<LOGN>(i1); }
    v[i0] = x.x;
    v[i1] = x.y;
}

// post-twist by r, flip (ntt.c:571-604) and store
template <int V, int LOGN>
__device__ __forceinline__ void store_inverse(int32_t *dst, int32_t *v, const int32_t *r, bool tw16, const RedConst &c)
{
    constexpr int N = 1 << LOGN;
    const int t = threadIdx.x;
    int2 o;
    {
        int i = 2 * t;                         // out[i] = fix(i ? v'[n-i] : -v'[0]),  v'[j] = pw(v[j], r[j])
        int j = (N - i) & (N - 1);
        int32_t x = tw16 ? Exact<V>::pw16(v[j], r[j], c) : Exact<V>::pw32(v[j], r[j], c);
        if (i == 0) x = (int32_t)(0u - (uint32_t)x);
        o.x = cond_fix(x, c.q);
    }
    {
        int i = 2 * t + 1;
        int j = N - i;
        int32_t x = tw16 ? Exact<V>::pw16(v[j], r[j], c) : Exact<V>::pw32(v[j], r[j], c);
        o.y = cond_fix(x, c.q);
    }
    reinterpret_cast<int2 *>(dst)[t] = o;
}

template <int LOGN>
__device__ __forceinline__ void store_row(int32_t *dst, const int32_t *v)
{
    const int t = threadIdx.x;
    reinterpret_cast<int2 *>(dst)[t] = make_int2(v[2 * t], v[2 * t + 1]);
}

template <int V, int LOGN>
__global__ void __launch_bounds__(1 << (LOGN - 1))
k_transform(ExactArgs g)
{
    constexpr int N = 1 << LOGN;
    __shared__ __align__(16) int32_t sw[N];
    __shared__ __align__(16) int32_t sr[N];
    __shared__ __align__(16) int32_t va[N];
    __shared__ __align__(16) int32_t vb[N];
    const int t = threadIdx.x;
    const RedConst c = g.rc;
    const bool tw16 = g.tw_bits == 16;
    sw[2 * t] = g.w[2 * t]; sw[2 * t + 1] = g.w[2 * t + 1];
    if (g.r) { sr[2 * t] = g.r[2 * t]; sr[2 * t + 1] = g.r[2 * t + 1]; }
    __syncthreads();

    for (size_t row = blockIdx.x; row < g.count; row += gridDim.x) {
        const int32_t *a = static_cast<const int32_t *>(g.a) + row * N;
        int32_t *out = g.out + row * N;
        switch (g.op) {
        case SCGPU_OP_FWD:
        case SCGPU_OP_FWD_LARGE:
            load_row<V, LOGN>(va, a, sw, true, tw16, true, c);
            __syncthreads();
            if (g.op == SCGPU_OP_FWD_LARGE) dit_fft<V, LOGN, true>(va, sw, c, tw16);
            else                            dit_fft<V, LOGN, false>(va, sw, c, tw16);
            store_row<LOGN>(out, va);
            break;
        case SCGPU_OP_FFT:
        case SCGPU_OP_FFT_LARGE:
            load_row<V, LOGN>(va, a, sw, false, tw16, false, c);
            __syncthreads();
            if (g.op == SCGPU_OP_FFT_LARGE) dit_fft<V, LOGN, true>(va, sw, c, tw16);
            else                            dit_fft<V, LOGN, false>(va, sw, c, tw16);
            store_row<LOGN>(out, va);
            break;
        case SCGPU_OP_INV:
        case SCGPU_OP_INV_LARGE:
            load_row<V, LOGN>(va, a, sw, false, tw16, true, c);
            __syncthreads();
            if (g.op == SCGPU_OP_INV_LARGE) dit_fft<V, LOGN, true>(va, sw, c, tw16);
            else                            dit_fft<V, LOGN, false>(va, sw, c, tw16);
            store_inverse<V, LOGN>(out, va, sr, tw16, c);
            break;
        case SCGPU_OP_POLYMUL: {
            const int32_t *b = static_cast<const int32_t *>(g.b) + row * g.b_stride;
            load_row<V, LOGN>(va, a, sw, true, tw16, true, c);
            load_row<V, LOGN>(vb, b, sw, true, tw16, true, c);
            __syncthreads();
            dit_fft<V, LOGN, false>(va, sw, c, tw16);
            dit_fft<V, LOGN, false>(vb, sw, c, tw16);
            // mul_32_pointwise, then the inverse transform's own shuffle
            int32_t p0 = Exact<V>::pw32(va[2 * t], vb[2 * t], c);
            int32_t p1 = Exact<V>::pw32(va[2 * t + 1], vb[2 * t + 1], c);
            __syncthreads();
            va[brev<LOGN>(2 * t)] = p0;
            va[brev<LOGN>(2 * t + 1)] = p1;
            __syncthreads();
            dit_fft<V, LOGN, false>(va, sw, c, tw16);
            store_inverse<V, LOGN>(out, va, sr, tw16, c);
        } break;
        case SCGPU_OP_TRIPLE16: {
            const int16_t *key = static_cast<const int16_t *>(g.b) + row * g.b_stride;
            load_row<V, LOGN>(va, a, sw, true, true, true, c);
            __syncthreads();
            dit_fft_t<V, LOGN, true, false>(va, sw, c);
            int32_t p0 = Exact<V>::pw16(va[2 * t], key[2 * t], c);
            int32_t p1 = Exact<V>::pw16(va[2 * t + 1], key[2 * t + 1], c);
            __syncthreads();
            va[brev<LOGN>(2 * t)] = p0;
            va[brev<LOGN>(2 * t + 1)] = p1;
            __syncthreads();
            dit_fft_t<V, LOGN, true, false>(va, sw, c);
            store_inverse<V, LOGN>(out, va, sr, true, c);
        } break;
        default: break;
        }
        __syncthreads();
    }
}

// ---- elementwise ops: 4 coefficients per thread, 128-bit accesses ------------------------------
template <int V>
__device__ __forceinline__ int32_t ew_one(int op, int32_t x, int32_t y, const RedConst &c, int32_t scalar)
{
    switch (op) {
    case SCGPU_OP_PW:        return Exact<V>::pw32(x, y, c);
    case SCGPU_OP_PW16:      return Exact<V>::pw16(x, y, c);
    case SCGPU_OP_NORMALIZE: return Exact<V>::normalize(x, c);
    case SCGPU_OP_CENTER:    return Exact<V>::center(x, c);
    case SCGPU_OP_MODN:      return Exact<V>::modn(x, c);
    case SCGPU_OP_MULN:      return Exact<V>::muln(x, y, c);
    case SCGPU_OP_SQRN:      return Exact<V>::muln(x, x, c);
    case SCGPU_OP_PWR:       return Exact<V>::pwr(x, y, c);
    case SCGPU_OP_SCALAR: {  // ntt.c:425-452 (AVX2 build): 64-bit product, +q if negative, truncate
        int64_t res = (int64_t)x * (int64_t)scalar;
        if (res < 0) res += c.q;
        return (int32_t)res;
    }
    default: return x;
    }
}

template <int V>
__global__ void __launch_bounds__(256)
k_elementwise(ExactArgs g, int n)
{
    const RedConst c = g.rc;
    const size_t total = g.count * (size_t)n;
    const bool has_b = g.b != nullptr;
    const bool b16 = g.op == SCGPU_OP_PW16;
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    // rows whose length is a multiple of 4 keep every 128-bit access inside one row of b
    const size_t total4 = (n % 4 == 0) ? total / 4 : 0;
    const int4 *a4 = static_cast<const int4 *>(g.a);
    int4 *o4 = reinterpret_cast<int4 *>(g.out);
    for (size_t i = tid; i < total4; i += nthreads) {
        int4 x = a4[i];
        int32_t y[4] = {0, 0, 0, 0};
        if (has_b) {
            size_t e = i * 4;
            size_t row = e / (size_t)n, col = e % (size_t)n;
            size_t off = row * g.b_stride + col;
            if (b16) {
                const short4 s = *reinterpret_cast<const short4 *>(static_cast<const int16_t *>(g.b) + off);
                y[0] = s.x; y[1] = s.y; y[2] = s.z; y[3] = s.w;
            } else {
                const int4 s = *reinterpret_cast<const int4 *>(static_cast<const int32_t *>(g.b) + off);
                y[0] = s.x; y[1] = s.y; y[2] = s.z; y[3] = s.w;
            }
        }
        x.x = ew_one<V>(g.op, x.x, y[0], c, g.scalar);
        x.y = ew_one<V>(g.op, x.y, y[1], c, g.scalar);
        x.z = ew_one<V>(g.op, x.z, y[2], c, g.scalar);
        x.w = ew_one<V>(g.op, x.w, y[3], c, g.scalar);
        o4[i] = x;
    }
    // ragged rows (normalize_32 / center_32 take any length; scalar members use length 1)
    for (size_t e = total4 * 4 + tid; e < total; e += nthreads) {
        int32_t y = 0;
        if (has_b) {
            size_t off = (e / (size_t)n) * g.b_stride + e % (size_t)n;
            y = b16 ? (int32_t)static_cast<const int16_t *>(g.b)[off] : static_cast<const int32_t *>(g.b)[off];
        }
        g.out[e] = ew_one<V>(g.op, static_cast<const int32_t *>(g.a)[e], y, c, g.scalar);
    }
}

// ---- row ops that need the whole row: flip, invert, div, sparse ---------------------------------
template <int V>
__global__ void __launch_bounds__(256)
k_rowop(ExactArgs g, int n)
{
    extern __shared__ int32_t sh[];          // n words of staging + 1
    __shared__ int first_zero;
    const RedConst c = g.rc;
    for (size_t row = blockIdx.x; row < g.count; row += gridDim.x) {
        int32_t *out = g.out + row * (size_t)n;
        if (g.op == SCGPU_OP_FLIP) {
            const int32_t *a = static_cast<const int32_t *>(g.a) + row * (size_t)n;
            for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = a[i];
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int32_t x = i ? sh[n - i] : (int32_t)(0u - (uint32_t)sh[0]);
                out[i] = cond_fix(x, c.q);
            }
        } else if (g.op == SCGPU_OP_INVERT || g.op == SCGPU_OP_DIV) {
            // invert_32 / div_32 stop at the first coefficient that reduces to zero and leave
            // the rest of the row untouched (:1732-1747, :1759-1774)
            const int32_t *a = static_cast<const int32_t *>(g.a) + row * (size_t)n;
            const int32_t *den = g.op == SCGPU_OP_DIV ? static_cast<const int32_t *>(g.b) + row * g.b_stride : a;
            if (threadIdx.x == 0) first_zero = n;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int32_t x = Exact<V>::modn(den[i], c);
                sh[i] = x;
                if (x == 0) atomicMin(&first_zero, i);
            }
            __syncthreads();
            const int stop = first_zero;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int32_t num = a[i];
                int32_t res = num;
                if (i < stop) {
                    int32_t inv = Exact<V>::pwr(sh[i], c.q - 2, c);
                    res = g.op == SCGPU_OP_DIV ? Exact<V>::muln(num, inv, c) : inv;
                }
                out[i] = res;
            }
            if (threadIdx.x == 0 && g.rcodes) g.rcodes[row] = stop < n ? 1 : 0;
        } else {  // sparse products, ntt.c:381-422: out[j] = sum_i (j < pos_i ? t[j+n-pos_i] : -t[j-pos_i])
            const int omega = g.scalar & 0xFFFF;
            const int32_t *u = static_cast<const int32_t *>(g.b) + row * g.b_stride;
            const bool t16 = g.op == SCGPU_OP_SPARSE16;
            for (int i = threadIdx.x; i < n; i += blockDim.x)
                sh[i] = t16 ? (int32_t)(static_cast<const int16_t *>(g.a) + row * (size_t)n)[i]
                            : (static_cast<const int32_t *>(g.a) + row * (size_t)n)[i];
            __syncthreads();
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                uint32_t acc = 0;
                for (int i = 0; i < omega; i++) {
                    int pos = u[i];
                    acc += (j < pos) ? (uint32_t)sh[j + n - pos] : 0u - (uint32_t)sh[j - pos];
                }
                out[j] = (int32_t)acc;
            }
        }
        __syncthreads();
    }
}

template <int V>
int launch_v(const NttPlanDev &plan, const ExactArgs &g, cudaStream_t st)
{
    const int n = plan.n;
    const int sms = plan.sm_count > 0 ? plan.sm_count : 148;
    switch (g.op) {
    case SCGPU_OP_FWD: case SCGPU_OP_INV: case SCGPU_OP_FWD_LARGE: case SCGPU_OP_INV_LARGE:
    case SCGPU_OP_FFT: case SCGPU_OP_FFT_LARGE: case SCGPU_OP_POLYMUL: case SCGPU_OP_TRIPLE16: {
        // resident CTAs per SM: 2048 threads / (n/2)
        const int per_sm = 2048 / (n / 2);
        size_t grid = (size_t)sms * per_sm;
        if (grid > g.count) grid = g.count;
        if (plan.logn == 8)       k_transform<V, 8><<<(unsigned)grid, 128, 0, st>>>(g);
        else if (plan.logn == 9)  k_transform<V, 9><<<(unsigned)grid, 256, 0, st>>>(g);
        else if (plan.logn == 10) k_transform<V, 10><<<(unsigned)grid, 512, 0, st>>>(g);
        else { set_error("unsupported n=%d (256, 512, 1024 only)", n); return SCGPU_ERR_UNSUPPORTED; }
    } break;
    case SCGPU_OP_PW: case SCGPU_OP_PW16: case SCGPU_OP_NORMALIZE: case SCGPU_OP_CENTER:
    case SCGPU_OP_MODN: case SCGPU_OP_MULN: case SCGPU_OP_SQRN: case SCGPU_OP_PWR: case SCGPU_OP_SCALAR: {
        size_t total4 = (g.count * (size_t)n + 3) / 4;
        size_t grid = (total4 + 255) / 256;
        size_t cap = (size_t)sms * 8 * 4;
        if (grid > cap) grid = cap;
        if (grid == 0) grid = 1;
        k_elementwise<V><<<(unsigned)grid, 256, 0, st>>>(g, n);
    } break;
    case SCGPU_OP_FLIP: case SCGPU_OP_INVERT: case SCGPU_OP_DIV: case SCGPU_OP_SPARSE32: case SCGPU_OP_SPARSE16: {
        size_t grid = (size_t)sms * 8;
        if (grid > g.count) grid = g.count;
        k_rowop<V><<<(unsigned)grid, 256, (size_t)n * sizeof(int32_t), st>>>(g, n);
    } break;
    default:
        set_error("unknown op %d", g.op);
        return SCGPU_ERR_ARG;
    }
    count_launch();
    SCGPU_CUDA_CHECK(cudaGetLastError());
    return SCGPU_OK;
}

}  // namespace

int launch_exact(const NttPlanDev &plan, const ExactArgs &g, cudaStream_t st)
{
    if (g.count == 0) return SCGPU_OK;
    if ((g.op == SCGPU_OP_FWD || g.op == SCGPU_OP_INV) && g.w == plan.w && g.r == plan.r) {
        // the warp-local schedule (ntt_exact_w32.cu); not applicable (unusual tables, unaligned rows): this file's kernel
        const int e = launch_exact_w32(plan, g.op, g.out, static_cast<const int32_t *>(g.a), g.count, st);
        if (e != SCGPU_ERR_UNSUPPORTED) return e;
    }
    switch (plan.variant) {
    case V_REFERENCE:  return launch_v<V_REFERENCE>(plan, g, st);
    case V_BARRETT:    return launch_v<V_BARRETT>(plan, g, st);
    case V_FP:         return launch_v<V_FP>(plan, g, st);
    case V_AVX:        return launch_v<V_AVX>(plan, g, st);
    case V_SOL7681:    return launch_v<V_SOL7681>(plan, g, st);
    case V_SOL8380417: return launch_v<V_SOL8380417>(plan, g, st);
    default:
        set_error("unsupported reduction variant %d", plan.variant);
        return SCGPU_ERR_UNSUPPORTED;
    }
}

}  // namespace scgpu
