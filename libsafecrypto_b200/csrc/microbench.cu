// microbench.cu -- measured INT32 / FP issue peaks on the device the library runs on.
// the INT roofline denominator "must be measured by a microbenchmark on the box".
// Each thread runs `iters` rounds of 8 independent dependent-chains of one instruction class;
// the result is written so nothing is optimised away.
#include "scgpu_internal.h"
#include "../../include/scgpu.h"

namespace scgpu {
namespace {

template <int KIND>
__global__ void __launch_bounds__(256) k_peak(int iters, int32_t seed, int32_t *sink)
{
    int32_t x[8];
    int32_t a = seed | 1, b = seed ^ 0x5bd1e995;
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i * 977 + seed;
    float f[8]; double d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { f[i] = (float)x[i]; d[i] = (double)x[i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (KIND == 0) x[i] = x[i] * a + b;                              // IMAD
                else if (KIND == 1) x[i] = __mulhi(x[i], a) + b;                 // IMAD.HI
                else if (KIND == 2) x[i] = x[i] + a + (b ^ x[(i + 1) & 7]);      // IADD3 (+LOP3)
                else if (KIND == 3) x[i] = (x[i] & a) ^ (b | x[(i + 1) & 7]);    // LOP3
                else if (KIND == 4) x[i] = __funnelshift_l(x[i], a, 7) ;         // SHF
                else if (KIND == 5) { long long w = (long long)x[i] * a; x[i] = (int32_t)w ^ (int32_t)(w >> 32); }  // IMAD.WIDE
                else if (KIND == 6) { x[i] = x[i] * a + b; x[i] = x[i] + a + x[(i + 1) & 7]; }  // IMAD + IADD3 pair
                else if (KIND == 7) f[i] = fmaf(f[i], 1.0000001f, 0.5f);         // FFMA
                else if (KIND == 8) d[i] = fma(d[i], 1.0000001, 0.5);            // DFMA
                else if (KIND == 10) {                                           // 32-bit Barrett butterfly (ntt_fast_sq.cu)
                    if (i < 4) {
                        int32_t p = x[i + 4] * a;
                        int32_t qe = __mulhi(p, 349496) ;
                        int32_t t = qe * -12289 + p;
                        x[i + 4] = x[i] - t; x[i] = x[i] + t;
                    }
                }
                else if (KIND == 11) {                                           // float-quotient butterfly (ntt_fast_fq.cu)
                    if (i < 4) {
                        float fq = fmaf(__int_as_float(x[i + 4]), __int_as_float(b), __int_as_float(a));
                        int32_t p = x[i + 4] * a + b;
                        int32_t t = __float_as_int(fq) * -12289 + p;
                        x[i + 4] = x[i] - t; x[i] = x[i] + t;
                    }
                }
                else if (KIND == 9) {                                            // Montgomery butterfly: 3 mul-class + 2 add-class
                    int32_t hi = __mulhi(x[i], a), lo = __mulhi(x[i] * b, 12289);
                    int32_t o = x[(i + 4) & 7];
                    x[i] = o - hi + lo; x[(i + 4) & 7] = o + hi - lo;
                }
            }
        }
    }
    int32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x[i] ^ __float_as_int(f[i]) ^ (int32_t)__double2ll_rz(d[i]);
    if (acc == 0x7fffffff) sink[0] = acc;
}

}  // namespace
}  // namespace scgpu

extern "C" double scgpu_int_peak_gops(int kind, int iters, int device)
{
    using namespace scgpu;
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    int32_t *sink = nullptr;
    if (cudaMalloc(&sink, 4) != cudaSuccess) return -1.0;
    const int grid = prop.multiProcessorCount * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto launch = [&](int n) {
        switch (kind) {
        case 0: k_peak<0><<<grid, block>>>(n, 12345, sink); break;
        case 1: k_peak<1><<<grid, block>>>(n, 12345, sink); break;
        case 2: k_peak<2><<<grid, block>>>(n, 12345, sink); break;
        case 3: k_peak<3><<<grid, block>>>(n, 12345, sink); break;
        case 4: k_peak<4><<<grid, block>>>(n, 12345, sink); break;
        case 5: k_peak<5><<<grid, block>>>(n, 12345, sink); break;
        case 6: k_peak<6><<<grid, block>>>(n, 12345, sink); break;
        case 7: k_peak<7><<<grid, block>>>(n, 12345, sink); break;
        case 8: k_peak<8><<<grid, block>>>(n, 12345, sink); break;
        case 10: k_peak<10><<<grid, block>>>(n, 12345, sink); break;
        case 11: k_peak<11><<<grid, block>>>(n, 12345, sink); break;
        default: k_peak<9><<<grid, block>>>(n, 12345, sink); break;
        }
        count_launch();
    };
    launch(iters / 8 + 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    // ops per inner statement: kinds 6 -> 2, 9 -> 5 (counted per butterfly: 4 per i-pair ... see bench.py)
    // kinds 10 / 11 report butterflies per second (4 butterflies per 8-slot round)
    double per = (kind == 6) ? 2.0 : (kind == 9 ? 5.0 : ((kind == 10 || kind == 11) ? 0.5 : 1.0));
    double ops = (double)grid * block * (double)iters * 4.0 * 8.0 * per;
    return ops / (ms * 1e-3) / 1e9;
}
