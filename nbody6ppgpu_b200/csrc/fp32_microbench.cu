// fp32_microbench.cu -- FP32 pipe microbenchmark for the roofline denominator.
// MEASURED_PEAKS.json has no FP32 (non-tensor) entry, and the regular-force kernel is bound by the
// FP32 FMA pipe, so the library measures the pipe itself: independent register-resident chains of
// scalar FFMA, packed FFMA2 / FADD2 / FMUL2 (f32x2, new on sm_100), MUFU.RSQ, and an FFMA2 + ALU mix.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../include/gpunb_b200.h"

namespace {
constexpr int NCHAIN = 16;

template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(int iters, float seed, float *out)
{
    float2 a[NCHAIN];
    const float2 b = make_float2(seed, seed * 0.5f), c = make_float2(1e-3f, 2e-3f);
#pragma unroll
    for (int k = 0; k < NCHAIN; k++) a[k] = make_float2(seed + k + threadIdx.x, seed - k);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < NCHAIN; k++) {
            if (MODE == 0) { a[k].x = fmaf(a[k].x, b.x, c.x); a[k].y = fmaf(a[k].y, b.y, c.y); }
            if (MODE == 1) a[k] = __ffma2_rn(a[k], b, c);
            if (MODE == 2) a[k] = __fadd2_rn(a[k], c);
            if (MODE == 3) a[k] = __fmul2_rn(a[k], b);
            if (MODE == 4) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k].x)); }
            if (MODE == 5) {       // 6 packed FMA-pipe ops : 2 ALU ops (FMNMX + FSEL-like), close to the regf mix
                a[k] = __ffma2_rn(a[k], b, c);
                if ((k & 3) == 3) { a[k].x = fminf(a[k].x, a[k - 1].y); a[k].y = (a[k].y < c.y) ? a[k - 2].x : a[k].y; }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCHAIN; k++) s += a[k].x + a[k].y;
    if (s == 12345.678f) out[0] = s;
}

template <int MODE> double run(int iters)
{
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    float *out; cudaMalloc(&out, 4);
    const int blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    pipe_kernel<MODE><<<blocks, 256>>>(iters / 8 + 1, 1.0001f, out);      // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    pipe_kernel<MODE><<<blocks, 256>>>(iters, 1.0001f, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "gpunb_b200: microbench launch failed\n"); abort(); }
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double lanes_ops = (double)blocks * 256 * (double)iters * NCHAIN;   // per-lane "chain steps"
    double per_step = 2.0;                     // scalar elements per chain step
    if (MODE == 4) per_step = 1.0;
    const double flop_per_elem = (MODE == 0 || MODE == 1 || MODE == 5) ? 2.0 : 1.0;
    return lanes_ops * per_step * flop_per_elem / (ms * 1e-3) * 1e-12;       // TFLOP/s (mode 4: Tera-ops/s)
}
}  // namespace

extern "C" double gpunb_b200_fp32_microbench(int mode, int iters)
{
    if (iters <= 0) iters = 4096;
    switch (mode) {
        case 0: return run<0>(iters);
        case 1: return run<1>(iters);
        case 2: return run<2>(iters);
        case 3: return run<3>(iters);
        case 4: return run<4>(iters);
        case 5: return run<5>(iters);
    }
    return -1.0;
}
