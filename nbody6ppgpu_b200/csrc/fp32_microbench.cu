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
    float2 a[NCHAIN], bb[NCHAIN], cc[NCHAIN];
    const float2 b = make_float2(seed, seed * 0.5f), c = make_float2(1e-3f, 2e-3f);
#pragma unroll
    for (int k = 0; k < NCHAIN; k++) {
        a[k] = make_float2(seed + k + threadIdx.x, seed - k);
        bb[k] = make_float2(seed * 0.25f + 1e-3f * k, seed * 0.125f - 1e-3f * k);     // all-distinct operand registers
        cc[k] = make_float2(1e-4f * (k + 1 + threadIdx.x), 2e-4f * (k + 1));
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < NCHAIN; k++) {
            if (MODE == 0) { a[k].x = fmaf(a[k].x, b.x, c.x); a[k].y = fmaf(a[k].y, b.y, c.y); }
            if (MODE == 1) a[k] = __ffma2_rn(a[k], b, c);
            if (MODE == 2) a[k] = __fadd2_rn(a[k], c);
            if (MODE == 3) a[k] = __fmul2_rn(a[k], b);
            if (MODE == 4) { asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k].x)); }
            if (MODE == 6) { a[k].x = fmaf(bb[k].x, cc[k].x, a[k].x); a[k].y = fmaf(bb[k].y, cc[k].y, a[k].y); }   // scalar, 3 distinct regs
            if (MODE == 7) a[k] = __ffma2_rn(bb[k], cc[k], a[k]);                      // packed, 3 distinct register pairs
            if (MODE == 8) a[k] = __ffma2_rn(bb[k], c, a[k]);                          // packed, one operand shared by all
            if (MODE == 9) a[k] = __fadd2_rn(a[k], bb[k]);                             // packed add, 2 distinct pairs
            if (MODE == 10) a[k] = __ffma2_rn(bb[k], bb[k], a[k]);                     // packed, square: 2 distinct pairs
            if (MODE == 11) { a[k] = __ffma2_rn(bb[k], cc[k], a[k]); if ((k & 3) == 3) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(cc[k].x)); }
            if (MODE == 12) { a[k].x = a[k].x + bb[k].x; a[k].y = a[k].y + bb[k].y; }                  // scalar FADD, 2 distinct regs
            if (MODE == 13) { a[k].x = a[k].x * bb[k].x; a[k].y = a[k].y * bb[k].y; }                  // scalar FMUL, 2 distinct regs
            if (MODE == 14) {                                                                           // scalar mix ~ regf far body (7 FADD : 7 FMUL : 13 FFMA)
                if ((k & 3) == 0) { a[k].x = a[k].x + bb[k].x; a[k].y = a[k].y + bb[k].y; }
                else if ((k & 3) == 1) { a[k].x = a[k].x * bb[k].x; a[k].y = a[k].y * bb[k].y; }
                else { a[k].x = fmaf(bb[k].x, cc[k].x, a[k].x); a[k].y = fmaf(bb[k].y, cc[k].y, a[k].y); }
            }
            if (MODE == 15) { a[k].x = fmaf(bb[k].x, cc[k].x, a[k].x); a[k].y = a[k].y + bb[k].y; }   // FFMA : FADD 1:1
            if (MODE == 5) {       // 6 packed FMA-pipe ops : 2 ALU ops (FMNMX + FSEL-like), close to the regf mix
                a[k] = __ffma2_rn(a[k], b, c);
                if ((k & 3) == 3) { a[k].x = fminf(a[k].x, a[k - 1].y); a[k].y = (a[k].y < c.y) ? a[k - 2].x : a[k].y; }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCHAIN; k++) s += a[k].x + a[k].y + bb[k].x + cc[k].x;
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
    const double flop_per_elem = (MODE == 0 || MODE == 1 || MODE == 5 || MODE == 6 || MODE == 7 || MODE == 8 || MODE == 10 || MODE == 11) ? 2.0 : 1.0;
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
        case 6: return run<6>(iters);
        case 7: return run<7>(iters);
        case 8: return run<8>(iters);
        case 9: return run<9>(iters);
        case 10: return run<10>(iters);
        case 11: return run<11>(iters);
        case 12: return run<12>(iters);
        case 13: return run<13>(iters);
        case 14: return run<14>(iters);
        case 15: return run<15>(iters);
    }
    return -1.0;
}

// ---------------------------------------------------------------------------------------------
// Far-body microbenchmark: the 27-op force-only pair body of regf_kernel in isolation (one smem tile re-read
// in a loop, no TMA, no classification), to separate what the instruction mix can deliver from the per-tile
// overheads.  mode bit0: replace MUFU.RSQ by an FMUL; bit1: operands from registers instead of LDS.128;
// bit2: two i-particles per lane.  Returns Gint/s.
// ---------------------------------------------------------------------------------------------
namespace {
struct FAcc { float ax, ay, az, p, jx, jy, jz; };
template <bool NOMUFU>
__device__ __forceinline__ void far1(FAcc &A, float cx, float cy, float cz, float nvx, float nvy, float nvz,
                                     float DX, float DY, float DZ, float VX, float VY, float VZ, float M)
{
    const float dx = DX + cx, dy = DY + cy, dz = DZ + cz;
    const float dvx = VX + nvx, dvy = VY + nvy, dvz = VZ + nvz;
    const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float rv = fmaf(dz, dvz, fmaf(dy, dvy, dx * dvx));
    float rinv;
    if (NOMUFU) rinv = r2 * 0.999f; else asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(r2));
    const float rinv2 = rinv * rinv, mrinv = M * rinv, mrinv3 = mrinv * rinv2, rv3 = rv * (rinv2 * -3.f);
    A.p += mrinv;
    A.ax = fmaf(mrinv3, dx, A.ax); A.ay = fmaf(mrinv3, dy, A.ay); A.az = fmaf(mrinv3, dz, A.az);
    A.jx = fmaf(mrinv3, fmaf(rv3, dx, dvx), A.jx);
    A.jy = fmaf(mrinv3, fmaf(rv3, dy, dvy), A.jy);
    A.jz = fmaf(mrinv3, fmaf(rv3, dz, dvz), A.jz);
}

template <int IT, bool NOMUFU, bool NOLDS>
__global__ void __launch_bounds__(128) farbody_kernel(int ntile_iters, float seed, float *out)
{
    __shared__ __align__(16) float tile[4][7 * 64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *tb = tile[warp];
    for (int k = lane; k < 7 * 64; k += 32) tb[k] = seed * (1.f + 0.001f * k) + 0.01f * warp;
    __syncwarp();
    FAcc A[IT][2];
    float cx[IT], cy[IT], cz[IT], nvx[IT], nvy[IT], nvz[IT];
#pragma unroll
    for (int k = 0; k < IT; k++) {
        A[k][0] = FAcc{0, 0, 0, 0, 0, 0, 0}; A[k][1] = A[k][0];
        cx[k] = seed + lane + k; cy[k] = seed - lane; cz[k] = 0.5f * seed + k;
        nvx[k] = 0.1f * lane; nvy[k] = -0.2f * lane; nvz[k] = 0.3f + k;
    }
    const float4 *c = reinterpret_cast<const float4 *>(tb);
    float4 R[7];
    if (NOLDS) {
#pragma unroll
        for (int q = 0; q < 7; q++) R[q] = c[q * 16 + (lane & 15)];
    }
    for (int t = 0; t < ntile_iters; t++) {
#pragma unroll 2
        for (int q = 0; q < 16; q++) {
            float4 DX, DY, DZ, VX, VY, VZ, M;
            if (NOLDS) { DX = R[0]; DY = R[1]; DZ = R[2]; VX = R[3]; VY = R[4]; VZ = R[5]; M = R[6]; }
            else { DX = c[q]; DY = c[16 + q]; DZ = c[32 + q]; VX = c[48 + q]; VY = c[64 + q]; VZ = c[80 + q]; M = c[96 + q]; }
#pragma unroll
            for (int k = 0; k < IT; k++) {
                far1<NOMUFU>(A[k][0], cx[k], cy[k], cz[k], nvx[k], nvy[k], nvz[k], DX.x, DY.x, DZ.x, VX.x, VY.x, VZ.x, M.x);
                far1<NOMUFU>(A[k][1], cx[k], cy[k], cz[k], nvx[k], nvy[k], nvz[k], DX.y, DY.y, DZ.y, VX.y, VY.y, VZ.y, M.y);
                far1<NOMUFU>(A[k][0], cx[k], cy[k], cz[k], nvx[k], nvy[k], nvz[k], DX.z, DY.z, DZ.z, VX.z, VY.z, VZ.z, M.z);
                far1<NOMUFU>(A[k][1], cx[k], cy[k], cz[k], nvx[k], nvy[k], nvz[k], DX.w, DY.w, DZ.w, VX.w, VY.w, VZ.w, M.w);
            }
            if (NOLDS) { R[0].x += 1e-7f; }
        }
#pragma unroll
        for (int k = 0; k < IT; k++) cx[k] += 1e-3f;
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < IT; k++)
        for (int h = 0; h < 2; h++) s += A[k][h].ax + A[k][h].ay + A[k][h].az + A[k][h].p + A[k][h].jx + A[k][h].jy + A[k][h].jz;
    if (s == 12345.678f) out[0] = s;
}

template <int IT, bool NOMUFU, bool NOLDS> double run_far(int iters, int ctas_per_sm)
{
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    float *out; cudaMalloc(&out, 4);
    const int blocks = prop.multiProcessorCount * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    farbody_kernel<IT, NOMUFU, NOLDS><<<blocks, 128>>>(iters / 8 + 1, 1.0001f, out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    farbody_kernel<IT, NOMUFU, NOLDS><<<blocks, 128>>>(iters, 1.0001f, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (cudaGetLastError() != cudaSuccess) { fprintf(stderr, "gpunb_b200: farbody microbench launch failed\n"); abort(); }
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double pairs = (double)blocks * 128 * (double)iters * 64 * IT;
    return pairs / (ms * 1e-3) * 1e-9;
}
}  // namespace

extern "C" double gpunb_b200_farbody_microbench(int mode, int iters, int ctas_per_sm)
{
    if (iters <= 0) iters = 2000;
    if (ctas_per_sm <= 0) ctas_per_sm = 4;
    switch (mode) {
        case 0: return run_far<1, false, false>(iters, ctas_per_sm);
        case 1: return run_far<1, true, false>(iters, ctas_per_sm);
        case 2: return run_far<1, false, true>(iters, ctas_per_sm);
        case 3: return run_far<1, true, true>(iters, ctas_per_sm);
        case 4: return run_far<2, false, false>(iters, ctas_per_sm);
        case 5: return run_far<2, true, false>(iters, ctas_per_sm);
        case 6: return run_far<2, false, true>(iters, ctas_per_sm);
    }
    return -1.0;
}
