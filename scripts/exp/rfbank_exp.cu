// rfbank_exp.cu -- what does a packed FFMA2 with three DISTINCT register-pair operands cost, as a function of WHICH registers?
// 16 independent chains  a[k] = fma2(bb[(k+S1)%16], cc[(k+S2)%16], a[k]); the shifts S1, S2 change the register numbers that meet
// in one instruction (ptxas allocates the three arrays contiguously); scripts/exp/rfbank_analyze.py reads the triples from the
// SASS and fits the cost per (A, B, C) residue class to the measured cycles per instruction printed here.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o rfbank_exp rfbank_exp.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int NCH = 16;
template <int S1, int S2, int MODE>
__global__ void __launch_bounds__(128) rf_kernel(int iters, float seed, float *out)
{
    float2 a[NCH], bb[NCH], cc[NCH];
    __shared__ __align__(16) float sm[128 * 4];
    if (MODE >= 7) { for (int k = threadIdx.x; k < 512; k += 128) sm[k] = seed * k; __syncthreads(); }
    float4 ld = make_float4(0.f, 0.f, 0.f, 0.f); float mu = seed + threadIdx.x; int iv = threadIdx.x;
#pragma unroll
    for (int k = 0; k < NCH; k++) {
        const float t = seed * (float)(threadIdx.x + 1);       // every component depends on run-time values: nothing to rematerialise in the loop
        a[k]  = make_float2(t + k, t - k);
        bb[k] = make_float2(t * 0.25f + 1e-3f * k, t * 0.125f - 1e-3f * k);
        cc[k] = make_float2(t * 1e-4f * (k + 1), t * 2e-4f + 1e-4f * k);
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < NCH; k++) {
            if (MODE == 0) a[k] = __ffma2_rn(bb[(k + S1) % NCH], cc[(k + S2) % NCH], a[k]);                 // acc += b * c : 3 distinct pairs
            if (MODE == 1) a[k] = __ffma2_rn(a[k], bb[(k + S1) % NCH], cc[(k + S2) % NCH]);                 // a = a * b + c : 3 distinct pairs
            if (MODE == 2) { a[k].x = fmaf(bb[(k + S1) % NCH].x, cc[(k + S2) % NCH].x, a[k].x);            // scalar twins
                             a[k].y = fmaf(bb[(k + S1) % NCH].y, cc[(k + S2) % NCH].y, a[k].y); }
            if (MODE == 3) a[k] = __ffma2_rn(bb[(k + S1) % NCH], bb[(k + S1) % NCH], a[k]);                 // 2 distinct
            if (MODE == 4) a[k] = __fadd2_rn(bb[(k + S1) % NCH], a[k]);                                     // 2 distinct, add
            if (MODE >= 7) {       // port sharing: 16 full-rate FFMA2 (one operand shared by all) + S1 other instructions per chain step group of 4
                a[k] = __ffma2_rn(a[k], bb[0], cc[0]);
                if ((k & 3) == 3) {
#pragma unroll
                    for (int e = 0; e < S1; e++) {
                        if (MODE == 7) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(mu));
                        if (MODE == 8) { float4 v; const unsigned ad = (unsigned)__cvta_generic_to_shared(&sm[((iv + it + k + e) & 127) * 4]);
                                         asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ad)); ld.x += v.x; }
                        if (MODE == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(iv) : "r"(it), "r"(k + e));
                    }
                }
            }
            if (MODE == 5) a[k] = __ffma2_rn(bb[((k / 4) * 4 + S1) % NCH], cc[(k + S2) % NCH], a[k]);      // groups of 4 share one multiplier (reuse cache)
            if (MODE == 6) a[k] = __ffma2_rn(bb[((k / 2) * 2 + S1) % NCH], cc[(k + S2) % NCH], a[k]);      // groups of 2 share one multiplier
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; k++) s += a[k].x + a[k].y + bb[k].x + cc[k].x + bb[k].y + cc[k].y;
    if (MODE >= 7) s += ld.x + ld.y + mu + (float)iv;
    if (s == 12345.678f) out[0] = s;
}
static int nsm = 148; static double clk_mhz = 1965.0;
template <int S1, int S2, int MODE> void run()
{
    float *out; cudaMalloc(&out, 4);
    const int iters = 4096, blocks = nsm * 8;             // 8 CTAs x 4 warps = 32 warps per SM: 8 per sub-partition
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        rf_kernel<S1, S2, MODE><<<blocks, 128>>>(iters, 1.0001f, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    // warp-instructions per sub-partition: blocks*4 warps / (nsm*4) * iters * NCH  (MODE 2: twice as many instructions)
    const double winstr = (double)blocks * 4 / (nsm * 4.0) * iters * NCH;
    const double cyc = best * 1e-3 * clk_mhz * 1e6 / winstr;
    printf("rf mode %d s1 %2d s2 %2d  %.3f ms  %.3f cycles per chain step and sub-partition\n", MODE, S1, S2, best, cyc);
    cudaFree(out);
}
template <int MODE> void sweep()
{
#define ROW(S1) run<S1, 0, MODE>(); run<S1, 1, MODE>(); run<S1, 2, MODE>(); run<S1, 3, MODE>(); run<S1, 4, MODE>(); run<S1, 5, MODE>(); run<S1, 6, MODE>(); run<S1, 7, MODE>();
    ROW(0) ROW(1) ROW(2) ROW(3) ROW(4) ROW(5) ROW(6) ROW(7)
#undef ROW
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); nsm = p.multiProcessorCount; clk_mhz = p.clockRate * 1e-3;
    printf("%s, %d SMs, %.0f MHz\n", p.name, nsm, clk_mhz);
    sweep<0>(); sweep<1>();
    run<0, 0, 2>(); run<1, 0, 2>(); run<0, 1, 2>(); run<1, 2, 2>(); run<3, 5, 2>();
    run<0, 0, 3>(); run<1, 0, 3>(); run<0, 0, 4>(); run<1, 0, 4>();
    run<0, 0, 5>(); run<1, 0, 5>(); run<0, 1, 5>(); run<1, 1, 5>(); run<2, 3, 5>(); run<3, 2, 5>();
    run<0, 0, 7>(); run<1, 0, 7>(); run<2, 0, 7>(); run<0, 0, 8>(); run<1, 0, 8>(); run<2, 0, 8>(); run<0, 0, 9>(); run<1, 0, 9>(); run<2, 0, 9>(); run<4, 0, 9>();
    run<0, 0, 6>(); run<1, 0, 6>(); run<0, 1, 6>(); run<1, 1, 6>(); run<2, 3, 6>(); run<3, 2, 6>();
    return 0;
}
