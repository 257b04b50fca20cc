// gpunb_b200.cu -- B200-native (sm_100a) regular-force library for NBODY6++GPU.
//
// Drop-in for the reference's gpunb.velocity.cu / gpupot.gpu.cu behind the same
// Fortran-callable C-ABI (include/gpunb_b200.h).  Written from scratch; the
// reference (file:line cited below) defines WHAT is computed, not how.
//
// Data layout in HBM (per device)
//   jraw   : the caller's fp64 snapshot, m[nj] | x[nj][3] | v[nj][3]                          56 B/j
//   jtile  : the device's j-shard, SORTED ALONG A HILBERT CURVE and cut into tiles of TJ=64.
//            One tile = 64 B header + 13 float arrays of 64                                  ~53 B/j
//              header : tile origin O = Oh + Ol (two floats per axis, centre of the tile's bounding box), box
//                       half-extents, velocity box (centre, half-extents), sqrt of the largest mass
//              arrays : dx,dy,dz = (float)(x - O)   tile-local offsets: |offset| <= tile extent, so the
//                                                   pair separation (O - x_i) + offset keeps a relative
//                                                   precision of 2^-24 no matter how large |x| is;
//                       vx,vy,vz, m                 fp32;
//                       xh,yh,zh = (float)x         exactly what the reference's FP32 cast sees
//                                                   (gpunb.velocity.cu:62-64): only read by tiles that
//                                                   can hold neighbours, to evaluate the reference's
//                                                   neighbour predicate bit for bit;
//                       xl,yl,zl = (float)(x - xh)  low words: NEAR pairs use the float-float separation.
//            A tile is contiguous: ONE 1-D TMA bulk copy brings header and data.
//   jidx   : sorted slot -> global j index (ghost slots of the last tile: -1)
//   per pipeline slot (struct Slot: the buffers one i-block or sub-block needs, and a lo/hi stream pair)
//     part   : per (j-slice s, local i-slot) partial sums, 7 doubles, + count                  [S][nloc]
//     seg    : per (local i-slot, s) neighbour-index segment, capacity segcap ints
//     res_f  : per i  acc[3] jrk[3] pot  (fp64)   ; res_list : [ni][lmax] int32 (the ABI layout); gpunb_regf_ results
//              go to MAPPED pinned host buffers of the same layout instead (written by the kernels over PCIe)
//   state  : device-resident predictor (body | x0 | x0dot | f | fdot | t0)
//
// Kernels
//   absmax/mortonkey/tilepack   fp64 snapshot -> sorted tiles (radix sort of the keys: CUB, plumbing)
//   isort_kernel   Morton order of the i-block, so that the 32 i-particles of a warp are close in space
//   regf_kernel    the O(ni*nj) pair kernel.  One WARP = one work item (32*IT i-particles, every S-th
//                  j-tile).  Tiles stream through warp-private shared memory by TMA bulk copies
//                  (cp.async.bulk + mbarrier, 3 stages); lanes own i-particles, j is broadcast from smem;
//                  packed f32x2 over j (see the note above accumulate()).
//                  Per (warp, tile) the bounding boxes decide: FAR tiles (no pair can satisfy the
//                  neighbour criterion, with margin) run a 27-op force-only body; NEAR tiles run the
//                  full body with the reference predicate.  FP32 chains are 32 terms, then fp64.
//   merge_kernel   per i: fp64 sum of the S partials in fixed order, gather of the S segments, sort
//                  ascending (lists must be strictly ascending: regcor_gpu.F:299-336), overflow
//                  encoding -(count) (reg.avx.cpp:320-321).
//                  Its last CTA publishes the shard result to the peers (flags over NVLink) when sharded.
//   combine_kernel multi-GPU exchange step over NVLink peer pointers (flags, peer pulls, acks).
//   pot_kernel     gpupot: tile-local dx (two-float separations for close tiles), packed f32x2, rsqrt + one
//                  Newton step, fp64 flush; pot_sum_kernel adds the shards' partials in rank order.
//   predict_kernel device-resident predictor (xbpredall.f restated, unfused fp64).
// Host side: resident sweeps cycle the i-blocks through pipeline slots; gpunb_regf_ splits its block into sub-blocks
// on the same slots.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <climits>
#include <vector>
#include <algorithm>
#include <sys/time.h>
#include <unistd.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <omp.h>
#include "../../include/gpunb_b200.h"
#include "internal.h"

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "gpunb_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, \
            cudaGetErrorString(e_)); abort(); } } while (0)
#define FATAL(...) do { fprintf(stderr, "gpunb_b200: " __VA_ARGS__); fprintf(stderr, "\n"); abort(); } while (0)

namespace {

constexpr int TJ          = 64;                   // j-particles per tile
constexpr int HDR         = 20;                   // header floats (16-17: bounding-sphere radius, |velocity half-extents|; 18-19 spare)
constexpr int NCOMP       = 13;                   // float arrays per tile
constexpr int TILE_FLOATS = HDR + TJ * NCOMP;     // 852
constexpr int TILE_BYTES  = TILE_FLOATS * 4;      // 3408 (multiple of 16: one TMA bulk copy)
enum { C_DX = 0, C_DY, C_DZ, C_VX, C_VY, C_VZ, C_M, C_XH, C_YH, C_ZH, C_XL, C_YL, C_ZL };
constexpr int NSTAGE      = 3;                    // smem stages per warp
constexpr int WARPS       = 4;                    // warps per CTA (warp-autonomous: no CTA-wide sync)
constexpr int NIMAX       = 2048;                 // capacity per call (reference: gpunb.velocity.cu:24)
constexpr int JOBCAP      = 16384;                // i-particles of one launch group: the union of all ranks' i-slices of a
                                                  // collective gpunb_regf_ (i-slice mode), or one block of a resident sweep
constexpr int PART_STRIDE = 8;                    // doubles per partial record (7 used)
constexpr int OVERSUB     = 1;                    // work items per resident warp slot of a resident SWEEP (GPUNB_B200_OVERSUB): 1 --
                                                  // consecutive launches fill each other's tails, finer items only add fixed costs
constexpr int REGF_OVERSUB = 4;                   // ... of a gpunb_regf_ call (GPUNB_B200_REGF_OVERSUB).  A launch that runs ALONE is one wave
                                                  // of 16 warps per SM which the warp scheduler does not serve evenly: work items end between
                                                  // 0.40 and 1.08 ms (profiles/r2s_tail_probe_per_sm.txt) and the last quarter of the kernel has
                                                  // one warp per sub-partition.  Four times as many, four times shorter work items keep the
                                                  // sub-partitions full until the last wave: 946 -> 1001 Gint/s per launch at ni = 1024,
                                                  // 906 -> 953 at ni = 256 (profiles/r2t_variant_probe.txt)
constexpr int MIN_TILES_PER_ITEM = 24;            // ... as long as a work item keeps this many j-tiles
constexpr int SORT_CAP    = 1024;
#ifndef FAR_UNROLL
#define FAR_UNROLL 2
#endif
constexpr int FAR_UNROLL_Q = FAR_UNROLL;          // quads of j per iteration of the FAR loop                 // merge_kernel sorts up to this many neighbours per i

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 dup2(float a) { return make_float2(a, a); }
// An FFMA2 whose three operands are three DIFFERENT register pairs runs at ~45 % of the FP32 rate on B200
// (register-file bandwidth; measured by gpunb_b200_fp32_microbench mode 7: 31 vs 70 TFLOP/s), while two scalar
// FFMAs run at full rate.  FFMA2 is kept where an operand repeats (squares, broadcast scalars, immediates).
__device__ __forceinline__ float2 fma2s(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Tile construction: |x|max -> Morton keys -> (CUB radix sort) -> tilepack
// ---------------------------------------------------------------------------------------------
// largest |coordinate| (as float bits; non-negative floats order like unsigned ints) + NaN check of the
// whole record (reference asserts on NaN input: gpunb.velocity.cu:72-78)
__global__ void absmax_kernel(int n, const double *__restrict__ m, const double *__restrict__ x,
                              const double *__restrict__ v, unsigned *__restrict__ out, int *__restrict__ nanflag)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    float a = 0.f;
    bool bad = false;
    if (j < n) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double xd = x[3 * (size_t)j + k];
            a = fmaxf(a, fabsf((float)xd));
            bad |= (xd != xd);
            if (v) { const double vd = v[3 * (size_t)j + k]; bad |= (vd != vd); }
        }
        const double md = m[j];
        bad |= (md != md);
    }
    a = warp_max(a);
    if ((threadIdx.x & 31) == 0 && a > 0.f) atomicMax(out, __float_as_uint(a));
    if (bad) *(volatile int *)nanflag = 1;            // mapped host memory: the host reads it after the stream sync
}

__device__ __forceinline__ unsigned long long spread21(unsigned v)
{   // 21 bits -> every third bit
    unsigned long long x = v & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8)  & 0x100f00f00f00f00full;
    x = (x | x << 4)  & 0x10c30c30c30c30c3ull;
    x = (x | x << 2)  & 0x1249249249249249ull;
    return x;
}
// Hilbert-curve key (Skilling's transpose algorithm, 21 bits per axis).  Unlike the Morton curve the Hilbert
// curve has no jumps: 64 consecutive particles form a compact tile, which is what keeps the tile-local
// offsets small (precision) and the bounding boxes tight (few NEAR tiles).
__device__ __forceinline__ unsigned long long morton_key(double px, double py, double pz, float H)
{
    const float s = 1048575.5f / H;      // 2^20 cells per half box
    unsigned X[3];
    X[0] = (unsigned)fminf(fmaxf((float)px * s + 1048576.f, 0.f), 2097151.f);
    X[1] = (unsigned)fminf(fmaxf((float)py * s + 1048576.f, 0.f), 2097151.f);
    X[2] = (unsigned)fminf(fmaxf((float)pz * s + 1048576.f, 0.f), 2097151.f);
    const unsigned Mtop = 1u << 20;
    for (unsigned Q = Mtop; Q > 1; Q >>= 1) {
        const unsigned P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) X[0] ^= P;
            else { const unsigned t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0]; X[2] ^= X[1];
    unsigned t = 0;
    for (unsigned Q = Mtop; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    return (spread21(X[0]) << 2) | (spread21(X[1]) << 1) | spread21(X[2]);
}
__global__ void mortonkey_kernel(int n, const double *__restrict__ x, const unsigned *__restrict__ hbits,
                                 unsigned long long *__restrict__ keys, int *__restrict__ vals)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float H = fmaxf(__uint_as_float(*hbits), 1e-30f);
    keys[j] = morton_key(x[3 * (size_t)j], x[3 * (size_t)j + 1], x[3 * (size_t)j + 2], H);
    vals[j] = j;
}

// One warp per tile: gathers its 64 particles through the sort permutation, finds the bounding boxes,
// writes header + arrays.  Ghost slots of the last tile replicate the tile's first particle with mass 0
// and index -1 (never listed, no force).  v may be NULL (gpupot tiles).
__global__ void __launch_bounds__(128) tilepack_kernel(int n, int t0, int tstride, int nloc, const double *__restrict__ m,
                                                        const double *__restrict__ x, const double *__restrict__ v,
                                                        const int *__restrict__ perm, float *__restrict__ tiles,
                                                        int *__restrict__ jidx, int *__restrict__ nanflag,
                                                        const unsigned *__restrict__ hbits, unsigned long long *__restrict__ qsum,
                                                        unsigned long long *__restrict__ q_host)
{   // packs tiles t = t0 + l * tstride (l < nloc) of the sorted order into tiles[l], jidx[l * TJ ..]
    // qsum (optional): [0] sum over the tiles of (hx + hy + hz) / |x|max in fixed point -- an ORDER-INDEPENDENT integer
    // sum, so that the decision it feeds (keep the Hilbert order for the next snapshot or sort again) is deterministic --,
    // [1] tiles done; the last tile to finish hands the sum to the host (mapped memory) and clears both.
    const int lane = threadIdx.x & 31;
    const int l = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (l >= nloc) return;
    const int t = t0 + l * tstride;
    double px[2][3], mn[3], mx[3];
    float pv[2][3], pm[2], vmn[3], vmx[3], mmax = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) { mn[c] = 1e300; mx[c] = -1e300; vmn[c] = 3e38f; vmx[c] = -3e38f; }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int p = t * TJ + h * 32 + lane;
        const bool real = p < n;
        const int src = perm[real ? p : t * TJ];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            px[h][c] = x[3 * (size_t)src + c];
            pv[h][c] = v ? (float)v[3 * (size_t)src + c] : 0.f;
            mn[c] = fmin(mn[c], px[h][c]); mx[c] = fmax(mx[c], px[h][c]);
            vmn[c] = fminf(vmn[c], pv[h][c]); vmx[c] = fmaxf(vmx[c], pv[h][c]);
        }
        pm[h] = real ? (float)m[src] : 0.f;
        if (pm[h] != pm[h] || px[h][0] != px[h][0] || px[h][1] != px[h][1] || px[h][2] != px[h][2] ||
            pv[h][0] != pv[h][0] || pv[h][1] != pv[h][1] || pv[h][2] != pv[h][2]) *(volatile int *)nanflag = 1;
        mmax = fmaxf(mmax, pm[h]);
        jidx[l * TJ + h * 32 + lane] = real ? src : -1;
    }
    float *tb = tiles + (size_t)l * TILE_FLOATS;
    double O[3];
    float Oh[3], Ol[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        mn[c] = warp_min(mn[c]); mx[c] = warp_max(mx[c]);
        vmn[c] = warp_min(vmn[c]); vmx[c] = warp_max(vmx[c]);
        // tile origin = centre of the bounding box, REPRESENTED AS TWO FLOATS (Oh + Ol): the pair kernel forms
        // (O - x_i) in fp32 as (Oh - xh_i) + (Ol - xl_i); the tile-local offsets are taken from exactly that O
        const double oc = 0.5 * (mn[c] + mx[c]);
        Oh[c] = (float)oc; Ol[c] = (float)(oc - (double)Oh[c]);
        O[c] = (double)Oh[c] + (double)Ol[c];
    }
    mmax = warp_max(mmax);
    if (lane == 0) {
        float hsum = 0.f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            tb[c] = Oh[c]; tb[3 + c] = Ol[c];
            // half-extents rounded up (conservative): fp32-rounded coordinates may sit 1 ulp outside
            const float hc = (float)fmax(mx[c] - O[c], O[c] - mn[c]) * 1.000002f + 1.2e-7f * (float)fmax(fabs(mn[c]), fabs(mx[c])) + 1e-30f;
            tb[6 + c]  = hc;
            hsum += hc;
            tb[9 + c]  = 0.5f * (vmn[c] + vmx[c]);
            tb[12 + c] = 0.5f * (vmx[c] - vmn[c]) * 1.000001f + 1.2e-7f * fmaxf(fabsf(vmn[c]), fabsf(vmx[c])) + 1e-30f;
        }
        tb[15] = sqrtf(mmax) * 1.000001f;          // sqrt of the largest mass (m_flag criterion h2*mj), rounded up
        // bounding sphere of the position box and length of the velocity half-extents, rounded up (quick FAR test)
        tb[16] = sqrtf(tb[6] * tb[6] + tb[7] * tb[7] + tb[8] * tb[8]) * 1.000001f;
        tb[17] = sqrtf(tb[12] * tb[12] + tb[13] * tb[13] + tb[14] * tb[14]) * 1.000001f;
        tb[18] = 0.f; tb[19] = 0.f;
        if (qsum) {
            const float H = fmaxf(__uint_as_float(*hbits), 1e-30f);
            const float e = fminf(hsum / H, 8.f);
            atomicAdd(&qsum[0], (unsigned long long)(e * 1048576.f));
            __threadfence();
            if (atomicAdd(&qsum[1], 1ull) == (unsigned long long)(nloc - 1)) {
                __threadfence();
                *q_host = atomicAdd(&qsum[0], 0ull);
                qsum[0] = 0ull; qsum[1] = 0ull;
                __threadfence();
            }
        }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k = h * 32 + lane;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            tb[HDR + (C_DX + c) * TJ + k] = (float)(px[h][c] - O[c]);
            tb[HDR + (C_VX + c) * TJ + k] = pv[h][c];
            const float xh = (float)px[h][c];
            tb[HDR + (C_XH + c) * TJ + k] = xh;
            tb[HDR + (C_XL + c) * TJ + k] = (float)(px[h][c] - (double)xh);
        }
        tb[HDR + C_M * TJ + k] = pm[h];
    }
}

// ---------------------------------------------------------------------------------------------
// isort_kernel: Morton order of the i-block, so that the 32 i-particles of a warp are close in space (fewer
// NEAR tiles per warp when the block is spatially correlated).  One CTA; every thread keeps its one or two
// composite keys (30-bit Morton code << 11 | index) in registers; bitonic strides < 32 are warp shuffles, larger
// strides go through shared memory.  ~6 us for 1024 particles (the 63-bit Hilbert version took 23 us).
// ---------------------------------------------------------------------------------------------
__global__ void iota_kernel(int n, int block, int *__restrict__ p, int *__restrict__ p_host)
{   // tuning (GPUNB_B200_NOISORT): every i-block in caller order
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { p[k] = k % block; if (p_host) p_host[k] = k % block; }
}
__device__ __forceinline__ unsigned spread10(unsigned v)
{   // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8))  & 0x0300f00fu;
    v = (v | (v << 4))  & 0x030c30c3u;
    v = (v | (v << 2))  & 0x09249249u;
    return v;
}
__device__ __forceinline__ unsigned long long isort_key(const double *__restrict__ xi, int k, int ni, float sc)
{
    if (k >= ni) return ~0ull;
    unsigned q[3];
#pragma unroll
    for (int c = 0; c < 3; c++) q[c] = (unsigned)fminf(fmaxf(fmaf((float)xi[3 * (size_t)k + c], sc, 512.f), 0.f), 1023.f);
    const unsigned mkey = (spread10(q[0]) << 2) | (spread10(q[1]) << 1) | spread10(q[2]);
    return ((unsigned long long)mkey << 11) | (unsigned)k;
}
__device__ __forceinline__ unsigned long long bitonic_pick(unsigned long long v, unsigned long long o, int e, int size, int stride)
{
    const bool up = (e & size) == 0, lower = (e & stride) == 0;
    return (lower == up) ? (v < o ? v : o) : (v < o ? o : v);
}
__global__ void __launch_bounds__(1024) isort_kernel(int ni_total, int block, int sortblk, const int *__restrict__ blk_off,
                                                      const double *__restrict__ xi_all, const unsigned *__restrict__ hbits,
                                                      int *__restrict__ iperm_all, int *__restrict__ iperm_host)
{   // One CTA sorts one SORT BLOCK (<= NIMAX particles); one launch orders every block of a resident sweep.
    //   blk_off == NULL: the i-set is cut into job blocks of `block` particles (one launch group each), every job block into
    //                    sort blocks of `sortblk`; CTA b = job * ceil(block / sortblk) + sub.
    //   blk_off != NULL: ONE job; sort block b covers [blk_off[b], blk_off[b+1]) (the ranks' i-slices of a collective call).
    // iperm holds indices LOCAL to the job block.  iperm_host (optional, mapped pinned memory): the same order for the
    // host side of gpunb_regf_, which receives its result rows in sorted order.
    __shared__ unsigned long long key[NIMAX];
    const int t = threadIdx.x;
    int off, ni, in_job;
    if (blk_off) {
        off = blk_off[blockIdx.x]; ni = blk_off[blockIdx.x + 1] - off; in_job = off;
    } else {
        const int spj = (block + sortblk - 1) / sortblk;
        const int job = blockIdx.x / spj, sub = blockIdx.x - job * spj;
        const int job_ni = min(block, ni_total - job * block);
        in_job = sub * sortblk;
        ni = min(sortblk, job_ni - in_job);
        off = job * block + in_job;
    }
    if (ni <= 0) return;
    const double *xi = xi_all + 3 * (size_t)off;
    int *iperm = iperm_all + off;
    const float sc = 511.5f / fmaxf(__uint_as_float(*hbits), 1e-30f);
    int n2 = 64;
    while (n2 < ni) n2 <<= 1;
    const bool two = n2 > 1024;                        // second element of this thread: t + 1024
    unsigned long long v0 = isort_key(xi, t, ni, sc), v1 = two ? isort_key(xi, t + 1024, ni, sc) : ~0ull;
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                key[t] = v0; if (two) key[t + 1024] = v1;
                __syncthreads();
                const unsigned long long o0 = key[(t ^ stride) & (NIMAX - 1)];
                const unsigned long long o1 = two ? key[((t + 1024) ^ stride) & (NIMAX - 1)] : ~0ull;
                __syncthreads();
                if (t < n2) v0 = bitonic_pick(v0, o0, t, size, stride);
                if (two) v1 = bitonic_pick(v1, o1, t + 1024, size, stride);
            } else {
                const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, v0, stride);
                v0 = bitonic_pick(v0, o0, t, size, stride);
                if (two) { const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, v1, stride); v1 = bitonic_pick(v1, o1, t + 1024, size, stride); }
            }
        }
    }
    if (t < ni) { const int i = in_job + (int)(v0 & 2047ull); iperm[t] = i; if (iperm_host) iperm_host[off + t] = i; }
    if (two && t + 1024 < ni) { const int i = in_job + (int)(v1 & 2047ull); iperm[t + 1024] = i; if (iperm_host) iperm_host[off + t + 1024] = i; }
}

// ---------------------------------------------------------------------------------------------
// regf_kernel
// ---------------------------------------------------------------------------------------------
struct RegfArgs {
    const float  *tiles;     // jtile
    const int    *jidx;      // sorted slot -> global j (or -1)
    int           ntiles;
    // i-particles (fp64, device): element i at h2[i], dtr[i], xi[3i..], vi[3i..]
    const double *h2, *dtr, *xi, *vi;
    const int    *iperm;     // Morton order of the i-block (indices local to the block)
    int           slot0, nloc;   // this launch covers the sorted slots [slot0, slot0 + nloc) of the block; everything
                                 // it writes is indexed by the LOCAL slot kl = slot - slot0 (merge maps kl -> i)
    int           n_itiles, S, n_items;
    double       *part;      // [S][nloc][PART_STRIDE]
    int          *cnt;       // [S][nloc]
    int          *seg;       // [nloc][S][segcap]
    int           segcap;
    int           force_near;  // debugging/tuning: classify every tile as NEAR
    int           itmap;       // tuning (GPUNB_B200_ITMAP=1): work item w -> (i-tile w % n_itiles, slice w / n_itiles)
    int           near_scalar; // tuning / A-B: NEAR tiles through the scalar body (the kernel before the packed NEAR body)
    unsigned long long *stats; // optional: [0] near tiles, [1] all tiles (per warp-tile visit), [2] exact quads of NEAR tiles
    unsigned long long *wtime; // optional: per work item start/end %globaltimer (tuning)
};

struct IState {               // loop invariants of one i-particle
    float cx, cy, cz;         // (O_tile - x_i), refreshed per tile
    float nxh, nyh, nzh;      // -(float)x_i : the reference's FP32 position
    float nxl, nyl, nzl;      // -(x_i - (float)x_i): low word, so that NEAR pairs get a float-float separation
    float nvx, nvy, nvz;      // -(float)v_i
    float dtr, h2;
    float rs, adtr, slack;    // classification: sqrt(h2) and |dtr| with their safety margins, fp32 rounding of x_i
};
struct Acc {                  // FP32 partial chains (scalar view: NEAR body)
    float ax, ay, az, p, jx, jy, jz;
    __device__ __forceinline__ void clear() { ax = ay = az = p = jx = jy = jz = 0.f; }
};
struct Acc2 {                 // the same two chains (.x: even j, .y: odd j) as aligned register pairs: FAR body (f32x2)
    float2 ax, ay, az, p, jx, jy, jz;
    __device__ __forceinline__ void clear() { ax = ay = az = p = jx = jy = jz = make_float2(0.f, 0.f); }
};

// Instruction shapes are chosen from measurements on B200 (scripts/exp/farbody_exp.cu, scripts/exp/rfbank_exp.cu;
// profiles/r01d_farbody_exp.txt, profiles/r2zd_rfbank_microbench.txt):
//   * a scalar FFMA with three distinct register operands costs 1.9 issue cycles; the scalar 27-op far body runs at 38
//     cycles per pair and warp, issue-bound;
//   * packed f32x2 (FFMA2/FADD2/FMUL2 on aligned register pairs, two j per instruction): <= 2 distinct operand pairs cost
//     the 2.08 cycles of the pipe slot, 3 distinct pairs 3.07 (whatever the register numbers); an operand held in the reuse
//     cache is not read.  The packed far body runs at 34.6 cycles per pair (27 x 2.08 / 2 + 6.5 three-read FFMA2 / 2 +
//     LDS / MUFU / loop) and needs half the issue slots.  The far body is therefore packed over j, and so is the NEAR body.
__device__ __forceinline__ void accumulate(Acc &A, float rinv, float m, float rv, float dx, float dy, float dz,
                                           float dvx, float dvy, float dvz)
{   // gpunb.velocity.cu:192-207
    const float rinv2  = rinv * rinv;
    const float mrinv  = m * rinv;
    const float mrinv3 = mrinv * rinv2;
    const float rv3    = rv * (rinv2 * -3.f);                 // -3 (r.v)/r^2
    A.p += mrinv;
    A.ax = fmaf(mrinv3, dx, A.ax);   A.ay = fmaf(mrinv3, dy, A.ay);   A.az = fmaf(mrinv3, dz, A.az);
    A.jx = fmaf(mrinv3, fmaf(rv3, dx, dvx), A.jx);
    A.jy = fmaf(mrinv3, fmaf(rv3, dy, dvy), A.jy);
    A.jz = fmaf(mrinv3, fmaf(rv3, dz, dvz), A.jz);
}

// FAR tile: the bounding boxes prove that no pair of (this warp's i-particles, this tile) can satisfy the
// neighbour criterion, so the body is the force alone: 27 FP32 ops + 1 MUFU per pair, two pairs per instruction.
// Order of the accumulating FFMA2s: the triples share their first operand (reuse cache -> 2 distinct pairs).
// far_force2: the part after the separation (dx, dy, dz) and r2 of two pairs are known.
template <bool NEWTON>
__device__ __forceinline__ void far_force2(const float2 dx, const float2 dy, const float2 dz, const float2 r2,
                                           const float2 nvx, const float2 nvy, const float2 nvz, Acc2 &A,
                                           float2 VX, float2 VY, float2 VZ, float2 M)
{
    const float2 dvx = add2(VX, nvx), dvy = add2(VY, nvy), dvz = add2(VZ, nvz);
    const float2 rv = fma2(dz, dvz, fma2(dy, dvy, mul2(dx, dvx)));
    float2 rinv = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
    if (NEWTON) {        // precision option (variant suffix "n"): one Newton step, y <- y - y/2 (r2 y^2 - 1), 4 packed ops
        const float2 e = fma2(mul2(r2, rinv), rinv, dup2(-1.f));
        rinv = fma2(mul2(rinv, e), dup2(-0.5f), rinv);
    }
    const float2 rinv2  = mul2(rinv, rinv);
    const float2 mrinv  = mul2(M, rinv);
    const float2 mrinv3 = mul2(mrinv, rinv2);
    const float2 rv3    = mul2(rv, mul2(rinv2, dup2(-3.f)));            // -3 (r.v)/r^2   (gpunb.velocity.cu:192-207)
    A.p = add2(A.p, mrinv);
    const float2 ix = fma2(rv3, dx, dvx), iy = fma2(rv3, dy, dvy), iz = fma2(rv3, dz, dvz);
    A.ax = fma2(mrinv3, dx, A.ax);   A.ay = fma2(mrinv3, dy, A.ay);   A.az = fma2(mrinv3, dz, A.az);
    A.jx = fma2(mrinv3, ix, A.jx);   A.jy = fma2(mrinv3, iy, A.jy);   A.jz = fma2(mrinv3, iz, A.jz);
}
template <bool NEWTON>
__device__ __forceinline__ void interact_far2(const float2 cx, const float2 cy, const float2 cz,
                                              const float2 nvx, const float2 nvy, const float2 nvz, Acc2 &A,
                                              float2 DX, float2 DY, float2 DZ, float2 VX, float2 VY, float2 VZ, float2 M)
{
    const float2 dx = add2(DX, cx), dy = add2(DY, cy), dz = add2(DZ, cz);
    const float2 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
    far_force2<NEWTON>(dx, dy, dz, r2, nvx, nvy, nvz, A, VX, VY, VZ, M);
}

// NEAR tile: full body.
//   Predicate: the reference's, bit for bit, on the fp32-rounded inputs (gpunb.velocity.cu:168-187,
//   :235 for m_flag), in the operation order nvcc gives the reference kernel (read off its SASS:
//   FMUL dy*dy; FFMA dx,dx; FFMA dz,dz): r2 = fma(dz,dz,fma(dx,dx,dy*dy)), dxp = fma(dtr,dvx,dx), r2p likewise,
//   min(r2,r2p) < h2 [*mj].
//   Force: from the float-float separation (xh_j - xh_i) + (xl_j - xl_i) -- exact to ~2^-48 whatever the tile
//   extent -- with a Newton-refined rsqrt (a close massive perturber can dominate the sum, so the single
//   term must hold ~1e-7).  Pairs at r2 == 0 (self) never contribute
//   (regint.f:40 skips J.EQ.I); the reference GPU code returns NaN for a self pair with h2 == 0.
// Returns true for a neighbour hit.
template <bool MFLAG>
__device__ __forceinline__ bool interact_near(const IState &I, Acc &A,
                                              float VX, float VY, float VZ, float M, float XH, float YH, float ZH,
                                              float XL, float YL, float ZL)
{
    const float dxr = XH + I.nxh, dyr = YH + I.nyh, dzr = ZH + I.nzh;
    const float dx = dxr + (XL + I.nxl), dy = dyr + (YL + I.nyl), dz = dzr + (ZL + I.nzl);
    const float dvx = VX + I.nvx, dvy = VY + I.nvy, dvz = VZ + I.nvz;

    const float r2r = fmaf(dzr, dzr, fmaf(dxr, dxr, dyr * dyr));
    const float dxp = fmaf(I.dtr, dvx, dxr), dyp = fmaf(I.dtr, dvy, dyr), dzp = fmaf(I.dtr, dvz, dzr);
    const float r2p = fmaf(dzp, dzp, fmaf(dxp, dxp, dyp * dyp));
    const float r2  = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    const float rv  = fmaf(dz, dvz, fmaf(dy, dvy, dx * dvx));

    const float lim = MFLAG ? M * I.h2 : I.h2;
    const bool nb = fminf(r2r, r2p) < lim;
    float rinv = (nb || !(r2 > 0.f)) ? 0.f : rsqrt_approx(r2);
    // one Newton step: y <- y - y/2 (r2 y^2 - 1)
    const float e = fmaf(r2 * rinv, rinv, -1.f);
    rinv = fmaf(rinv * e, -0.5f, rinv);
    accumulate(A, rinv, M, rv, dx, dy, dz, dvx, dvy, dvz);
    return nb;
}

// The same body PACKED over two j (f32x2): every operation is the component-wise IEEE operation of the scalar body
// above, in the same order, so predicate and sums are bit-for-bit those of interact_near on the pair's two chains --
// at ~26 issue slots per pair instead of ~50.  Returns bit 0 / bit 1: neighbour hit of the .x / .y pair.
template <bool MFLAG>
__device__ __forceinline__ unsigned interact_near2(const IState &I, Acc2 &A, float2 VX, float2 VY, float2 VZ, float2 M,
                                                   float2 XH, float2 YH, float2 ZH, float2 XL, float2 YL, float2 ZL)
{
    const float2 dxr = add2(XH, dup2(I.nxh)), dyr = add2(YH, dup2(I.nyh)), dzr = add2(ZH, dup2(I.nzh));
    const float2 dx = add2(dxr, add2(XL, dup2(I.nxl))), dy = add2(dyr, add2(YL, dup2(I.nyl))),
                 dz = add2(dzr, add2(ZL, dup2(I.nzl)));
    const float2 dvx = add2(VX, dup2(I.nvx)), dvy = add2(VY, dup2(I.nvy)), dvz = add2(VZ, dup2(I.nvz));

    const float2 r2r = fma2(dzr, dzr, fma2(dxr, dxr, mul2(dyr, dyr)));
    const float2 dtr = dup2(I.dtr);
    const float2 dxp = fma2(dtr, dvx, dxr), dyp = fma2(dtr, dvy, dyr), dzp = fma2(dtr, dvz, dzr);
    const float2 r2p = fma2(dzp, dzp, fma2(dxp, dxp, mul2(dyp, dyp)));
    const float2 r2  = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
    const float2 rv  = fma2(dz, dvz, fma2(dy, dvy, mul2(dx, dvx)));

    const float2 lim = MFLAG ? mul2(M, dup2(I.h2)) : dup2(I.h2);
    const bool nb0 = fminf(r2r.x, r2p.x) < lim.x, nb1 = fminf(r2r.y, r2p.y) < lim.y;
    float2 rinv;
    rinv.x = (nb0 || !(r2.x > 0.f)) ? 0.f : rsqrt_approx(r2.x);
    rinv.y = (nb1 || !(r2.y > 0.f)) ? 0.f : rsqrt_approx(r2.y);
    // one Newton step: y <- y - y/2 (r2 y^2 - 1)
    const float2 e = fma2(mul2(r2, rinv), rinv, dup2(-1.f));
    rinv = fma2(mul2(rinv, e), dup2(-0.5f), rinv);
    // accumulate(): gpunb.velocity.cu:192-207
    const float2 rinv2  = mul2(rinv, rinv);
    const float2 mrinv  = mul2(M, rinv);
    const float2 mrinv3 = mul2(mrinv, rinv2);
    const float2 rv3    = mul2(rv, mul2(rinv2, dup2(-3.f)));
    A.p = add2(A.p, mrinv);
    A.ax = fma2(mrinv3, dx, A.ax);   A.ay = fma2(mrinv3, dy, A.ay);   A.az = fma2(mrinv3, dz, A.az);
    A.jx = fma2(mrinv3, fma2(rv3, dx, dvx), A.jx);
    A.jy = fma2(mrinv3, fma2(rv3, dy, dvy), A.jy);
    A.jz = fma2(mrinv3, fma2(rv3, dz, dvz), A.jz);
    return (nb0 ? 1u : 0u) | (nb1 ? 2u : 0u);
}

// OPT bit 0 (variant suffix "n"): Newton step in the FAR body (jerk-precision option, +4 packed ops per two pairs).
// OPT bit 1 (variant suffix "t", IT = 1): NEAR tiles in which only a few lanes fail the FAR test are TRANSPOSED -- every
//   lane runs the cheap FAR body, the failing lanes throw their tile sums away and get them back from the whole warp:
//   their i-state is broadcast by shuffles, the 32 lanes take two j of the tile each through the exact NEAR body, the
//   seven sums are reduced by a butterfly and neighbour hits are appended by the hitting lanes.  When the i-particles of a
//   warp are far apart (the rule for an i-block picked by time step, not by position) a NEAR tile holds neighbours of one or
//   two lanes only, and the other ~30 lanes no longer pay the exact body for nothing.
template <int IT, bool MFLAG, int MINB, int OPT>
__global__ void __launch_bounds__(WARPS * 32, MINB) regf_kernel(const RegfArgs a)
{
    constexpr int ITILE = 32 * IT;
    constexpr bool NEWTON = (OPT & 1) != 0;
    constexpr bool TRANSPOSE = (OPT & 2) != 0 && IT == 1;
    constexpr bool QUICK  = (OPT & 4) != 0 && IT == 1;   // "q": bounding-sphere FAR test first, the box test only when a lane fails it
    constexpr bool FLUSH2 = (OPT & 8) != 0;              // "f": FP32 chains of 64 terms (flush every second tile)
    constexpr int  UQ     = (OPT & 16) ? 4 : FAR_UNROLL_Q;   // "u": four quads of j per iteration of the FAR loop
    constexpr int  TRANSPOSE_MAX = 8;                 // more failing lanes than this: the whole warp runs the NEAR body
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * WARPS + warp;
    if (w >= a.n_items) return;                       // warp-uniform; no CTA-wide barrier is used below
    float    *buf  = reinterpret_cast<float *>(smem_raw) + warp * NSTAGE * TILE_FLOATS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + WARPS * NSTAGE * TILE_BYTES) + warp * NSTAGE;

    // this warp visits tiles s, s+S, s+2S, ... for i-tile it; itmap: the warps of a CTA take DIFFERENT i-tiles and the same s
    const int it = a.itmap ? w % a.n_itiles : w / a.S, s = a.itmap ? w / a.n_itiles : w - (w / a.S) * a.S;
    if (it >= a.n_itiles || s >= a.S) return;
    if (a.wtime && lane == 0) { unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); a.wtime[3 * w] = t0; }

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NSTAGE; k++) mbar_init(&bars[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NSTAGE; k++)
            if (s + k * a.S < a.ntiles) {
                mbar_expect_tx(&bars[k], TILE_BYTES);
                tma_bulk_g2s(buf + k * TILE_FLOATS, a.tiles + (size_t)(s + k * a.S) * TILE_FLOATS, TILE_BYTES, &bars[k]);
            }
    }

    // i-particles of this lane (Morton-ordered block: the warp's i-particles are close in space)
    IState I[IT];
    Acc2   P[IT];             // two chains per quantity (.x even / .y odd j): 32-term FP32 chains
    double D[IT][7];
    int    cnt[IT], iidx[IT];     // iidx: local slot kl of the lane's i-particle (-1: padding lane)
    int   *segp[IT];
#pragma unroll
    for (int k = 0; k < IT; k++) {
        const int kl = it * ITILE + k * 32 + lane;
        const bool valid = kl < a.nloc;
        const int i = valid ? a.iperm[a.slot0 + kl] : -1;
        iidx[k] = valid ? kl : -1;
        double xd[3] = {0, 0, 0}, v[3] = {0, 0, 0}, h2 = 0, dtr = 0;
        if (valid) {
            h2 = a.h2[i]; dtr = a.dtr[i];
#pragma unroll
            for (int c = 0; c < 3; c++) { xd[c] = a.xi[3 * (size_t)i + c]; v[c] = a.vi[3 * (size_t)i + c]; }
        }
        const float xh[3] = {(float)xd[0], (float)xd[1], (float)xd[2]};
        const float vf[3] = {(float)v[0], (float)v[1], (float)v[2]};
        I[k].nxh = -xh[0]; I[k].nyh = -xh[1]; I[k].nzh = -xh[2];
        I[k].nxl = -(float)(xd[0] - (double)xh[0]); I[k].nyl = -(float)(xd[1] - (double)xh[1]);
        I[k].nzl = -(float)(xd[2] - (double)xh[2]);
        I[k].nvx = -vf[0]; I[k].nvy = -vf[1]; I[k].nvz = -vf[2];
        I[k].dtr = (float)dtr;
        I[k].h2  = valid ? (float)h2 : 0.f;
        I[k].slack = 1.2e-7f * fmaxf(fabsf(xh[0]), fmaxf(fabsf(xh[1]), fabsf(xh[2])));
        I[k].rs    = sqrtf(fmaxf(I[k].h2, 0.f)) * 1.0001f;
        I[k].adtr  = fabsf(I[k].dtr) * 1.00002f;
        P[k].clear();
#pragma unroll
        for (int c = 0; c < 7; c++) D[k][c] = 0.0;
        cnt[k]  = 0;
        segp[k] = a.seg + ((size_t)(valid ? kl : 0) * a.S + s) * a.segcap;
    }
    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < IT; k++) {
            D[k][0] += (double)(P[k].ax.x + P[k].ax.y);
            D[k][1] += (double)(P[k].ay.x + P[k].ay.y);
            D[k][2] += (double)(P[k].az.x + P[k].az.y);
            D[k][3] += (double)(P[k].jx.x + P[k].jx.y);
            D[k][4] += (double)(P[k].jy.x + P[k].jy.y);
            D[k][5] += (double)(P[k].jz.x + P[k].jz.y);
            D[k][6] += (double)(P[k].p.x + P[k].p.y);
            P[k].clear();
        }
    };

    unsigned n_near = 0, n_all = 0, n_tr = 0;
    int n = 0;
    for (int t = s; t < a.ntiles; t += a.S, n++) {
        const int st = n % NSTAGE;
        const uint32_t parity = (n / NSTAGE) & 1;
        mbar_wait(&bars[st], parity);
        const float *tb = buf + st * TILE_FLOATS;
        // ---- header: classify the (warp, tile) pair (branch-free, fp32 only) -------------------------
        const float4 h0 = reinterpret_cast<const float4 *>(tb)[0];      // Oh.xyz | Ol.x
        const float4 h1 = reinterpret_cast<const float4 *>(tb)[1];      // Ol.yz  | hx hy
        const float4 h2v = reinterpret_cast<const float4 *>(tb)[2];     // hz | vcx vcy vcz
        const float4 h3 = reinterpret_cast<const float4 *>(tb)[3];      // hvx hvy hvz | sqrt(mmax)
        // Separation (tile origin - x_i) from the two-float representations: |c| ~ pair distance, so c + offset
        // keeps a relative precision of ~2^-23 whatever |x| is.
#pragma unroll
        for (int k = 0; k < IT; k++) {
            I[k].cx = (h0.x + I[k].nxh) + (h0.w + I[k].nxl);
            I[k].cy = (h0.y + I[k].nyh) + (h1.x + I[k].nyl);
            I[k].cz = (h0.z + I[k].nzh) + (h1.y + I[k].nzl);
        }
        // Per-lane test of i against the tile's boxes: FAR iff no j of the tile can satisfy the reference
        // criterion min(|dx|^2, |dx + dtr dv|^2) < h2 [* mj]:  gap(i, box) > sqrt(h2 [* mmax]) + |dtr| max|dv|,
        // with margins that dominate every fp32 rounding (positions/velocities rounded to fp32 by the predicate,
        // this arithmetic itself, the approximate sqrt).
        bool lane_far = true;
        bool need_box = true;
        if (QUICK) {
            // Bounding-sphere version of the test below (a proof of FAR by the same argument with coarser bounds: gap to the
            // sphere <= gap to the box, |dv| <= |v_tile centre - v_i| + |velocity half-extents|): ~17 instructions instead of
            // ~35; the box test runs only for the (warp, tile) visits in which some lane fails it.
            const float4 h4 = reinterpret_cast<const float4 *>(tb)[4];      // R | VR | - | -
            const float c2 = fmaf(I[0].cz, I[0].cz, fmaf(I[0].cy, I[0].cy, I[0].cx * I[0].cx));
            float dist; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(dist) : "f"(c2));
            const float g = fmaf(dist, 0.999999f, -(h4.x + 2.f * I[0].slack));
            const float wx = h2v.y + I[0].nvx, wy = h2v.z + I[0].nvy, wz = h2v.w + I[0].nvz;
            float wn; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(wn) : "f"(fmaf(wz, wz, fmaf(wy, wy, wx * wx))));
            const float reach = fmaf(I[0].adtr, fmaf(wn, 1.000001f, h4.y), MFLAG ? I[0].rs * h3.w : I[0].rs);
            const bool qf = (g > reach) && (g > 0.5f * fmaf(dist, 1.000001f, h4.x));
            need_box = __any_sync(0xffffffffu, !(qf || iidx[0] < 0));
        }
        if (need_box) {
            const float jh[3] = {h1.z, h1.w, h2v.x};
            const float jvc[3] = {h2v.y, h2v.z, h2v.w};
            const float jvh[3] = {h3.x, h3.y, h3.z};
#pragma unroll
            for (int k = 0; k < IT; k++) {
                const float cf[3] = {I[k].cx, I[k].cy, I[k].cz};
                const float nv[3] = {I[k].nvx, I[k].nvy, I[k].nvz};
                float d2 = 0.f, dv2 = 0.f, sreach = 0.f;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const float g = fmaxf(fmaf(fabsf(cf[c]), 0.9999996f, -(jh[c] + I[k].slack)), 0.f);
                    d2 = fmaf(g, g, d2);
                    const float u = fabsf(jvc[c] + nv[c]) + jvh[c];
                    dv2 = fmaf(u, u, dv2);
                    sreach = fmaxf(sreach, fabsf(cf[c]) + jh[c]);
                }
                float sq; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(dv2));
                const float rr = fmaf(I[k].adtr, sq, MFLAG ? I[k].rs * h3.w : I[k].rs);
                // precision: the FAR body forms dx = c + offset in fp32 (error ~2^-23 (|c|+h)); keep that below
                // ~1e-7 of the smallest separation by sending tiles closer than half their own reach NEAR
                const bool f = (d2 > rr * rr) && (d2 > 0.25f * sreach * sreach);
                lane_far &= (f || iidx[k] < 0);
            }
        }
        const unsigned nearmask = __ballot_sync(0xffffffffu, !lane_far);
        const bool transposed = TRANSPOSE && nearmask != 0u && __popc(nearmask) <= TRANSPOSE_MAX && !a.force_near;
        const bool far = (nearmask == 0u && !a.force_near) || transposed;
        n_all++;
        const float4 *c = reinterpret_cast<const float4 *>(tb + HDR);
        float2 cx2[IT], cy2[IT], cz2[IT], nvx2[IT], nvy2[IT], nvz2[IT];
#pragma unroll
        for (int k = 0; k < IT; k++) {
            cx2[k] = dup2(I[k].cx); cy2[k] = dup2(I[k].cy); cz2[k] = dup2(I[k].cz);
            nvx2[k] = dup2(I[k].nvx); nvy2[k] = dup2(I[k].nvy); nvz2[k] = dup2(I[k].nvz);
        }
        if (FLUSH2 && TRANSPOSE && transposed) flush();   // the transposed path replaces this tile's sums of the failing lanes
        if (far) {
#pragma unroll UQ
            for (int q = 0; q < TJ / 4; q++) {
                const float4 DX = c[C_DX * 16 + q], DY = c[C_DY * 16 + q], DZ = c[C_DZ * 16 + q];
                const float4 VX = c[C_VX * 16 + q], VY = c[C_VY * 16 + q], VZ = c[C_VZ * 16 + q];
                const float4 M  = c[C_M * 16 + q];
#pragma unroll
                for (int k = 0; k < IT; k++) {
                    interact_far2<NEWTON>(cx2[k], cy2[k], cz2[k], nvx2[k], nvy2[k], nvz2[k], P[k], make_float2(DX.x, DX.y),
                                  make_float2(DY.x, DY.y), make_float2(DZ.x, DZ.y), make_float2(VX.x, VX.y),
                                  make_float2(VY.x, VY.y), make_float2(VZ.x, VZ.y), make_float2(M.x, M.y));
                    interact_far2<NEWTON>(cx2[k], cy2[k], cz2[k], nvx2[k], nvy2[k], nvz2[k], P[k], make_float2(DX.z, DX.w),
                                  make_float2(DY.z, DY.w), make_float2(DZ.z, DZ.w), make_float2(VX.z, VX.w),
                                  make_float2(VY.z, VY.w), make_float2(VZ.z, VZ.w), make_float2(M.z, M.w));
                }
            }
            if (TRANSPOSE && transposed) {
                // The failing lanes' FAR sums of this tile are void (neighbours inside, or the tile too close for the
                // tile-local separations): cleared here, recomputed exactly by the whole warp below.
                n_near++; n_tr++;
                if (!lane_far) P[0].clear();
                const float2 *c2 = reinterpret_cast<const float2 *>(tb + HDR);      // this lane's two j: 2 lane, 2 lane + 1
                const float2 jVX = c2[C_VX * 32 + lane], jVY = c2[C_VY * 32 + lane], jVZ = c2[C_VZ * 32 + lane], jM = c2[C_M * 32 + lane];
                const float2 jXH = c2[C_XH * 32 + lane], jYH = c2[C_YH * 32 + lane], jZH = c2[C_ZH * 32 + lane];
                const float2 jXL = c2[C_XL * 32 + lane], jYL = c2[C_YL * 32 + lane], jZL = c2[C_ZL * 32 + lane];
                for (unsigned rest = nearmask; rest; rest &= rest - 1u) {
                    const int L = __ffs(rest) - 1;
                    IState B;                          // i-state of lane L (only the fields the NEAR body reads)
                    B.nxh = __shfl_sync(0xffffffffu, I[0].nxh, L); B.nyh = __shfl_sync(0xffffffffu, I[0].nyh, L);
                    B.nzh = __shfl_sync(0xffffffffu, I[0].nzh, L); B.nxl = __shfl_sync(0xffffffffu, I[0].nxl, L);
                    B.nyl = __shfl_sync(0xffffffffu, I[0].nyl, L); B.nzl = __shfl_sync(0xffffffffu, I[0].nzl, L);
                    B.nvx = __shfl_sync(0xffffffffu, I[0].nvx, L); B.nvy = __shfl_sync(0xffffffffu, I[0].nvy, L);
                    B.nvz = __shfl_sync(0xffffffffu, I[0].nvz, L); B.dtr = __shfl_sync(0xffffffffu, I[0].dtr, L);
                    B.h2  = __shfl_sync(0xffffffffu, I[0].h2, L);
                    Acc2 Q; Q.clear();
                    const unsigned hit = interact_near2<MFLAG>(B, Q, jVX, jVY, jVZ, jM, jXH, jYH, jZH, jXL, jYL, jZL);
                    // neighbour hits: appended by the hitting lanes at positions handed out by ballots (ghost slots of the
                    // last tile are never neighbours); lane L keeps the count
                    bool v0 = false, v1 = false;
                    int jg0 = -1, jg1 = -1;
                    if (hit) {
                        const int2 jg = reinterpret_cast<const int2 *>(a.jidx + (size_t)t * TJ)[lane];
                        jg0 = jg.x; jg1 = jg.y;
                        v0 = (hit & 1u) && jg0 >= 0; v1 = (hit & 2u) && jg1 >= 0;
                    }
                    const unsigned b0 = __ballot_sync(0xffffffffu, v0), b1 = __ballot_sync(0xffffffffu, v1);
                    if (b0 | b1) {
                        const int cntL = __shfl_sync(0xffffffffu, cnt[0], L);
                        const unsigned lt = (1u << lane) - 1u;
                        int pos = cntL + __popc(b0 & lt) + __popc(b1 & lt);
                        int *sp = a.seg + ((size_t)(it * ITILE + L) * a.S + s) * a.segcap;
                        if (v0) { if (pos < a.segcap) sp[pos] = jg0; pos++; }
                        if (v1) { if (pos < a.segcap) sp[pos] = jg1; }
                        if (lane == L) cnt[0] = cntL + __popc(b0) + __popc(b1);
                    }
                    float r[7] = {Q.ax.x + Q.ax.y, Q.ay.x + Q.ay.y, Q.az.x + Q.az.y, Q.p.x + Q.p.y,
                                  Q.jx.x + Q.jx.y, Q.jy.x + Q.jy.y, Q.jz.x + Q.jz.y};
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                        for (int q = 0; q < 7; q++) r[q] += __shfl_xor_sync(0xffffffffu, r[q], o);
                    }
                    if (lane == L) {
                        P[0].ax.x = r[0]; P[0].ay.x = r[1]; P[0].az.x = r[2]; P[0].p.x = r[3];
                        P[0].jx.x = r[4]; P[0].jy.x = r[5]; P[0].jz.x = r[6];
                    }
                }
            }
        } else {
            // NEAR tile: the full body (reference predicate on the fp32-rounded positions, two-float separation,
            // Newton-refined rsqrt), packed over j like the FAR body.  A two-pass variant (packed far body first, exact
            // fix-up of the quads in which any lane has a pair inside its neighbour bound) was measured and dropped:
            // 45 % of the quads of NEAR tiles need the fix-up, because a tile is smaller than a neighbour sphere
            // (profiles/r01t_pipeline_probe_1gpu.txt).
            n_near++;
#pragma unroll 1
            for (int q = 0; q < TJ / 4; q++) {
                const float4 VX = c[C_VX * 16 + q], VY = c[C_VY * 16 + q], VZ = c[C_VZ * 16 + q];
                const float4 M  = c[C_M * 16 + q];
                const float4 XH = c[C_XH * 16 + q], YH = c[C_YH * 16 + q], ZH = c[C_ZH * 16 + q];
                const float4 XL = c[C_XL * 16 + q], YL = c[C_YL * 16 + q], ZL = c[C_ZL * 16 + q];
                unsigned hit = 0;
#ifdef NEAR_SCALAR_AB
                if (a.near_scalar) {                   // A/B build: the scalar body, bit-for-bit the same results
#pragma unroll
                    for (int k = 0; k < IT; k++) {
                        Acc A0 = Acc{P[k].ax.x, P[k].ay.x, P[k].az.x, P[k].p.x, P[k].jx.x, P[k].jy.x, P[k].jz.x};
                        Acc A1 = Acc{P[k].ax.y, P[k].ay.y, P[k].az.y, P[k].p.y, P[k].jx.y, P[k].jy.y, P[k].jz.y};
                        const bool h0 = interact_near<MFLAG>(I[k], A0, VX.x, VY.x, VZ.x, M.x, XH.x, YH.x, ZH.x, XL.x, YL.x, ZL.x);
                        const bool h1b = interact_near<MFLAG>(I[k], A1, VX.y, VY.y, VZ.y, M.y, XH.y, YH.y, ZH.y, XL.y, YL.y, ZL.y);
                        const bool h2b = interact_near<MFLAG>(I[k], A0, VX.z, VY.z, VZ.z, M.z, XH.z, YH.z, ZH.z, XL.z, YL.z, ZL.z);
                        const bool h3b = interact_near<MFLAG>(I[k], A1, VX.w, VY.w, VZ.w, M.w, XH.w, YH.w, ZH.w, XL.w, YL.w, ZL.w);
                        hit |= ((h0 ? 1u : 0u) | (h1b ? 2u : 0u) | (h2b ? 4u : 0u) | (h3b ? 8u : 0u)) << (4 * k);
                        P[k].ax = make_float2(A0.ax, A1.ax); P[k].ay = make_float2(A0.ay, A1.ay);
                        P[k].az = make_float2(A0.az, A1.az); P[k].p  = make_float2(A0.p,  A1.p);
                        P[k].jx = make_float2(A0.jx, A1.jx); P[k].jy = make_float2(A0.jy, A1.jy);
                        P[k].jz = make_float2(A0.jz, A1.jz);
                    }
                } else
#endif
#pragma unroll
                for (int k = 0; k < IT; k++) {
                    const unsigned ha = interact_near2<MFLAG>(I[k], P[k], make_float2(VX.x, VX.y), make_float2(VY.x, VY.y),
                        make_float2(VZ.x, VZ.y), make_float2(M.x, M.y), make_float2(XH.x, XH.y), make_float2(YH.x, YH.y),
                        make_float2(ZH.x, ZH.y), make_float2(XL.x, XL.y), make_float2(YL.x, YL.y), make_float2(ZL.x, ZL.y));
                    const unsigned hb = interact_near2<MFLAG>(I[k], P[k], make_float2(VX.z, VX.w), make_float2(VY.z, VY.w),
                        make_float2(VZ.z, VZ.w), make_float2(M.z, M.w), make_float2(XH.z, XH.w), make_float2(YH.z, YH.w),
                        make_float2(ZH.z, ZH.w), make_float2(XL.z, XL.w), make_float2(YL.z, YL.w), make_float2(ZL.z, ZL.w));
                    hit |= (ha | (hb << 2)) << (4 * k);
                }
                if (hit) {                             // rare: ~2e-4 of pairs are neighbours
                    const int pb = t * TJ + q * 4;
#pragma unroll
                    for (int k = 0; k < IT; k++) {
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            if ((hit >> (4 * k + b)) & 1u) {
                                const int jg = a.jidx[pb + b];
                                if (jg >= 0) {         // ghost slots of the last tile are never neighbours
                                    if (cnt[k] < a.segcap) segp[k][cnt[k]] = jg;
                                    cnt[k]++;
                                }
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();                                  // every lane is done reading stage st
        if (lane == 0 && t + NSTAGE * a.S < a.ntiles) {
            mbar_expect_tx(&bars[st], TILE_BYTES);
            tma_bulk_g2s(buf + st * TILE_FLOATS, a.tiles + (size_t)(t + NSTAGE * a.S) * TILE_FLOATS, TILE_BYTES, &bars[st]);
        }
        if (!FLUSH2 || (n & 1)) flush();               // FP32 chains of 32 (64) terms -> fp64
    }
    if (FLUSH2) flush();

#pragma unroll
    for (int k = 0; k < IT; k++) {
        const int kl = iidx[k];
        if (kl >= 0) {
            double *o = a.part + ((size_t)s * a.nloc + kl) * PART_STRIDE;
#pragma unroll
            for (int c = 0; c < 7; c++) o[c] = D[k][c];
            a.cnt[(size_t)s * a.nloc + kl] = cnt[k];
        }
    }
    if (a.stats && lane == 0) { atomicAdd(&a.stats[0], (unsigned long long)n_near); atomicAdd(&a.stats[1], (unsigned long long)n_all); atomicAdd(&a.stats[2], (unsigned long long)n_tr); }
    if (a.wtime && lane == 0) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); a.wtime[3 * w + 1] = t1;
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.wtime[3 * w + 2] = (unsigned long long)n_near | ((unsigned long long)smid << 32); }
}

// ---------------------------------------------------------------------------------------------
// Warp-wide ascending sort of `total` ints held in shared memory sb[0..n2) (n2 = power of two >= 32,
// entries >= total set to INT_MAX by the caller).  Up to 512 elements are sorted in REGISTERS (element
// e = lane + 32 k lives in v[k] of lane): bitonic strides < 32 are warp shuffles, larger strides are
// register swaps -- no shared-memory round trips or warp barriers inside the network.  Larger rows fall
// back to the shared-memory network.
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void warp_bitonic_regs(int *sb, int lane)
{
    int v[K];
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = sb[lane + 32 * k];
#pragma unroll
    for (int size = 2; size <= 32 * K; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int ks = stride >> 5;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if ((k & ks) == 0) {
                        const bool up = (((lane + 32 * k) & size) == 0);
                        const int x = v[k], y = v[k | ks];
                        const bool sw = (x > y) == up;
                        v[k] = sw ? y : x; v[k | ks] = sw ? x : y;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int o = __shfl_xor_sync(0xffffffffu, v[k], stride);
                    const bool up = (((lane + 32 * k) & size) == 0);
                    const bool lower = (lane & stride) == 0;
                    v[k] = (lower == up) ? min(v[k], o) : max(v[k], o);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) sb[lane + 32 * k] = v[k];
}

__device__ __noinline__ void warp_sort_smem(int *sb, int n2, int lane)
{
    __syncwarp();
    switch (n2) {
        case 32:  warp_bitonic_regs<1>(sb, lane); break;
        case 64:  warp_bitonic_regs<2>(sb, lane); break;
        case 128: warp_bitonic_regs<4>(sb, lane); break;
        case 256: warp_bitonic_regs<8>(sb, lane); break;
        case 512: warp_bitonic_regs<16>(sb, lane); break;
        default:
            for (int size = 2; size <= n2; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int k = lane; k < (n2 >> 1); k += 32) {
                        const int lo = 2 * k - (k & (stride - 1));
                        const int hi = lo + stride;
                        const bool up = (lo & size) == 0;
                        const int x = sb[lo], y = sb[hi];
                        if ((x > y) == up) { sb[lo] = y; sb[hi] = x; }
                    }
                    __syncwarp();
                }
            }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// merge_kernel: one warp per i-particle.
//   fp64 sum over the S slices in a fixed order (deterministic); gather of the S neighbour segments into
//   shared memory; bitonic sort ascending -- tiles are visited in Morton order, the caller needs strictly
//   ascending j (its two-pointer list diff: regcor_gpu.F:299-336).
//   count > nnbmax  ->  list[0] = -count, no entries written (reg.avx.cpp:320-321).
// ---------------------------------------------------------------------------------------------
constexpr int MAX_RANKS = 16;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Loads of peer memory that another GPU wrote during this kernel's lifetime: system-scope, never served from a
// stale L1 line.  (ld.global.cv / .cg / ld.volatile measured the same on NVLink 5.)
__device__ __forceinline__ double peer_ld(const double *p)
{
    double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}
__device__ __forceinline__ int peer_ld(const int *p)
{
    int v; asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}

// In-kernel publication to the peers (one process per GPU; see run_job).  The LAST CTA of a launch to finish
// stores `seq` into this rank's entry of every peer's flag (merge: "my shard rows of call seq are complete") or ack
// (combine: "I have finished reading everybody's rows of call seq") array over NVLink -- no extra launch.
struct ExchSignal {
    unsigned *done_ctr;                        // device-local count of finished CTAs (reset by the last one); NULL: off
    unsigned long long *peer[MAX_RANKS];       // &array_of_peer_r[this rank]
    int R;
    unsigned long long seq;
};
__device__ __forceinline__ void publish_when_last(const ExchSignal &g)
{   // called by EVERY thread of EVERY CTA, after its last global write / peer read
    if (!g.done_ctr) return;
    __shared__ int is_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(g.done_ctr, 1u);
        is_last = (prev == gridDim.x - 1);
        if (is_last) { *g.done_ctr = 0; __threadfence(); }
    }
    __syncthreads();
    if (is_last && (int)threadIdx.x < g.R) {
        __threadfence_system();
        st_release_sys(g.peer[threadIdx.x], g.seq);
    }
}
// Bounded: a peer that died or fell out of step (different calls on different ranks) must not wedge the other ranks inside
// a kernel for ever.  After g_spin_timeout_ns (default 60 s, GPUNB_B200_SPIN_TIMEOUT_S) the waiting kernel reports which
// rank's flag is stuck and traps; the host side then aborts with the CUDA error like for any other device fault.
__device__ unsigned long long g_spin_timeout_ns = 60000000000ull;
__device__ __forceinline__ void wait_all_ranks(const unsigned long long *flags, int R, long long need, int lane)
{   // every warp for itself: lanes < R spin on the R entries of a flag array in LOCAL memory (written by the peers)
    if (lane < R) {
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while ((long long)ld_acquire_sys(flags + lane) < need) {
            if ((++spins & 0xfffu) == 0u) {
                unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t;
                else if (t - t0 > g_spin_timeout_ns) {
                    printf("gpunb_b200: exchange step %lld: the flag of rank %d is stuck at %lld -- a peer rank died or the ranks "
                           "do not make identical calls\n", need, lane, (long long)ld_acquire_sys(flags + lane));
                    __trap();
                }
            }
        }
    }
    __syncwarp();
}

struct MergeArgs {
    const double *part; const int *cnt; const int *seg;
    int nloc, S, segcap, lmax, nnbmax;
    int kl0, kl1;       // this launch merges the local slots [kl0, kl1) (the whole job unless the delivery is cut into parts)
    const int *iperm;   // output row of local slot kl: i = iperm[kl] (the caller passes iperm + slot0); NULL: i = kl
    double *res_f;      // [.][f_stride]; f_stride = 8 stores the (signed) count in slot 7 for the shard combine
    double *abi_acc, *abi_jrk, *abi_pot;   // non-NULL: final sums go straight to the caller's acc[.][3] / jrk[.][3] / pot[.]
                                           // (host arrays the caller has pinned, gpunb_b200_pin_host_) instead of res_f
    int     f_stride;
    int    *res_list;   // [.][lmax]
    int    *res_list2;  // optional second copy of the final rows in DEVICE memory (gpunb_b200_regcor_last_ reads it); NULL: none
    int     sort;       // 0: leave the row in arrival order (a shard row: combine_kernel sorts the union)
    // exchange slot reuse (one process per GPU): wait until every peer has acknowledged reading the previous
    // contents (acks[r] >= ack_need) before overwriting res_f / res_list.  NULL: no wait.
    const unsigned long long *acks; long long ack_need; int R;
    ExchSignal sig;
};

__device__ __forceinline__ void merge_row(const MergeArgs &a, int kl, int lane, int *sb)
{
    double f[7] = {0, 0, 0, 0, 0, 0, 0};
    int total = 0;
#pragma unroll 4
    for (int s = lane; s < a.S; s += 32) {             // loads of several rounds in flight
        const double *p = a.part + ((size_t)s * a.nloc + kl) * PART_STRIDE;
#pragma unroll
        for (int c = 0; c < 7; c++) f[c] += p[c];
        total += a.cnt[(size_t)s * a.nloc + kl];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 7; c++) f[c] += __shfl_xor_sync(0xffffffffu, f[c], o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    const int i = a.iperm ? a.iperm[kl] : kl;
    if (a.abi_acc) {
        if (lane < 3) a.abi_acc[3 * (size_t)i + lane] = f[lane];
        else if (lane < 6) a.abi_jrk[3 * (size_t)i + lane - 3] = f[lane];
        else if (lane == 6) a.abi_pot[i] = f[6];
    } else if (lane < 7) a.res_f[(size_t)i * a.f_stride + lane] = f[lane];   // f[] is uniform after the butterfly
    if (a.f_stride == 8 && lane == 7) a.res_f[(size_t)i * 8 + 7] = (double)(total > a.nnbmax ? -total : total);
    int *row = a.res_list + (size_t)i * a.lmax;
    int *row2 = a.res_list2 ? a.res_list2 + (size_t)i * a.lmax : nullptr;
    if (total > a.nnbmax) { if (lane == 0) { row[0] = -total; if (row2) row2[0] = -total; } return; }
    if (lane == 0) { row[0] = total; if (row2) row2[0] = total; }
    if (total == 0) return;
    int base = 0;
    for (int s0 = 0; s0 < a.S; s0 += 32) {
        const int s = s0 + lane;
        const int n = (s < a.S) ? a.cnt[(size_t)s * a.nloc + kl] : 0;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        const int off = base + incl - n;
        const int *src = a.seg + ((size_t)kl * a.S + s) * a.segcap;
        for (int k = 0; k < n; k++) sb[off + k] = src[k];
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (a.sort) {
        int n2 = 32;
        while (n2 < total) n2 <<= 1;
        for (int k = total + lane; k < n2; k += 32) sb[k] = INT_MAX;
        warp_sort_smem(sb, n2, lane);
    } else {
        __syncwarp();
    }
    for (int k = lane; k < total; k += 32) row[1 + k] = sb[k];
    if (row2) for (int k = lane; k < total; k += 32) row2[1 + k] = sb[k];
}

__global__ void __launch_bounds__(128) merge_kernel(const MergeArgs a)
{
    __shared__ int sbuf[4][SORT_CAP];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    int kl = a.kl0 + blockIdx.x * 4 + wq;              // grid-stride over the rows (the grid is capped when the kernel may spin)
    if (kl < a.kl1 && a.acks) wait_all_ranks(a.acks, a.R, a.ack_need, lane);
    for (; kl < a.kl1; kl += gridDim.x * 4) {
        merge_row(a, kl, lane, sbuf[wq]);
        __syncwarp();
    }
    publish_when_last(a.sig);
}

// ---------------------------------------------------------------------------------------------
// combine_kernel: j-shard exchange step (multi-GPU).  Every shard r (a GPU holding every R-th tile of the
// Hilbert-sorted j-set, see shard_tiles) has produced, per i-particle, 7 fp64 partial sums + its signed
// neighbour count (fr[r][.][8]) and a row of GLOBAL j indices (rows[r][.][lmax]).  One warp per i:
// fp64 sum over shards in rank order (the reference sums GPUs in fp64 on the host, :823-845), counts scanned in
// rank order, rows gathered and sorted ascending (the reference's index-range shards only need concatenating,
// :852-871; spatial shards interleave in j).
// fr[r] / rows[r] are PEER pointers (NVLink P2P: cudaIpc-mapped across processes, or peer-enabled
// devices of one process): the kernel pulls only the `count` valid entries of each remote row, so the
// exchange moves ~4*nnb bytes per i instead of whole rows.  Overflow: any shard negative or
// total > nnbmax -> -(sum |count_r|) (reg.avx.cpp:320-321 encoding of the true count).
// Shard records are indexed by the local sorted slot kl (every rank computes the same order); the combined row goes
// to output row iperm[kl].
// ---------------------------------------------------------------------------------------------
struct CombineArgs {
    int nloc, R, lmax, nnbmax;
    int kl0, kl1;                    // this launch combines the local slots [kl0, kl1) of the job: everything (replicated
                                     // results) or the slots of this rank's own i-slice (i-slice mode)
    int row_base;                    // output row of slot kl: iperm[kl] - row_base
    const double *fr[MAX_RANKS];     // [nloc][8]
    const int    *rows[MAX_RANKS];   // [nloc][lmax]
    const int    *iperm;             // output row of kl (NULL: kl)
    double *res_f;                   // [.][7]
    double *abi_acc, *abi_jrk, *abi_pot;   // non-NULL: the caller's own (pinned) arrays instead of res_f
    int    *res_list;                // [.][lmax]
    int    *res_list2;               // optional device copy of the final rows (see MergeArgs)
    // one process per GPU: flags[r] (in THIS rank's exchange buffer) is set to `seq` by rank r, over NVLink, once
    // its fr/rows of this call are complete (last CTA of its merge_kernel).  NULL: ordering is done with stream events.
    const unsigned long long *flags;
    unsigned long long seq;
    ExchSignal ack;                  // tells every peer that this rank is done reading call `seq`
};

__device__ __forceinline__ void combine_row(const CombineArgs &a, int kl, int lane, int *sb, int *offs, int *cnts)
{
    // records are requested four shards at a time before any is used: R/4 NVLink round trips, not R
    double f = 0.0;
    int total = 0;
    bool over = false;
    for (int r0 = 0; r0 < a.R; r0 += 4) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
            v[q] = (r0 + q < a.R && lane < 8) ? peer_ld(a.fr[r0 + q] + (size_t)kl * 8 + lane) : 0.0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (r0 + q < a.R) {
                f += v[q];                             // rank order: deterministic
                const int c = (int)__shfl_sync(0xffffffffu, v[q], 7);
                over |= c < 0;
                const int ca = c < 0 ? -c : c;
                if (lane == 0) { offs[r0 + q] = total; cnts[r0 + q] = ca; }
                total += ca;
            }
        }
    }
    const int i = (a.iperm ? a.iperm[kl] : kl) - a.row_base;
    if (a.abi_acc) {
        if (lane < 3) a.abi_acc[3 * (size_t)i + lane] = f;
        else if (lane < 6) a.abi_jrk[3 * (size_t)i + lane - 3] = f;
        else if (lane == 6) a.abi_pot[i] = f;
    } else if (lane < 7) a.res_f[(size_t)i * 7 + lane] = f;
    int *row = a.res_list + (size_t)i * a.lmax;
    int *row2 = a.res_list2 ? a.res_list2 + (size_t)i * a.lmax : nullptr;
    if (over || total > a.nnbmax) { if (lane == 0) { row[0] = -total; if (row2) row2[0] = -total; } return; }
    if (lane == 0) { row[0] = total; if (row2) row2[0] = total; }
    if (total == 0) return;
    __syncwarp();
    // flat gather: entry e of the union comes from the shard whose [off, off + cnt) contains it
    {
        int r = 0, o = offs[0], c = cnts[0];
        for (int e = lane; e < total; e += 32) {
            while (e >= o + c) { r++; o = offs[r]; c = cnts[r]; }
            sb[e] = peer_ld(a.rows[r] + (size_t)kl * a.lmax + 1 + (e - o));
        }
    }
    __syncwarp();
    int n2 = 32;
    while (n2 < total) n2 <<= 1;
    for (int k = total + lane; k < n2; k += 32) sb[k] = INT_MAX;
    warp_sort_smem(sb, n2, lane);
    for (int k = lane; k < total; k += 32) row[1 + k] = sb[k];
    if (row2) for (int k = lane; k < total; k += 32) row2[1 + k] = sb[k];
}

__global__ void __launch_bounds__(128) combine_kernel(const CombineArgs a)
{
    __shared__ int sbuf[4][SORT_CAP];
    __shared__ int soff[4][MAX_RANKS], scnt[4][MAX_RANKS];
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    int kl = a.kl0 + blockIdx.x * 4 + wq;              // grid-stride over the rows
    if (kl < a.kl1 && a.flags) wait_all_ranks(a.flags, a.R, (long long)a.seq, lane);    // every shard has published this call
    for (; kl < a.kl1; kl += gridDim.x * 4) {
        combine_row(a, kl, lane, sbuf[wq], soff[wq], scnt[wq]);
        __syncwarp();
    }
    publish_when_last(a.ack);
}

// ---------------------------------------------------------------------------------------------
// pot_kernel (gpupot): phi_i = sum_{j, r>0} m_j / r_ij (gpupot.gpu.cu:32-57).  Same sorted tiles
// (velocities unused).  Work item = warp: 32 i-particles x every S-th tile; separation from the
// tile-local offsets, rsqrt.approx + one Newton step (the AVX twin does the same, pot.avx.cpp:22-25),
// FP32 chains of 64 terms flushed to fp64.  A self pair gives dx = fl(O-x_i) + fl(x_i-O) = 0 exactly and
// is skipped by r2 > 0 like the reference (:52).  Partial sums are combined by pot_merge_kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pot_kernel(const float *__restrict__ tiles, int ntiles, int S,
                                                   const double *__restrict__ x, int i0, int ni, double *__restrict__ part)
{
    __shared__ __align__(16) float sb[4][HDR + 4 * TJ];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * 4 + warp;
    const int n_it = (ni + 31) / 32;
    if (w >= n_it * S) return;
    const int it = w / S, s = w - it * S;
    const int ii = it * 32 + lane;
    const int i = i0 + (ii < ni ? ii : ni - 1);
    const double xd = x[3 * (size_t)i], yd = x[3 * (size_t)i + 1], zd = x[3 * (size_t)i + 2];
    // two-float position of i: tiles closer than half their own reach use float-float separations (below)
    const float xh = (float)xd, yh = (float)yd, zh = (float)zd;
    const float xl = (float)(xd - (double)xh), yl = (float)(yd - (double)yh), zl = (float)(zd - (double)zh);
    double phi = 0.0;
    float *b = sb[warp];
    for (int t = s; t < ntiles; t += S) {
        const float *tp = tiles + (size_t)t * TILE_FLOATS;
        __syncwarp();
        if (lane < HDR) b[lane] = tp[lane];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int jj = lane + 32 * h;
            b[HDR + jj]          = tp[HDR + C_DX * TJ + jj];
            b[HDR + TJ + jj]     = tp[HDR + C_DY * TJ + jj];
            b[HDR + 2 * TJ + jj] = tp[HDR + C_DZ * TJ + jj];
            b[HDR + 3 * TJ + jj] = tp[HDR + C_M * TJ + jj];
        }
        __syncwarp();
        // origin = Oh + Ol exactly (tilepack takes the offsets from this sum): one rounding of the exact separation
        const float cx = (float)(((double)b[0] + (double)b[3]) - xd), cy = (float)(((double)b[1] + (double)b[4]) - yd),
                    cz = (float)(((double)b[2] + (double)b[5]) - zd);
        const float4 *c = reinterpret_cast<const float4 *>(b + HDR);
        float acc = 0.f;
        // dx = c + offset carries an absolute error of ~2^-23 (|c| + h): fine while the pair distance is comparable to
        // the tile's reach (same rule as regf_kernel's FAR tiles).  Tiles closer than that -- the lane's own tile and
        // its neighbours, where a close pair can carry most of phi_i -- take the separation from the two-float
        // positions, (xh_j - xh_i) + (xl_j - xl_i), exact to ~2^-48 (the reference keeps float2 positions for the
        // same reason, gpupot.gpu.cu:15-30).
        float d2 = 0.f, sreach = 0.f;
        {
            const float cf[3] = {cx, cy, cz};
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const float g = fmaxf(fabsf(cf[q]) - b[6 + q], 0.f);
                d2 = fmaf(g, g, d2);
                sreach = fmaxf(sreach, fabsf(cf[q]) + b[6 + q]);
            }
        }
        if (!__all_sync(0xffffffffu, d2 > 0.25f * sreach * sreach)) {
            const float *g = tp + HDR;
#pragma unroll 4
            for (int j = 0; j < TJ; j++) {
                const float dx = (g[C_XH * TJ + j] - xh) + (g[C_XL * TJ + j] - xl);
                const float dy = (g[C_YH * TJ + j] - yh) + (g[C_YL * TJ + j] - yl);
                const float dz = (g[C_ZH * TJ + j] - zh) + (g[C_ZL * TJ + j] - zl);
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                float y = rsqrt_approx(r2);
                y = y * fmaf(-0.5f * r2 * y, y, 1.5f);
                acc += (r2 > 0.f) ? b[HDR + 3 * TJ + j] * y : 0.f;
            }
            phi += (double)acc;
            continue;
        }
        // packed f32x2 over j: 7 packed FP32 ops + 2 MUFU.RSQ per two pairs, two chains of 32 terms.  No pair of a tile
        // that passed the test above is a self pair or a coincident pair (their tiles' boxes contain x_i: d2 = 0), so
        // r2 > 0 needs no guard here; and the raw MUFU.RSQ (relative error 2^-22.9 per term, a random walk over the
        // sum) is enough for the 1e-6 bar of a FAR term -- the Newton step of pot.avx.cpp:22-25 stays in the exact path
        // above, where one close pair can carry most of phi_i.  The body is then bound by the MUFU pipe (16 lanes per
        // clock and SM), not by the FMA pipe.
        const float2 cx2 = dup2(cx), cy2 = dup2(cy), cz2 = dup2(cz);
        float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll 4
        for (int q = 0; q < TJ / 4; q++) {
            const float4 DX = c[q], DY = c[16 + q], DZ = c[32 + q], M = c[48 + q];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float2 dx = add2(h ? make_float2(DX.z, DX.w) : make_float2(DX.x, DX.y), cx2);
                const float2 dy = add2(h ? make_float2(DY.z, DY.w) : make_float2(DY.x, DY.y), cy2);
                const float2 dz = add2(h ? make_float2(DZ.z, DZ.w) : make_float2(DZ.x, DZ.y), cz2);
                const float2 m2 = h ? make_float2(M.z, M.w) : make_float2(M.x, M.y);
                const float2 r2 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                const float2 y = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                acc2 = fma2(m2, y, acc2);
            }
        }
        phi += (double)(acc2.x + acc2.y);
    }
    if (ii < ni) part[(size_t)s * ni + ii] = phi;
}

__global__ void pot_merge_kernel(const double *__restrict__ part, int S, int ni, double *__restrict__ out)
{
    const int ii = blockIdx.x * blockDim.x + threadIdx.x;
    if (ii >= ni) return;
    double p = 0.0;
    for (int s = 0; s < S; s++) p += part[(size_t)s * ni + ii];
    out[ii] = p;
}

// gpupot over j-shards: phi_i = sum over shards, in rank order (deterministic), of the shard partials.  parts[r] may
// be peer pointers (one process driving several GPUs) or rows of an all-gathered buffer (one process per GPU).
struct PotParts { const double *p[MAX_RANKS]; };
__global__ void pot_sum_kernel(const PotParts parts, int R, int ni, double *__restrict__ out)
{
    const int ii = blockIdx.x * blockDim.x + threadIdx.x;
    if (ii >= ni) return;
    double p = 0.0;
    for (int r = 0; r < R; r++) p += parts.p[r][ii];
    out[ii] = p;
}

// ---------------------------------------------------------------------------------------------
// Device-resident predictor.  The state of the j-set -- BODY, X0, X0DOT, F (= force / 2), FDOT
// (= derivative / 6) and T0 in the integrator's own conventions -- lives on the device; predict_kernel restates
// xbpredall.f:17-26 in fp64,
//     S = TIME - T0;  X = ((FDOT*S + F)*S + X0DOT)*S + X0;  XDOT = (FDOT*(1.5 S) + F)*(2 S) + X0DOT,
// and writes the m | x | v snapshot gpunb_send_ would have uploaded.  Products and sums are rounded separately
// (__dmul_rn / __dadd_rn: no FMA contraction), i.e. the arithmetic of an unfused host build, so that the
// predicted snapshot is bit-for-bit the one a host predictor followed by gpunb_send_ produces.
// State layout: body[cap] | x0[3 cap] | x0dot[3 cap] | f[3 cap] | fdot[3 cap] | t0[cap].
// ---------------------------------------------------------------------------------------------
__global__ void predict_kernel(int n, int cap, const double *__restrict__ st, double time, double *__restrict__ jraw)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double *x0 = st + cap, *v0 = st + 4 * (size_t)cap, *f = st + 7 * (size_t)cap, *fd = st + 10 * (size_t)cap;
    const double s = __dsub_rn(time, st[13 * (size_t)cap + j]);
    const double s1 = __dmul_rn(1.5, s), s2 = __dmul_rn(2.0, s);
    jraw[j] = st[j];
    double *x = jraw + n, *v = jraw + 4 * (size_t)n;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const size_t q = 3 * (size_t)j + c;
        const double F = f[q], FD = fd[q], V0 = v0[q];
        x[q] = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(FD, s), F), s), V0), s), x0[q]);
        v[q] = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(FD, s1), F), s2), V0);
    }
}
// out[k] = x[idx[k]] | v[idx[k]] (6 doubles) from the m | x | v snapshot
// The same predictor from particle RECORDS owned by somebody else on this device -- the irregular-force library's table
// (irr_b200.cu: 16 doubles per particle, x0[3] m | v0[3] t0 | a2[3] . | j6[3] ., a2 = F/2, j6 = FDOT/6), which the
// integrator refreshes after every corrector anyway: one copy of the state serves both libraries and nothing is uploaded
// twice.  Same unfused operations as predict_kernel.
__global__ void predict_records_kernel(int n, const double *__restrict__ rec, int stride, double time, double *__restrict__ jraw)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double *r = rec + (size_t)j * stride;
    const double2 a = reinterpret_cast<const double2 *>(r)[0], b = reinterpret_cast<const double2 *>(r)[1];
    const double2 c2 = reinterpret_cast<const double2 *>(r)[2], d = reinterpret_cast<const double2 *>(r)[3];
    const double2 e = reinterpret_cast<const double2 *>(r)[4], f2 = reinterpret_cast<const double2 *>(r)[5];
    const double2 g = reinterpret_cast<const double2 *>(r)[6], h = reinterpret_cast<const double2 *>(r)[7];
    const double X0[3] = {a.x, a.y, b.x}, V0[3] = {c2.x, c2.y, d.x}, F[3] = {e.x, e.y, f2.x}, FD[3] = {g.x, g.y, h.x};
    const double s = __dsub_rn(time, d.y);
    const double s1 = __dmul_rn(1.5, s), s2 = __dmul_rn(2.0, s);
    jraw[j] = b.y;
    double *x = jraw + n, *v = jraw + 4 * (size_t)n;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const size_t q = 3 * (size_t)j + c;
        x[q] = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(FD[c], s), F[c]), s), V0[c]), s), X0[c]);
        v[q] = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(FD[c], s1), F[c]), s2), V0[c]);
    }
}

// Scattered gpunb_send_ (one process per GPU): every rank uploads 1/R of the snapshot over its own PCIe link into its chunk
// of `tmp` (per rank: m[chunk] | x[chunk][3] | v[chunk][3]), an all-gather over NVLink completes `tmp` on every GPU, and
// this kernel lays it out as the packed snapshot m[nj] | x[nj][3] | v[nj][3].
__global__ void send_unpack_kernel(int nj, int chunk, const double *__restrict__ tmp, double *__restrict__ jraw)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nj) return;
    const int r = j / chunk, o = j - r * chunk;
    const double *src = tmp + (size_t)r * 7 * chunk;
    jraw[j] = src[o];
    double *x = jraw + nj, *v = jraw + 4 * (size_t)nj;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        x[3 * (size_t)j + c] = src[chunk + 3 * (size_t)o + c];
        v[3 * (size_t)j + c] = src[4 * (size_t)chunk + 3 * (size_t)o + c];
    }
}

__global__ void snapshot_gather_kernel(int n, int nj, const int *__restrict__ idx, const double *__restrict__ jraw,
                                       double *__restrict__ out, int *__restrict__ bad)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int j = idx[k];
    if (j < 0 || j >= nj) { atomicExch(bad, 1); return; }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        out[6 * (size_t)k + c] = jraw[nj + 3 * (size_t)j + c];
        out[6 * (size_t)k + 3 + c] = jraw[4 * (size_t)nj + 3 * (size_t)j + c];
    }
}
// rec: n records of 14 doubles (body, x0[3], x0dot[3], f[3], fdot[3], t0) for the particles idx[k] (0-based)
__global__ void state_scatter_kernel(int n, int cap, int nj, const int *__restrict__ idx, const double *__restrict__ rec,
                                     double *__restrict__ st, int *__restrict__ bad)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int j = idx[k];
    if (j < 0 || j >= nj) { atomicExch(bad, 1); return; }
    const double *r = rec + 14 * (size_t)k;
    st[j] = r[0];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        st[cap + 3 * (size_t)j + c] = r[1 + c];
        st[4 * (size_t)cap + 3 * (size_t)j + c] = r[4 + c];
        st[7 * (size_t)cap + 3 * (size_t)j + c] = r[7 + c];
        st[10 * (size_t)cap + 3 * (size_t)j + c] = r[10 + c];
    }
    st[13 * (size_t)cap + j] = r[13];
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
double wtime() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + 1e-6 * tv.tv_usec; }

// Kernel variants (i-particles per lane x minimum resident CTAs per SM); GPUNB_B200_VARIANT selects one by
// name for tuning runs, the default is the fastest measured on B200 (see profiles/).
typedef void (*RegfKernel)(const RegfArgs);
struct Variant { const char *name; int it; RegfKernel k[2]; };
const Variant VARIANTS[] = {
    {"it2",    2, {regf_kernel<2, false, 1, 0>, regf_kernel<2, true, 1, 0>}},
    {"it2b3",  2, {regf_kernel<2, false, 3, 0>, regf_kernel<2, true, 3, 0>}},
    {"it1",    1, {regf_kernel<1, false, 1, 0>, regf_kernel<1, true, 1, 0>}},
    {"it1b5",  1, {regf_kernel<1, false, 5, 0>, regf_kernel<1, true, 5, 0>}},
    {"it1b4",  1, {regf_kernel<1, false, 4, 0>, regf_kernel<1, true, 4, 0>}},
    {"it1b3",  1, {regf_kernel<1, false, 3, 0>, regf_kernel<1, true, 3, 0>}},
    {"it1b4n", 1, {regf_kernel<1, false, 4, 1>, regf_kernel<1, true, 4, 1>}},      // + Newton step in the FAR body
    {"it1b4t", 1, {regf_kernel<1, false, 4, 2>, regf_kernel<1, true, 4, 2>}},      // + transposed NEAR lanes
    {"it1b4nt", 1, {regf_kernel<1, false, 4, 3>, regf_kernel<1, true, 4, 3>}},
    {"it1b4tq", 1, {regf_kernel<1, false, 4, 2 + 4>, regf_kernel<1, true, 4, 2 + 4>}},         // + bounding-sphere FAR test first
    {"it1b4tf", 1, {regf_kernel<1, false, 4, 2 + 8>, regf_kernel<1, true, 4, 2 + 8>}},         // + FP32 chains of 64 terms
    {"it1b4tu", 1, {regf_kernel<1, false, 4, 2 + 16>, regf_kernel<1, true, 4, 2 + 16>}},       // + FAR loop unrolled 4 quads
    {"it1b4tqf", 1, {regf_kernel<1, false, 4, 2 + 4 + 8>, regf_kernel<1, true, 4, 2 + 4 + 8>}},
    {"it1b4tqu", 1, {regf_kernel<1, false, 4, 2 + 4 + 16>, regf_kernel<1, true, 4, 2 + 4 + 16>}},
};
constexpr int NVARIANTS = sizeof(VARIANTS) / sizeof(VARIANTS[0]);
constexpr int DEFAULT_VARIANT = 9;     // it1b4tq (it1b4t + the bounding-sphere FAR test first: +0.7 %, identical results, profiles/r2r_variant_probe.txt); it1b4t: 4 CTAs/SM (<= 128 registers) + transposed NEAR lanes: +2 % at ni = 1024, +8 % at 256,
                                       // +18 % at 32 over it1b4 at N = 1M (profiles/r2a_variant_probe.txt); the Newton step in the
                                       // FAR body ("n") costs 10 % and leaves the strict jerk error where it is (2.9e-6 vs 3.1e-6)
constexpr int ROWS_LMAX_CAP = 1024;    // row stride capacity of the IPC-exported shard rows (NCCL mode)

// Pipeline slot: the work buffers one i-block (or sub-block) needs from the pair kernel to the merged result, and the
// two streams it runs on.  Consecutive blocks go to different slots so that the pair kernel of block b+1 (low
// priority stream `lo`) fills the SMs as the CTAs of block b retire, while merge / exchange of block b (high
// priority stream `hi`, latency-bound kernels) run beside it.
constexpr int MAX_SLOTS = 4;
constexpr int DEFAULT_NSLOT = 3;       // slots a sharded resident sweep cycles through (GPUNB_B200_NSLOT); one GPU: 2
constexpr int DEFAULT_NSUB  = 4;       // sub-blocks of one gpunb_regf_ call (GPUNB_B200_NSUB)
struct Slot {
    cudaStream_t lo = nullptr, hi = nullptr;
    cudaEvent_t ev_regf = nullptr, ev_done = nullptr, ev_start = nullptr;
    bool used = false;                 // ev_done has been recorded at least once
    int items_cap = 0;
    double *part = nullptr; int *cnt = nullptr;
    int *seg = nullptr; size_t seg_ints = 0;
    double *res_f = nullptr; int *res_list = nullptr; size_t res_list_ints = 0;   // device-resident results (sweeps)
    unsigned *done_ctr = nullptr;      // [2]: finished-CTA counters of merge / combine (in-kernel publication)
};

struct Dev {
    int id = -1;
    cudaStream_t st = nullptr;
    Slot slots[MAX_SLOTS];
    cudaEvent_t ev_fork = nullptr;
    int *iperm_all = nullptr; size_t iperm_all_n = 0;   // Morton order of every block of a resident sweep
    double *state = nullptr; int state_cap = 0, state_n = 0;   // device-resident predictor state (14 doubles per particle)
    double *upd_rec = nullptr; int *upd_idx = nullptr; int upd_cap = 0; int *upd_bad = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, evs0 = nullptr, evs1 = nullptr, evdone = nullptr, evpush = nullptr;
    int nsm = 0, warps_resident = 0, variant = DEFAULT_VARIANT, itile = 32, oversub = OVERSUB;
    // j: full fp64 snapshot (m | x | v, packed for the current nj_total) and the tiles of this device's shard
    int raw_cap = 0, tile_cap = 0;
    double *jraw = nullptr;
    float *jtile = nullptr;
    double *radii = nullptr;      // h2[raw_cap] | dtr[raw_cap]  (resident sweeps)
    int nj_total = 0, nj = 0, ntiles = 0;
    int *jidx = nullptr;          // sorted slot -> global j
    // Morton sort scratch
    unsigned *hbits = nullptr;    // largest |coordinate| of the shard (float bits)
    unsigned long long *keys_in = nullptr, *keys_out = nullptr;
    int *vals_in = nullptr, *perm = nullptr; int sort_cap = 0;
    int perm_n = 0;               // perm holds the order of the regf snapshot of this many particles (0: invalid)
    unsigned long long *qsum = nullptr;   // [2] tile-extent sum / tiles done of the running tilepack (adaptive re-sort)
    void *cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
    int *iperm = nullptr;         // Morton order of the current i-block
    int *iperm_identity = nullptr;   // 0, 1, 2, ... (blocks of a single i-tile are not sorted)
    unsigned long long *stats = nullptr;   // [0] near (warp,tile) visits, [1] all visits (GPUNB_B200_STATS=1)
    unsigned long long *wtime = nullptr;   // per work item start/end timestamps (GPUNB_B200_STATS=2)
    double *ibuf = nullptr;       // 8*JOBCAP doubles: h2 | dtr | x | v (+ 64: slice offsets of a collective call)
    int segcap = 0;
    double *fr = nullptr;         // [NIMAX][8] shard partial + count (multi-GPU)
    int *rows = nullptr; size_t rows_ints = 0;                                     // shard rows (in-process multi-GPU)
    double *send_tmp = nullptr; size_t send_tmp_n = 0;                             // scattered gpunb_send_: [R][7 chunk] (NCCL mode)
    int *last_rows = nullptr; size_t last_rows_ints = 0;                           // root: device copy of the rows of the last gpunb_regf_
    int *nanflag = nullptr;       // device alias of this device's entry of L.h_nan (mapped pinned host memory)
    double *pot_part = nullptr, *pot_out = nullptr; size_t pot_part_n = 0, pot_out_n = 0;
    // gpupot's own snapshot (m | x) and tiles of this device's shard, so that gpupot never disturbs the regf j-set
    double *pot_jraw = nullptr; float *pot_jtile = nullptr; int *pot_jidx = nullptr; int pot_cap = 0;
    double *pot_gather = nullptr; size_t pot_gather_n = 0;      // [R][ni] partial potentials of all ranks (NCCL mode)
};

// NCCL types / entry points, resolved with dlopen so that the library has no link-time NCCL dependency
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int (*pfn_ncclGetUniqueId)(ncclUniqueId *);
typedef int (*pfn_ncclCommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
typedef int (*pfn_ncclAllGather)(const void *, void *, size_t, int /*datatype*/, ncclComm_t, cudaStream_t);
typedef int (*pfn_ncclCommDestroy)(ncclComm_t);
typedef const char *(*pfn_ncclGetErrorString)(int);
constexpr int NCCL_INT8 = 0, NCCL_FLOAT64 = 8;

struct Shard {                     // one process per GPU, j sharded over ranks
    bool on = false;
    int rank = 0, R = 1;
    void *dl = nullptr;
    pfn_ncclGetUniqueId getid = nullptr; pfn_ncclCommInitRank init = nullptr; pfn_ncclAllGather allgather = nullptr;
    pfn_ncclCommDestroy destroy = nullptr; pfn_ncclGetErrorString errstr = nullptr;
    ncclComm_t comm = nullptr;
    // Exchange buffer of this rank (one allocation, exported with cudaIpc, mapped by every peer):
    //   [XSLOTS] x { fr[JOBCAP][8] doubles | rows[JOBCAP][ROWS_LMAX_CAP] ints }  |  flags[XSLOTS][MAX_RANKS] u64
    //   | acks[XSLOTS][MAX_RANKS] u64
    unsigned char *xbuf = nullptr;
    unsigned char *xbuf_peer[MAX_RANKS] = {nullptr};
    unsigned long long seq = 0;    // exchange steps so far (every rank makes the same calls)
    double *scratch = nullptr;     // bootstrap / finalize all-gathers
    // i-slice mode (gpunb_b200_set_islice): gpunb_regf_ is a COLLECTIVE call in which every rank passes its OWN i-slice
    // (what NBODY6++'s MPI build does: each rank integrates its share of the block, intgrt.F:982-1231) and receives the
    // results of that slice only.  The slices (<= 64 B x 2048 per rank) meet in a POSIX shared-memory segment of the node.
    bool islice = false;
    struct ShmSeg *shm = nullptr;
    unsigned long long icall = 0;  // collective calls so far
};
struct ShmSlice { int ni, lmax, nnbmax, m_flag; int pad[12]; double data[8 * NIMAX]; };     // h2 | dtr | x | v of one rank
struct ShmSeg {
    unsigned long long seq[2][MAX_RANKS];      // call number whose slice sits in slice[parity][rank]
    ShmSlice slice[2][MAX_RANKS];
};

struct Lib {
    bool devinit = false, is_open = false;
    std::vector<Dev> devs;
    Shard sh;
    int nbmax = 0, nbody = 0;
    double *h_j = nullptr; size_t h_j_n = 0;          // pinned staging: 7*nj doubles
    double *h_upd = nullptr; int *h_upd_idx = nullptr; int h_upd_cap = 0;    // pinned staging of state updates
    double *h_i = nullptr;                            // 8*JOBCAP + 64 doubles (pinned staging of the i-block)
    // results of gpunb_regf_: MAPPED pinned host memory that merge / combine write straight over PCIe (zero copy),
    // and the device-side aliases of those buffers
    double *h_f = nullptr, *h_f_dev = nullptr;        // [NIMAX][7]
    int *h_list = nullptr, *h_list_dev = nullptr; size_t h_list_n = 0;
    int *h_iperm = nullptr, *h_iperm_dev = nullptr;   // [NIMAX] sorted slot -> i of the current block
    int *h_flag = nullptr;
    int *h_nan = nullptr;          // [MAX_RANKS] NaN flags written by the tile kernels straight into host memory
    int nslot = DEFAULT_NSLOT, nsub = DEFAULT_NSUB;
    int host_threads = 0;          // host-side OpenMP team size; 0 until lib_devinit: the caller's default team (internal.h), or GPUNB_B200_HOST_THREADS
    int regf_oversub = REGF_OVERSUB;
    int send_scatter_min = 40000;  // one process per GPU: snapshots of at least this many particles are uploaded in R slices and
                                   // all-gathered over NVLink (GPUNB_B200_SEND_SCATTER_MIN; < 0: never)
    int resort_every = 0;          // Hilbert order refreshed every k-th snapshot (GPUNB_B200_RESORT_EVERY); 1 = always;
                                   // 0 (default) = adaptive: kept while the tiles stay compact (one GPU; sharded runs always sort)
    unsigned long long *h_q = nullptr;                 // mapped: tile-extent sum of the last tilepack (device 0)
    double q_ref = 0.0, q_last = 0.0;                  // ... right after the last sort / of the last snapshot
    double sub_pairs = 1.5e8;      // pairs a sub-block of gpunb_regf_ must keep (GPUNB_B200_SUB_PAIRS)
    double isort_pairs = 2.5e7;    // gpunb_regf_ calls with fewer pairs skip the Morton sort of the i-block (GPUNB_B200_ISORT_PAIRS)
    int snapshots_since_sort = 0;
    bool taper = false;            // tapering sub-block sizes (GPUNB_B200_TAPER=1); measured: no gain at 4 sub-blocks
    bool enqueue_threads = true;   // one process driving G > 1 GPUs: one host thread per device enqueues that device's copies and
                                   // kernels (GPUNB_B200_ENQUEUE_THREADS=0: one thread for all, the A/B arm)
    bool nslot_auto = true;        // nslot not chosen by the caller (environment / gpunb_b200_set_tuning)
    int near_exact = -1;           // >= 0: overrides GPUNB_B200_NEAR_EXACT (gpunb_b200_set_near_exact)
    bool nsub_forced = false;      // tests: split even when the pair kernels would be too short to be worth it
    int last_slot = 0; bool last_on_host = false;
    double time_send = 0, time_grav = 0, time_reduce = 0, time_out = 0;      // reference: gpunb.velocity.cu:557-559
    long long numInter = 0; int icall = 0, ini = 0, isend = 0;
    double ctr[GPUNB_B200_CTR_COUNT] = {0};
    int last_ni = 0, last_lmax = 0;
    int last_rows_ni = 0, last_rows_lmax = 0;      // rows of the last gpunb_regf_ held in devs[0].last_rows (0: none)
} L;

template <class T> void dev_alloc(T *&p, size_t n) { CUDA_CHECK(cudaMalloc((void **)&p, n * sizeof(T))); }
template <class T> void dev_free(T *&p) { if (p) CUDA_CHECK(cudaFree(p)); p = nullptr; }
template <class T> void host_alloc(T *&p, size_t n) { CUDA_CHECK(cudaMallocHost((void **)&p, n * sizeof(T))); }
template <class T> void host_free(T *&p) { if (p) CUDA_CHECK(cudaFreeHost(p)); p = nullptr; }

void set_dev(const Dev &d) { CUDA_CHECK(cudaSetDevice(d.id)); }
// counters touched by the per-device enqueue threads of the in-process multi-GPU mode
inline void ctr_add(int k, double v)
{
#pragma omp atomic
    L.ctr[k] += v;
}
// One host thread per device for the per-device enqueue loops (the reference drives its GPUs from one OpenMP thread each as
// well, gpunb.velocity.cu:607-613,756): with one thread the launches of device g start ~15 us x g after those of device 0.
// The team has the size of the library's other host-side teams (staging copies, row delivery): libgomp re-docks its threads when
// consecutive parallel regions ask for different sizes, which cost more than the parallel enqueue saved (2 GPUs, pageable
// arrays: 692 -> 797 us per call with teams of 2 and 4 alternating).
int host_team() { return L.host_threads > 0 ? L.host_threads : std::max(1, std::min(omp_get_max_threads(), 64)); }
inline int enqueue_team()
{
    const int G = (int)L.devs.size();
    if (G <= 1 || !L.enqueue_threads) return 1;
    if (host_team() < G) L.host_threads = G;
    return host_team();
}
int  total_ranks() { return L.sh.on ? L.sh.R : (int)L.devs.size(); }

void lib_devinit(int irank)
{
    if (L.devinit) return;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        FATAL("no CUDA device available (%s). This library has no CPU fallback.", cudaGetErrorString(e));
    std::vector<int> ids;
    const char *gl = getenv("GPU_LIST");               // reference: gpunb.velocity.cu:582-591
    if (gl && *gl) {
        char *tmp = strdup(gl);
        for (char *p = strtok(tmp, " ,"); p; p = strtok(nullptr, " ,")) ids.push_back(atoi(p));
        free(tmp);
    } else {
        for (int i = 0; i < ndev; i++) ids.push_back(i);
    }
    // One process driving several GPUs (the reference's gpunb.velocity.cu model, j split across them) is
    // enabled with GPUNB_B200_MULTI=1; default is the first listed device (one process per GPU, as under
    // MPI or torchrun, where ranks are joined with gpunb_b200_nccl_init).
    const char *multi = getenv("GPUNB_B200_MULTI");
    if (!(multi && atoi(multi) > 0) && ids.size() > 1) ids.resize(1);
    if ((int)ids.size() > MAX_RANKS) ids.resize(MAX_RANKS);
    char host[150] = {0};
    gethostname(host, 149);
    for (size_t k = 0; k < ids.size(); k++) {
        if (ids[k] < 0 || ids[k] >= ndev) FATAL("GPU_LIST names device %d but only %d are visible", ids[k], ndev);
        Dev d; d.id = ids[k];
        cudaDeviceProp prop; CUDA_CHECK(cudaGetDeviceProperties(&prop, d.id));
        if (prop.major < 10) FATAL("device %d (%s) is sm_%d%d; this library is built for sm_100a only", d.id, prop.name, prop.major, prop.minor);
        d.nsm = prop.multiProcessorCount;
        CUDA_CHECK(cudaSetDevice(d.id));
        CUDA_CHECK(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
        cudaEvent_t *evs[] = {&d.ev0, &d.ev1, &d.ev2, &d.ev3, &d.evs0, &d.evs1};
        for (cudaEvent_t *ev : evs) CUDA_CHECK(cudaEventCreate(ev));
        CUDA_CHECK(cudaEventCreateWithFlags(&d.evdone, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&d.evpush, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming));
        int prio_least = 0, prio_greatest = 0;
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        // hi streams: highest priority; lo streams: lowest, all equal (descending priorities by slot were measured:
        // they do not fix the start order of sub-blocks and cost 2 % in 3-slot sweeps, profiles/r01p_pipeline_probe_1gpu.txt)
        for (int q = 0; q < MAX_SLOTS; q++) {
            Slot &sl = d.slots[q];
            CUDA_CHECK(cudaStreamCreateWithPriority(&sl.lo, cudaStreamNonBlocking, prio_least));
            CUDA_CHECK(cudaStreamCreateWithPriority(&sl.hi, cudaStreamNonBlocking, prio_greatest));
            CUDA_CHECK(cudaEventCreateWithFlags(&sl.ev_regf, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&sl.ev_start, cudaEventDisableTiming));
        }
        const int smem = WARPS * NSTAGE * TILE_BYTES + WARPS * NSTAGE * 8;
        const char *vn = getenv("GPUNB_B200_VARIANT");
        if (vn && *vn) {
            int f = -1;
            for (int q = 0; q < NVARIANTS; q++) if (!strcmp(vn, VARIANTS[q].name)) f = q;
            if (f < 0) FATAL("GPUNB_B200_VARIANT=%s is not a kernel variant", vn);
            d.variant = f;
        }
        const Variant &V = VARIANTS[d.variant];
        d.itile = 32 * V.it;
        int nb = 1 << 30;
        for (int q = 0; q < 2; q++) {
            int nbq = 0;
            CUDA_CHECK(cudaFuncSetAttribute(V.k[q], cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbq, V.k[q], WARPS * 32, smem));
            if (nbq < nb) nb = nbq;
        }
        if (nb < 1) FATAL("regf_kernel does not fit on an SM");
        d.warps_resident = d.nsm * nb * WARPS;
        { const char *e = getenv("GPUNB_B200_OVERSUB"); if (e && atoi(e) >= 1 && atoi(e) <= 32) d.oversub = atoi(e); }
        { const char *e = getenv("GPUNB_B200_REGF_OVERSUB"); if (e && atoi(e) >= 1 && atoi(e) <= 32) L.regf_oversub = atoi(e); }
        { const char *e = getenv("GPUNB_B200_SEND_SCATTER_MIN"); if (e) L.send_scatter_min = atoi(e); }
        fprintf(stderr, "# GPU initialization - rank: %d; HOST %s; NGPU %d; device: %d %s; B200-native regf[%s]: %d SMs x %d CTAs x %d warps\n",
                irank, host, (int)ids.size(), d.id, prop.name, V.name, d.nsm, nb, WARPS);
        L.devs.push_back(d);
    }
    // in-process multi-GPU: the root device pulls shard results over NVLink P2P, and every device pushes its slice of a
    // scattered gpunb_send_ to every other one
    for (size_t a = 0; a < L.devs.size(); a++)
        for (size_t g = 0; g < L.devs.size(); g++) {
            if (a == g) continue;
            int can = 0;
            CUDA_CHECK(cudaDeviceCanAccessPeer(&can, L.devs[a].id, L.devs[g].id));
            if (!can) FATAL("device %d cannot access device %d (P2P needed for the j-shard combine)", L.devs[a].id, L.devs[g].id);
            CUDA_CHECK(cudaSetDevice(L.devs[a].id));
            cudaError_t pe = cudaDeviceEnablePeerAccess(L.devs[g].id, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(pe);
            (void)cudaGetLastError();
        }
    CUDA_CHECK(cudaSetDevice(L.devs[0].id));
    host_alloc(L.h_flag, 16);
    memset(L.h_flag, 0, 16 * sizeof(int));
    CUDA_CHECK(cudaHostAlloc((void **)&L.h_nan, MAX_RANKS * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(L.h_nan, 0, MAX_RANKS * sizeof(int));
    for (size_t g = 0; g < L.devs.size(); g++) {
        CUDA_CHECK(cudaSetDevice(L.devs[g].id));
        CUDA_CHECK(cudaHostGetDevicePointer((void **)&L.devs[g].nanflag, (void *)(L.h_nan + g), 0));
    }
    CUDA_CHECK(cudaSetDevice(L.devs[0].id));
    { const char *e = getenv("GPUNB_B200_NSLOT"); if (e && atoi(e) >= 1 && atoi(e) <= MAX_SLOTS) { L.nslot = atoi(e); L.nslot_auto = false; } }
    { const char *e = getenv("GPUNB_B200_ENQUEUE_THREADS"); if (e) L.enqueue_threads = atoi(e) != 0; }
    { const char *e = getenv("GPUNB_B200_NSUB");  if (e && atoi(e) >= 1 && atoi(e) <= MAX_SLOTS) L.nsub = atoi(e); }
    { const char *e = getenv("GPUNB_B200_TAPER"); if (e) L.taper = atoi(e) != 0; }
    { const char *e = getenv("GPUNB_B200_SUB_PAIRS"); if (e && atof(e) >= 1.0) L.sub_pairs = atof(e); }
    { const char *e = getenv("GPUNB_B200_ISORT_PAIRS"); if (e && atof(e) >= 0.0) L.isort_pairs = atof(e); }
    { const char *e = getenv("GPUNB_B200_RESORT_EVERY"); if (e && atoi(e) >= 0) L.resort_every = atoi(e); }
    CUDA_CHECK(cudaHostAlloc((void **)&L.h_q, sizeof(unsigned long long), cudaHostAllocMapped | cudaHostAllocPortable));
    *L.h_q = 0ull;
    L.host_threads = std::max(1, std::min(omp_get_max_threads(), 64));
    { const char *e = getenv("GPUNB_B200_HOST_THREADS"); if (e && atoi(e) >= 1 && atoi(e) <= 64) L.host_threads = atoi(e); }
    L.devinit = true;
}

// j-shards.  The reference cuts the j ARRAY into contiguous index ranges (joff[id] = id*nbody/numGPU,
// gpunb.velocity.cu:713-715).  Here the WHOLE j-set is Hilbert-sorted and cut into tiles of TJ, and shard r owns
// the tiles t = r, r+R, r+2R, ... of the T = ceil(nj/TJ) tiles.  Every tile is as compact as on one GPU (an
// index-range shard is a random 1/R subsample of space: measured 15 % NEAR tiles at R = 8 against 5 % for the
// whole set), and every shard samples every region of the cluster, so the NEAR work of any i-block is spread evenly
// over the shards (contiguous curve ranges left one shard 6 % slower than the others at R = 4).  The result is the
// same function of the input; neighbour rows of different shards interleave in j and are merged by sorting.
void shard_tiles(int r, int R, int nj, int &nloc)
{
    const int T = (nj + TJ - 1) / TJ;
    nloc = r < T ? (T - r + R - 1) / R : 0;
}

void ensure_j_capacity(Dev &d, int nj_total, int shard_n)
{
    set_dev(d);
    if (nj_total > d.raw_cap) {
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        dev_free(d.jraw); dev_free(d.radii);
        d.raw_cap = nj_total + 64;
        dev_alloc(d.jraw, (size_t)7 * d.raw_cap);
        dev_alloc(d.radii, (size_t)2 * d.raw_cap);
    }
    const int tiles = (shard_n + TJ - 1) / TJ + 1;
    if (tiles > d.tile_cap) {
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        dev_free(d.jtile); dev_free(d.jidx);
        d.tile_cap = tiles;
        dev_alloc(d.jtile, (size_t)tiles * TILE_FLOATS);
        dev_alloc(d.jidx, (size_t)tiles * TJ);
    }
}

int hilbert_bits(int n)
{
    int lg = 0;
    while ((1ll << lg) < n) lg++;
    int b = (lg + 2) / 3 + 5;
    return b < 8 ? 8 : (b > 21 ? 21 : b);
}

void ensure_sort_capacity(Dev &d, int n)
{
    set_dev(d);
    if (!d.hbits) dev_alloc(d.hbits, 1);
    if (n <= d.sort_cap) return;
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    dev_free(d.keys_in); dev_free(d.keys_out); dev_free(d.vals_in); dev_free(d.perm);
    d.perm_n = 0;
    if (d.cub_tmp) CUDA_CHECK(cudaFree(d.cub_tmp)); d.cub_tmp = nullptr;
    d.sort_cap = n + 1024;
    dev_alloc(d.keys_in, d.sort_cap); dev_alloc(d.keys_out, d.sort_cap);
    dev_alloc(d.vals_in, d.sort_cap); dev_alloc(d.perm, d.sort_cap);
    size_t bytes = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d.keys_in, d.keys_out, d.vals_in, d.perm, d.sort_cap, 0, 63, d.st));
    d.cub_tmp_bytes = bytes;
    CUDA_CHECK(cudaMalloc(&d.cub_tmp, bytes));
}

// fp64 particles (m, x, v or NULL; n of them, device pointers) -> Hilbert-sorted tiles + slot->index map.
// ALL n particles are sorted; tiles t0, t0 + tstride, ... (nloc of them) of the sorted order are packed (the tiles
// of one j-shard, see shard_tiles).  The radix sort of the 63-bit keys is CUB (plumbing, not the hot path).
// reuse_order: keep the permutation of the previous snapshot of the same n (GPUNB_B200_RESORT_EVERY > 1); the tiles are
// re-packed from the current positions, so boxes, offsets and results stay exact -- only the compactness of the tiles
// ages with the particles' motion.
void build_tiles(Dev &d, int n, const double *m, const double *x, const double *v, float *tiles, int *jidx,
                 int t0, int tstride, int nloc, bool reuse_order = false, unsigned long long *qsum = nullptr,
                 unsigned long long *q_host = nullptr)
{
    if (n <= 0) return;
    ensure_sort_capacity(d, n);
    if (reuse_order && d.perm_n == n) {
        if (nloc > 0) tilepack_kernel<<<(nloc + 3) / 4, 128, 0, d.st>>>(n, t0, tstride, nloc, m, x, v, d.perm, tiles, jidx, d.nanflag, d.hbits, qsum, q_host);
        CUDA_CHECK(cudaGetLastError());
        ctr_add(GPUNB_B200_CTR_LAUNCHES, 1);
        return;
    }
    CUDA_CHECK(cudaMemsetAsync(d.hbits, 0, sizeof(unsigned), d.st));
    absmax_kernel<<<(n + 255) / 256, 256, 0, d.st>>>(n, m, x, v, d.hbits, d.nanflag);
    mortonkey_kernel<<<(n + 255) / 256, 256, 0, d.st>>>(n, x, d.hbits, d.keys_in, d.vals_in);
    d.perm_n = n;
    size_t bytes = d.cub_tmp_bytes;
    // Only the leading 3*b bits of the 63-bit keys are sorted (b = bits per axis for ~32 cells per particle along the
    // curve); ties keep their index order (stable).  5 radix passes at N = 10^6 and 4 at N = 10^4 instead of 8: at
    // small N the sort is nothing but launch latency.  Mirror: sharding.hilbert_order().
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d.cub_tmp, bytes, d.keys_in, d.keys_out, d.vals_in, d.perm, n,
                                               63 - 3 * hilbert_bits(n), 63, d.st));
    if (nloc > 0) tilepack_kernel<<<(nloc + 3) / 4, 128, 0, d.st>>>(n, t0, tstride, nloc, m, x, v, d.perm, tiles, jidx, d.nanflag, d.hbits, qsum, q_host);
    CUDA_CHECK(cudaGetLastError());
    ctr_add(GPUNB_B200_CTR_LAUNCHES, 4);       // + the CUB sort passes (library code, not counted)
}

// Host ranges the caller has pinned with gpunb_b200_pin_host_ (cudaHostRegister, mapped): gpunb_send_ uploads from them
// without the staging copy, and gpunb_regf_ lets the kernels write results straight into them.
struct PinnedRange { const char *base; size_t bytes; };
std::vector<PinnedRange> g_pinned;
template <class T> T *pinned_alias(const T *p, size_t n)
{   // device-side alias of p[0..n) when the whole range lies inside one pinned range, else NULL
    const char *b = reinterpret_cast<const char *>(p);
    for (const PinnedRange &r : g_pinned)
        if (b >= r.base && b + n * sizeof(T) <= r.base + r.bytes) {
            void *d = nullptr;
            if (cudaHostGetDevicePointer(&d, const_cast<char *>(b), 0) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
            return reinterpret_cast<T *>(d);
        }
    return nullptr;
}

template <class T> void host_alloc_mapped(T *&h, T *&dptr, size_t n)
{
    CUDA_CHECK(cudaHostAlloc((void **)&h, n * sizeof(T), cudaHostAllocMapped | cudaHostAllocPortable));
    CUDA_CHECK(cudaHostGetDevicePointer((void **)&dptr, (void *)h, 0));
}

// Work buffers of the first `nslots` pipeline slots of device d, and (root) the mapped host result buffers.
void ensure_work_buffers(Dev &d, int lmax, int nnbmax, bool is_root, int nslots, bool device_results, int job_rows = NIMAX)
{
    set_dev(d);
    // n_itiles * S never exceeds warps_resident (+ n_itiles when S = 1)
    const int items = d.warps_resident * std::max(d.oversub, L.regf_oversub) + JOBCAP / d.itile;
    const int segcap = ((nnbmax > 0 ? nnbmax : 1) + 3) & ~3;
    const size_t rl = (size_t)NIMAX * lmax;                  // host-side results: one rank's call
    const size_t rl_job = (size_t)(job_rows > NIMAX ? job_rows : NIMAX) * lmax;      // device-side results of a sweep block
    for (int q = 0; q < nslots; q++) {
        Slot &sl = d.slots[q];
        if (items > sl.items_cap) {
            CUDA_CHECK(cudaDeviceSynchronize());
            dev_free(sl.part); dev_free(sl.cnt);
            sl.items_cap = items;
            dev_alloc(sl.part, (size_t)items * d.itile * PART_STRIDE);
            dev_alloc(sl.cnt, (size_t)items * d.itile);
        }
        const size_t seg_need = (size_t)sl.items_cap * d.itile * segcap;
        if (seg_need > sl.seg_ints) {
            CUDA_CHECK(cudaDeviceSynchronize());
            dev_free(sl.seg);
            sl.seg_ints = seg_need;
            dev_alloc(sl.seg, seg_need);
        }
        if (!sl.done_ctr) { dev_alloc(sl.done_ctr, 2); CUDA_CHECK(cudaMemset(sl.done_ctr, 0, 2 * sizeof(unsigned))); }
        if (is_root && device_results) {
            if (!sl.res_f) dev_alloc(sl.res_f, (size_t)7 * JOBCAP);
            if (rl_job > sl.res_list_ints) {
                CUDA_CHECK(cudaDeviceSynchronize());
                dev_free(sl.res_list);
                sl.res_list_ints = rl_job;
                dev_alloc(sl.res_list, rl_job);
            }
        }
    }
    d.segcap = segcap;
    if (L.devs.size() > 1 && rl > d.rows_ints) {
        CUDA_CHECK(cudaDeviceSynchronize());
        dev_free(d.rows);
        d.rows_ints = rl;
        dev_alloc(d.rows, rl);
    }
    if (is_root && !device_results && rl > d.last_rows_ints) {
        CUDA_CHECK(cudaDeviceSynchronize());
        dev_free(d.last_rows);
        d.last_rows_ints = rl;
        dev_alloc(d.last_rows, rl);
    }
    if (is_root && rl > L.h_list_n) {
        CUDA_CHECK(cudaDeviceSynchronize());
        host_free(L.h_list);
        L.h_list_n = rl;
        host_alloc_mapped(L.h_list, L.h_list_dev, rl);
    }
    if (L.sh.on && lmax > ROWS_LMAX_CAP) FATAL("lmax=%d exceeds the shard-row capacity %d of the NCCL mode", lmax, ROWS_LMAX_CAP);
}

void lib_open(int nbmax, int irank)
{
    L.time_send = L.time_grav = L.time_reduce = L.time_out = 0.0;      // reference: :629-632
    L.numInter = 0; L.icall = L.ini = L.isend = 0;
    lib_devinit(irank);
    if (L.is_open) { fprintf(stderr, "gpunb: it is already open\n"); return; }   // reference: :636-639
    L.is_open = true;
    L.nbmax = nbmax;
    const int R = total_ranks();
    for (Dev &d : L.devs) {
        set_dev(d);
        ensure_j_capacity(d, nbmax, nbmax / R + TJ);
        if (!d.ibuf)    dev_alloc(d.ibuf, (size_t)8 * JOBCAP + 64);          // + the slice offsets of a collective call
        if (!d.fr)      dev_alloc(d.fr, (size_t)8 * NIMAX);
        if (!d.iperm)   dev_alloc(d.iperm, (size_t)JOBCAP);
        if (!d.iperm_identity) {
            dev_alloc(d.iperm_identity, (size_t)JOBCAP);
            iota_kernel<<<(JOBCAP + 255) / 256, 256, 0, d.st>>>(JOBCAP, JOBCAP, d.iperm_identity, nullptr);
            CUDA_CHECK(cudaGetLastError());
        }
        if (!d.stats && getenv("GPUNB_B200_STATS")) { dev_alloc(d.stats, 4); CUDA_CHECK(cudaMemsetAsync(d.stats, 0, 32, d.st)); }
        if (!d.wtime && getenv("GPUNB_B200_STATS") && atoi(getenv("GPUNB_B200_STATS")) >= 2) dev_alloc(d.wtime, (size_t)3 * 65536);
    }
    const size_t hj = (size_t)7 * ((size_t)nbmax + 64);
    if (hj > L.h_j_n) { host_free(L.h_j); L.h_j_n = hj; host_alloc(L.h_j, hj); }
    if (!L.h_i) host_alloc(L.h_i, (size_t)8 * JOBCAP + 64);
    if (!L.h_f) host_alloc_mapped(L.h_f, L.h_f_dev, (size_t)7 * NIMAX);
    if (!L.h_iperm) host_alloc_mapped(L.h_iperm, L.h_iperm_dev, (size_t)NIMAX);
    fprintf(stderr, "# Open GPU regular force - rank: %d; nbmax: %d\n", irank, nbmax);
}

void lib_close()
{
    if (!L.is_open) { fprintf(stderr, "gpunb: it is already close\n"); return; }   // reference: :669-672
    L.is_open = false;
    if (!L.devs.empty()) { set_dev(L.devs[0]); CUDA_CHECK(cudaDeviceSynchronize()); gpunb_b200_internal_regcor_close(); }
    for (Dev &d : L.devs) {
        set_dev(d);
        CUDA_CHECK(cudaDeviceSynchronize());
        dev_free(d.jraw); dev_free(d.jtile); dev_free(d.radii); d.raw_cap = d.tile_cap = 0; d.nj = d.ntiles = d.nj_total = 0;
        dev_free(d.ibuf);
        for (Slot &sl : d.slots) {
            dev_free(sl.part); dev_free(sl.cnt); sl.items_cap = 0;
            dev_free(sl.seg); sl.seg_ints = 0; dev_free(sl.res_f); dev_free(sl.res_list); sl.res_list_ints = 0;
            dev_free(sl.done_ctr); sl.used = false;
        }
        dev_free(d.iperm_all); d.iperm_all_n = 0;
        dev_free(d.state); d.state_cap = d.state_n = 0; dev_free(d.upd_rec); dev_free(d.upd_idx); dev_free(d.upd_bad); d.upd_cap = 0;
        dev_free(d.fr); dev_free(d.rows); d.rows_ints = 0; dev_free(d.last_rows); d.last_rows_ints = 0;
        dev_free(d.send_tmp); d.send_tmp_n = 0; dev_free(d.iperm); dev_free(d.iperm_identity); dev_free(d.stats); dev_free(d.wtime);
        dev_free(d.jidx);
        dev_free(d.qsum); d.perm_n = 0;
    }
    host_free(L.h_j); L.h_j_n = 0; host_free(L.h_i); host_free(L.h_f); host_free(L.h_list); L.h_list_n = 0;
    host_free(L.h_iperm); host_free(L.h_upd); host_free(L.h_upd_idx); L.h_upd_cap = 0;
    L.nbmax = 0; L.q_ref = L.q_last = 0.0; L.snapshots_since_sort = 0;
}

// The snapshot is staged in pinned memory (m | x | v) in chunks: a few host threads copy chunk c+1 while the copy
// engine uploads chunk c to every local device, then each device converts and tiles its own j-shard.  The reference
// converts fp64->fp32 on ONE host thread per GPU (gpunb.velocity.cu:721-724) and uses a blocking copy (:726).
void finish_send(int nj, double wt0, const char *who);
void set_shards(int nj)
{
    const int R = total_ranks();
    for (size_t g = 0; g < L.devs.size(); g++) {
        Dev &d = L.devs[g];
        set_dev(d);
        int nloc;
        shard_tiles(L.sh.on ? L.sh.rank : (int)g, R, nj, nloc);
        ensure_j_capacity(d, nj, nloc * TJ);
        d.nj_total = nj; d.ntiles = nloc; d.nj = nloc * TJ;
    }
}
void threaded_copy(double *dst, const double *src, size_t n)
{
    const int T = host_team();
    if (T <= 1 || n < 65536) { memcpy(dst, src, sizeof(double) * n); return; }
#pragma omp parallel for num_threads(T) schedule(static)
    for (int t = 0; t < T; t++) {
        const size_t lo = n * t / T, hi = n * (t + 1) / T;
        memcpy(dst + lo, src + lo, sizeof(double) * (hi - lo));
    }
}

void lib_send(int nj, const double *mj, const double *xj, const double *vj)
{
    if (!L.is_open) FATAL("gpunb_send called while the library is closed");
    if (nj > L.nbmax) FATAL("gpunb_send: nj=%d exceeds nbmax=%d given to gpunb_open", nj, L.nbmax);
    const double wt0 = wtime();
    L.time_send -= wt0;
    L.isend++;
    L.nbody = nj;
    double *h = L.h_j;
    set_shards(nj);
    constexpr int CHUNK = 1 << 17;                    // particles per chunk (7 MB)
    const bool direct = pinned_alias(mj, (size_t)nj) && pinned_alias(xj, (size_t)3 * nj) && pinned_alias(vj, (size_t)3 * nj);
    if (L.sh.on && L.send_scatter_min >= 0 && nj >= L.send_scatter_min && L.sh.R > 1) {
        // "send scatters instead of broadcasts".  Every rank holds the whole host snapshot (replicated-data callers)
        // but uploads only its 1/R slice over its own PCIe link; the slices meet on every GPU by one all-gather over NVLink
        // (the reference splits the j array the same way, gpunb.velocity.cu:713-715 -- there the slices stay apart, here
        // every GPU needs every position for the Hilbert order that defines the tile ownership).
        Dev &d = L.devs[0];
        set_dev(d);
        const int R = L.sh.R, r = L.sh.rank;
        const int chunk = (nj + R - 1) / R;
        const size_t lo = (size_t)std::min(nj, r * chunk), hi = (size_t)std::min(nj, (r + 1) * chunk), nm = hi - lo;
        if ((size_t)R * 7 * chunk > d.send_tmp_n) {
            CUDA_CHECK(cudaStreamSynchronize(d.st));
            dev_free(d.send_tmp);
            d.send_tmp_n = (size_t)R * 7 * (chunk + 1024);
            dev_alloc(d.send_tmp, d.send_tmp_n);
        }
        double *mine = d.send_tmp + (size_t)r * 7 * chunk;
        if (nm > 0) {
            const double *sm = mj + lo, *sx = xj + 3 * lo, *sv = vj + 3 * lo;
            if (!direct) {             // pageable caller arrays: only the slice goes through the pinned staging buffer
                threaded_copy(h, sm, nm); threaded_copy(h + nm, sx, 3 * nm); threaded_copy(h + 4 * nm, sv, 3 * nm);
                sm = h; sx = h + nm; sv = h + 4 * nm;
            }
            CUDA_CHECK(cudaMemcpyAsync(mine, sm, sizeof(double) * nm, cudaMemcpyHostToDevice, d.st));
            CUDA_CHECK(cudaMemcpyAsync(mine + chunk, sx, sizeof(double) * 3 * nm, cudaMemcpyHostToDevice, d.st));
            CUDA_CHECK(cudaMemcpyAsync(mine + 4 * (size_t)chunk, sv, sizeof(double) * 3 * nm, cudaMemcpyHostToDevice, d.st));
        }
        const int rc = L.sh.allgather(mine, d.send_tmp, (size_t)7 * chunk, NCCL_FLOAT64, L.sh.comm, d.st);
        if (rc != 0) FATAL("gpunb_send: ncclAllGather failed: %s", L.sh.errstr ? L.sh.errstr(rc) : "?");
        send_unpack_kernel<<<(nj + 255) / 256, 256, 0, d.st>>>(nj, chunk, d.send_tmp, d.jraw);
        CUDA_CHECK(cudaGetLastError());
        L.ctr[GPUNB_B200_CTR_LAUNCHES] += 2;
        L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * 7.0 * nm;
        finish_send(nj, wt0, "gpunb_send");
        return;
    }
    // (pageable caller arrays on fewer than four devices keep the whole-snapshot path below: its staging copy runs on four
    // host threads under the uploads, a slice staged by ONE thread per device is slower there -- 2 GPUs: 2.4 vs 3.7 ms)
    if (!L.sh.on && L.devs.size() > 1 && L.send_scatter_min >= 0 && nj >= L.send_scatter_min && (direct || L.devs.size() >= 4)) {
        // One process driving G GPUs, the same scatter: device g uploads slice g over ITS PCIe link (pageable caller arrays:
        // staged by the device's own host thread) and pushes it to every peer over NVLink; every device then lays the G
        // slices out as m | x | v.  Instead of G uploads of the whole snapshot out of the same host memory.
        const int G = (int)L.devs.size();
        const int chunk = (nj + G - 1) / G;
        for (int g = 0; g < G; g++) {                 // the buffers the peers push into exist before anybody pushes
            Dev &d = L.devs[g];
            if ((size_t)G * 7 * chunk > d.send_tmp_n) {
                set_dev(d);
                CUDA_CHECK(cudaStreamSynchronize(d.st));
                dev_free(d.send_tmp);
                d.send_tmp_n = (size_t)G * 7 * (chunk + 1024);
                dev_alloc(d.send_tmp, d.send_tmp_n);
            }
        }
#pragma omp parallel for num_threads(enqueue_team()) schedule(static, 1) if (enqueue_team() > 1)
        for (int g = 0; g < G; g++) {
            Dev &d = L.devs[g];
            set_dev(d);
            const size_t lo = (size_t)std::min(nj, g * chunk), hi = (size_t)std::min(nj, (g + 1) * chunk), nm = hi - lo;
            double *mine = d.send_tmp + (size_t)g * 7 * chunk;
            if (nm > 0) {
                const double *sm = mj + lo, *sx = xj + 3 * lo, *sv = vj + 3 * lo;
                if (direct) {
                    CUDA_CHECK(cudaMemcpyAsync(mine, sm, sizeof(double) * nm, cudaMemcpyHostToDevice, d.st));
                    CUDA_CHECK(cudaMemcpyAsync(mine + chunk, sx, sizeof(double) * 3 * nm, cudaMemcpyHostToDevice, d.st));
                    CUDA_CHECK(cudaMemcpyAsync(mine + 4 * (size_t)chunk, sv, sizeof(double) * 3 * nm, cudaMemcpyHostToDevice, d.st));
                } else {
                    // pageable: this slice's part of the pinned staging buffer, in pieces of 32768 particles whose uploads
                    // run under the staging copy of the next piece
                    double *hs = h + 7 * lo;
                    for (size_t c = 0; c < nm; c += 32768) {
                        const size_t n = std::min<size_t>(32768, nm - c);
                        memcpy(hs + c, sm + c, sizeof(double) * n);
                        memcpy(hs + nm + 3 * c, sx + 3 * c, sizeof(double) * 3 * n);
                        memcpy(hs + 4 * nm + 3 * c, sv + 3 * c, sizeof(double) * 3 * n);
                        CUDA_CHECK(cudaMemcpyAsync(mine + c, hs + c, sizeof(double) * n, cudaMemcpyHostToDevice, d.st));
                        CUDA_CHECK(cudaMemcpyAsync(mine + chunk + 3 * c, hs + nm + 3 * c, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, d.st));
                        CUDA_CHECK(cudaMemcpyAsync(mine + 4 * (size_t)chunk + 3 * c, hs + 4 * nm + 3 * c, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, d.st));
                    }
                }
                for (int q = 1; q < G; q++) {         // staggered: at any moment the G devices push to G different peers
                    Dev &peer = L.devs[(g + q) % G];
                    CUDA_CHECK(cudaMemcpyPeerAsync(peer.send_tmp + (size_t)g * 7 * chunk, peer.id, mine, d.id, sizeof(double) * 7 * (size_t)chunk, d.st));
                }
            }
            CUDA_CHECK(cudaEventRecord(d.evpush, d.st));
        }
#pragma omp parallel for num_threads(enqueue_team()) schedule(static, 1) if (enqueue_team() > 1)
        for (int g = 0; g < G; g++) {
            Dev &d = L.devs[g];
            set_dev(d);
            for (int q = 1; q < G; q++) CUDA_CHECK(cudaStreamWaitEvent(d.st, L.devs[(g + q) % G].evpush, 0));
            send_unpack_kernel<<<(nj + 255) / 256, 256, 0, d.st>>>(nj, chunk, d.send_tmp, d.jraw);
            CUDA_CHECK(cudaGetLastError());
            ctr_add(GPUNB_B200_CTR_LAUNCHES, 1);
        }
        set_dev(L.devs[0]);
        L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * 7.0 * nj;
        finish_send(nj, wt0, "gpunb_send");
        return;
    }
    if (direct) {                                     // the caller's arrays are pinned: DMA from them, no staging copy
#pragma omp parallel for num_threads(enqueue_team()) schedule(static, 1) if (enqueue_team() > 1)
        for (int g = 0; g < (int)L.devs.size(); g++) {
            Dev &d = L.devs[g];
            set_dev(d);
            CUDA_CHECK(cudaMemcpyAsync(d.jraw, mj, sizeof(double) * nj, cudaMemcpyHostToDevice, d.st));
            CUDA_CHECK(cudaMemcpyAsync(d.jraw + nj, xj, sizeof(double) * 3 * nj, cudaMemcpyHostToDevice, d.st));
            CUDA_CHECK(cudaMemcpyAsync(d.jraw + 4 * (size_t)nj, vj, sizeof(double) * 3 * nj, cudaMemcpyHostToDevice, d.st));
        }
    }
    if (!direct && nj <= CHUNK) {
        // small snapshot: one chunk, whose staging layout (m | x | v) IS the device layout -- two uploads, the first in
        // flight while the velocities are staged
        const size_t n4 = 4 * (size_t)nj, n3 = 3 * (size_t)nj;
        memcpy(h, mj, sizeof(double) * nj);
        memcpy(h + nj, xj, sizeof(double) * n3);
        for (Dev &d : L.devs) { set_dev(d); CUDA_CHECK(cudaMemcpyAsync(d.jraw, h, sizeof(double) * n4, cudaMemcpyHostToDevice, d.st)); }
        memcpy(h + n4, vj, sizeof(double) * n3);
        for (Dev &d : L.devs) { set_dev(d); CUDA_CHECK(cudaMemcpyAsync(d.jraw + n4, h + n4, sizeof(double) * n3, cudaMemcpyHostToDevice, d.st)); }
    }
    for (int c0 = 0; c0 < nj && !direct && nj > CHUNK; c0 += CHUNK) {
        const size_t c = (size_t)c0, n = (size_t)((nj - c0 < CHUNK) ? nj - c0 : CHUNK);
        threaded_copy(h + c, mj + c, n);
        threaded_copy(h + nj + 3 * c, xj + 3 * c, 3 * n);
        threaded_copy(h + 4 * (size_t)nj + 3 * c, vj + 3 * c, 3 * n);
        for (Dev &d : L.devs) {
            set_dev(d);
            CUDA_CHECK(cudaMemcpyAsync(d.jraw + c, h + c, sizeof(double) * n, cudaMemcpyHostToDevice, d.st));
            CUDA_CHECK(cudaMemcpyAsync(d.jraw + nj + 3 * c, h + nj + 3 * c, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, d.st));
            CUDA_CHECK(cudaMemcpyAsync(d.jraw + 4 * (size_t)nj + 3 * c, h + 4 * (size_t)nj + 3 * c, sizeof(double) * 3 * n,
                                       cudaMemcpyHostToDevice, d.st));
        }
    }
    L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * 7.0 * nj * L.devs.size();
    finish_send(nj, wt0, "gpunb_send");
}

// Tiles of every device's j-shard from the m | x | v snapshot already in d.jraw (uploaded or predicted on the device).
void finish_send(int nj, double wt0, const char *who)
{
    const int R = total_ranks();
    const double wt1 = wtime();
    // Keep the Hilbert order of the previous snapshot?  k > 1: for k-1 of k snapshots.  Adaptive (default, one GPU): while
    // the tiles are still about as compact as right after the last sort -- the sum of their half-extents (an exact integer
    // sum, so the decision is reproducible) has grown by less than 10 % -- and for at most 64 snapshots.  The tiles are
    // re-packed from the current positions either way: lists stay bit-exact, forces within the same bars; only the
    // summation order follows the kept permutation.  Sharded runs always sort: every rank must cut the same tiles.
    const bool adaptive = L.resort_every == 0 && R == 1;
    bool reuse = false;
    if (L.resort_every > 1) reuse = (L.snapshots_since_sort % L.resort_every) != 0;
    else if (adaptive) reuse = L.devs[0].perm_n == nj && L.q_ref > 0.0 && L.q_last <= 1.10 * L.q_ref && L.snapshots_since_sort < 64;
    if (reuse && L.devs[0].perm_n != nj) reuse = false;
    L.snapshots_since_sort = reuse ? L.snapshots_since_sort + 1 : 1;
#pragma omp parallel for num_threads(enqueue_team()) schedule(static, 1) if (enqueue_team() > 1)
    for (int g = 0; g < (int)L.devs.size(); g++) {
        Dev &d = L.devs[g];
        set_dev(d);
        if (adaptive && !d.qsum) { dev_alloc(d.qsum, 2); CUDA_CHECK(cudaMemsetAsync(d.qsum, 0, 2 * sizeof(unsigned long long), d.st)); }
        unsigned long long *q_dev = nullptr;
        if (adaptive) CUDA_CHECK(cudaHostGetDevicePointer((void **)&q_dev, (void *)L.h_q, 0));
        CUDA_CHECK(cudaEventRecord(d.evs0, d.st));
        build_tiles(d, nj, d.jraw, d.jraw + nj, d.jraw + 4 * (size_t)nj, d.jtile, d.jidx, L.sh.on ? L.sh.rank : (int)g, R, d.ntiles, reuse,
                    adaptive ? d.qsum : nullptr, q_dev);
        CUDA_CHECK(cudaEventRecord(d.evs1, d.st));
    }
    for (size_t g = 0; g < L.devs.size(); g++) {
        set_dev(L.devs[g]);
        CUDA_CHECK(cudaStreamSynchronize(L.devs[g].st));
        if (L.h_nan[g]) FATAL("%s: NaN in j-particle data (reference asserts here, gpunb.velocity.cu:72-78)", who);
    }
    if (adaptive) {
        L.q_last = (double)*L.h_q;
        if (!reuse) L.q_ref = L.q_last;
    }
    if (reuse) L.ctr[GPUNB_B200_CTR_SENDS_ORDER_KEPT] += 1;
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, L.devs[0].evs0, L.devs[0].evs1));
    const double wt2 = wtime();
    L.ctr[GPUNB_B200_CTR_SEND_TILES_MS] += ms;
    L.ctr[GPUNB_B200_CTR_SEND_STAGE_MS] += (wt1 - wt0) * 1e3;
    L.ctr[GPUNB_B200_CTR_SEND_MS] += (wt2 - wt0) * 1e3;
    L.ctr[GPUNB_B200_CTR_SENDS] += 1;
    L.time_send += wt2;
}

// ---- device-resident predictor: host side ------------------------------------------------------
void ensure_state(Dev &d, int nj)
{
    set_dev(d);
    if (nj > d.state_cap) {
        if (d.state_n > 0) FATAL("gpunb_b200_state: nj=%d exceeds the state capacity %d (call gpunb_b200_state_all_ with the new set)", nj, d.state_cap);
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        dev_free(d.state);
        d.state_cap = (nj > L.nbmax ? nj : L.nbmax) + 64;
        dev_alloc(d.state, (size_t)14 * d.state_cap);
    }
    if (!d.upd_bad) { dev_alloc(d.upd_bad, 1); CUDA_CHECK(cudaMemsetAsync(d.upd_bad, 0, sizeof(int), d.st)); }
}

// Full state of the j-set (once, and again whenever the caller's particle table is re-ordered: KS creation /
// termination, escapers).  Arrays start at the first j-particle (IFIRST), like the arguments of gpunb_send_.
void lib_state_all(int nj, const double *body, const double *x0, const double *x0dot, const double *f, const double *fdot,
                   const double *t0)
{
    if (!L.is_open) FATAL("gpunb_b200_state_all called while the library is closed");
    if (nj > L.nbmax) FATAL("gpunb_b200_state_all: nj=%d exceeds nbmax=%d given to gpunb_open", nj, L.nbmax);
    const double *src[6] = {body, x0, x0dot, f, fdot, t0};
    const int width[6] = {1, 3, 3, 3, 3, 1};
    for (Dev &d : L.devs) { d.state_n = 0; ensure_state(d, nj); }
    // staged through the pinned snapshot buffer (7 (nbmax+64) doubles) in two halves: {body, x0, x0dot} then {f, fdot, t0}
    for (int half = 0; half < 2; half++) {
        double *h = L.h_j;
        size_t off = 0;
        for (int q = 3 * half; q < 3 * half + 3; q++) {
            threaded_copy(h + off, src[q], (size_t)width[q] * nj);
            off += (size_t)width[q] * nj;
        }
        for (Dev &d : L.devs) {
            set_dev(d);
            size_t o = 0;
            const size_t dst_off[6] = {0, (size_t)d.state_cap, 4 * (size_t)d.state_cap, 7 * (size_t)d.state_cap,
                                       10 * (size_t)d.state_cap, 13 * (size_t)d.state_cap};
            for (int q = 3 * half; q < 3 * half + 3; q++) {
                CUDA_CHECK(cudaMemcpyAsync(d.state + dst_off[q], h + o, sizeof(double) * width[q] * nj, cudaMemcpyHostToDevice, d.st));
                o += (size_t)width[q] * nj;
            }
        }
        for (Dev &d : L.devs) { set_dev(d); CUDA_CHECK(cudaStreamSynchronize(d.st)); }     // h is reused
    }
    for (Dev &d : L.devs) d.state_n = nj;
    L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * 14.0 * nj * L.devs.size();
}

// State of the n particles idx[k] (0-based, relative to the j array like the neighbour indices) that the integrator
// has just advanced: entry k of every array belongs to particle idx[k].
void lib_state_update(int n, const int *idx, const double *body, const double *x0, const double *x0dot, const double *f,
                      const double *fdot, const double *t0)
{
    if (!L.is_open) FATAL("gpunb_b200_state_update called while the library is closed");
    if (n <= 0) return;
    if (L.devs[0].state_n <= 0) FATAL("gpunb_b200_state_update before gpunb_b200_state_all_");
    if (n > L.h_upd_cap) {
        for (Dev &d : L.devs) { set_dev(d); CUDA_CHECK(cudaStreamSynchronize(d.st)); }
        host_free(L.h_upd); host_free(L.h_upd_idx);
        L.h_upd_cap = n + 4096;
        host_alloc(L.h_upd, (size_t)14 * L.h_upd_cap);
        host_alloc(L.h_upd_idx, (size_t)L.h_upd_cap);
    }
    for (int k = 0; k < n; k++) {
        double *r = L.h_upd + 14 * (size_t)k;
        r[0] = body[k];
        for (int c = 0; c < 3; c++) {
            r[1 + c] = x0[3 * (size_t)k + c]; r[4 + c] = x0dot[3 * (size_t)k + c];
            r[7 + c] = f[3 * (size_t)k + c];  r[10 + c] = fdot[3 * (size_t)k + c];
        }
        r[13] = t0[k];
        L.h_upd_idx[k] = idx[k];
    }
    for (size_t g = 0; g < L.devs.size(); g++) {
        Dev &d = L.devs[g];
        set_dev(d);
        if (n > d.upd_cap) {
            CUDA_CHECK(cudaStreamSynchronize(d.st));
            dev_free(d.upd_rec); dev_free(d.upd_idx);
            d.upd_cap = L.h_upd_cap;
            dev_alloc(d.upd_rec, (size_t)14 * d.upd_cap); dev_alloc(d.upd_idx, (size_t)d.upd_cap);
        }
        CUDA_CHECK(cudaMemcpyAsync(d.upd_rec, L.h_upd, sizeof(double) * 14 * n, cudaMemcpyHostToDevice, d.st));
        CUDA_CHECK(cudaMemcpyAsync(d.upd_idx, L.h_upd_idx, sizeof(int) * n, cudaMemcpyHostToDevice, d.st));
        state_scatter_kernel<<<(n + 127) / 128, 128, 0, d.st>>>(n, d.state_cap, d.state_n, d.upd_idx, d.upd_rec, d.state, d.upd_bad);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(L.h_flag + g, d.upd_bad, sizeof(int), cudaMemcpyDeviceToHost, d.st));
    }
    for (size_t g = 0; g < L.devs.size(); g++) {      // the pinned staging is reused by the next update
        set_dev(L.devs[g]);
        CUDA_CHECK(cudaStreamSynchronize(L.devs[g].st));
        if (L.h_flag[g]) FATAL("gpunb_b200_state_update: particle index outside [0, %d)", L.devs[g].state_n);
    }
    L.ctr[GPUNB_B200_CTR_H2D_BYTES] += (sizeof(double) * 14.0 + sizeof(int)) * n * L.devs.size();
    L.ctr[GPUNB_B200_CTR_LAUNCHES] += (double)L.devs.size();
}

// xbpredall + gpunb_send_ in one call, without the PCIe upload: predict every j-particle to `time` on the device and
// rebuild the tiles.  nj may be smaller than the state (the caller's NTOT shrinks); it cannot exceed it.
void lib_predict_send(int nj, double time)
{
    if (!L.is_open) FATAL("gpunb_b200_predict_send called while the library is closed");
    const double wt0 = wtime();
    L.time_send -= wt0;
    L.isend++;
    L.nbody = nj;
    set_shards(nj);
    for (Dev &d : L.devs) {
        if (nj > d.state_n) FATAL("gpunb_b200_predict_send: nj=%d but the device state holds %d particles", nj, d.state_n);
        set_dev(d);
        predict_kernel<<<(nj + 255) / 256, 256, 0, d.st>>>(nj, d.state_cap, d.state, time, d.jraw);
        CUDA_CHECK(cudaGetLastError());
    }
    L.ctr[GPUNB_B200_CTR_LAUNCHES] += (double)L.devs.size();
    finish_send(nj, wt0, "gpunb_b200_predict_send");
}

// predict_send from particle records another library keeps on this device (see predict_records_kernel).  The caller
// guarantees that the records are complete (irr_b200_flush_) and stay untouched during the call.
void lib_predict_send_records(int nj, double time, const double *records_dev, int stride)
{
    if (!L.is_open) FATAL("gpunb_b200_predict_send_records called while the library is closed");
    if (nj > L.nbmax) FATAL("gpunb_b200_predict_send_records: nj=%d exceeds nbmax=%d given to gpunb_open", nj, L.nbmax);
    if (L.devs.size() != 1 || L.sh.on) FATAL("gpunb_b200_predict_send_records: one process, one GPU (the records live on one device)");
    if (stride < 16 || (stride & 1) || ((uintptr_t)records_dev & 15)) FATAL("gpunb_b200_predict_send_records: records of %d doubles at %p", stride, (const void *)records_dev);
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, records_dev) != cudaSuccess || attr.type != cudaMemoryTypeDevice || attr.device != L.devs[0].id)
        FATAL("gpunb_b200_predict_send_records: %p is not memory of device %d", (const void *)records_dev, L.devs[0].id);
    const double wt0 = wtime();
    L.time_send -= wt0;
    L.isend++;
    L.nbody = nj;
    set_shards(nj);
    Dev &d = L.devs[0];
    set_dev(d);
    predict_records_kernel<<<(nj + 255) / 256, 256, 0, d.st>>>(nj, records_dev, stride, time, d.jraw);
    CUDA_CHECK(cudaGetLastError());
    L.ctr[GPUNB_B200_CTR_LAUNCHES] += 1;
    finish_send(nj, wt0, "gpunb_b200_predict_send_records");
}

// Predicted x / xdot of the particles idx[k] from the snapshot built by the last predict_send (the caller needs them
// for the i-block it passes to gpunb_regf_ when it does not run its own full predictor).
void lib_get_predicted(int n, const int *idx, double *x, double *xdot)
{
    if (!L.is_open) FATAL("gpunb_b200_get_predicted called while the library is closed");
    if (n <= 0) return;
    Dev &d = L.devs[0];
    set_dev(d);
    const int nj = d.nj_total;
    // reuses the staging buffers of state_update: idx up, 6 doubles per particle down
    if (n > L.h_upd_cap) {
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        host_free(L.h_upd); host_free(L.h_upd_idx);
        L.h_upd_cap = n + 4096;
        host_alloc(L.h_upd, (size_t)14 * L.h_upd_cap);
        host_alloc(L.h_upd_idx, (size_t)L.h_upd_cap);
    }
    if (n > d.upd_cap) {
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        dev_free(d.upd_rec); dev_free(d.upd_idx);
        d.upd_cap = L.h_upd_cap;
        dev_alloc(d.upd_rec, (size_t)14 * d.upd_cap); dev_alloc(d.upd_idx, (size_t)d.upd_cap);
    }
    if (!d.upd_bad) { dev_alloc(d.upd_bad, 1); CUDA_CHECK(cudaMemsetAsync(d.upd_bad, 0, sizeof(int), d.st)); }
    memcpy(L.h_upd_idx, idx, sizeof(int) * n);
    CUDA_CHECK(cudaMemcpyAsync(d.upd_idx, L.h_upd_idx, sizeof(int) * n, cudaMemcpyHostToDevice, d.st));
    snapshot_gather_kernel<<<(n + 127) / 128, 128, 0, d.st>>>(n, nj, d.upd_idx, d.jraw, d.upd_rec, d.upd_bad);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(L.h_upd, d.upd_rec, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, d.st));
    CUDA_CHECK(cudaMemcpyAsync(L.h_flag, d.upd_bad, sizeof(int), cudaMemcpyDeviceToHost, d.st));
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    if (L.h_flag[0]) FATAL("gpunb_b200_get_predicted: particle index outside [0, %d)", nj);
    for (int k = 0; k < n; k++)
        for (int c = 0; c < 3; c++) { x[3 * (size_t)k + c] = L.h_upd[6 * (size_t)k + c]; xdot[3 * (size_t)k + c] = L.h_upd[6 * (size_t)k + 3 + c]; }
    L.ctr[GPUNB_B200_CTR_D2H_BYTES] += sizeof(double) * 6.0 * n;
    L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(int) * (double)n;
    L.ctr[GPUNB_B200_CTR_LAUNCHES] += 1;
}

struct Plan { int n_itiles, S, n_items; };
Plan make_plan(const Dev &d, int nloc, int oversub)
{
    Plan p;
    p.n_itiles = (nloc + d.itile - 1) / d.itile;
    const int S1 = std::max(1, d.warps_resident / p.n_itiles);
    while (oversub > 1 && d.ntiles < (long long)MIN_TILES_PER_ITEM * S1 * oversub) oversub--;
    int S = (d.warps_resident * oversub) / p.n_itiles;
    if (S < 1) S = 1;
    if (S > d.ntiles) S = d.ntiles > 0 ? d.ntiles : 1;
    p.S = S;
    p.n_items = p.n_itiles * S;
    return p;
}

struct IBlock { const double *h2, *dtr, *xi, *vi; };

// exchange-buffer layout (one process per GPU).  XSLOTS exchange slots are used round robin by call number, so that
// several i-blocks can be in flight (pipelined sweeps); a slot is overwritten only after every peer has acknowledged
// reading its previous contents (acks, checked inside merge_kernel).
constexpr int    XSLOTS        = 8;
constexpr size_t XB_FR_BYTES   = (size_t)JOBCAP * 8 * sizeof(double);
constexpr size_t XB_ROWS_BYTES = (size_t)JOBCAP * 1024 /*ROWS_LMAX_CAP*/ * sizeof(int);
constexpr size_t XB_SLOT       = XB_FR_BYTES + XB_ROWS_BYTES;
constexpr size_t XB_FLAGS_OFF  = XSLOTS * XB_SLOT;
constexpr size_t XB_ACKS_OFF   = XB_FLAGS_OFF + (size_t)XSLOTS * 16 /*MAX_RANKS*/ * sizeof(unsigned long long);
constexpr size_t XB_BYTES      = XB_ACKS_OFF + (size_t)XSLOTS * 16 /*MAX_RANKS*/ * sizeof(unsigned long long);
inline double *xb_fr(unsigned char *b, int xs)   { return reinterpret_cast<double *>(b + xs * XB_SLOT); }
inline int    *xb_rows(unsigned char *b, int xs) { return reinterpret_cast<int *>(b + xs * XB_SLOT + XB_FR_BYTES); }
inline unsigned long long *xb_flags(unsigned char *b, int xs) { return reinterpret_cast<unsigned long long *>(b + XB_FLAGS_OFF) + xs * 16; }
inline unsigned long long *xb_acks(unsigned char *b, int xs)  { return reinterpret_cast<unsigned long long *>(b + XB_ACKS_OFF) + xs * 16; }

void launch_isort(Dev &d, cudaStream_t st, int ni_total, int block, const double *xi, int *iperm, int *iperm_host,
                  const int *blk_off = nullptr, int nblk_off = 0)
{   // blk_off (device, nblk_off + 1 entries): one job whose sort blocks are the ranks' i-slices; else uniform job blocks
    static int noisort = -1;
    if (noisort < 0) { const char *e = getenv("GPUNB_B200_NOISORT"); noisort = (e && atoi(e) > 0) ? 1 : 0; }
    const int sortblk = block <= NIMAX ? block : 1024;
    const int spj = (block + sortblk - 1) / sortblk;
    const int njobs = (ni_total + block - 1) / block;
    if (noisort) iota_kernel<<<(ni_total + 255) / 256, 256, 0, st>>>(ni_total, block, iperm, iperm_host);
    else if (blk_off) isort_kernel<<<nblk_off, 1024, 0, st>>>(ni_total, ni_total, NIMAX, blk_off, xi, d.hbits, iperm, iperm_host);
    else isort_kernel<<<njobs * spj, 1024, 0, st>>>(ni_total, block, sortblk, nullptr, xi, d.hbits, iperm, iperm_host);
    CUDA_CHECK(cudaGetLastError());
    ctr_add(GPUNB_B200_CTR_LAUNCHES, 1);
}

// What one launch group works on: the sorted slots [slot0, slot0 + nloc) of an i-block whose Morton order is iperm.
struct Job {
    int slot0, nloc;
    int lmax, nnbmax, m_flag;
    double *out_f; int *out_list;      // final results, row i = iperm[slot] (device memory or mapped host memory)
    int *out_list2 = nullptr;          // device copy of the final rows (gpunb_regf_: kept for gpunb_b200_regcor_last_)
    double *abi_acc = nullptr, *abi_jrk = nullptr, *abi_pot = nullptr;   // or the caller's pinned arrays (with out_list)
    // i-slice mode (collective gpunb_regf_, one process per GPU): this rank combines and receives only the sorted slots
    // [own0, own1) of the block -- its own i-slice -- and its result rows are numbered from row_base
    int own0 = 0, own1 = INT_MAX, row_base = 0;
    int merge_parts = 1;               // single GPU: merge launches (delivery parts) of this job
    int oversub = 1;                   // work items per resident warp slot (see OVERSUB / REGF_OVERSUB)
};

// Pair kernel (stream lo) + shard-local merge (stream hi) of one job on device d, pipeline slot sl.
void launch_regf(Dev &d, Slot &sl, cudaStream_t lo, cudaStream_t hi, const Job &j, const IBlock &ib, const int *iperm,
                 MergeArgs m, bool time_it, cudaEvent_t *tl)
{   // tl (optional): [2] after regf, [3] after merge.  m arrives with its outputs / exchange fields set.
    const Plan p = make_plan(d, j.nloc, j.oversub);
    RegfArgs a;
    a.tiles = d.jtile; a.jidx = d.jidx; a.ntiles = d.ntiles; a.iperm = iperm;
    { static int fn = -1; if (fn < 0) { const char *e = getenv("GPUNB_B200_FORCE_NEAR"); fn = e ? atoi(e) : 0; } a.force_near = fn; }
    { static int im = -1; if (im < 0) { const char *e = getenv("GPUNB_B200_ITMAP"); im = e ? atoi(e) : 0; } a.itmap = im; }
    { static int ne = -1; if (ne < 0) { const char *e = getenv("GPUNB_B200_NEAR_EXACT"); ne = e ? atoi(e) : 0; } a.near_scalar = L.near_exact >= 0 ? L.near_exact : ne; }
    a.stats = d.stats; a.wtime = d.wtime;
    a.h2 = ib.h2; a.dtr = ib.dtr; a.xi = ib.xi; a.vi = ib.vi;
    a.slot0 = j.slot0; a.nloc = j.nloc; a.n_itiles = p.n_itiles; a.S = p.S; a.n_items = p.n_items;
    a.part = sl.part; a.cnt = sl.cnt; a.seg = sl.seg; a.segcap = d.segcap;
    const int smem = WARPS * NSTAGE * TILE_BYTES + WARPS * NSTAGE * 8;
    const int blocks = (p.n_items + WARPS - 1) / WARPS;
    if (time_it) CUDA_CHECK(cudaEventRecord(d.ev0, lo));
    VARIANTS[d.variant].k[j.m_flag ? 1 : 0]<<<blocks, WARPS * 32, smem, lo>>>(a);
    CUDA_CHECK(cudaGetLastError());
    if (time_it) CUDA_CHECK(cudaEventRecord(d.ev1, lo));
    if (tl) CUDA_CHECK(cudaEventRecord(tl[2], lo));
    if (lo != hi) {
        CUDA_CHECK(cudaEventRecord(sl.ev_regf, lo));
        CUDA_CHECK(cudaStreamWaitEvent(hi, sl.ev_regf, 0));
    }
    m.part = sl.part; m.cnt = sl.cnt; m.seg = sl.seg; m.nloc = j.nloc; m.S = p.S; m.segcap = d.segcap;
    m.lmax = j.lmax; m.nnbmax = j.nnbmax;
    // delivery in parts (single GPU, staged results): one merge launch per part, an event behind each, so that the host
    // copies the rows of part p into the caller's arrays while part p+1 is being merged
    const int parts = (j.merge_parts > 1 && !m.sig.done_ctr) ? j.merge_parts : 1;
    for (int q = 0; q < parts; q++) {
        m.kl0 = (int)((long long)j.nloc * q / parts) & ~3;
        m.kl1 = q == parts - 1 ? j.nloc : (int)((long long)j.nloc * (q + 1) / parts) & ~3;
        // A kernel that may SPIN on peer flags never gets more than a few CTAs per SM: pair kernels of equal-priority
        // streams do not start in submission order, and a machine full of spinning CTAs of block b+1 would starve the
        // pair kernel of block b whose merge the peers are waiting for (seen at 8 ranks with 9472-row blocks).
        int mgrid = (m.kl1 - m.kl0 + 3) / 4;
        if (m.acks && mgrid > 2 * d.nsm) mgrid = 2 * d.nsm;
        if (m.kl1 > m.kl0) merge_kernel<<<mgrid, 128, 0, hi>>>(m);
        if (parts > 1) CUDA_CHECK(cudaEventRecord(d.slots[q].ev_start, hi));        // idle events of the unused pipeline slots
    }
    CUDA_CHECK(cudaGetLastError());
    ctr_add(GPUNB_B200_CTR_LAUNCHES, parts - 1);
    if (time_it) CUDA_CHECK(cudaEventRecord(d.ev2, hi));
    if (tl) CUDA_CHECK(cudaEventRecord(tl[3], hi));
    ctr_add(GPUNB_B200_CTR_LAUNCHES, 2);
}

MergeArgs merge_defaults()
{
    MergeArgs m;
    memset(&m, 0, sizeof(m));
    return m;
}

// One job on all shards + the exchange step, in pipeline slot q.  ib[g] / iperm[g] are DEVICE pointers valid on local
// device g.  own_streams: the job runs on the slot's own (lo, hi) stream pair and slot.ev_done marks its end;
// otherwise everything is queued on the device's main stream.  On completion j.out_f / j.out_list hold the combined
// result rows of the job's i-particles.
void run_job(const Job &j, const IBlock *ib, const int *const *iperm, int q, bool own_streams, bool time_it,
             cudaEvent_t *tl = nullptr)
{
    const int G = (int)L.devs.size();
    Dev &root = L.devs[0];
    Slot &sl = root.slots[q];
    set_dev(root);
    cudaStream_t lo = own_streams ? sl.lo : root.st, hi = own_streams ? sl.hi : root.st;
    if (!L.sh.on && G == 1) {          // single GPU: the shard-local merge IS the final result
        MergeArgs m = merge_defaults();
        m.iperm = iperm[0] + j.slot0; m.res_f = j.out_f; m.f_stride = 7; m.res_list = j.out_list; m.res_list2 = j.out_list2; m.sort = 1;
        m.abi_acc = j.abi_acc; m.abi_jrk = j.abi_jrk; m.abi_pot = j.abi_pot;
        launch_regf(root, sl, lo, hi, j, ib[0], iperm[0], m, time_it, tl);
    } else {
        CombineArgs c;
        memset(&c, 0, sizeof(c));
        c.nloc = j.nloc; c.lmax = j.lmax; c.nnbmax = j.nnbmax; c.res_f = j.out_f; c.res_list = j.out_list; c.res_list2 = j.out_list2;
        c.iperm = iperm[0] + j.slot0;
        c.kl0 = j.own0 > j.slot0 ? j.own0 - j.slot0 : 0;
        c.kl1 = j.own1 - j.slot0 < j.nloc ? j.own1 - j.slot0 : j.nloc;
        if (c.kl1 < c.kl0) c.kl1 = c.kl0;          // nothing of this rank in the job: the launch still acknowledges
        c.row_base = j.row_base;
        c.abi_acc = j.abi_acc; c.abi_jrk = j.abi_jrk; c.abi_pot = j.abi_pot;
        if (L.sh.on) {
            // One process per GPU.  No collective call in the data path: merge_kernel writes the shard result into
            // this rank's exchange slot and its last CTA raises this rank's flag in every peer's buffer over NVLink;
            // combine_kernel -- after seeing all R flags of this call -- pulls partial sums and the valid part of the
            // neighbour rows straight from the peers' HBM and its last CTA acknowledges to every peer.  Exchange
            // slots are reused every XSLOTS calls: merge_kernel first waits for every peer's ack of call seq-XSLOTS.
            Shard &sh = L.sh;
            const unsigned long long seq = ++sh.seq;
            const int xs = (int)(seq % XSLOTS);
            MergeArgs m = merge_defaults();
            m.iperm = nullptr; m.res_f = xb_fr(sh.xbuf, xs); m.f_stride = 8; m.res_list = xb_rows(sh.xbuf, xs); m.sort = 0;
            m.R = sh.R;
            if (seq > (unsigned long long)XSLOTS) { m.acks = xb_acks(sh.xbuf, xs); m.ack_need = (long long)seq - XSLOTS; }
            m.sig.done_ctr = sl.done_ctr; m.sig.R = sh.R; m.sig.seq = seq;
            c.ack.done_ctr = sl.done_ctr + 1; c.ack.R = sh.R; c.ack.seq = seq;
            for (int r = 0; r < sh.R; r++) {
                m.sig.peer[r] = xb_flags(sh.xbuf_peer[r], xs) + sh.rank;
                c.ack.peer[r] = xb_acks(sh.xbuf_peer[r], xs) + sh.rank;
                c.fr[r] = xb_fr(sh.xbuf_peer[r], xs);
                c.rows[r] = xb_rows(sh.xbuf_peer[r], xs);
            }
            launch_regf(root, sl, lo, hi, j, ib[0], iperm[0], m, time_it, tl);
            c.R = sh.R;
            c.flags = xb_flags(sh.xbuf, xs); c.seq = seq;
        } else {                       // one process, G GPUs: root waits for every shard, pulls over P2P
            if (own_streams) FATAL("internal: in-process multi-GPU jobs run on the main streams");
            // one enqueue thread per device; the root stream's waits on the shards follow behind the team (enqueued from a
            // shard's thread they could land in front of the root's own pair kernel)
#pragma omp parallel for num_threads(enqueue_team()) schedule(static, 1) if (enqueue_team() > 1)
            for (int g = 0; g < G; g++) {
                Dev &d = L.devs[g];
                set_dev(d);
                if (g > 0) CUDA_CHECK(cudaStreamWaitEvent(d.st, root.evdone, 0));     // previous combine has consumed d.fr / d.rows
                MergeArgs m = merge_defaults();
                m.iperm = nullptr; m.res_f = d.fr; m.f_stride = 8; m.res_list = d.rows; m.sort = 0;
                launch_regf(d, d.slots[q], d.st, d.st, j, ib[g], iperm[g], m, time_it && g == 0, g == 0 ? tl : nullptr);
                if (g > 0) CUDA_CHECK(cudaEventRecord(d.evdone, d.st));
                c.fr[g] = d.fr; c.rows[g] = d.rows;
            }
            c.R = G;
            set_dev(root);
            for (int g = 1; g < G; g++) CUDA_CHECK(cudaStreamWaitEvent(root.st, L.devs[g].evdone, 0));
        }
        const int ncomb = c.kl1 - c.kl0;
        int cgrid = ncomb > 0 ? (ncomb + 3) / 4 : 1;                             // 4 warps per CTA: one i each
        if (c.flags && cgrid > root.nsm) cgrid = root.nsm;                       // spins on peer flags: see launch_regf (one CTA per SM and pipeline slot)
        combine_kernel<<<cgrid, 128, 0, hi>>>(c);
        CUDA_CHECK(cudaGetLastError());
        L.ctr[GPUNB_B200_CTR_LAUNCHES] += 1;
        if (!L.sh.on) CUDA_CHECK(cudaEventRecord(root.evdone, root.st));
    }
    if (time_it) CUDA_CHECK(cudaEventRecord(root.ev3, hi));
    if (tl) CUDA_CHECK(cudaEventRecord(tl[4], hi));
    if (own_streams) { CUDA_CHECK(cudaEventRecord(sl.ev_done, hi)); sl.used = true; }
}

// Host side of gpunb_regf_: rows [k0, k1) of the sorted order from the mapped staging buffers (indexed by i) to the
// caller's arrays.  order == NULL: identity.
void scatter_rows(const int *order, int k0, int k1, int lmax, double *acc, double *jrk, double *pot, int *list)
{   // the rows were just written by the device, so they are cold for the CPU: a few host threads hide the DRAM latency
    // (the reference's host side is OpenMP as well, gpunb.velocity.cu:607-613,756)
    double bytes = 0;
#pragma omp parallel for num_threads(host_team()) schedule(static) reduction(+ : bytes) if (host_team() > 1 && k1 - k0 >= 64)
    for (int k = k0; k < k1; k++) {
        const int i = order ? order[k] : k;
        const double *f = L.h_f + 7 * (size_t)i;
        acc[3 * i] = f[0]; acc[3 * i + 1] = f[1]; acc[3 * i + 2] = f[2];
        jrk[3 * i] = f[3]; jrk[3 * i + 1] = f[4]; jrk[3 * i + 2] = f[5];
        pot[i] = f[6];
        const int *src = L.h_list + (size_t)i * lmax;
        int *dst = list + (size_t)i * lmax;
        const int n = src[0];
        dst[0] = n;
        if (n > 0) memcpy(dst + 1, src + 1, sizeof(int) * n);
        bytes += 56.0 + 4.0 * (1 + (n > 0 ? n : 0));
    }
    L.ctr[GPUNB_B200_CTR_D2H_BYTES] += bytes;      // what the kernels wrote over PCIe into the mapped buffers
}

void lib_regf_islice(int ni, const double *h2, const double *dtr, const double *xi, const double *vi,
                     double *acc, double *jrk, double *pot, int lmax, int nnbmax, int *list, int m_flag);

void lib_regf(int ni, const double *h2, const double *dtr, const double *xi, const double *vi,
              double *acc, double *jrk, double *pot, int lmax, int nnbmax, int *list, int m_flag)
{
    if (!L.is_open) FATAL("gpunb_regf called while the library is closed");
    if (L.sh.on && L.sh.islice) { lib_regf_islice(ni, h2, dtr, xi, vi, acc, jrk, pot, lmax, nnbmax, list, m_flag); return; }
    if (!(0 < ni && ni <= NIMAX)) FATAL("gpunb_regf: ni=%d out of range (0, %d]", ni, NIMAX);
    if (nnbmax + 1 > lmax) FATAL("gpunb_regf: nnbmax=%d does not fit rows of lmax=%d", nnbmax, lmax);
    if (nnbmax > SORT_CAP) FATAL("gpunb_regf: nnbmax=%d exceeds the list capacity %d of this build", nnbmax, SORT_CAP);
    const double wt_in = wtime();
    L.numInter += (long long)ni * L.nbody;       // reference counts every pair, self included (:747)
    L.ctr[GPUNB_B200_CTR_INTERACTIONS] += (double)ni * L.nbody;
    L.ini += ni; L.icall++;

    // NaN check + pack of the i-block (reference asserts per element, gpunb.velocity.cu:109-115)
    double *h = L.h_i;
    memcpy(h, h2, sizeof(double) * ni);
    memcpy(h + ni, dtr, sizeof(double) * ni);
    memcpy(h + 2 * (size_t)ni, xi, sizeof(double) * 3 * ni);
    memcpy(h + 5 * (size_t)ni, vi, sizeof(double) * 3 * ni);
    for (int k = 0; k < 8 * ni; k++) if (h[k] != h[k]) FATAL("gpunb_regf: NaN in i-particle data");
    const double wt_packed = wtime();
    double t_wait = 0.0;
    const int G = (int)L.devs.size();
    Dev &root = L.devs[0];
    // One call = nsub sub-blocks of the Morton-sorted i-block, each in its own pipeline slot: the pair kernel of
    // sub-block q+1 fills the SMs as the CTAs of sub-block q retire, and merge, the PCIe writes of the result rows
    // and the host-side copy into the caller's arrays of sub-block q all run beside it.
    // A sub-block must keep the pair kernel busy for >~150 us, or its fixed costs (launch, merge, exchange step) show.
    int nsub = (G == 1) ? L.nsub : 1;
    if (nsub > ni / 256) nsub = ni / 256;
    // decided from rank-invariant quantities only: every rank of a sharded run must cut the call into the same
    // sub-blocks (each is one exchange step), and shard sizes differ by a tile between ranks
    const double pairs = (double)ni * L.nbody / total_ranks();
    if (!L.nsub_forced && nsub > (int)(pairs / L.sub_pairs)) nsub = (int)(pairs / L.sub_pairs);
    if (nsub < 1) nsub = 1;
    IBlock ib[MAX_RANKS];
    const int *ipm[MAX_RANKS];
    const bool i_sorted = !(ni <= root.itile || (double)ni * L.nbody < L.isort_pairs);
#pragma omp parallel for num_threads(enqueue_team()) schedule(static, 1) if (enqueue_team() > 1)
    for (int g = 0; g < G; g++) {
        Dev &d = L.devs[g];
        set_dev(d);
        ensure_work_buffers(d, lmax, nnbmax, g == 0, nsub, false);
        CUDA_CHECK(cudaMemcpyAsync(d.ibuf, h, sizeof(double) * 8 * ni, cudaMemcpyHostToDevice, d.st));
        ctr_add(GPUNB_B200_CTR_H2D_BYTES, sizeof(double) * 8.0 * ni);
        ib[g] = IBlock{d.ibuf, d.ibuf + ni, d.ibuf + 2 * (size_t)ni, d.ibuf + 5 * (size_t)ni};
        // one i-tile: the order does not matter.  Small calls (ni x nj below ~2.5e7 pairs: the pair kernel is shorter
        // than the 12 us the sort adds to the critical path of a synchronous call) keep the caller's order too:
        // measured at N = 10^4, 76 vs 89 us per call at ni = 256, 118 vs 126 us at ni = 1024 (profiles/r2d_small_n.txt).
        // Rank-invariant (ni and the global nj), like every decision that shapes the exchange.
        if (!i_sorted) {
            ipm[g] = d.iperm_identity;
        } else {
            launch_isort(d, d.st, ni, ni, ib[g].xi, d.iperm, g == 0 ? L.h_iperm_dev : nullptr);
            ipm[g] = d.iperm;
        }
    }
    set_dev(root);
    Job j;
    j.lmax = lmax; j.nnbmax = nnbmax; j.m_flag = m_flag; j.out_f = L.h_f_dev; j.out_list = L.h_list_dev;
    j.oversub = 1;                     // set below: only a call that is ONE launch is oversubscribed
    j.out_list2 = root.last_rows;      // the rows also stay on the device, for the list bookkeeping that follows (regcor_b200.cu)
    L.last_rows_ni = 0;
    // Output arrays the caller has pinned (gpunb_b200_pin_host_): merge / combine write the ABI layout straight into
    // them over PCIe and the host-side copy of the rows disappears.
    bool direct_out = false;
    {
        double *a_acc = pinned_alias(acc, (size_t)3 * ni), *a_jrk = pinned_alias(jrk, (size_t)3 * ni), *a_pot = pinned_alias(pot, (size_t)ni);
        int *a_list = pinned_alias(list, (size_t)ni * lmax);
        if (a_acc && a_jrk && a_pot && a_list) {
            direct_out = true;
            j.abi_acc = a_acc; j.abi_jrk = a_jrk; j.abi_pot = a_pot; j.out_list = a_list; j.out_f = nullptr;
        }
    }
    auto deliver = [&](int k0, int k1) {
        if (!direct_out) { scatter_rows(i_sorted ? L.h_iperm : nullptr, k0, k1, lmax, acc, jrk, pot, list); return; }
        if (k0 == 0) {                 // bytes the kernels wrote over PCIe (counted once per call)
            double bytes = 0;
            for (int i = 0; i < ni; i++) { const int c = list[(size_t)i * lmax]; bytes += 56.0 + 4.0 * (1 + (c > 0 ? c : 0)); }
            L.ctr[GPUNB_B200_CTR_D2H_BYTES] += bytes;
        }
    };
    double t_scatter = 0.0;
    if (nsub == 1) {
        j.slot0 = 0; j.nloc = ni;
        j.oversub = L.regf_oversub;
        // staged results of a large block on one GPU: delivered in parts, the host copy of part p beside the merge of p+1
        // (measured at N = 10^4, ni = 1024: 162 us per call with 4 parts against 152 us with one merge launch -- the extra
        // launches and event waits cost more than the overlapped host copy saves; kept for tuning: GPUNB_B200_MERGE_PARTS)
        static int mp = -1;
        if (mp < 0) { const char *e = getenv("GPUNB_B200_MERGE_PARTS"); mp = (e && atoi(e) >= 1 && atoi(e) <= MAX_SLOTS) ? atoi(e) : 1; }
        const int parts = (G == 1 && !L.sh.on && !direct_out && ni >= 512) ? mp : 1;
        j.merge_parts = parts;
        run_job(j, ib, ipm, 0, false, true);
        L.ctr[GPUNB_B200_CTR_HOST_ENQUEUE_MS] += (wtime() - wt_packed) * 1e3;
        if (parts > 1) {
            for (int q = 0; q < parts; q++) {
                const int k0 = (int)((long long)ni * q / parts) & ~3;
                const int k1 = q == parts - 1 ? ni : (int)((long long)ni * (q + 1) / parts) & ~3;
                const double tw = wtime();
                CUDA_CHECK(cudaEventSynchronize(root.slots[q].ev_start));
                const double t0 = wtime();
                t_wait += t0 - tw;
                if (k1 > k0) deliver(k0, k1);
                t_scatter += wtime() - t0;
            }
            CUDA_CHECK(cudaStreamSynchronize(root.st));
        } else {
            const double tw = wtime();
            CUDA_CHECK(cudaStreamSynchronize(root.st));
            const double t0 = wtime();
            t_wait = t0 - tw;
            deliver(0, ni);
            t_scatter = wtime() - t0;
        }
    } else {
        // Whole i-tiles per sub-block, equal sizes by default.  Tapering sizes (weights 7:5:3:1 for four, so that what
        // stays exposed at the end of the call -- the unfilled tail of the last pair kernel, its merge and the host copy
        // of its rows -- shrinks with the last sub-block) were measured: 1206 us per call either way at 4 sub-blocks,
        // because the hardware does not start the pair kernels of equal-priority streams in submission order (with two
        // sub-blocks the SMALL one ran first).  The streams are still released in slot order.
        int off[MAX_SLOTS + 1] = {0};
        int nq = 0;
        {
            const int wsum = nsub * nsub;                       // sum of 2(nsub-q)-1
            for (int q = 0; q < nsub && off[nq] < ni; q++) {
                int sz = L.taper ? (int)((long long)ni * (2 * (nsub - q) - 1) / wsum) : (ni + nsub - 1) / nsub;
                sz = (sz + 31) & ~31;
                if (sz < 32) sz = 32;
                if (q == nsub - 1 || off[nq] + sz > ni) sz = ni - off[nq];
                off[nq + 1] = off[nq] + sz;
                nq++;
            }
        }
        CUDA_CHECK(cudaEventRecord(root.ev_fork, root.st));
        for (int q = 0; q < nq; q++) {
            Slot &sl = root.slots[q];
            CUDA_CHECK(cudaStreamWaitEvent(sl.lo, q == 0 ? root.ev_fork : root.slots[q - 1].ev_start, 0));
            CUDA_CHECK(cudaEventRecord(sl.ev_start, sl.lo));
            j.slot0 = off[q]; j.nloc = off[q + 1] - off[q];
            // sub-blocks are not oversubscribed: each pair kernel is followed by the next one, and for the last one the longer
            // merge (4x the partial records, on the critical path of the call) costs more than the shorter tail saves
            // (measured on one box: 1123 us per call without, 1152 us with the last sub-block at 4x, 1180 us with all)
            if (q == 0) CUDA_CHECK(cudaEventRecord(root.ev0, sl.lo));       // "grav" span: start of the first pair kernel ...
            run_job(j, ib, ipm, q, true, false);
        }
        CUDA_CHECK(cudaEventRecord(root.ev1, root.slots[nq - 1].lo));       // ... to the end of the last one
        CUDA_CHECK(cudaEventRecord(root.ev3, root.slots[nq - 1].hi));
        L.ctr[GPUNB_B200_CTR_HOST_ENQUEUE_MS] += (wtime() - wt_packed) * 1e3;
        for (int q = 0; q < nq; q++) {
            const double tw = wtime();
            CUDA_CHECK(cudaEventSynchronize(root.slots[q].ev_done));
            const double t0 = wtime();
            t_wait += t0 - tw;
            if (!direct_out) deliver(off[q], off[q + 1]);
            else if (q == nq - 1) deliver(0, ni);
            t_scatter += wtime() - t0;
        }
        for (int q = 0; q < nq; q++) CUDA_CHECK(cudaStreamWaitEvent(root.st, root.slots[q].ev_done, 0));
        CUDA_CHECK(cudaEventSynchronize(root.ev3));
        CUDA_CHECK(cudaEventSynchronize(root.ev1));      // recorded on a lo stream: not implied by ev3 / ev_done
    }
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, root.ev0, root.ev1)); L.ctr[GPUNB_B200_CTR_GRAV_MS] += ms;
    CUDA_CHECK(cudaEventElapsedTime(&ms, root.ev1, root.ev3)); L.ctr[GPUNB_B200_CTR_MERGE_MS] += ms;
    L.ctr[GPUNB_B200_CTR_GRAV_LAUNCHES] += 1;
    const double wt = wtime();
    L.time_grav += (wt - wt_in) - t_scatter; L.time_reduce += t_scatter;      // reference buckets: grav(s), nb(s)
    L.ctr[GPUNB_B200_CTR_HOST_PACK_MS] += (wt_packed - wt_in) * 1e3;
    L.ctr[GPUNB_B200_CTR_HOST_WAIT_MS] += t_wait * 1e3;
    L.ctr[GPUNB_B200_CTR_HOST_SCATTER_MS] += t_scatter * 1e3;
    L.last_ni = ni; L.last_lmax = lmax; L.last_on_host = true;
    L.last_rows_ni = ni; L.last_rows_lmax = lmax;
}

// gpunb_regf_ in i-slice mode (one process per GPU, gpunb_b200_set_islice): a collective call.  Every rank passes its own
// i-slice (ni may differ between ranks and may be 0); the slices meet in shared memory, every rank computes the UNION
// against its j-shard in one launch group (the pair kernel runs on ni_total x nj/R pairs, so its fixed costs are paid
// once per R slices), merge publishes the shard rows of the whole union, and combine_kernel on rank r pulls and delivers
// only the rows of rank r's own slice: result traffic over NVLink and PCIe is partitioned, not replicated.
void lib_regf_islice(int ni, const double *h2, const double *dtr, const double *xi, const double *vi,
                     double *acc, double *jrk, double *pot, int lmax, int nnbmax, int *list, int m_flag)
{
    Shard &sh = L.sh;
    Dev &root = L.devs[0];
    const int R = sh.R, me = sh.rank;
    L.last_rows_ni = 0;
    if (!(0 <= ni && ni <= NIMAX)) FATAL("gpunb_regf (i-slice mode): ni=%d out of range [0, %d]", ni, NIMAX);
    if (nnbmax + 1 > lmax) FATAL("gpunb_regf: nnbmax=%d does not fit rows of lmax=%d", nnbmax, lmax);
    if (nnbmax > SORT_CAP) FATAL("gpunb_regf: nnbmax=%d exceeds the list capacity %d of this build", nnbmax, SORT_CAP);
    const double wt_in = wtime();
    L.numInter += (long long)ni * L.nbody;
    L.ctr[GPUNB_B200_CTR_INTERACTIONS] += (double)ni * L.nbody;
    L.ini += ni; L.icall++;
    // ---- rendezvous: publish this rank's slice, wait for everybody's
    const unsigned long long c = ++sh.icall;
    const int p = (int)(c & 1ull);
    {
        ShmSlice &mine = sh.shm->slice[p][me];
        mine.ni = ni; mine.lmax = lmax; mine.nnbmax = nnbmax; mine.m_flag = m_flag;
        double *h = mine.data;
        memcpy(h, h2, sizeof(double) * ni);
        memcpy(h + ni, dtr, sizeof(double) * ni);
        memcpy(h + 2 * (size_t)ni, xi, sizeof(double) * 3 * ni);
        memcpy(h + 5 * (size_t)ni, vi, sizeof(double) * 3 * ni);
        for (int k = 0; k < 8 * ni; k++) if (h[k] != h[k]) FATAL("gpunb_regf: NaN in i-particle data");
        __atomic_store_n(&sh.shm->seq[p][me], c, __ATOMIC_RELEASE);
    }
    int off[MAX_RANKS + 1];
    off[0] = 0;
    for (int q = 0; q < R; q++) {
        const double tw = wtime();
        unsigned spins = 0;
        while (__atomic_load_n(&sh.shm->seq[p][q], __ATOMIC_ACQUIRE) < c) {
            if ((++spins & 0x3ffu) == 0u) {
                if (wtime() - tw > 120.0) FATAL("gpunb_regf (i-slice mode): rank %d has not joined collective call %llu after 120 s", q, c);
                sched_yield();
            }
        }
        const ShmSlice &sl = sh.shm->slice[p][q];
        if (sl.lmax != lmax || sl.nnbmax != nnbmax || sl.m_flag != m_flag)
            FATAL("gpunb_regf (i-slice mode): rank %d passed lmax/nnbmax/m_flag = %d/%d/%d, this rank %d/%d/%d", q, sl.lmax, sl.nnbmax, sl.m_flag, lmax, nnbmax, m_flag);
        off[q + 1] = off[q] + sl.ni;
    }
    const int tot = off[R];
    const double wt_packed0 = wtime();
    if (tot == 0) return;
    if (tot > JOBCAP) FATAL("gpunb_regf (i-slice mode): %d i-particles in one collective call exceed the capacity %d", tot, JOBCAP);
    // ---- union of the slices: [slice offsets | h2 | dtr | x | v] in pinned staging, one upload
    int *hoff = reinterpret_cast<int *>(L.h_i);
    for (int q = 0; q <= R; q++) hoff[q] = off[q];
    double *h = L.h_i + 64;
    for (int q = 0; q < R; q++) {
        const ShmSlice &sl = sh.shm->slice[p][q];
        const int n = sl.ni;
        memcpy(h + off[q], sl.data, sizeof(double) * n);
        memcpy(h + tot + off[q], sl.data + n, sizeof(double) * n);
        memcpy(h + 2 * (size_t)tot + 3 * (size_t)off[q], sl.data + 2 * (size_t)n, sizeof(double) * 3 * n);
        memcpy(h + 5 * (size_t)tot + 3 * (size_t)off[q], sl.data + 5 * (size_t)n, sizeof(double) * 3 * n);
    }
    const double wt_packed = wtime();
    int nsub = L.nsub;
    if (nsub > tot / 256) nsub = tot / 256;
    const double pairs = (double)tot * L.nbody / R;          // rank-invariant: every rank cuts the same sub-blocks
    if (!L.nsub_forced && nsub > (int)(pairs / L.sub_pairs)) nsub = (int)(pairs / L.sub_pairs);
    if (nsub < 1) nsub = 1;
    set_dev(root);
    ensure_work_buffers(root, lmax, nnbmax, true, nsub, false);
    CUDA_CHECK(cudaMemcpyAsync(root.ibuf, L.h_i, sizeof(double) * (64 + 8 * (size_t)tot), cudaMemcpyHostToDevice, root.st));
    L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * (64.0 + 8.0 * tot);
    const double *ib0 = root.ibuf + 64;
    IBlock ib[1] = {IBlock{ib0, ib0 + tot, ib0 + 2 * (size_t)tot, ib0 + 5 * (size_t)tot}};
    const int *ipm[1];
    if (tot <= root.itile) {
        ipm[0] = root.iperm_identity;
    } else {
        launch_isort(root, root.st, tot, tot, ib[0].xi, root.iperm, nullptr, reinterpret_cast<const int *>(root.ibuf), R);
        ipm[0] = root.iperm;
    }
    Job j;
    j.lmax = lmax; j.nnbmax = nnbmax; j.m_flag = m_flag; j.out_f = L.h_f_dev; j.out_list = L.h_list_dev;
    j.own0 = off[me]; j.own1 = off[me + 1]; j.row_base = off[me];
    j.oversub = L.regf_oversub;
    bool direct_out = false;
    if (ni > 0) {
        double *a_acc = pinned_alias(acc, (size_t)3 * ni), *a_jrk = pinned_alias(jrk, (size_t)3 * ni), *a_pot = pinned_alias(pot, (size_t)ni);
        int *a_list = pinned_alias(list, (size_t)ni * lmax);
        if (a_acc && a_jrk && a_pot && a_list) {
            direct_out = true;
            j.abi_acc = a_acc; j.abi_jrk = a_jrk; j.abi_pot = a_pot; j.out_list = a_list; j.out_f = nullptr;
        }
    }
    int soff[MAX_SLOTS + 1] = {0};
    int nq = 0;
    for (int q = 0; q < nsub && soff[nq] < tot; q++) {
        int sz = (tot + nsub - 1) / nsub;
        sz = (sz + 31) & ~31;
        if (q == nsub - 1 || soff[nq] + sz > tot) sz = tot - soff[nq];
        soff[nq + 1] = soff[nq] + sz;
        nq++;
    }
    CUDA_CHECK(cudaEventRecord(root.ev_fork, root.st));
    for (int q = 0; q < nq; q++) {
        Slot &sl = root.slots[q];
        CUDA_CHECK(cudaStreamWaitEvent(sl.lo, q == 0 ? root.ev_fork : root.slots[q - 1].ev_start, 0));
        CUDA_CHECK(cudaEventRecord(sl.ev_start, sl.lo));
        j.slot0 = soff[q]; j.nloc = soff[q + 1] - soff[q];
        if (q == 0) CUDA_CHECK(cudaEventRecord(root.ev0, sl.lo));
        run_job(j, ib, ipm, q, true, false);
    }
    CUDA_CHECK(cudaEventRecord(root.ev1, root.slots[nq - 1].lo));
    CUDA_CHECK(cudaEventRecord(root.ev3, root.slots[nq - 1].hi));
    L.ctr[GPUNB_B200_CTR_HOST_ENQUEUE_MS] += (wtime() - wt_packed) * 1e3;
    const double tw = wtime();
    for (int q = 0; q < nq; q++) CUDA_CHECK(cudaEventSynchronize(root.slots[q].ev_done));
    for (int q = 0; q < nq; q++) CUDA_CHECK(cudaStreamWaitEvent(root.st, root.slots[q].ev_done, 0));
    CUDA_CHECK(cudaEventSynchronize(root.ev3));
    CUDA_CHECK(cudaEventSynchronize(root.ev1));
    const double t0 = wtime();
    if (ni > 0) {
        if (!direct_out) scatter_rows(nullptr, 0, ni, lmax, acc, jrk, pot, list);
        else {
            double bytes = 0;
            for (int i = 0; i < ni; i++) { const int cnt = list[(size_t)i * lmax]; bytes += 56.0 + 4.0 * (1 + (cnt > 0 ? cnt : 0)); }
            L.ctr[GPUNB_B200_CTR_D2H_BYTES] += bytes;
        }
    }
    const double wt = wtime();
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, root.ev0, root.ev1)); L.ctr[GPUNB_B200_CTR_GRAV_MS] += ms;
    CUDA_CHECK(cudaEventElapsedTime(&ms, root.ev1, root.ev3)); L.ctr[GPUNB_B200_CTR_MERGE_MS] += ms;
    L.ctr[GPUNB_B200_CTR_GRAV_LAUNCHES] += 1;
    L.time_grav += (wt - wt_in) - (wt - t0); L.time_reduce += wt - t0;
    L.ctr[GPUNB_B200_CTR_HOST_PACK_MS] += (wt_packed - wt_in) * 1e3;
    L.ctr[GPUNB_B200_CTR_HOST_RENDEZVOUS_MS] += (wt_packed0 - wt_in) * 1e3;
    L.ctr[GPUNB_B200_CTR_HOST_WAIT_MS] += (t0 - tw) * 1e3;
    L.ctr[GPUNB_B200_CTR_HOST_SCATTER_MS] += (wt - t0) * 1e3;
    L.last_ni = ni; L.last_lmax = lmax; L.last_on_host = true;
}

void lib_profile(int irank)
{
    if (L.icall) {        // same line format as the reference (gpunb.velocity.cu:894)
        fprintf(stderr, "[R.%d GPU Reg.F ] Nsend %d  Ngrav %d  <Ni> %d   send(s) %f grav(s) %f  nb(s) %f  out(s) %f  Perf.(Gflops) %f\n",
                irank, L.isend, L.icall, L.isend ? L.ini / L.isend : L.ini, L.time_send, L.time_grav, L.time_reduce, L.time_out,
                60.e-9 * L.numInter / L.time_grav);
    }
    L.time_send = L.time_grav = L.time_reduce = L.time_out = 0.0;
    L.numInter = 0; L.icall = L.ini = L.isend = 0;
}

// gpupot: may be called with the library closed (reference: gpupot.gpu.cu:69 only needs devinit).
// Multi-GPU: the j-set is sharded like regf's -- every device tiles and sums over every R-th Hilbert tile
// for the whole i-range -- and the partial potentials are added in rank order: over P2P by a kernel on the first
// device (one process, several GPUs) or after ONE ncclAllGather of the ni partials (one process per GPU; this is the
// path's real exchange step for gpupot, 8 B per i and rank).
void lib_pot(int irank, int istart, int ni, int n, const double *m, const double *x, double *pot)
{
    lib_devinit(irank);
    const double t0 = wtime();
    if (ni <= 0) return;
    if (istart < 1 || istart - 1 + ni > n) FATAL("gpupot: istart=%d ni=%d outside 1..n=%d", istart, ni, n);
    // pinned staging: the snapshot buffer of gpunb_send_ when the library is open and large enough (ADJUST calls gpupot_ in
    // the middle of a run: no 32 MB pinned allocation on the first energy check), else a buffer of its own
    static double *hpin_own = nullptr; static size_t hpin_own_n = 0;
    double *hpin = nullptr;
    if (L.is_open && L.h_j && (size_t)4 * n + ni <= L.h_j_n) hpin = L.h_j;
    else {
        if ((size_t)4 * n + ni > hpin_own_n) { host_free(hpin_own); hpin_own_n = (size_t)4 * n + ni + 1024; host_alloc(hpin_own, hpin_own_n); }
        hpin = hpin_own;
    }
    threaded_copy(hpin, m, (size_t)n);
    threaded_copy(hpin + n, x, (size_t)3 * n);
    const int G = (int)L.devs.size(), R = total_ranks();
    const int n_it = (ni + 31) / 32;
    Dev &root = L.devs[0];
    for (int g = 0; g < G; g++) {
        Dev &d = L.devs[g];
        set_dev(d);
        int nloc;
        shard_tiles(L.sh.on ? L.sh.rank : g, R, n, nloc);
        const int ntiles_all = (n + TJ - 1) / TJ;
        if (ntiles_all * TJ > d.pot_cap) {
            CUDA_CHECK(cudaStreamSynchronize(d.st));
            dev_free(d.pot_jraw); dev_free(d.pot_jtile); dev_free(d.pot_jidx);
            d.pot_cap = ntiles_all * TJ;
            dev_alloc(d.pot_jraw, (size_t)4 * d.pot_cap); dev_alloc(d.pot_jtile, (size_t)(ntiles_all + 2) * TILE_FLOATS);
            dev_alloc(d.pot_jidx, (size_t)(ntiles_all + 2) * TJ);        // sized for R = 1: R may change between calls
        }
        CUDA_CHECK(cudaMemcpyAsync(d.pot_jraw, hpin, sizeof(double) * 4 * n, cudaMemcpyHostToDevice, d.st));
        build_tiles(d, n, d.pot_jraw, d.pot_jraw + n, nullptr, d.pot_jtile, d.pot_jidx, L.sh.on ? L.sh.rank : g, R, nloc);
        d.perm_n = 0;                              // the sort scratch now holds gpupot's order
        int S = (d.nsm * 16 * 4) / n_it; if (S < 1) S = 1; if (S > nloc) S = nloc > 0 ? nloc : 1;
        if ((size_t)S * ni > d.pot_part_n) { CUDA_CHECK(cudaStreamSynchronize(d.st)); dev_free(d.pot_part); d.pot_part_n = (size_t)S * ni; dev_alloc(d.pot_part, d.pot_part_n); }
        if ((size_t)ni > d.pot_out_n) { CUDA_CHECK(cudaStreamSynchronize(d.st)); dev_free(d.pot_out); d.pot_out_n = ni; dev_alloc(d.pot_out, d.pot_out_n); }
        if (g == 0) CUDA_CHECK(cudaEventRecord(d.evs0, d.st));
        if (g > 0) CUDA_CHECK(cudaStreamWaitEvent(d.st, root.evdone, 0));        // the root has consumed d.pot_out of the previous call
        pot_kernel<<<(n_it * S + 3) / 4, 128, 0, d.st>>>(d.pot_jtile, nloc, S, d.pot_jraw + n, istart - 1, ni, d.pot_part);
        CUDA_CHECK(cudaGetLastError());
        pot_merge_kernel<<<(ni + 127) / 128, 128, 0, d.st>>>(d.pot_part, S, ni, d.pot_out);
        CUDA_CHECK(cudaGetLastError());
        L.ctr[GPUNB_B200_CTR_LAUNCHES] += 2;
        L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * 4.0 * n;
        if (g > 0) {
            CUDA_CHECK(cudaEventRecord(d.evdone, d.st));
            CUDA_CHECK(cudaStreamWaitEvent(root.st, d.evdone, 0));
        }
    }
    set_dev(root);
    const double *result = root.pot_out;
    if (R > 1) {
        if ((size_t)(R + 1) * ni > root.pot_gather_n) {
            CUDA_CHECK(cudaStreamSynchronize(root.st));
            dev_free(root.pot_gather); root.pot_gather_n = (size_t)(R + 1) * ni; dev_alloc(root.pot_gather, root.pot_gather_n);
        }
        PotParts pp;
        if (L.sh.on) {
            const int rc = L.sh.allgather(root.pot_out, root.pot_gather, (size_t)ni, NCCL_FLOAT64, L.sh.comm, root.st);
            if (rc != 0) FATAL("gpupot: ncclAllGather failed: %s", L.sh.errstr ? L.sh.errstr(rc) : "?");
            for (int r = 0; r < R; r++) pp.p[r] = root.pot_gather + (size_t)r * ni;
        } else {
            for (int g = 0; g < G; g++) pp.p[g] = L.devs[g].pot_out;
        }
        double *sum = root.pot_gather + (size_t)R * ni;
        pot_sum_kernel<<<(ni + 127) / 128, 128, 0, root.st>>>(pp, R, ni, sum);
        CUDA_CHECK(cudaGetLastError());
        if (!L.sh.on) CUDA_CHECK(cudaEventRecord(root.evdone, root.st));
        L.ctr[GPUNB_B200_CTR_LAUNCHES] += 1;
        result = sum;
    }
    CUDA_CHECK(cudaEventRecord(root.evs1, root.st));
    double *hout = hpin + 4 * (size_t)n;
    CUDA_CHECK(cudaMemcpyAsync(hout, result, sizeof(double) * ni, cudaMemcpyDeviceToHost, root.st));
    for (int g = G - 1; g >= 0; g--) {
        set_dev(L.devs[g]);
        CUDA_CHECK(cudaStreamSynchronize(L.devs[g].st));
        if (L.h_nan[g]) FATAL("gpupot: NaN in particle data");
    }
    float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, root.evs0, root.evs1));
    L.ctr[GPUNB_B200_CTR_POT_MS] += ms;
    L.ctr[GPUNB_B200_CTR_D2H_BYTES] += sizeof(double) * (double)ni;
    memcpy(pot, hout, sizeof(double) * ni);
    const double t1 = wtime();
    fprintf(stderr, "[R.%d GPU Pot.A] Ni %d  NTOT %d  pot(s) %f\n", irank, ni, n, t1 - t0);   // reference: gpupot.gpu.cu:112
}

// ---- NCCL mode -------------------------------------------------------------------------------
}  // namespace

// regcor_b200.cu works on the snapshot the last send left on the root device
int gpunb_b200_internal_host_team() { return host_team(); }

bool gpunb_b200_internal_snapshot(GpunbSnapshotView *out)
{
    if (!L.is_open || L.devs.empty() || L.devs[0].nj_total <= 0 || !L.devs[0].jraw) return false;
    const Dev &d = L.devs[0];
    out->device = d.id; out->stream = d.st; out->nj = d.nj_total; out->nbmax = L.nbmax;
    out->m = d.jraw; out->x = d.jraw + d.nj_total; out->v = d.jraw + 4 * (size_t)d.nj_total;
    out->counters = L.ctr;
    out->last_rows = L.last_rows_ni > 0 ? d.last_rows : nullptr; out->last_rows_ni = L.last_rows_ni; out->last_rows_lmax = L.last_rows_lmax;
    out->pinned_alias = [](const void *p, size_t bytes) -> void * { return pinned_alias(reinterpret_cast<const char *>(p), bytes); };
    return true;
}

namespace {

void nccl_load()
{
    Shard &sh = L.sh;
    if (sh.dl) return;
    sh.dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);            // the copy torch already loaded, if any
    if (!sh.dl) sh.dl = dlopen("libnccl.so.2", RTLD_NOW);
    if (!sh.dl) sh.dl = dlopen("libnccl.so", RTLD_NOW);
    if (!sh.dl) FATAL("cannot load libnccl.so.2: %s", dlerror());
    sh.getid = (pfn_ncclGetUniqueId)dlsym(sh.dl, "ncclGetUniqueId");
    sh.init = (pfn_ncclCommInitRank)dlsym(sh.dl, "ncclCommInitRank");
    sh.allgather = (pfn_ncclAllGather)dlsym(sh.dl, "ncclAllGather");
    sh.destroy = (pfn_ncclCommDestroy)dlsym(sh.dl, "ncclCommDestroy");
    sh.errstr = (pfn_ncclGetErrorString)dlsym(sh.dl, "ncclGetErrorString");
    if (!sh.getid || !sh.init || !sh.allgather || !sh.destroy) FATAL("libnccl.so.2 lacks the expected entry points");
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

void gpunb_devinit_(int *irank) { lib_devinit(*irank); }
void gpunb_open_(int *nbmax, int *irank) { lib_open(*nbmax, *irank); }
void gpunb_close_(void) { lib_close(); }
void gpunb_send_(int *nj, double mj[], double xj[][3], double vj[][3]) { lib_send(*nj, mj, &xj[0][0], &vj[0][0]); }
void gpunb_regf_(int *ni, double h2[], double dtr[], double xi[][3], double vi[][3], double acc[][3], double jrk[][3],
                 double pot[], int *lmax, int *nnbmax, int *list, int *m_flag)
{
    lib_regf(*ni, h2, dtr, &xi[0][0], &vi[0][0], &acc[0][0], &jrk[0][0], pot, *lmax, *nnbmax, list, *m_flag);
}
void gpunb_profile_(int *irank) { lib_profile(*irank); }
void gpupot_(int *irank, int *istart, int *ni, int *n, double m[], double x[][3], double pot[])
{
    lib_pot(*irank, *istart, *ni, *n, m, &x[0][0], pot);
}

int gpunb_b200_version(void) { return 200; }
int gpunb_b200_has_near_scalar_ab(void)
{
#ifdef NEAR_SCALAR_AB
    return 1;
#else
    return 0;
#endif
}
const char *gpunb_b200_build_info(void)
{
    return "gpunb_b200 sm_100a: regf_kernel<TMA bulk Hilbert tiles TJ=64, packed f32x2 FAR body, scalar float-float NEAR body>, isort, merge (register bitonic), combine (NVLink peer pulls + flags), pot, tilepack";
}
int gpunb_b200_num_devices(void) { return (int)L.devs.size(); }
int gpunb_b200_resident_warps(void) { return L.devs.empty() ? 0 : L.devs[0].warps_resident; }
void gpunb_b200_get_counters(double out[GPUNB_B200_CTR_COUNT])
{
    if (L.devinit && !L.devs.empty() && L.devs[0].stats) {
        Dev &d = L.devs[0];
        set_dev(d);
        unsigned long long h[4];
        CUDA_CHECK(cudaMemcpyAsync(h, d.stats, 32, cudaMemcpyDeviceToHost, d.st));
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        L.ctr[GPUNB_B200_CTR_NEAR_TILES] = (double)h[0];
        L.ctr[GPUNB_B200_CTR_ALL_TILES] = (double)h[1];
        L.ctr[GPUNB_B200_CTR_TRANSPOSED_TILES] = (double)h[2];

    }
    for (int k = 0; k < GPUNB_B200_CTR_COUNT; k++) out[k] = L.ctr[k];
}
void gpunb_b200_reset_counters(void)
{
    for (int k = 0; k < GPUNB_B200_CTR_COUNT; k++) L.ctr[k] = 0;
    if (L.devinit && !L.devs.empty() && L.devs[0].stats) {
        set_dev(L.devs[0]);
        CUDA_CHECK(cudaMemsetAsync(L.devs[0].stats, 0, 32, L.devs[0].st));
    }
}

void gpunb_b200_set_radii(int *njp, double h2[], double dtr[])
{
    if (!L.is_open) FATAL("gpunb_b200_set_radii: library closed");
    const int nj = *njp;
    for (Dev &d : L.devs) {
        if (nj != d.nj_total) FATAL("gpunb_b200_set_radii: nj=%d differs from the snapshot (%d)", nj, d.nj_total);
        set_dev(d);
        CUDA_CHECK(cudaMemcpyAsync(d.radii, h2, sizeof(double) * nj, cudaMemcpyHostToDevice, d.st));
        CUDA_CHECK(cudaMemcpyAsync(d.radii + d.raw_cap, dtr, sizeof(double) * nj, cudaMemcpyHostToDevice, d.st));
        CUDA_CHECK(cudaStreamSynchronize(d.st));
    }
}

float gpunb_b200_sweep_resident(int *i0p, int *nip, int *blockp, int *lmaxp, int *nnbmaxp, int *m_flagp)
{
    if (!L.is_open) FATAL("gpunb_b200_sweep_resident: library closed");
    Dev &root = L.devs[0];
    const int G = (int)L.devs.size();
    const int i0 = *i0p, ni = *nip, block = *blockp;
    if (i0 < 0 || i0 + ni > root.nj_total || block < 1 || block > JOBCAP) FATAL("gpunb_b200_sweep_resident: bad range");
    if (block > NIMAX && G > 1) FATAL("gpunb_b200_sweep_resident: blocks of more than %d i-particles need one process per GPU", NIMAX);
    static int timeline = -1;
    if (timeline < 0) { const char *e = getenv("GPUNB_B200_TIMELINE"); timeline = (e && atoi(e) > 0 && L.devs.size() == 1) ? 1 : 0; }
    // Pipelined (default, one GPU per process): every block's Morton order comes from ONE batched isort launch, then
    // the blocks cycle through nslot pipeline slots (see Slot).  Sequential (GPUNB_B200_NSLOT=1, the per-kernel
    // timeline, or one process driving several GPUs): one block after the other on the main stream.
    // measured (profiles/r2c_sweep_probe.txt): with 2 slots the stagger of consecutive launches can collapse (945 ... 1042
    // Gint/s between runs of the same sweep), 3 and 4 slots hold 1035-1039 on one GPU and hide the exchange step when sharded
    const int nslot = (G == 1 && !timeline) ? L.nslot : 1;
    const bool pipelined = nslot > 1;
    for (int g = 0; g < G; g++) ensure_work_buffers(L.devs[g], *lmaxp, *nnbmaxp, g == 0, nslot, true, block);
    set_dev(root);
    const int nblocks = (ni + block - 1) / block;
    static std::vector<cudaEvent_t> tlev;
    if (timeline && (int)tlev.size() < 5 * nblocks) {
        const size_t old = tlev.size();
        tlev.resize((size_t)5 * nblocks);
        for (size_t k = old; k < tlev.size(); k++) CUDA_CHECK(cudaEventCreate(&tlev[k]));
    }
    Job j;
    j.lmax = *lmaxp; j.nnbmax = *nnbmaxp; j.m_flag = *m_flagp; j.slot0 = 0;
    j.oversub = root.oversub;
    auto iblock_of = [&](Dev &d, int b0) {
        const double *x = d.jraw + d.nj_total, *v = d.jraw + 4 * (size_t)d.nj_total;
        return IBlock{d.radii + b0, d.radii + d.raw_cap + b0, x + 3 * (size_t)b0, v + 3 * (size_t)b0};
    };
    CUDA_CHECK(cudaEventRecord(root.evs0, root.st));
    int nlaunch = 0, last = 0;
    if (pipelined) {
        if ((size_t)ni > root.iperm_all_n) {
            CUDA_CHECK(cudaDeviceSynchronize());
            dev_free(root.iperm_all);
            root.iperm_all_n = (size_t)ni + 4096;
            dev_alloc(root.iperm_all, root.iperm_all_n);
        }
        launch_isort(root, root.st, ni, block, root.jraw + root.nj_total + 3 * (size_t)i0, root.iperm_all, nullptr);
        CUDA_CHECK(cudaEventRecord(root.ev_fork, root.st));
        for (int q = 0; q < nslot && q < nblocks; q++) CUDA_CHECK(cudaStreamWaitEvent(root.slots[q].lo, root.ev_fork, 0));
        for (int b = 0; b < nblocks; b++) {
            const int q = b % nslot;
            Slot &sl = root.slots[q];
            const int b0 = i0 + b * block;
            const int n = (i0 + ni - b0 < block) ? i0 + ni - b0 : block;
            // the slot's buffers are free once merge / combine of its previous block are done
            if (sl.used) CUDA_CHECK(cudaStreamWaitEvent(sl.lo, sl.ev_done, 0));
            IBlock ib[1] = {iblock_of(root, b0)};
            const int *ipm[1] = {root.iperm_all + (size_t)b * block};
            j.nloc = n; j.out_f = sl.res_f; j.out_list = sl.res_list;
            run_job(j, ib, ipm, q, true, false);
            L.ctr[GPUNB_B200_CTR_INTERACTIONS] += (double)n * root.nj_total;
            nlaunch++; last = n; L.last_slot = q;
        }
        for (int q = 0; q < nslot && q < nblocks; q++) CUDA_CHECK(cudaStreamWaitEvent(root.st, root.slots[q].ev_done, 0));
    } else {
        for (int b0 = i0; b0 < i0 + ni; b0 += block) {
            const int n = (i0 + ni - b0 < block) ? i0 + ni - b0 : block;
            IBlock ib[MAX_RANKS];
            const int *ipm[MAX_RANKS];
            cudaEvent_t *tl = timeline ? &tlev[(size_t)5 * nlaunch] : nullptr;
            if (tl) CUDA_CHECK(cudaEventRecord(tl[0], root.st));
            for (int g = 0; g < G; g++) {
                Dev &d = L.devs[g];
                set_dev(d);
                ib[g] = iblock_of(d, b0);
                launch_isort(d, d.st, n, n, ib[g].xi, d.iperm, nullptr);
                ipm[g] = d.iperm;
            }
            set_dev(root);
            if (tl) CUDA_CHECK(cudaEventRecord(tl[1], root.st));
            j.nloc = n; j.out_f = root.slots[0].res_f; j.out_list = root.slots[0].res_list;
            run_job(j, ib, ipm, 0, false, false, tl);
            L.ctr[GPUNB_B200_CTR_INTERACTIONS] += (double)n * root.nj_total;
            nlaunch++; last = n; L.last_slot = 0;
        }
    }
    set_dev(root);
    CUDA_CHECK(cudaEventRecord(root.evs1, root.st));
    CUDA_CHECK(cudaStreamSynchronize(root.st));
    float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, root.evs0, root.evs1));
    if (timeline) {
        for (int k = 0; k < nlaunch; k++) {
            float t[4];
            for (int q = 0; q < 4; q++) CUDA_CHECK(cudaEventElapsedTime(&t[q], tlev[(size_t)5 * k + q], tlev[(size_t)5 * k + q + 1]));
            L.ctr[GPUNB_B200_CTR_TL_ISORT_MS] += t[0]; L.ctr[GPUNB_B200_CTR_TL_REGF_MS] += t[1];
            L.ctr[GPUNB_B200_CTR_TL_MERGE_MS] += t[2]; L.ctr[GPUNB_B200_CTR_TL_EXCH_MS] += t[3];
        }
        L.ctr[GPUNB_B200_CTR_TL_BLOCKS] += nlaunch;
    }
    L.ctr[GPUNB_B200_CTR_GRAV_LAUNCHES] += nlaunch;
    L.last_ni = last; L.last_lmax = *lmaxp; L.last_on_host = false;
    return ms;
}

void gpunb_b200_fetch_last(int *n_last, double acc[][3], double jrk[][3], double pot[], int *lmaxp, int *list)
{
    const int ni = L.last_ni, lmax = L.last_lmax;
    *n_last = ni; *lmaxp = lmax;
    if (ni <= 0) return;
    if (ni > NIMAX) FATAL("gpunb_b200_fetch_last: the last block holds %d rows, the host buffers %d", ni, NIMAX);
    if (!L.last_on_host) {         // a resident sweep leaves its last block in the slot's device buffers
        Dev &root = L.devs[0];
        set_dev(root);
        Slot &sl = root.slots[L.last_slot];
        CUDA_CHECK(cudaMemcpyAsync(L.h_f, sl.res_f, sizeof(double) * 7 * ni, cudaMemcpyDeviceToHost, root.st));
        CUDA_CHECK(cudaMemcpyAsync(L.h_list, sl.res_list, sizeof(int) * (size_t)ni * lmax, cudaMemcpyDeviceToHost, root.st));
        CUDA_CHECK(cudaStreamSynchronize(root.st));
        L.last_on_host = true;
    }
    scatter_rows(nullptr, 0, ni, lmax, &acc[0][0], &jrk[0][0], pot, list);
}

void gpunb_b200_state_all_(int *nj, double body[], double x0[][3], double x0dot[][3], double f[][3], double fdot[][3], double t0[])
{
    lib_state_all(*nj, body, &x0[0][0], &x0dot[0][0], &f[0][0], &fdot[0][0], t0);
}
void gpunb_b200_state_update_(int *n, int idx[], double body[], double x0[][3], double x0dot[][3], double f[][3],
                              double fdot[][3], double t0[])
{
    lib_state_update(*n, idx, body, &x0[0][0], &x0dot[0][0], &f[0][0], &fdot[0][0], t0);
}
void gpunb_b200_predict_send_(int *nj, double *time) { lib_predict_send(*nj, *time); }
void gpunb_b200_predict_send_records_(int *nj, double *time, const double *records_dev, int *stride) { lib_predict_send_records(*nj, *time, records_dev, *stride); }
void gpunb_b200_get_predicted_(int *n, int idx[], double x[][3], double xdot[][3]) { lib_get_predicted(*n, idx, &x[0][0], &xdot[0][0]); }

void gpunb_b200_set_near_exact(int on) { L.near_exact = on; }
int gpunb_b200_pin_host_(void *ptr, long long *bytes)
{
    if (!L.devinit) FATAL("gpunb_b200_pin_host_ before gpunb_devinit_");
    if (!ptr || *bytes <= 0) return 1;
    set_dev(L.devs[0]);
    const cudaError_t e = cudaHostRegister(ptr, (size_t)*bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        fprintf(stderr, "gpunb_b200: cannot pin %lld bytes at %p (%s); the staged path stays in use\n", *bytes, ptr, cudaGetErrorString(e));
        return 2;
    }
    g_pinned.push_back(PinnedRange{reinterpret_cast<const char *>(ptr), (size_t)*bytes});
    return 0;
}
void gpunb_b200_unpin_host_(void *ptr)
{
    for (size_t k = 0; k < g_pinned.size(); k++)
        if (g_pinned[k].base == reinterpret_cast<const char *>(ptr)) {
            for (Dev &d : L.devs) { set_dev(d); CUDA_CHECK(cudaDeviceSynchronize()); }
            CUDA_CHECK(cudaHostUnregister(ptr));
            g_pinned.erase(g_pinned.begin() + (long)k);
            return;
        }
}
void gpunb_b200_set_taper(int on) { L.taper = on != 0; }
void gpunb_b200_set_send_scatter(int min_nj) { L.send_scatter_min = min_nj; }
void gpunb_b200_set_regf_oversub(int k) { if (k >= 1 && k <= REGF_OVERSUB) L.regf_oversub = k; }
void gpunb_b200_set_sub_pairs(double pairs) { if (pairs >= 1.0) L.sub_pairs = pairs; }
void gpunb_b200_set_isort_pairs(double pairs) { if (pairs >= 0.0) L.isort_pairs = pairs; }
void gpunb_b200_set_islice(int on)
{
    if (on && !L.sh.on) FATAL("gpunb_b200_set_islice: the i-slice mode needs one process per GPU (gpunb_b200_nccl_init)");
    L.sh.islice = on != 0;
}
void gpunb_b200_set_resort_every(int k) { if (k >= 0) { L.resort_every = k; L.snapshots_since_sort = 0; L.q_ref = 0.0; } }

void gpunb_b200_set_tuning(int nslot, int nsub)
{
    if (nslot >= 1 && nslot <= MAX_SLOTS) { L.nslot = nslot; L.nslot_auto = false; }
    if (nsub >= 1 && nsub <= MAX_SLOTS) { L.nsub = nsub; L.nsub_forced = false; }
    if (nsub <= -1 && nsub >= -MAX_SLOTS) { L.nsub = -nsub; L.nsub_forced = true; }
}

// tuning aid: per-work-item start/end timestamps (ns) of the LAST regf_kernel launch (GPUNB_B200_STATS=2)
int gpunb_b200_debug_wtimes(unsigned long long *out, int max_items)
{
    if (!L.devinit || L.devs.empty() || !L.devs[0].wtime) return 0;
    Dev &d = L.devs[0];
    set_dev(d);
    const int n = max_items < 65536 ? max_items : 65536;
    CUDA_CHECK(cudaMemcpyAsync(out, d.wtime, sizeof(unsigned long long) * 3 * n, cudaMemcpyDeviceToHost, d.st));
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    return n;
}

int gpunb_b200_nccl_unique_id(unsigned char id128[128])
{
    nccl_load();
    ncclUniqueId id;
    const int rc = L.sh.getid(&id);
    if (rc != 0) return rc;
    memcpy(id128, id.internal, 128);
    return 0;
}

// Join `nranks` processes (one GPU each) into one j-sharded force library.  Call after gpunb_devinit_
// and before gpunb_open_.  Every rank then makes IDENTICAL calls (same snapshot, same i-blocks) and every
// rank receives the complete result; rank r sums over the tiles r, r+R, r+2R, ... of the Hilbert-sorted j-set (shard_tiles).
int gpunb_b200_nccl_init(int rank, int nranks, const unsigned char id128[128])
{
    if (!L.devinit) FATAL("gpunb_b200_nccl_init before gpunb_devinit_");
    if (L.devs.size() != 1) FATAL("NCCL mode drives one GPU per process (unset GPUNB_B200_MULTI)");
    if (nranks < 1 || nranks > MAX_RANKS) FATAL("nranks=%d outside 1..%d", nranks, MAX_RANKS);
    if (L.is_open) FATAL("gpunb_b200_nccl_init while the library is open");
    nccl_load();
    Shard &sh = L.sh;
    Dev &d = L.devs[0];
    set_dev(d);
    ncclUniqueId id; memcpy(id.internal, id128, 128);
    int rc = sh.init(&sh.comm, nranks, id, rank);
    if (rc != 0) FATAL("ncclCommInitRank failed: %s", sh.errstr ? sh.errstr(rc) : "?");
    sh.rank = rank; sh.R = nranks; sh.seq = 0;
    { const char *e = getenv("GPUNB_B200_SPIN_TIMEOUT_S");
      if (e && atof(e) > 0) { const unsigned long long ns = (unsigned long long)(atof(e) * 1e9); CUDA_CHECK(cudaMemcpyToSymbol(g_spin_timeout_ns, &ns, sizeof(ns))); } }
    for (Slot &sl : d.slots) if (sl.done_ctr) CUDA_CHECK(cudaMemsetAsync(sl.done_ctr, 0, 2 * sizeof(unsigned), d.st));
    CUDA_CHECK(cudaMalloc((void **)&sh.xbuf, XB_BYTES));
    CUDA_CHECK(cudaMemsetAsync(sh.xbuf, 0, XB_BYTES, d.st));
    // exchange cudaIpc handles of the exchange buffers with an all-gather (bootstrap only), then map every peer's
    cudaIpcMemHandle_t mine;
    CUDA_CHECK(cudaIpcGetMemHandle(&mine, sh.xbuf));
    unsigned char *hbuf = nullptr; dev_alloc(hbuf, (size_t)(nranks + 1) * sizeof(cudaIpcMemHandle_t));
    CUDA_CHECK(cudaMemcpyAsync(hbuf + (size_t)nranks * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, d.st));
    rc = sh.allgather(hbuf + (size_t)nranks * sizeof(mine), hbuf, sizeof(mine), NCCL_INT8, sh.comm, d.st);
    if (rc != 0) FATAL("ncclAllGather(ipc handles) failed: %s", sh.errstr ? sh.errstr(rc) : "?");
    std::vector<cudaIpcMemHandle_t> all(nranks);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), hbuf, (size_t)nranks * sizeof(mine), cudaMemcpyDeviceToHost, d.st));
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    dev_free(hbuf);
    for (int r = 0; r < nranks; r++) {
        if (r == rank) { sh.xbuf_peer[r] = sh.xbuf; continue; }
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) FATAL("cudaIpcOpenMemHandle(rank %d) failed: %s (NVLink P2P between ranks is required)", r, cudaGetErrorString(e));
        sh.xbuf_peer[r] = (unsigned char *)p;
    }
    // shared-memory segment of the node for the i-slice rendezvous (named after the unique id; unlinked once every
    // rank has mapped it, so nothing is left behind if a rank dies)
    char shm_name[64];
    {
        unsigned long long hsh = 1469598103934665603ull;
        for (int k = 0; k < 128; k++) hsh = (hsh ^ id128[k]) * 1099511628211ull;
        snprintf(shm_name, sizeof(shm_name), "/gpunb_b200_%016llx", hsh);
        const int fd = shm_open(shm_name, O_CREAT | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)sizeof(ShmSeg)) != 0) FATAL("cannot create the shared-memory segment %s of the i-slice mode", shm_name);
        void *m = mmap(nullptr, sizeof(ShmSeg), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (m == MAP_FAILED) FATAL("cannot map the shared-memory segment %s", shm_name);
        sh.shm = reinterpret_cast<ShmSeg *>(m);
        sh.icall = 0;
        const char *e = getenv("GPUNB_B200_ISLICE");
        sh.islice = e && atoi(e) > 0;
    }
    // every rank has zeroed its flags and mapped its peers before anybody can signal: one more all-gather as barrier
    dev_alloc(sh.scratch, 2 * (size_t)MAX_RANKS);
    rc = sh.allgather(sh.scratch + MAX_RANKS, sh.scratch, 1, NCCL_FLOAT64, sh.comm, d.st);
    if (rc != 0) FATAL("ncclAllGather(barrier) failed: %s", sh.errstr ? sh.errstr(rc) : "?");
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    if (rank == 0) shm_unlink(shm_name);
    sh.on = true;
    // R processes share the node's cores: the host-side teams shrink accordingly (never below the 4 threads a staging copy
    // needs to keep a PCIe link busy) unless GPUNB_B200_HOST_THREADS fixed their size
    if (!getenv("GPUNB_B200_HOST_THREADS")) L.host_threads = std::max(1, std::min(L.host_threads, std::max(4, omp_get_max_threads() / nranks)));
    return 0;
}

void gpunb_b200_nccl_finalize(void)
{
    Shard &sh = L.sh;
    if (!sh.on) return;
    Dev &d = L.devs[0];
    set_dev(d);
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    // nobody unmaps or frees while a peer may still be pulling: all-gather as barrier, before and after the unmap
    sh.allgather(sh.scratch + MAX_RANKS, sh.scratch, 1, NCCL_FLOAT64, sh.comm, d.st);
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    for (int r = 0; r < sh.R; r++)
        if (r != sh.rank && sh.xbuf_peer[r]) { CUDA_CHECK(cudaIpcCloseMemHandle(sh.xbuf_peer[r])); sh.xbuf_peer[r] = nullptr; }
    sh.allgather(sh.scratch + MAX_RANKS, sh.scratch, 1, NCCL_FLOAT64, sh.comm, d.st);
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    dev_free(sh.scratch);
    CUDA_CHECK(cudaFree(sh.xbuf)); sh.xbuf = nullptr;
    sh.destroy(sh.comm);
    if (sh.shm) { munmap(sh.shm, sizeof(ShmSeg)); sh.shm = nullptr; }
    sh.islice = false;
    sh.comm = nullptr; sh.on = false; sh.R = 1; sh.rank = 0;
}

}  // extern "C"
