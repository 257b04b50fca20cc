// gpunb_b200.cu -- B200-native (sm_100a) regular-force library for NBODY6++GPU.
//
// Drop-in for the reference's gpunb.velocity.cu / gpupot.gpu.cu behind the same
// Fortran-callable C-ABI (include/gpunb_b200.h).  Written from scratch; the
// reference (file:line cited below) defines WHAT is computed, not how.
//
// Data layout in HBM (per device)
//   jraw   : the caller's fp64 snapshot, m[nj] | x[nj][3] | v[nj][3]            56 B/j
//   jtile  : AoSoA tiles of TJ=64 j-particles, 10 float arrays per tile
//            {xh,yh,zh, xl,yl,zl, vx,vy,vz, m}[64]  = 2560 B/tile                40 B/j
//            xh=(float)x, xl=(float)(x-xh): float-float positions, so that close
//            pairs keep ~48 bits in dx; xh alone is what the reference's FP32
//            cast sees (gpunb.velocity.cu:62-64), so its neighbour predicate can
//            be evaluated bit-for-bit.  One tile = ONE 1-D TMA bulk copy.
//   part   : per (j-slice s, i) partial sums, 7 doubles, + count               [S][ni]
//   seg    : per (i, s) neighbour-index segment, capacity segcap ints
//   res_f  : per i  acc[3] jrk[3] pot  (fp64)   ; res_list : [ni][lmax] int32 (the ABI layout)
//
// Kernels
//   jpack_kernel   fp64 snapshot -> jtile (+ NaN check, reference asserts: gpunb.velocity.cu:72-78)
//   regf_kernel    the O(ni*nj) pair kernel.  One WARP = one work item (i-tile of 32*IT
//                  i-particles, contiguous range of j-tiles).  j-tiles are staged through
//                  warp-private shared memory by TMA bulk copies (cp.async.bulk + mbarrier,
//                  double buffered); lanes own i-particles, j is broadcast from smem; two
//                  j-particles are processed per instruction with packed f32x2 FMA/ADD/MUL.
//                  FP32 chains are 128 terms long, then flushed into fp64 accumulators.
//   merge_kernel   per i: fp64 sum of the S partials in fixed order, exclusive scan of the S
//                  segment counts, concatenation in slice order (= ascending j), overflow
//                  encoding -(count) (reg.avx.cpp:320-321).
//   pot_kernel     gpupot: float-float dx, FP32 rsqrt + one Newton step, fp64 flush.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <vector>
#include <sys/time.h>
#include <unistd.h>
#include <dlfcn.h>
#include "../../include/gpunb_b200.h"

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "gpunb_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, \
            cudaGetErrorString(e_)); abort(); } } while (0)
#define FATAL(...) do { fprintf(stderr, "gpunb_b200: " __VA_ARGS__); fprintf(stderr, "\n"); abort(); } while (0)

namespace {

constexpr int TJ          = 64;               // j-particles per tile
constexpr int NCOMP       = 10;               // float arrays per tile
constexpr int TILE_FLOATS = TJ * NCOMP;       // 640
constexpr int TILE_BYTES  = TILE_FLOATS * 4;  // 2560
constexpr int NSTAGE      = 2;                // smem stages per warp
constexpr int WARPS       = 4;                // warps per CTA (warp-autonomous: no CTA-wide sync)
constexpr int ITILE_MAX   = 64;               // largest i-tile of any kernel variant (32 lanes * IT)
constexpr int FLUSH_TILES = 1;                // FP32 chains: 1 tile * 64 j / 2 (f32x2 halves) = 32 terms, then fp64
constexpr int NIMAX       = 2048;             // capacity per call (reference: gpunb.velocity.cu:24)
constexpr int PART_STRIDE = 8;                // doubles per partial record (7 used)

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
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

// ---------------------------------------------------------------------------------------------
// jpack_kernel: fp64 snapshot -> float-float AoSoA tiles.  Ghost slots (j >= nj) get mass 0 and a
// far-away position; they are additionally excluded from lists by index.
// ---------------------------------------------------------------------------------------------
__global__ void jpack_kernel(int nj, int ntiles, const double *__restrict__ m, const double *__restrict__ x,
                             const double *__restrict__ v, float *__restrict__ tiles, int *__restrict__ nanflag)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ntiles * TJ) return;
    float *t = tiles + (size_t)(j / TJ) * TILE_FLOATS + (j % TJ);
    if (j < nj) {
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double xd = x[3 * (size_t)j + k], vd = v[3 * (size_t)j + k];
            float hi = (float)xd;
            float lo = (float)(xd - (double)hi);
            t[(0 + k) * TJ] = hi;
            t[(3 + k) * TJ] = lo;
            t[(6 + k) * TJ] = (float)vd;
            bad |= (xd != xd) | (vd != vd);
        }
        double md = m[j];
        t[9 * TJ] = (float)md;
        bad |= (md != md);
        if (bad) atomicExch(nanflag, 1);
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) { t[k * TJ] = 1.0e6f; t[(3 + k) * TJ] = 0.f; t[(6 + k) * TJ] = 0.f; }
        t[9 * TJ] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// regf_kernel
// ---------------------------------------------------------------------------------------------
struct RegfArgs {
    const float  *tiles;     // jtile
    int           ntiles, nj;
    int           joff;      // global index of this device's first j (multi-GPU shards)
    // i-particles (fp64, device): element i at h2[i], dtr[i], xi[3i..], vi[3i..]
    const double *h2, *dtr, *xi, *vi;
    int           ni, n_itiles, S, n_items;
    double       *part;      // [S][ni][PART_STRIDE]
    int          *cnt;       // [S][ni]
    int          *seg;       // [ni][S][segcap]
    int           segcap;
    int           flush_tiles;   // FP32 chains are flushed into fp64 every flush_tiles j-tiles
};

struct IState {               // loop invariants of one i-particle, duplicated for f32x2 operands
    float2 nxh, nyh, nzh;     // -x_i (hi)
    float2 nxl, nyl, nzl;     // -x_i (lo)
    float2 nvx, nvy, nvz;     // -v_i
    float2 dtr, h2;
};
struct Acc {                  // FP32 partial chains (f32x2: one chain per j parity)
    float2 ax, ay, az, p, jx, jy, jz;
    __device__ __forceinline__ void clear() { ax = ay = az = p = jx = jy = jz = make_float2(0.f, 0.f); }
};

// One i-particle against a packed pair of j-particles.
//   Predicate: the reference's, bit for bit, on the fp32-rounded inputs (gpunb.velocity.cu:168-187,
//   :235 for m_flag): r2 = fma(dz,dz,fma(dy,dy,dx*dx)), dxp = fma(dtr,dvx,dx), min(r2,r2p) < h2 [*mj].
//   Force: same formula (gpunb.velocity.cu:192-207) but from the float-float dx and a Newton-refined rsqrt.
//   Pairs at r2 == 0 (self) never contribute (regint.f:40 skips J.EQ.I); the reference GPU code
//   returns NaN for a self pair with h2 == 0.
// Returns a 2-bit mask of neighbour hits.
template <bool MFLAG>
__device__ __forceinline__ unsigned interact(const IState &I, Acc &A,
                                             float2 XH, float2 YH, float2 ZH, float2 XL, float2 YL, float2 ZL,
                                             float2 VX, float2 VY, float2 VZ, float2 M)
{
    const float2 dxr = add2(XH, I.nxh), dyr = add2(YH, I.nyh), dzr = add2(ZH, I.nzh);
    const float2 dx = add2(dxr, add2(XL, I.nxl));
    const float2 dy = add2(dyr, add2(YL, I.nyl));
    const float2 dz = add2(dzr, add2(ZL, I.nzl));
    const float2 dvx = add2(VX, I.nvx), dvy = add2(VY, I.nvy), dvz = add2(VZ, I.nvz);

    const float2 r2r = fma2(dzr, dzr, fma2(dyr, dyr, mul2(dxr, dxr)));
    const float2 dxp = fma2(I.dtr, dvx, dxr), dyp = fma2(I.dtr, dvy, dyr), dzp = fma2(I.dtr, dvz, dzr);
    const float2 r2p = fma2(dzp, dzp, fma2(dyp, dyp, mul2(dxp, dxp)));
    const float2 r2  = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
    const float2 rv  = fma2(dz, dvz, fma2(dy, dvy, mul2(dx, dvx)));

    float2 lim = I.h2;
    if (MFLAG) lim = mul2(M, I.h2);
    const bool nb0 = fminf(r2r.x, r2p.x) < lim.x;
    const bool nb1 = fminf(r2r.y, r2p.y) < lim.y;
    float2 rinv;
    rinv.x = (nb0 || !(r2.x > 0.f)) ? 0.f : rsqrt_approx(r2.x);
    rinv.y = (nb1 || !(r2.y > 0.f)) ? 0.f : rsqrt_approx(r2.y);

    // one Newton step: y <- y - y/2 (r2 y^2 - 1); rsqrt.approx alone (2^-22.9) leaves 5 ulp on the r^-5 term
    const float2 e = fma2(mul2(r2, rinv), rinv, dup2(-1.f));
    rinv = fma2(mul2(rinv, e), dup2(-0.5f), rinv);
    const float2 rinv2  = mul2(rinv, rinv);
    const float2 mrinv  = mul2(M, rinv);
    const float2 mrinv3 = mul2(mrinv, rinv2);
    const float2 rv3    = mul2(rv, mul2(rinv2, dup2(-3.f)));          // -3 (r.v)/r^2
    A.p  = add2(A.p, mrinv);
    A.ax = fma2(mrinv3, dx, A.ax);   A.ay = fma2(mrinv3, dy, A.ay);   A.az = fma2(mrinv3, dz, A.az);
    A.jx = fma2(mrinv3, fma2(rv3, dx, dvx), A.jx);
    A.jy = fma2(mrinv3, fma2(rv3, dy, dvy), A.jy);
    A.jz = fma2(mrinv3, fma2(rv3, dz, dvz), A.jz);
    return (nb0 ? 1u : 0u) | (nb1 ? 2u : 0u);
}

template <int IT, bool MFLAG, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) regf_kernel(const RegfArgs a)
{
    constexpr int ITILE = 32 * IT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * WARPS + warp;
    if (w >= a.n_items) return;                       // warp-uniform; no CTA-wide barrier is used below
    float    *buf  = reinterpret_cast<float *>(smem_raw) + warp * NSTAGE * TILE_FLOATS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + WARPS * NSTAGE * TILE_BYTES) + warp * NSTAGE;

    const int it = w / a.S, s = w - it * a.S;
    const int t0 = (int)(((long long)s * a.ntiles) / a.S);
    const int t1 = (int)(((long long)(s + 1) * a.ntiles) / a.S);

    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NSTAGE; k++) mbar_init(&bars[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NSTAGE; k++)
            if (t0 + k < t1) {
                mbar_expect_tx(&bars[k], TILE_BYTES);
                tma_bulk_g2s(buf + k * TILE_FLOATS, a.tiles + (size_t)(t0 + k) * TILE_FLOATS, TILE_BYTES, &bars[k]);
            }
    }

    // i-particles of this lane
    IState I[IT];
    Acc    A[IT];
    double D[IT][7];
    int    cnt[IT];
    int   *segp[IT];
    bool   valid[IT];
#pragma unroll
    for (int k = 0; k < IT; k++) {
        const int i = it * ITILE + k * 32 + lane;
        valid[k] = i < a.ni;
        double x[3] = {0, 0, 0}, v[3] = {0, 0, 0}, h2 = 0, dtr = 0;
        if (valid[k]) {
            h2 = a.h2[i]; dtr = a.dtr[i];
#pragma unroll
            for (int c = 0; c < 3; c++) { x[c] = a.xi[3 * (size_t)i + c]; v[c] = a.vi[3 * (size_t)i + c]; }
        }
        float xh[3], xl[3];
#pragma unroll
        for (int c = 0; c < 3; c++) { xh[c] = (float)x[c]; xl[c] = (float)(x[c] - (double)xh[c]); }
        I[k].nxh = dup2(-xh[0]); I[k].nyh = dup2(-xh[1]); I[k].nzh = dup2(-xh[2]);
        I[k].nxl = dup2(-xl[0]); I[k].nyl = dup2(-xl[1]); I[k].nzl = dup2(-xl[2]);
        I[k].nvx = dup2(-(float)v[0]); I[k].nvy = dup2(-(float)v[1]); I[k].nvz = dup2(-(float)v[2]);
        I[k].dtr = dup2((float)dtr);
        I[k].h2  = dup2(valid[k] ? (float)h2 : 0.f);
        A[k].clear();
#pragma unroll
        for (int c = 0; c < 7; c++) D[k][c] = 0.0;
        cnt[k]  = 0;
        segp[k] = a.seg + ((size_t)(valid[k] ? i : 0) * a.S + s) * a.segcap;
    }

    auto flush = [&]() {
#pragma unroll
        for (int k = 0; k < IT; k++) {
            D[k][0] += (double)(A[k].ax.x + A[k].ax.y);
            D[k][1] += (double)(A[k].ay.x + A[k].ay.y);
            D[k][2] += (double)(A[k].az.x + A[k].az.y);
            D[k][3] += (double)(A[k].jx.x + A[k].jx.y);
            D[k][4] += (double)(A[k].jy.x + A[k].jy.y);
            D[k][5] += (double)(A[k].jz.x + A[k].jz.y);
            D[k][6] += (double)(A[k].p.x + A[k].p.y);
            A[k].clear();
        }
    };

    int since_flush = 0;
    for (int t = t0; t < t1; t++) {
        const int st = (t - t0) % NSTAGE;
        const uint32_t parity = ((t - t0) / NSTAGE) & 1;
        mbar_wait(&bars[st], parity);
        const float4 *c = reinterpret_cast<const float4 *>(buf + st * TILE_FLOATS);
        const int jtile0 = t * TJ;
#pragma unroll 1
        for (int q = 0; q < TJ / 4; q++) {
            const float4 XH = c[0 * 16 + q], YH = c[1 * 16 + q], ZH = c[2 * 16 + q];
            const float4 XL = c[3 * 16 + q], YL = c[4 * 16 + q], ZL = c[5 * 16 + q];
            const float4 VX = c[6 * 16 + q], VY = c[7 * 16 + q], VZ = c[8 * 16 + q];
            const float4 M  = c[9 * 16 + q];
            unsigned hit = 0;
#pragma unroll
            for (int k = 0; k < IT; k++) {
                const unsigned h0 = interact<MFLAG>(I[k], A[k],
                    make_float2(XH.x, XH.y), make_float2(YH.x, YH.y), make_float2(ZH.x, ZH.y),
                    make_float2(XL.x, XL.y), make_float2(YL.x, YL.y), make_float2(ZL.x, ZL.y),
                    make_float2(VX.x, VX.y), make_float2(VY.x, VY.y), make_float2(VZ.x, VZ.y),
                    make_float2(M.x, M.y));
                const unsigned h1 = interact<MFLAG>(I[k], A[k],
                    make_float2(XH.z, XH.w), make_float2(YH.z, YH.w), make_float2(ZH.z, ZH.w),
                    make_float2(XL.z, XL.w), make_float2(YL.z, YL.w), make_float2(ZL.z, ZL.w),
                    make_float2(VX.z, VX.w), make_float2(VY.z, VY.w), make_float2(VZ.z, VZ.w),
                    make_float2(M.z, M.w));
                hit |= (h0 | (h1 << 2)) << (4 * k);
            }
            if (hit) {                                 // rare: ~2e-4 of pairs are neighbours
                const int jb = jtile0 + q * 4;
#pragma unroll
                for (int k = 0; k < IT; k++) {
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        if ((hit >> (4 * k + b)) & 1u) {
                            const int j = jb + b;
                            if (j < a.nj) {            // ghosts of the last tile are never neighbours
                                if (cnt[k] < a.segcap) segp[k][cnt[k]] = a.joff + j;
                                cnt[k]++;
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();                                  // every lane is done reading stage st
        if (lane == 0 && t + NSTAGE < t1) {
            mbar_expect_tx(&bars[st], TILE_BYTES);
            tma_bulk_g2s(buf + st * TILE_FLOATS, a.tiles + (size_t)(t + NSTAGE) * TILE_FLOATS, TILE_BYTES, &bars[st]);
        }
        if (++since_flush == a.flush_tiles) { flush(); since_flush = 0; }
    }
    flush();

#pragma unroll
    for (int k = 0; k < IT; k++) {
        const int i = it * ITILE + k * 32 + lane;
        if (valid[k]) {
            double *o = a.part + ((size_t)s * a.ni + i) * PART_STRIDE;
#pragma unroll
            for (int c = 0; c < 7; c++) o[c] = D[k][c];
            a.cnt[(size_t)s * a.ni + i] = cnt[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// merge_kernel: one warp per i-particle.
//   fp64 sum over the S slices in a fixed order; scan of the S counts; concatenation of the
//   segments in slice order, which IS ascending j because slices are contiguous j ranges and
//   each warp visits its j ascending (the caller's two-pointer list diff needs strictly
//   ascending order, regcor_gpu.F:299-336).
//   count > nnbmax  ->  list[0] = -count, no entries written (reg.avx.cpp:320-321).
// ---------------------------------------------------------------------------------------------
struct MergeArgs {
    const double *part; const int *cnt; const int *seg;
    int ni, S, segcap, lmax, nnbmax;
    double *res_f;      // [ni][f_stride]; f_stride = 8 stores the (signed) count in slot 7 for the shard combine
    int     f_stride;
    int    *res_list;   // [ni][lmax]
};

__global__ void __launch_bounds__(128) merge_kernel(const MergeArgs a)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= a.ni) return;
    double f[7] = {0, 0, 0, 0, 0, 0, 0};
    int total = 0;
    for (int s = lane; s < a.S; s += 32) {
        const double *p = a.part + ((size_t)s * a.ni + i) * PART_STRIDE;
#pragma unroll
        for (int c = 0; c < 7; c++) f[c] += p[c];
        total += a.cnt[(size_t)s * a.ni + i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 7; c++) f[c] += __shfl_xor_sync(0xffffffffu, f[c], o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
    }
    if (lane < 7) a.res_f[(size_t)i * a.f_stride + lane] = f[lane];   // f[] is uniform after the butterfly
    if (a.f_stride == 8 && lane == 7) a.res_f[(size_t)i * 8 + 7] = (double)(total > a.nnbmax ? -total : total);
    int *row = a.res_list + (size_t)i * a.lmax;
    if (total > a.nnbmax) { if (lane == 0) row[0] = -total; return; }
    if (lane == 0) row[0] = total;
    int base = 1;
    for (int s0 = 0; s0 < a.S; s0 += 32) {
        const int s = s0 + lane;
        const int n = (s < a.S) ? a.cnt[(size_t)s * a.ni + i] : 0;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        const int off = base + incl - n;
        const int *src = a.seg + ((size_t)i * a.S + s) * a.segcap;
        for (int k = 0; k < n; k++) row[off + k] = src[k];
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// ---------------------------------------------------------------------------------------------
// combine_kernel: j-shard exchange step (multi-GPU).  Every shard r (a GPU holding j in
// [r*nj/R, (r+1)*nj/R), reference split: gpunb.velocity.cu:713-715) has produced, per i-particle,
// 7 fp64 partial sums + its signed neighbour count (fr[r][i][8]) and an ascending row of GLOBAL j
// indices (rows[r][i][lmax]).  One warp per i: fp64 sum over shards in rank order (the reference
// sums GPUs in fp64 on the host, :823-845), counts scanned in rank order, rows concatenated -- rank
// order is ascending j, so the concatenation is the index-ordered merge (:852-871).
// fr[r] / rows[r] are PEER pointers (NVLink P2P: cudaIpc-mapped across processes, or peer-enabled
// devices of one process): the kernel pulls only the `count` valid entries of each remote row, so the
// exchange moves ~4*nnb bytes per i instead of whole rows.  Overflow: any shard negative or
// total > nnbmax -> -(sum |count_r|) (reg.avx.cpp:320-321 encoding of the true count).
// ---------------------------------------------------------------------------------------------
constexpr int MAX_RANKS = 16;
struct CombineArgs {
    int ni, R, lmax, nnbmax;
    const double *fr[MAX_RANKS];     // [ni][8]
    const int    *rows[MAX_RANKS];   // [ni][lmax]
    double *res_f;                   // [ni][7]
    int    *res_list;                // [ni][lmax]
};

__global__ void __launch_bounds__(128) combine_kernel(const CombineArgs a)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= a.ni) return;
    double f = 0.0;
    int total = 0;
    bool over = false;
    int off[MAX_RANKS], cnt[MAX_RANKS];
    for (int r = 0; r < a.R; r++) {
        const double *p = a.fr[r] + (size_t)i * 8;
        if (lane < 7) f += p[lane];
        const int c = (int)p[7];
        over |= c < 0;
        cnt[r] = c < 0 ? -c : c;
        off[r] = 1 + total;
        total += cnt[r];
    }
    if (lane < 7) a.res_f[(size_t)i * 7 + lane] = f;
    int *row = a.res_list + (size_t)i * a.lmax;
    if (over || total > a.nnbmax) { if (lane == 0) row[0] = -total; return; }
    if (lane == 0) row[0] = total;
    for (int r = 0; r < a.R; r++) {
        const int *src = a.rows[r] + (size_t)i * a.lmax + 1;
        for (int k = lane; k < cnt[r]; k += 32) row[off[r] + k] = src[k];
    }
}

// ---------------------------------------------------------------------------------------------
// pot_kernel (gpupot): phi_i = sum_{j, r>0} m_j / r_ij.  Reuses the jtile layout (velocities
// unused).  Lanes own i-particles, j broadcast from smem tiles staged by the CTA; rsqrt.approx +
// one Newton step (the AVX twin does the same, pot.avx.cpp:22-25), FP32 chains of 64 terms
// flushed to fp64.  Work item = (32 i, j-slice); partial sums combined by pot_merge_kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pot_kernel(const float *__restrict__ tiles, int ntiles, int S,
                                                   int i0, int ni, double *__restrict__ part)
{
    __shared__ __align__(16) float sb[4][4 * TJ];        // per warp: xh|yh|zh|... staged 64 j at a time
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * 4 + warp;
    const int n_it = (ni + 31) / 32;
    if (w >= n_it * S) return;
    const int it = w / S, s = w - it * S;
    const int t0 = (int)(((long long)s * ntiles) / S), t1 = (int)(((long long)(s + 1) * ntiles) / S);
    const int ii = it * 32 + lane;
    const int i = i0 + (ii < ni ? ii : 0);
    const float *ti = tiles + (size_t)(i / TJ) * TILE_FLOATS + (i % TJ);
    const float xh = ti[0], yh = ti[TJ], zh = ti[2 * TJ], xl = ti[3 * TJ], yl = ti[4 * TJ], zl = ti[5 * TJ];
    double phi = 0.0;
    float *b = sb[warp];
    for (int t = t0; t < t1; t++) {
        const float *tp = tiles + (size_t)t * TILE_FLOATS;
        __syncwarp();
        // stage: 7 arrays of 64 floats needed (xh yh zh xl yl zl m); 64 floats = 2 per lane
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int jj = lane + 32 * c;
            b[jj]          = tp[jj];                 // xh
            b[TJ + jj]     = tp[TJ + jj];            // yh
            b[2 * TJ + jj] = tp[2 * TJ + jj];        // zh
            b[3 * TJ + jj] = tp[9 * TJ + jj];        // m
        }
        float lo[6];
#pragma unroll
        for (int c = 0; c < 2; c++) { lo[3*c] = tp[3 * TJ + lane + 32 * c]; lo[3*c+1] = tp[4 * TJ + lane + 32 * c]; lo[3*c+2] = tp[5 * TJ + lane + 32 * c]; }
        __syncwarp();
        float acc = 0.f;
#pragma unroll 8
        for (int jj = 0; jj < TJ; jj++) {
            const float jxl = __shfl_sync(0xffffffffu, lo[3 * (jj >> 5)],     jj & 31);
            const float jyl = __shfl_sync(0xffffffffu, lo[3 * (jj >> 5) + 1], jj & 31);
            const float jzl = __shfl_sync(0xffffffffu, lo[3 * (jj >> 5) + 2], jj & 31);
            const float dx = (b[jj] - xh) + (jxl - xl);
            const float dy = (b[TJ + jj] - yh) + (jyl - yl);
            const float dz = (b[2 * TJ + jj] - zh) + (jzl - zl);
            const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            float y = rsqrt_approx(r2);
            y = y * fmaf(-0.5f * r2 * y, y, 1.5f);      // one Newton step
            acc += (r2 > 0.f) ? b[3 * TJ + jj] * y : 0.f;
        }
        phi += (double)acc;
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
// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
double wtime() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + 1e-6 * tv.tv_usec; }

// Kernel variants (i-particles per lane x minimum resident CTAs per SM); GPUNB_B200_VARIANT selects one by
// name for tuning runs, the default is the fastest measured on B200 (see profiles/).
typedef void (*RegfKernel)(const RegfArgs);
struct Variant { const char *name; int it; RegfKernel k[2]; };
const Variant VARIANTS[] = {
    {"it2",   2, {regf_kernel<2, false, 1>, regf_kernel<2, true, 1>}},
    {"it2b3", 2, {regf_kernel<2, false, 3>, regf_kernel<2, true, 3>}},
    {"it1",   1, {regf_kernel<1, false, 1>, regf_kernel<1, true, 1>}},
    {"it1b5", 1, {regf_kernel<1, false, 5>, regf_kernel<1, true, 5>}},
};
constexpr int NVARIANTS = sizeof(VARIANTS) / sizeof(VARIANTS[0]);
constexpr int DEFAULT_VARIANT = 2;     // it1: best at small ni, equal at ni=1024 (profiles/r01b_variants.txt)
constexpr int ROWS_LMAX_CAP = 1024;    // row stride capacity of the IPC-exported shard rows (NCCL mode)

struct Dev {
    int id = -1;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, evs0 = nullptr, evs1 = nullptr, evdone = nullptr;
    int nsm = 0, warps_resident = 0, variant = DEFAULT_VARIANT, itile = 32;
    // j: full fp64 snapshot (m | x | v, packed for the current nj_total) and the tiles of this device's shard
    int raw_cap = 0, tile_cap = 0;
    double *jraw = nullptr;
    float *jtile = nullptr;
    double *radii = nullptr;      // h2[raw_cap] | dtr[raw_cap]  (resident sweeps)
    int nj_total = 0, j0 = 0, nj = 0, ntiles = 0;
    double *ibuf = nullptr;       // 8*NIMAX doubles: h2 | dtr | x | v
    int items_cap = 0;
    double *part = nullptr; int *cnt = nullptr;
    int *seg = nullptr; int segcap = 0; size_t seg_ints = 0;
    double *res_f = nullptr; int *res_list = nullptr; size_t res_list_ints = 0;   // final results (root)
    double *fr = nullptr;         // [NIMAX][8] shard partial + count (multi-GPU)
    int *rows = nullptr; size_t rows_ints = 0;                                     // shard rows (in-process multi-GPU)
    int *nanflag = nullptr;
    double *pot_part = nullptr, *pot_out = nullptr; size_t pot_part_n = 0, pot_out_n = 0;
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
    double *fr_all = nullptr;      // [R][NIMAX*8]
    int *rows_local = nullptr;     // [2][NIMAX][ROWS_LMAX_CAP], exported with cudaIpc
    int *rows_peer[MAX_RANKS] = {nullptr};
    int parity = 0;
};

struct Lib {
    bool devinit = false, is_open = false;
    std::vector<Dev> devs;
    Shard sh;
    int nbmax = 0, nbody = 0;
    double *h_j = nullptr; size_t h_j_n = 0;          // pinned staging: 7*nj doubles
    double *h_i = nullptr, *h_f = nullptr;            // 8*NIMAX, 7*NIMAX doubles
    int *h_list = nullptr; size_t h_list_n = 0;
    int *h_flag = nullptr;
    double time_send = 0, time_grav = 0, time_reduce = 0, time_out = 0;      // reference: gpunb.velocity.cu:557-559
    long long numInter = 0; int icall = 0, ini = 0, isend = 0;
    double ctr[GPUNB_B200_CTR_COUNT] = {0};
    int last_ni = 0, last_lmax = 0;
} L;

template <class T> void dev_alloc(T *&p, size_t n) { CUDA_CHECK(cudaMalloc((void **)&p, n * sizeof(T))); }
template <class T> void dev_free(T *&p) { if (p) CUDA_CHECK(cudaFree(p)); p = nullptr; }
template <class T> void host_alloc(T *&p, size_t n) { CUDA_CHECK(cudaMallocHost((void **)&p, n * sizeof(T))); }
template <class T> void host_free(T *&p) { if (p) CUDA_CHECK(cudaFreeHost(p)); p = nullptr; }

void set_dev(const Dev &d) { CUDA_CHECK(cudaSetDevice(d.id)); }
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
        fprintf(stderr, "# GPU initialization - rank: %d; HOST %s; NGPU %d; device: %d %s; B200-native regf[%s]: %d SMs x %d CTAs x %d warps\n",
                irank, host, (int)ids.size(), d.id, prop.name, V.name, d.nsm, nb, WARPS);
        L.devs.push_back(d);
    }
    // in-process multi-GPU: the root device pulls shard results over NVLink P2P
    for (size_t g = 1; g < L.devs.size(); g++) {
        int can = 0;
        CUDA_CHECK(cudaDeviceCanAccessPeer(&can, L.devs[0].id, L.devs[g].id));
        if (!can) FATAL("device %d cannot access device %d (P2P needed for the j-shard combine)", L.devs[0].id, L.devs[g].id);
        CUDA_CHECK(cudaSetDevice(L.devs[0].id));
        cudaError_t pe = cudaDeviceEnablePeerAccess(L.devs[g].id, 0);
        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CUDA_CHECK(pe);
        (void)cudaGetLastError();
    }
    CUDA_CHECK(cudaSetDevice(L.devs[0].id));
    host_alloc(L.h_flag, 16);
    memset(L.h_flag, 0, 16 * sizeof(int));
    L.devinit = true;
}

void shard_range(int r, int R, int nj, int &j0, int &j1)
{   // reference split: joff[id] = id*nbody/numGPU (gpunb.velocity.cu:713-715)
    j0 = (int)(((long long)r * nj) / R);
    j1 = (int)(((long long)(r + 1) * nj) / R);
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
        dev_free(d.jtile);
        d.tile_cap = tiles;
        dev_alloc(d.jtile, (size_t)tiles * TILE_FLOATS);
    }
}

void ensure_work_buffers(Dev &d, int lmax, int nnbmax, bool is_root)
{
    set_dev(d);
    // n_itiles * S never exceeds warps_resident (+ n_itiles when S = 1)
    const int items = d.warps_resident + NIMAX / d.itile;
    const int segcap = ((nnbmax > 0 ? nnbmax : 1) + 3) & ~3;
    if (items > d.items_cap) {
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        dev_free(d.part); dev_free(d.cnt);
        d.items_cap = items;
        dev_alloc(d.part, (size_t)items * d.itile * PART_STRIDE);
        dev_alloc(d.cnt, (size_t)items * d.itile);
    }
    const size_t seg_need = (size_t)d.items_cap * d.itile * segcap;
    if (seg_need > d.seg_ints) {
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        dev_free(d.seg);
        d.seg_ints = seg_need;
        dev_alloc(d.seg, seg_need);
    }
    d.segcap = segcap;
    const size_t rl = (size_t)NIMAX * lmax;
    if (is_root && rl > d.res_list_ints) {
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        dev_free(d.res_list);
        d.res_list_ints = rl;
        dev_alloc(d.res_list, rl);
    }
    if (L.devs.size() > 1 && rl > d.rows_ints) {
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        dev_free(d.rows);
        d.rows_ints = rl;
        dev_alloc(d.rows, rl);
    }
    if (is_root && rl > L.h_list_n) {
        host_free(L.h_list);
        L.h_list_n = rl;
        host_alloc(L.h_list, rl);
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
        if (!d.ibuf)    dev_alloc(d.ibuf, (size_t)8 * NIMAX);
        if (!d.res_f)   dev_alloc(d.res_f, (size_t)7 * NIMAX);
        if (!d.fr)      dev_alloc(d.fr, (size_t)8 * NIMAX);
        if (!d.nanflag) { dev_alloc(d.nanflag, 1); CUDA_CHECK(cudaMemsetAsync(d.nanflag, 0, sizeof(int), d.st)); }
    }
    const size_t hj = (size_t)7 * ((size_t)nbmax + 64);
    if (hj > L.h_j_n) { host_free(L.h_j); L.h_j_n = hj; host_alloc(L.h_j, hj); }
    if (!L.h_i) host_alloc(L.h_i, (size_t)8 * NIMAX);
    if (!L.h_f) host_alloc(L.h_f, (size_t)7 * NIMAX);
    fprintf(stderr, "# Open GPU regular force - rank: %d; nbmax: %d\n", irank, nbmax);
}

void lib_close()
{
    if (!L.is_open) { fprintf(stderr, "gpunb: it is already close\n"); return; }   // reference: :669-672
    L.is_open = false;
    for (Dev &d : L.devs) {
        set_dev(d);
        CUDA_CHECK(cudaStreamSynchronize(d.st));
        dev_free(d.jraw); dev_free(d.jtile); dev_free(d.radii); d.raw_cap = d.tile_cap = 0; d.nj = d.ntiles = d.nj_total = 0;
        dev_free(d.ibuf); dev_free(d.part); dev_free(d.cnt); d.items_cap = 0;
        dev_free(d.seg); d.seg_ints = 0; dev_free(d.res_f); dev_free(d.res_list); d.res_list_ints = 0;
        dev_free(d.fr); dev_free(d.rows); d.rows_ints = 0;
        dev_free(d.nanflag);
    }
    host_free(L.h_j); L.h_j_n = 0; host_free(L.h_i); host_free(L.h_f); host_free(L.h_list); L.h_list_n = 0;
    L.nbmax = 0;
}

// The snapshot is staged once in pinned memory (m | x | v), copied to every local device asynchronously and
// converted there; each device tiles only its own j-shard.  The reference converts fp64->fp32 on ONE host
// thread per GPU (gpunb.velocity.cu:721-724) and uses a blocking copy (:726).
void lib_send(int nj, const double *mj, const double *xj, const double *vj)
{
    if (!L.is_open) FATAL("gpunb_send called while the library is closed");
    if (nj > L.nbmax) FATAL("gpunb_send: nj=%d exceeds nbmax=%d given to gpunb_open", nj, L.nbmax);
    L.time_send -= wtime();
    L.isend++;
    L.nbody = nj;
    double *h = L.h_j;
    memcpy(h, mj, sizeof(double) * nj);
    memcpy(h + nj, xj, sizeof(double) * 3 * nj);
    memcpy(h + 4 * (size_t)nj, vj, sizeof(double) * 3 * nj);
    const int R = total_ranks();
    for (size_t g = 0; g < L.devs.size(); g++) {
        Dev &d = L.devs[g];
        set_dev(d);
        int j0, j1;
        shard_range(L.sh.on ? L.sh.rank : (int)g, R, nj, j0, j1);
        ensure_j_capacity(d, nj, j1 - j0);
        d.nj_total = nj; d.j0 = j0; d.nj = j1 - j0; d.ntiles = (d.nj + TJ - 1) / TJ;
        CUDA_CHECK(cudaMemcpyAsync(d.jraw, h, sizeof(double) * 7 * nj, cudaMemcpyHostToDevice, d.st));
        L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * 7.0 * nj;
        if (d.ntiles > 0) {
            const int threads = 256, blocks = (d.ntiles * TJ + threads - 1) / threads;
            jpack_kernel<<<blocks, threads, 0, d.st>>>(d.nj, d.ntiles, d.jraw + j0, d.jraw + nj + 3 * (size_t)j0,
                                                       d.jraw + 4 * (size_t)nj + 3 * (size_t)j0, d.jtile, d.nanflag);
            CUDA_CHECK(cudaGetLastError());
            L.ctr[GPUNB_B200_CTR_LAUNCHES] += 1;
        }
        CUDA_CHECK(cudaMemcpyAsync(L.h_flag + g, d.nanflag, sizeof(int), cudaMemcpyDeviceToHost, d.st));
    }
    for (size_t g = 0; g < L.devs.size(); g++) {
        set_dev(L.devs[g]);
        CUDA_CHECK(cudaStreamSynchronize(L.devs[g].st));
        if (L.h_flag[g]) FATAL("gpunb_send: NaN in j-particle data (reference asserts here, gpunb.velocity.cu:72-78)");
    }
    L.time_send += wtime();
}

struct Plan { int n_itiles, S, n_items; };
Plan make_plan(const Dev &d, int ni)
{
    Plan p;
    p.n_itiles = (ni + d.itile - 1) / d.itile;
    int S = d.warps_resident / p.n_itiles;
    if (S < 1) S = 1;
    if (S > d.ntiles) S = d.ntiles > 0 ? d.ntiles : 1;
    p.S = S;
    p.n_items = p.n_itiles * S;
    return p;
}

struct IBlock { const double *h2, *dtr, *xi, *vi; };

// Pair kernel + shard-local merge for one i-block on device d (async on d.st).
void launch_regf(Dev &d, int ni, const IBlock &ib, int lmax, int nnbmax, int m_flag, bool time_it,
                 double *out_f, int f_stride, int *out_rows)
{
    const Plan p = make_plan(d, ni);
    RegfArgs a;
    a.tiles = d.jtile; a.ntiles = d.ntiles; a.nj = d.nj; a.joff = d.j0;
    a.h2 = ib.h2; a.dtr = ib.dtr; a.xi = ib.xi; a.vi = ib.vi;
    a.ni = ni; a.n_itiles = p.n_itiles; a.S = p.S; a.n_items = p.n_items;
    a.part = d.part; a.cnt = d.cnt; a.seg = d.seg; a.segcap = d.segcap;
    { static int ft = 0; if (!ft) { const char *e = getenv("GPUNB_B200_FLUSH"); ft = e ? atoi(e) : FLUSH_TILES; if (ft < 1) ft = 1; } a.flush_tiles = ft; }
    const int smem = WARPS * NSTAGE * TILE_BYTES + WARPS * NSTAGE * 8;
    const int blocks = (p.n_items + WARPS - 1) / WARPS;
    if (time_it) CUDA_CHECK(cudaEventRecord(d.ev0, d.st));
    VARIANTS[d.variant].k[m_flag ? 1 : 0]<<<blocks, WARPS * 32, smem, d.st>>>(a);
    CUDA_CHECK(cudaGetLastError());
    if (time_it) CUDA_CHECK(cudaEventRecord(d.ev1, d.st));
    MergeArgs m;
    m.part = d.part; m.cnt = d.cnt; m.seg = d.seg; m.ni = ni; m.S = p.S; m.segcap = d.segcap;
    m.lmax = lmax; m.nnbmax = nnbmax; m.res_f = out_f; m.f_stride = f_stride; m.res_list = out_rows;
    merge_kernel<<<(ni + 3) / 4, 128, 0, d.st>>>(m);
    CUDA_CHECK(cudaGetLastError());
    if (time_it) CUDA_CHECK(cudaEventRecord(d.ev2, d.st));
    L.ctr[GPUNB_B200_CTR_LAUNCHES] += 2;
}

// One i-block on all shards + the exchange step.  ib[g] are DEVICE pointers valid on local device g.
// On return (asynchronously, on the root stream) root.res_f / root.res_list hold the combined result.
void regf_block(int ni, const IBlock *ib, int lmax, int nnbmax, int m_flag, bool time_it)
{
    const int G = (int)L.devs.size();
    Dev &root = L.devs[0];
    if (!L.sh.on && G == 1) {          // single GPU: the shard-local merge IS the final result
        set_dev(root);
        launch_regf(root, ni, ib[0], lmax, nnbmax, m_flag, time_it, root.res_f, 7, root.res_list);
        if (time_it) CUDA_CHECK(cudaEventRecord(root.ev3, root.st));
        return;
    }
    CombineArgs c;
    c.ni = ni; c.lmax = lmax; c.nnbmax = nnbmax; c.res_f = root.res_f; c.res_list = root.res_list;
    if (L.sh.on) {                     // one process per GPU: NCCL all-gather of the 64 B/i partials, P2P pull of rows
        Shard &sh = L.sh;
        set_dev(root);
        int *rows = sh.rows_local + (size_t)sh.parity * NIMAX * ROWS_LMAX_CAP;
        launch_regf(root, ni, ib[0], lmax, nnbmax, m_flag, time_it, root.fr, 8, rows);
        int rc = sh.allgather(root.fr, sh.fr_all, (size_t)ni * 8, NCCL_FLOAT64, sh.comm, root.st);
        if (rc != 0) FATAL("ncclAllGather failed: %s", sh.errstr ? sh.errstr(rc) : "?");
        c.R = sh.R;
        for (int r = 0; r < sh.R; r++) {
            c.fr[r] = sh.fr_all + (size_t)r * ni * 8;
            c.rows[r] = sh.rows_peer[r] + (size_t)sh.parity * NIMAX * ROWS_LMAX_CAP;
        }
        sh.parity ^= 1;                // rows are double buffered: a peer may still be pulling the previous block
    } else {                           // one process, G GPUs: root waits for every shard, pulls over P2P
        for (int g = 0; g < G; g++) {
            Dev &d = L.devs[g];
            set_dev(d);
            if (g > 0) CUDA_CHECK(cudaStreamWaitEvent(d.st, root.evdone, 0));     // previous combine has consumed d.fr / d.rows
            launch_regf(d, ni, ib[g], lmax, nnbmax, m_flag, time_it && g == 0, d.fr, 8, d.rows);
            if (g > 0) {
                CUDA_CHECK(cudaEventRecord(d.evdone, d.st));
                CUDA_CHECK(cudaStreamWaitEvent(root.st, d.evdone, 0));
            }
            c.fr[g] = d.fr; c.rows[g] = d.rows;
        }
        c.R = G;
        set_dev(root);
    }
    combine_kernel<<<(ni + 3) / 4, 128, 0, root.st>>>(c);
    CUDA_CHECK(cudaGetLastError());
    L.ctr[GPUNB_B200_CTR_LAUNCHES] += 1;
    if (!L.sh.on) CUDA_CHECK(cudaEventRecord(root.evdone, root.st));
    if (time_it) CUDA_CHECK(cudaEventRecord(root.ev3, root.st));
}

void fetch_results(int ni, int lmax, double *acc, double *jrk, double *pot, int *list)
{
    Dev &root = L.devs[0];
    set_dev(root);
    CUDA_CHECK(cudaMemcpyAsync(L.h_f, root.res_f, sizeof(double) * 7 * ni, cudaMemcpyDeviceToHost, root.st));
    CUDA_CHECK(cudaMemcpyAsync(L.h_list, root.res_list, sizeof(int) * (size_t)ni * lmax, cudaMemcpyDeviceToHost, root.st));
    L.ctr[GPUNB_B200_CTR_D2H_BYTES] += sizeof(double) * 7.0 * ni + sizeof(int) * (double)ni * lmax;
    CUDA_CHECK(cudaStreamSynchronize(root.st));
    for (int i = 0; i < ni; i++) {
        const double *f = L.h_f + 7 * (size_t)i;
        acc[3 * i] = f[0]; acc[3 * i + 1] = f[1]; acc[3 * i + 2] = f[2];
        jrk[3 * i] = f[3]; jrk[3 * i + 1] = f[4]; jrk[3 * i + 2] = f[5];
        pot[i] = f[6];
        const int *src = L.h_list + (size_t)i * lmax;
        int *dst = list + (size_t)i * lmax;
        const int n = src[0];
        dst[0] = n;
        if (n > 0) memcpy(dst + 1, src + 1, sizeof(int) * n);
    }
}

void lib_regf(int ni, const double *h2, const double *dtr, const double *xi, const double *vi,
              double *acc, double *jrk, double *pot, int lmax, int nnbmax, int *list, int m_flag)
{
    if (!L.is_open) FATAL("gpunb_regf called while the library is closed");
    if (!(0 < ni && ni <= NIMAX)) FATAL("gpunb_regf: ni=%d out of range (0, %d]", ni, NIMAX);
    if (nnbmax + 1 > lmax) FATAL("gpunb_regf: nnbmax=%d does not fit rows of lmax=%d", nnbmax, lmax);
    L.time_grav -= wtime();
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
    IBlock ib[MAX_RANKS];
    for (size_t g = 0; g < L.devs.size(); g++) {
        Dev &d = L.devs[g];
        set_dev(d);
        ensure_work_buffers(d, lmax, nnbmax, g == 0);
        CUDA_CHECK(cudaMemcpyAsync(d.ibuf, h, sizeof(double) * 8 * ni, cudaMemcpyHostToDevice, d.st));
        L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * 8.0 * ni;
        ib[g] = IBlock{d.ibuf, d.ibuf + ni, d.ibuf + 2 * (size_t)ni, d.ibuf + 5 * (size_t)ni};
    }
    regf_block(ni, ib, lmax, nnbmax, m_flag, true);
    const double wt0 = wtime();
    fetch_results(ni, lmax, acc, jrk, pot, list);
    Dev &root = L.devs[0];
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, root.ev0, root.ev1)); L.ctr[GPUNB_B200_CTR_GRAV_MS] += ms;
    CUDA_CHECK(cudaEventElapsedTime(&ms, root.ev1, root.ev3)); L.ctr[GPUNB_B200_CTR_MERGE_MS] += ms;
    L.ctr[GPUNB_B200_CTR_GRAV_LAUNCHES] += 1;
    const double wt = wtime();
    L.time_grav += wt0; L.time_reduce += wt - wt0;
    L.last_ni = ni; L.last_lmax = lmax;
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
void lib_pot(int irank, int istart, int ni, int n, const double *m, const double *x, double *pot)
{
    lib_devinit(irank);
    const double t0 = wtime();
    if (ni <= 0) return;
    if (istart < 1 || istart - 1 + ni > n) FATAL("gpupot: istart=%d ni=%d outside 1..n=%d", istart, ni, n);
    Dev &d = L.devs[0];
    set_dev(d);
    // private buffers so that gpupot never disturbs the regf j-snapshot
    static double *jraw = nullptr; static float *jtile = nullptr; static int cap = 0; static double *hpin = nullptr; static size_t hpin_n = 0;
    const int ntiles = (n + TJ - 1) / TJ;
    if (ntiles * TJ > cap) {
        dev_free(jraw); dev_free(jtile);
        cap = ntiles * TJ;
        dev_alloc(jraw, (size_t)4 * cap); dev_alloc(jtile, (size_t)ntiles * TILE_FLOATS);
    }
    if ((size_t)4 * n + ni > hpin_n) { host_free(hpin); hpin_n = (size_t)4 * n + ni + 1024; host_alloc(hpin, hpin_n); }
    if (!d.nanflag) { dev_alloc(d.nanflag, 1); CUDA_CHECK(cudaMemsetAsync(d.nanflag, 0, sizeof(int), d.st)); }
    memcpy(hpin, m, sizeof(double) * n);
    memcpy(hpin + n, x, sizeof(double) * 3 * n);
    CUDA_CHECK(cudaMemcpyAsync(jraw, hpin, sizeof(double) * 4 * n, cudaMemcpyHostToDevice, d.st));
    // velocities are not needed: pass the position array as "v" (finite, ignored by pot_kernel)
    jpack_kernel<<<(ntiles * TJ + 255) / 256, 256, 0, d.st>>>(n, ntiles, jraw, jraw + n, jraw + n, jtile, d.nanflag);
    CUDA_CHECK(cudaGetLastError());
    const int n_it = (ni + 31) / 32;
    int S = (d.nsm * 16 * 4) / n_it; if (S < 1) S = 1; if (S > ntiles) S = ntiles;
    if ((size_t)S * ni > d.pot_part_n) { dev_free(d.pot_part); d.pot_part_n = (size_t)S * ni; dev_alloc(d.pot_part, d.pot_part_n); }
    if ((size_t)ni > d.pot_out_n) { dev_free(d.pot_out); d.pot_out_n = ni; dev_alloc(d.pot_out, d.pot_out_n); }
    CUDA_CHECK(cudaEventRecord(d.evs0, d.st));
    pot_kernel<<<(n_it * S + 3) / 4, 128, 0, d.st>>>(jtile, ntiles, S, istart - 1, ni, d.pot_part);
    CUDA_CHECK(cudaGetLastError());
    pot_merge_kernel<<<(ni + 127) / 128, 128, 0, d.st>>>(d.pot_part, S, ni, d.pot_out);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaEventRecord(d.evs1, d.st));
    double *hout = hpin + 4 * (size_t)n;
    CUDA_CHECK(cudaMemcpyAsync(hout, d.pot_out, sizeof(double) * ni, cudaMemcpyDeviceToHost, d.st));
    CUDA_CHECK(cudaMemcpyAsync(L.h_flag, d.nanflag, sizeof(int), cudaMemcpyDeviceToHost, d.st));
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    if (L.h_flag[0]) FATAL("gpupot: NaN in particle data");
    float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, d.evs0, d.evs1));
    L.ctr[GPUNB_B200_CTR_POT_MS] += ms;
    L.ctr[GPUNB_B200_CTR_LAUNCHES] += 3;
    L.ctr[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * 4.0 * n;
    L.ctr[GPUNB_B200_CTR_D2H_BYTES] += sizeof(double) * (double)ni;
    memcpy(pot, hout, sizeof(double) * ni);
    const double t1 = wtime();
    fprintf(stderr, "[R.%d GPU Pot.A] Ni %d  NTOT %d  pot(s) %f\n", irank, ni, n, t1 - t0);   // reference: gpupot.gpu.cu:112
}

// ---- NCCL mode -------------------------------------------------------------------------------
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

int gpunb_b200_version(void) { return 101; }
const char *gpunb_b200_build_info(void)
{
    return "gpunb_b200 sm_100a: regf_kernel<f32x2, TMA bulk j-tiles, TJ=64>, merge_kernel, combine_kernel (P2P), pot_kernel, jpack_kernel";
}
int gpunb_b200_num_devices(void) { return (int)L.devs.size(); }
void gpunb_b200_get_counters(double out[GPUNB_B200_CTR_COUNT]) { for (int k = 0; k < GPUNB_B200_CTR_COUNT; k++) out[k] = L.ctr[k]; }
void gpunb_b200_reset_counters(void) { for (int k = 0; k < GPUNB_B200_CTR_COUNT; k++) L.ctr[k] = 0; }

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
    const int i0 = *i0p, ni = *nip, block = *blockp;
    if (i0 < 0 || i0 + ni > root.nj_total || block < 1 || block > NIMAX) FATAL("gpunb_b200_sweep_resident: bad range");
    for (size_t g = 0; g < L.devs.size(); g++) ensure_work_buffers(L.devs[g], *lmaxp, *nnbmaxp, g == 0);
    set_dev(root);
    CUDA_CHECK(cudaEventRecord(root.evs0, root.st));
    int nlaunch = 0, last = 0;
    for (int b = i0; b < i0 + ni; b += block) {
        const int n = (i0 + ni - b < block) ? i0 + ni - b : block;
        IBlock ib[MAX_RANKS];
        for (size_t g = 0; g < L.devs.size(); g++) {
            Dev &d = L.devs[g];
            const double *x = d.jraw + d.nj_total, *v = d.jraw + 4 * (size_t)d.nj_total;
            ib[g] = IBlock{d.radii + b, d.radii + d.raw_cap + b, x + 3 * (size_t)b, v + 3 * (size_t)b};
        }
        regf_block(n, ib, *lmaxp, *nnbmaxp, *m_flagp, false);
        L.ctr[GPUNB_B200_CTR_INTERACTIONS] += (double)n * root.nj_total;
        nlaunch++; last = n;
    }
    set_dev(root);
    CUDA_CHECK(cudaEventRecord(root.evs1, root.st));
    CUDA_CHECK(cudaStreamSynchronize(root.st));
    float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, root.evs0, root.evs1));
    L.ctr[GPUNB_B200_CTR_GRAV_LAUNCHES] += nlaunch;
    L.last_ni = last; L.last_lmax = *lmaxp;
    return ms;
}

void gpunb_b200_fetch_last(int *n_last, double acc[][3], double jrk[][3], double pot[], int *lmaxp, int *list)
{
    const int ni = L.last_ni, lmax = L.last_lmax;
    *n_last = ni; *lmaxp = lmax;
    if (ni <= 0) return;
    fetch_results(ni, lmax, &acc[0][0], &jrk[0][0], pot, list);
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
// rank receives the complete result; rank r sums over j in [r*nj/R, (r+1)*nj/R).
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
    sh.rank = rank; sh.R = nranks; sh.parity = 0;
    dev_alloc(sh.fr_all, (size_t)nranks * NIMAX * 8);
    dev_alloc(sh.rows_local, (size_t)2 * NIMAX * ROWS_LMAX_CAP);
    // exchange cudaIpc handles of the row buffers with an all-gather, then map every peer's buffer
    cudaIpcMemHandle_t mine;
    CUDA_CHECK(cudaIpcGetMemHandle(&mine, sh.rows_local));
    unsigned char *hbuf = nullptr; dev_alloc(hbuf, (size_t)(nranks + 1) * sizeof(cudaIpcMemHandle_t));
    CUDA_CHECK(cudaMemcpyAsync(hbuf + (size_t)nranks * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, d.st));
    rc = sh.allgather(hbuf + (size_t)nranks * sizeof(mine), hbuf, sizeof(mine), NCCL_INT8, sh.comm, d.st);
    if (rc != 0) FATAL("ncclAllGather(ipc handles) failed: %s", sh.errstr ? sh.errstr(rc) : "?");
    std::vector<cudaIpcMemHandle_t> all(nranks);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), hbuf, (size_t)nranks * sizeof(mine), cudaMemcpyDeviceToHost, d.st));
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    dev_free(hbuf);
    for (int r = 0; r < nranks; r++) {
        if (r == rank) { sh.rows_peer[r] = sh.rows_local; continue; }
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) FATAL("cudaIpcOpenMemHandle(rank %d) failed: %s (NVLink P2P between ranks is required)", r, cudaGetErrorString(e));
        sh.rows_peer[r] = (int *)p;
    }
    sh.on = true;
    return 0;
}

void gpunb_b200_nccl_finalize(void)
{
    Shard &sh = L.sh;
    if (!sh.on) return;
    Dev &d = L.devs[0];
    set_dev(d);
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    for (int r = 0; r < sh.R; r++)
        if (r != sh.rank && sh.rows_peer[r]) CUDA_CHECK(cudaIpcCloseMemHandle(sh.rows_peer[r]));
    dev_free(sh.fr_all);
    // the exported buffer is released only after every peer closed its mapping: a final all-gather is the barrier
    dev_alloc(sh.fr_all, 2 * (size_t)sh.R);
    sh.allgather(sh.fr_all + sh.R, sh.fr_all, 1, NCCL_FLOAT64, sh.comm, d.st);
    CUDA_CHECK(cudaStreamSynchronize(d.st));
    dev_free(sh.fr_all);
    dev_free(sh.rows_local);
    sh.destroy(sh.comm);
    sh.comm = nullptr; sh.on = false; sh.R = 1; sh.rank = 0;
}

}  // extern "C"
