// regcor_b200.cu -- the neighbour-list bookkeeping that follows every gpunb_regf_, batched on the device
// (part of libgpunb_b200.so, entry points in include/gpunb_b200.h part 3).
//
// What the reference does on the host, one i-particle at a time (serial two-pointer merges and up to 2 NNB fp64 pair
// forces per particle, src/Main/regcor_gpu.F:267-470, after the index shift / self removal of src/Main/util_gpu.F:102-111):
//   NLIST  <- row of gpunb_regf_ + IFIRST, self dropped
//   NBLOSS / NBGAIN / JJLIST <- old list LIST(:,I) against NLIST (both ascending)
//   lost members with a small step inside 2 RS are put back (ordered insertion; FREG / FDR corrected)
//   DFIRR / DFD <- - sum over lost + sum over gained of the fp64 pair force / derivative at the predicted positions
// Here: one WARP per row.  Both lists sit in shared memory; membership is a binary search per lane and the lost / gained
// members are compacted in order with ballots, which yields JJLIST in exactly the ascending order of the Fortran walk.
// The pair terms are evaluated by the lanes in parallel (the 56 B of particle J come from the snapshot gpunb_send_ left
// on the device -- X, XDOT, BODY of regcor are the predicted values of the block, i.e. that snapshot) and ADDED IN LIST
// ORDER by every lane redundantly (shuffles), every operation a single IEEE fp64 operation (__dmul_rn, __dadd_rn, ...:
// never contracted), so DFIRR / DFD / FREG / FDR are bit for bit what an unfused host build of the Fortran computes.
// The rare retention step is inherently sequential (ordered insertion, JJLIST edited in place): lane 0 runs it as the
// Fortran states it.  Old lists may come from the caller or from a device-resident list store (one row per particle,
// committed by this kernel itself), so that in steady state no list is uploaded at all.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <sys/time.h>
#include "../../include/gpunb_b200.h"
#include "internal.h"

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "gpunb_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, \
            cudaGetErrorString(e_)); abort(); } } while (0)
#define FATAL(...) do { fprintf(stderr, "gpunb_b200: " __VA_ARGS__); fprintf(stderr, "\n"); abort(); } while (0)

namespace {

constexpr int RC_ROWS  = 2048;     // rows per launch (larger batches are cut)
constexpr int RC_WARPS = 4;        // warps (rows) per CTA

struct RegcorArgs {
    int ni, ifirst, n, ntot, lmax, nnbmax, nj;
    const int *index_i;                 // [ni] particle number I of each row (Fortran numbering)
    const int *new_rows, *new_off;      // rows of gpunb_regf_ = [count, 0-based j ...]: packed, row r at new_rows + new_off[r]; or
                                        // (new_off == NULL) at new_rows + r lmax: the device copy the last gpunb_regf_ left
    const int *old_rows, *old_off;      // packed old lists [NNB0, members ...] (Fortran numbering), or NULL: list store
    int       *store; int store_stride; // resident list store: row I - 1 (NULL: none); the final NLIST is committed to it
    const double *m, *x, *v;            // snapshot: particle J at index J - ifirst
    const double *step; double smin;    // STEP of the snapshot particles (NULL: nothing is ever retained)
    const double *rs2;                  // [ni] RS(I)^2 at entry of regcor (regcor_gpu.F:41)
    const double *fio_in;               // [ni][12] FREG | FDR | DFIRR | DFD as passed
    double *fio_out;                    // [ni][12] the same, updated
    int *out_nlist;                     // [ni][lmax]   NLIST = [NNB, members ...]
    int *out_jj;                        // [ni][2 lmax] JJLIST: lost at [0, NBLOSS), gained at [NNB0, NNB0 + NBGAIN)
    int *out_cnt;                       // [ni][4]      NBLOSS, NBGAIN, members retained, NNB0
};

struct Pair { double f[3], fd[3]; };
// regcor_gpu.F:393-404 (= :428-438, :450-460), one IEEE operation per Fortran operation, left to right
__device__ __forceinline__ Pair pair_terms(const double xi[3], const double vi[3], const double *__restrict__ x,
                                           const double *__restrict__ v, const double *__restrict__ m, int j)
{
    Pair p;
    const double a1 = __dsub_rn(x[3 * (size_t)j], xi[0]), a2 = __dsub_rn(x[3 * (size_t)j + 1], xi[1]), a3 = __dsub_rn(x[3 * (size_t)j + 2], xi[2]);
    const double d1 = __dsub_rn(v[3 * (size_t)j], vi[0]), d2 = __dsub_rn(v[3 * (size_t)j + 1], vi[1]), d3 = __dsub_rn(v[3 * (size_t)j + 2], vi[2]);
    const double rij2 = __dadd_rn(__dadd_rn(__dmul_rn(a1, a1), __dmul_rn(a2, a2)), __dmul_rn(a3, a3));
    const double dr2i = __ddiv_rn(1.0, rij2);
    const double dr3i = __dmul_rn(__dmul_rn(m[j], dr2i), __dsqrt_rn(dr2i));
    const double drdv = __dadd_rn(__dadd_rn(__dmul_rn(a1, d1), __dmul_rn(a2, d2)), __dmul_rn(a3, d3));
    const double drdp = __dmul_rn(__dmul_rn(3.0, drdv), dr2i);
    p.f[0] = __dmul_rn(a1, dr3i); p.f[1] = __dmul_rn(a2, dr3i); p.f[2] = __dmul_rn(a3, dr3i);
    p.fd[0] = __dmul_rn(__dsub_rn(d1, __dmul_rn(a1, drdp)), dr3i);
    p.fd[1] = __dmul_rn(__dsub_rn(d2, __dmul_rn(a2, drdp)), dr3i);
    p.fd[2] = __dmul_rn(__dsub_rn(d3, __dmul_rn(a3, drdp)), dr3i);
    return p;
}

__device__ __forceinline__ bool contains(const int *__restrict__ a, int n, int key)
{   // a[0..n) strictly ascending
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo < n && a[lo] == key;
}

__global__ void __launch_bounds__(RC_WARPS * 32) regcor_kernel(const RegcorArgs a)
{
    extern __shared__ int smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * RC_WARPS + warp;
    if (r >= a.ni) return;                                        // warp-uniform; no CTA-wide barrier below
    int *NL = smem + (size_t)warp * (4 * a.lmax + 8);             // NL[0..nnb): new members (room for one retained member per lost one)
    int *OL = NL + a.lmax + 2;                                    // OL[0..nnb0): old members
    int *JJ = OL + a.lmax + 2;                                    // JJLIST(k) = JJ[k-1]
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1u;
    const int I = a.index_i[r];
    const int *nrow = a.new_off ? a.new_rows + a.new_off[r] : a.new_rows + (size_t)r * a.lmax;
    int *onl = a.out_nlist + (size_t)r * a.lmax;
    int *ocnt = a.out_cnt + 4 * (size_t)r;
    double *fo = a.fio_out + 12 * (size_t)r;
    const int cnt = nrow[0];
    if (cnt < 0) {                                                // overflow row: the caller's retry business (util_gpu.F:71-97)
        if (lane == 0) { onl[0] = cnt; ocnt[0] = ocnt[1] = ocnt[2] = ocnt[3] = 0; }
        if (lane < 12) fo[lane] = a.fio_in[12 * (size_t)r + lane];
        return;
    }
    // ---- 1. util_gpu.F:102-111: + IFIRST, self dropped ------------------------------------------------------------
    int nnb = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int k = base + lane;
        const int p = k < cnt ? nrow[1 + k] + a.ifirst : I;
        const bool keep = k < cnt && p != I;
        const unsigned mk = __ballot_sync(FULL, keep);
        if (keep) NL[nnb + __popc(mk & lt)] = p;
        nnb += __popc(mk);
    }
    // ---- old list ---------------------------------------------------------------------------------------------------
    const int *orow = a.old_rows ? a.old_rows + a.old_off[r] : a.store + (size_t)(I - 1) * a.store_stride;
    const int nnb0 = orow[0];
    for (int k = lane; k < nnb0; k += 32) OL[k] = orow[1 + k];
    __syncwarp();
    // ---- 2. regcor_gpu.F:267-336 ------------------------------------------------------------------------------------
    int nbloss = 0, nbgain = 0;
    bool small_step = false;
    if (nnb0 == 0) {                                              // :271-283: everything is gained, JJLIST(L) = NLIST(L+1)
        nbgain = nnb;
        for (int k = lane; k < nnb; k += 32) JJ[k] = NL[k];
    } else {
        for (int base = 0; base < nnb0; base += 32) {             // lost = old \ new, ascending
            const int k = base + lane;
            const int o = k < nnb0 ? OL[k] : 0;
            const bool lost = k < nnb0 && !contains(NL, nnb, o);
            const unsigned mk = __ballot_sync(FULL, lost);
            if (lost) {
                JJ[nbloss + __popc(mk & lt)] = o;
                if (a.step && a.step[o - a.ifirst] < a.smin) small_step = true;      // :317 (JMIN)
            }
            nbloss += __popc(mk);
        }
        for (int base = 0; base < nnb; base += 32) {              // gained = new \ old, ascending, filed behind NNB0
            const int k = base + lane;
            const int g = k < nnb ? NL[k] : 0;
            const bool gained = k < nnb && !contains(OL, nnb0, g);
            const unsigned mk = __ballot_sync(FULL, gained);
            if (gained) JJ[nnb0 + nbgain + __popc(mk & lt)] = g;
            nbgain += __popc(mk);
        }
    }
    const bool jmin = __any_sync(FULL, small_step);
    __syncwarp();
    double xi[3], vi[3];
#pragma unroll
    for (int c = 0; c < 3; c++) { xi[c] = a.x[3 * (size_t)(I - a.ifirst) + c]; vi[c] = a.v[3 * (size_t)(I - a.ifirst) + c]; }
    double F[12];                                                  // FREG | FDR | DFIRR | DFD (every lane keeps the same copy)
#pragma unroll
    for (int c = 0; c < 12; c++) F[c] = a.fio_in[12 * (size_t)r + c];
    // ---- 3. regcor_gpu.F:338-420 (rare): lane 0, as the Fortran states it ----------------------------------------
    int nbsmin = 0;
    if (jmin) {
        if (lane == 0) {
            const double rs2 = a.rs2[r];
            int k = 1;
            while (k <= nbloss) {
                if (nnb > a.nnbmax || I > a.n) break;                                  // :342
                const int j = JJ[k - 1];
                bool keep = !(a.step[j - a.ifirst] > a.smin || j < a.ifirst || j > a.n);   // :345
                if (keep) {
                    const double *xj = a.x + 3 * (size_t)(j - a.ifirst);
                    const double e1 = __dsub_rn(xi[0], xj[0]), e2 = __dsub_rn(xi[1], xj[1]), e3 = __dsub_rn(xi[2], xj[2]);
                    const double rij2 = __dadd_rn(__dadd_rn(__dmul_rn(e1, e1), __dmul_rn(e2, e2)), __dmul_rn(e3, e3));
                    if (rij2 > __dmul_rn(4.0, rs2)) keep = false;                      // :347-348
                }
                if (!keep) { k++; continue; }
                if (nnb == 0) {
                    // :351-358 with NNB = 0: the Fortran compares against NLIST(1), which holds scratch (the last old member,
                    // :304), and enters THAT instead of J unless it is smaller -- reproduced as written
                    const int s = OL[nnb0 - 1];
                    NL[0] = s < j ? j : s;
                } else {
                    int l2 = nnb - 1;                                                  // :351-358 ordered insertion
                    while (l2 >= 0 && !(NL[l2] < j)) { NL[l2 + 1] = NL[l2]; l2--; }
                    NL[l2 + 1] = j;
                }
                nnb++; nbloss--; nbsmin++;
                const Pair p = pair_terms(xi, vi, a.x, a.v, a.m, j - a.ifirst);       // :367-392
#pragma unroll
                for (int c = 0; c < 3; c++) { F[c] = __dsub_rn(F[c], p.f[c]); F[3 + c] = __dsub_rn(F[3 + c], p.fd[c]); }
                if (k > nbloss) break;                                                 // :408
                for (int l3 = k; l3 <= nbloss; l3++) JJ[l3 - 1] = JJ[l3];              // :409-411
            }
        }
        nnb = __shfl_sync(FULL, nnb, 0); nbloss = __shfl_sync(FULL, nbloss, 0); nbsmin = __shfl_sync(FULL, nbsmin, 0);
#pragma unroll
        for (int c = 0; c < 6; c++) F[c] = __shfl_sync(FULL, F[c], 0);
        __syncwarp();
    }
    // ---- 4. regcor_gpu.F:425-470: DFIRR / DFD, lost (-) then gained (+), in list order ---------------------------
    const int nchg = nbloss + nbgain;
    for (int base = 0; base < nchg; base += 32) {
        const int e = base + lane;
        Pair p;
#pragma unroll
        for (int c = 0; c < 3; c++) { p.f[c] = 0.0; p.fd[c] = 0.0; }
        if (e < nchg) {
            const int j = e < nbloss ? JJ[e] : JJ[nnb0 + (e - nbloss)];
            p = pair_terms(xi, vi, a.x, a.v, a.m, j - a.ifirst);
        }
        const int m = min(32, nchg - base);
        for (int k = 0; k < m; k++) {
            const bool minus = base + k < nbloss;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double tf = __shfl_sync(FULL, p.f[c], k), td = __shfl_sync(FULL, p.fd[c], k);
                F[6 + c] = minus ? __dsub_rn(F[6 + c], tf) : __dadd_rn(F[6 + c], tf);
                F[9 + c] = minus ? __dsub_rn(F[9 + c], td) : __dadd_rn(F[9 + c], td);
            }
        }
    }
    // ---- results ---------------------------------------------------------------------------------------------------
    if (lane < 12) {
        double val = F[0];
#pragma unroll
        for (int c = 1; c < 12; c++) if (lane == c) val = F[c];
        fo[lane] = val;
    }
    if (lane == 0) { onl[0] = nnb; ocnt[0] = nbloss; ocnt[1] = nbgain; ocnt[2] = nbsmin; ocnt[3] = nnb0; }
    for (int k = lane; k < nnb; k += 32) onl[1 + k] = NL[k];
    int *ojj = a.out_jj + 2 * (size_t)r * a.lmax;
    for (int k = lane; k < nbloss; k += 32) ojj[k] = JJ[k];
    for (int k = lane; k < nbgain; k += 32) ojj[nnb0 + k] = JJ[nnb0 + k];
    if (a.store) {                                                // the list the integrator now holds for particle I
        int *srow = a.store + (size_t)(I - 1) * a.store_stride;
        if (lane == 0) srow[0] = nnb;
        for (int k = lane; k < nnb; k += 32) srow[1 + k] = NL[k];
    }
}

// rows[k] (stride lmax, [count, members ...]) -> store row index_i[k] - 1
__global__ void store_put_kernel(int n, int lmax, const int *__restrict__ index_i, const int *__restrict__ rows,
                                 const int *__restrict__ off, int *__restrict__ store, int stride)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const int *s = rows + off[w];
    int *d = store + (size_t)(index_i[w] - 1) * stride;
    const int c = s[0];
    for (int k = lane; k <= c; k += 32) d[k] = s[k];
}
__global__ void store_get_kernel(int n, int lmax, const int *__restrict__ index_i, const int *__restrict__ store, int stride,
                                 int *__restrict__ rows)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const int *s = store + (size_t)(index_i[w] - 1) * stride;
    int *d = rows + (size_t)w * lmax;
    const int c = min(max(s[0], 0), lmax - 1);
    for (int k = lane; k <= c; k += 32) d[k] = s[k];
}
__global__ void step_scatter_kernel(int n, const int *__restrict__ idx, const double *__restrict__ val, double *__restrict__ step, int cap)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && idx[k] >= 0 && idx[k] < cap) step[idx[k]] = val[k];
}

struct Regcor {
    int lmax = 0;
    // pinned staging (inputs), device copies, mapped pinned outputs
    int *h_int = nullptr, *d_int = nullptr; size_t int_cap = 0;
    double *h_dbl = nullptr, *d_dbl = nullptr;
    int *o_nlist = nullptr, *o_nlist_d = nullptr, *o_jj = nullptr, *o_jj_d = nullptr, *o_cnt = nullptr, *o_cnt_d = nullptr;
    double *o_f = nullptr, *o_f_d = nullptr;
    double *d_step = nullptr; int step_cap = 0; bool step_resident = false;
    int *store = nullptr; int store_rows = 0, store_stride = 0;
    bool smem_attr = false;
    // GPUNB_B200_REGCOR_PROFILE=1: host buckets and kernel time per call (us), printed by gpunb_close_
    bool profile = false; cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double t_pack = 0, t_enqueue = 0, t_wait = 0, t_unpack = 0, t_kernel = 0; long long n_prof = 0;
} RC;

double now_us() { struct timeval tv; gettimeofday(&tv, nullptr); return 1e6 * tv.tv_sec + tv.tv_usec; }

template <class T> void pinned(T *&p, size_t n) { CUDA_CHECK(cudaMallocHost((void **)&p, n * sizeof(T))); }
template <class T> void mapped(T *&h, T *&d, size_t n)
{
    CUDA_CHECK(cudaHostAlloc((void **)&h, n * sizeof(T), cudaHostAllocMapped | cudaHostAllocPortable));
    CUDA_CHECK(cudaHostGetDevicePointer((void **)&d, (void *)h, 0));
}

void free_buffers()
{
    if (RC.h_int) cudaFreeHost(RC.h_int);
    if (RC.h_dbl) cudaFreeHost(RC.h_dbl);
    if (RC.d_int) cudaFree(RC.d_int);
    if (RC.d_dbl) cudaFree(RC.d_dbl);
    if (RC.o_nlist) cudaFreeHost(RC.o_nlist);
    if (RC.o_jj) cudaFreeHost(RC.o_jj);
    if (RC.o_cnt) cudaFreeHost(RC.o_cnt);
    if (RC.o_f) cudaFreeHost(RC.o_f);
    RC.h_int = RC.d_int = RC.o_nlist = RC.o_jj = RC.o_cnt = nullptr; RC.h_dbl = RC.d_dbl = RC.o_f = nullptr;
    RC.lmax = 0;
}

void ensure_buffers(int lmax)
{
    if (lmax == RC.lmax) return;
    free_buffers();
    RC.lmax = lmax;
    RC.int_cap = (size_t)RC_ROWS * (2 * (size_t)lmax + 3) + 8;   // index | new_off | old_off | packed new | packed old
    pinned(RC.h_int, RC.int_cap);
    CUDA_CHECK(cudaMalloc((void **)&RC.d_int, RC.int_cap * sizeof(int)));
    pinned(RC.h_dbl, (size_t)RC_ROWS * 13);                      // rs2 | fio
    CUDA_CHECK(cudaMalloc((void **)&RC.d_dbl, (size_t)RC_ROWS * 13 * sizeof(double)));
    mapped(RC.o_nlist, RC.o_nlist_d, (size_t)RC_ROWS * lmax);
    mapped(RC.o_jj, RC.o_jj_d, (size_t)RC_ROWS * 2 * lmax);
    mapped(RC.o_cnt, RC.o_cnt_d, (size_t)RC_ROWS * 4);
    mapped(RC.o_f, RC.o_f_d, (size_t)RC_ROWS * 12);
}

void ensure_store(const GpunbSnapshotView &S, int lmax, int max_index)
{
    if (!RC.store) {
        RC.store_stride = lmax;
        RC.store_rows = 2 * S.nbmax + 64;                          // particle numbers reach NTOT = nj + 2 NPAIRS
        CUDA_CHECK(cudaMalloc((void **)&RC.store, sizeof(int) * (size_t)RC.store_rows * RC.store_stride));
        CUDA_CHECK(cudaMemsetAsync(RC.store, 0, sizeof(int) * (size_t)RC.store_rows * RC.store_stride, S.stream));
    }
    if (lmax != RC.store_stride) FATAL("list store holds rows of %d entries, this call uses lmax = %d", RC.store_stride, lmax);
    if (max_index > RC.store_rows) FATAL("particle number %d outside the list store (%d rows)", max_index, RC.store_rows);
}

void ensure_step(const GpunbSnapshotView &S)
{
    if (RC.step_cap < S.nbmax + 64) {
        CUDA_CHECK(cudaStreamSynchronize(S.stream));
        double *old = RC.d_step;
        const int cap = S.nbmax + 64;
        CUDA_CHECK(cudaMalloc((void **)&RC.d_step, sizeof(double) * (size_t)cap));
        CUDA_CHECK(cudaMemset(RC.d_step, 0, sizeof(double) * (size_t)cap));
        if (old) { CUDA_CHECK(cudaMemcpy(RC.d_step, old, sizeof(double) * (size_t)RC.step_cap, cudaMemcpyDeviceToDevice)); cudaFree(old); }
        RC.step_cap = cap;
    }
}

GpunbSnapshotView snapshot_or_die(const char *who)
{
    GpunbSnapshotView S;
    if (!gpunb_b200_internal_snapshot(&S)) FATAL("%s needs an open library and a snapshot (gpunb_send_ / gpunb_b200_predict_send_)", who);
    CUDA_CHECK(cudaSetDevice(S.device));
    return S;
}

// packs rows [count, entries ...] of stride lmax into dst, offsets into off[0..n]; returns ints used
size_t pack_rows(int n, int lmax, const int *rows, int *dst, int *off, int base)
{
    size_t used = 0;
    for (int r = 0; r < n; r++) {
        const int c0 = rows[(size_t)r * lmax], c = c0 < 0 ? 0 : c0;
        if (c + 1 > lmax) FATAL("list row %d holds %d members, lmax = %d", r, c, lmax);
        off[r] = base + (int)used;
        used += (size_t)c + 1;
    }
    off[n] = base + (int)used;
#pragma omp parallel for num_threads(gpunb_b200_internal_host_team()) schedule(static) if (n >= 256)
    for (int r = 0; r < n; r++)
        memcpy(dst + (off[r] - base), rows + (size_t)r * lmax, sizeof(int) * (size_t)(off[r + 1] - off[r]));
    return used;
}

}  // namespace

void gpunb_b200_internal_regcor_close()
{
    if (RC.profile && RC.n_prof)
        fprintf(stderr, "gpunb_b200_regcor: %lld launches, us per launch: pack %.1f enqueue %.1f wait %.1f (kernel %.1f) unpack %.1f\n", RC.n_prof,
                RC.t_pack / RC.n_prof, RC.t_enqueue / RC.n_prof, RC.t_wait / RC.n_prof, RC.t_kernel / RC.n_prof, RC.t_unpack / RC.n_prof);
    if (RC.ev0) { cudaEventDestroy(RC.ev0); cudaEventDestroy(RC.ev1); }
    free_buffers();
    if (RC.d_step) cudaFree(RC.d_step);
    if (RC.store) cudaFree(RC.store);
    RC = Regcor();
}

extern "C" {

static void regcor_impl(bool last, int *nip, int index_i[], int *ifirstp, int *np, int *ntotp, int *lmaxp, int new_list[], int old_list[],
                        double rs2[], double step[], double *sminp, int *nnbmaxp, double freg[][3], double fdr[][3],
                        double dfirr[][3], double dfd[][3], int nbloss[], int nbgain[], int jjlist[], int *nbsmin)
{
    const GpunbSnapshotView S = snapshot_or_die(last ? "gpunb_b200_regcor_last_" : "gpunb_b200_regcor_");
    if (last && (S.last_rows == nullptr || S.last_rows_ni != *nip || S.last_rows_lmax != *lmaxp))
        FATAL("gpunb_b200_regcor_last_: the last gpunb_regf_ call left %d rows of lmax %d on the device, this call names %d rows of lmax %d "
              "(i-slice mode keeps none)", S.last_rows_ni, S.last_rows_lmax, *nip, *lmaxp);
    struct timeval tv0; gettimeofday(&tv0, nullptr);
    const int ni = *nip, ifirst = *ifirstp, lmax = *lmaxp;
    if (lmax < 4) FATAL("gpunb_b200_regcor_: lmax = %d", lmax);
    ensure_buffers(lmax);
    if (!RC.smem_attr) {
        CUDA_CHECK(cudaFuncSetAttribute(regcor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        RC.smem_attr = true;
        const char *e = getenv("GPUNB_B200_REGCOR_PROFILE");
        RC.profile = e && atoi(e) > 0;
        if (RC.profile) { CUDA_CHECK(cudaEventCreate(&RC.ev0)); CUDA_CHECK(cudaEventCreate(&RC.ev1)); }
    }
    const size_t smem = sizeof(int) * (size_t)RC_WARPS * (4 * (size_t)lmax + 8);
    if (smem > 200 * 1024) FATAL("gpunb_b200_regcor_: lmax = %d needs %zu bytes of shared memory per CTA", lmax, smem);
    const double *d_step = nullptr;
    if (step) {                                     // STEP(IFIRST ..) of the snapshot particles, uploaded for this call
        ensure_step(S);
        CUDA_CHECK(cudaMemcpyAsync(RC.d_step, step, sizeof(double) * (size_t)S.nj, cudaMemcpyHostToDevice, S.stream));
        S.counters[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * (double)S.nj;
        RC.step_resident = false;
        d_step = RC.d_step;
    } else if (RC.step_resident) d_step = RC.d_step;
    int total_smin = 0;
    for (int r0 = 0; r0 < ni; r0 += RC_ROWS) {
        const int nr = ni - r0 < RC_ROWS ? ni - r0 : RC_ROWS;
        int maxI = 0;
        for (int r = 0; r < nr; r++) {
            const int I = index_i[r0 + r];
            if (I < ifirst || I - ifirst >= S.nj) FATAL("gpunb_b200_regcor_: particle %d outside the snapshot [%d, %d)", I, ifirst, ifirst + S.nj);
            maxI = I > maxI ? I : maxI;
        }
        if (!old_list && !RC.store) FATAL("gpunb_b200_regcor_: no old lists passed and no resident list store (gpunb_b200_lists_put_)");
        // the store is read (old_list == NULL) and committed to; a call that brings its own old lists at another lmax than the
        // store's rows simply does not commit
        const bool use_store = RC.store && (!old_list || lmax == RC.store_stride);
        if (use_store) ensure_store(S, lmax, maxI);
        // ---- pack and upload -------------------------------------------------------------------------------------
        int *hi = RC.h_int;
        // layout sized by this call's rows: index | new_off | old_off | packed new | packed old
        int *h_index = hi, *h_noff = hi + nr, *h_ooff = h_noff + nr + 1, *h_pack = h_ooff + nr + 1;
        const int pack_base = (int)(h_pack - hi);
        const double tp0 = now_us();
        memcpy(h_index, index_i + r0, sizeof(int) * (size_t)nr);
        size_t used = last ? 0 : pack_rows(nr, lmax, new_list + (size_t)r0 * lmax, h_pack, h_noff, pack_base);
        // overflow rows keep their negative count in the packed copy (pack_rows copies row[0] as it is)
        if (old_list) used += pack_rows(nr, lmax, old_list + (size_t)(r0) * lmax, h_pack + used, h_ooff, pack_base + (int)used);
        double *hd = RC.h_dbl;
        for (int r = 0; r < nr; r++) {
            hd[r] = rs2[r0 + r];
            double *f = hd + nr + 12 * (size_t)r;
            for (int c = 0; c < 3; c++) {
                f[c] = freg[r0 + r][c]; f[3 + c] = fdr[r0 + r][c]; f[6 + c] = dfirr[r0 + r][c]; f[9 + c] = dfd[r0 + r][c];
            }
        }
        const size_t nint = (size_t)pack_base + used;
        CUDA_CHECK(cudaMemcpyAsync(RC.d_int, hi, sizeof(int) * nint, cudaMemcpyHostToDevice, S.stream));
        const double tp1 = now_us();
        CUDA_CHECK(cudaMemcpyAsync(RC.d_dbl, hd, sizeof(double) * 13 * (size_t)nr, cudaMemcpyHostToDevice, S.stream));
        RegcorArgs a;
        a.ni = nr; a.ifirst = ifirst; a.n = *np; a.ntot = *ntotp; a.lmax = lmax; a.nnbmax = *nnbmaxp; a.nj = S.nj;
        a.index_i = RC.d_int; a.new_rows = last ? S.last_rows : RC.d_int; a.new_off = last ? nullptr : RC.d_int + nr;
        a.old_rows = old_list ? RC.d_int : nullptr; a.old_off = RC.d_int + 2 * nr + 1;
        a.store = use_store ? RC.store : nullptr; a.store_stride = RC.store_stride;
        a.m = S.m; a.x = S.x; a.v = S.v;
        a.step = d_step; a.smin = *sminp;
        a.rs2 = RC.d_dbl; a.fio_in = RC.d_dbl + nr; a.fio_out = RC.o_f_d;
        a.out_nlist = RC.o_nlist_d; a.out_jj = RC.o_jj_d; a.out_cnt = RC.o_cnt_d;
        if (RC.profile) CUDA_CHECK(cudaEventRecord(RC.ev0, S.stream));
        regcor_kernel<<<(nr + RC_WARPS - 1) / RC_WARPS, RC_WARPS * 32, smem, S.stream>>>(a);
        CUDA_CHECK(cudaGetLastError());
        if (RC.profile) CUDA_CHECK(cudaEventRecord(RC.ev1, S.stream));
        const double tp2 = now_us();
        CUDA_CHECK(cudaStreamSynchronize(S.stream));
        const double tp3 = now_us();
        // ---- results: only the entries in use cross PCIe (the kernel wrote them into mapped pinned memory) ----------
        size_t out_ints = 0;
#pragma omp parallel for num_threads(gpunb_b200_internal_host_team()) schedule(static) reduction(+ : out_ints, total_smin) if (nr >= 256)
        for (int r = 0; r < nr; r++) {
            const int *oc = RC.o_cnt + 4 * (size_t)r;
            const int *onl = RC.o_nlist + (size_t)r * lmax;
            int *nl = new_list + (size_t)(r0 + r) * lmax;
            const int nnb = onl[0];
            if (nnb >= 0) memcpy(nl, onl, sizeof(int) * (size_t)(nnb + 1));
            else nl[0] = nnb;
            nbloss[r0 + r] = oc[0]; nbgain[r0 + r] = oc[1]; total_smin += oc[2];
            const int *ojj = RC.o_jj + 2 * (size_t)r * lmax;
            int *jj = jjlist + 2 * (size_t)(r0 + r) * lmax;
            memcpy(jj, ojj, sizeof(int) * (size_t)oc[0]);
            memcpy(jj + oc[3], ojj + oc[3], sizeof(int) * (size_t)oc[1]);
            const double *f = RC.o_f + 12 * (size_t)r;
            for (int c = 0; c < 3; c++) {
                freg[r0 + r][c] = f[c]; fdr[r0 + r][c] = f[3 + c]; dfirr[r0 + r][c] = f[6 + c]; dfd[r0 + r][c] = f[9 + c];
            }
            out_ints += (size_t)(nnb > 0 ? nnb : 0) + 1 + oc[0] + oc[1] + 4;
        }
        if (RC.profile) {
            float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, RC.ev0, RC.ev1));
            RC.t_pack += tp1 - tp0; RC.t_enqueue += tp2 - tp1; RC.t_wait += tp3 - tp2; RC.t_unpack += now_us() - tp3; RC.t_kernel += 1e3 * ms; RC.n_prof++;
        }
        S.counters[GPUNB_B200_CTR_H2D_BYTES] += sizeof(int) * (double)nint + sizeof(double) * 13.0 * nr;
        S.counters[GPUNB_B200_CTR_D2H_BYTES] += sizeof(int) * (double)out_ints + sizeof(double) * 12.0 * nr;
        S.counters[GPUNB_B200_CTR_LAUNCHES] += 1;
    }
    *nbsmin = total_smin;
    struct timeval tv1; gettimeofday(&tv1, nullptr);
    S.counters[GPUNB_B200_CTR_REGCOR_MS] += 1e3 * (tv1.tv_sec - tv0.tv_sec) + 1e-3 * (tv1.tv_usec - tv0.tv_usec);
    S.counters[GPUNB_B200_CTR_REGCOR_ROWS] += ni;
}

void gpunb_b200_regcor_(int *ni, int index_i[], int *ifirst, int *n, int *ntot, int *lmax, int new_list[], int old_list[],
                        double rs2[], double step[], double *smin, int *nnbmax, double freg[][3], double fdr[][3],
                        double dfirr[][3], double dfd[][3], int nbloss[], int nbgain[], int jjlist[], int *nbsmin)
{
    regcor_impl(false, ni, index_i, ifirst, n, ntot, lmax, new_list, old_list, rs2, step, smin, nnbmax, freg, fdr, dfirr, dfd, nbloss, nbgain, jjlist, nbsmin);
}
// The same for the rows of the LAST gpunb_regf_ call, which are still on the device: new_list is output only (NLIST) and no
// list is uploaded -- none at all when the old lists come from the resident store.
void gpunb_b200_regcor_last_(int *ni, int index_i[], int *ifirst, int *n, int *ntot, int *lmax, int new_list[], int old_list[],
                             double rs2[], double step[], double *smin, int *nnbmax, double freg[][3], double fdr[][3],
                             double dfirr[][3], double dfd[][3], int nbloss[], int nbgain[], int jjlist[], int *nbsmin)
{
    regcor_impl(true, ni, index_i, ifirst, n, ntot, lmax, new_list, old_list, rs2, step, smin, nnbmax, freg, fdr, dfirr, dfd, nbloss, nbgain, jjlist, nbsmin);
}

// Rows of the resident list store: lists[k] (stride lmax) = LIST(1:LMAX, index_i[k]) = [NNB, members ...], Fortran numbering.
void gpunb_b200_lists_put_(int *np, int index_i[], int *lmaxp, int lists[])
{
    const GpunbSnapshotView S = snapshot_or_die("gpunb_b200_lists_put_");
    const int n = *np, lmax = *lmaxp;
    ensure_buffers(lmax);
    for (int r0 = 0; r0 < n; r0 += RC_ROWS) {
        const int nr = n - r0 < RC_ROWS ? n - r0 : RC_ROWS;
        int maxI = 0;
        for (int r = 0; r < nr; r++) { if (index_i[r0 + r] < 1) FATAL("gpunb_b200_lists_put_: particle number %d", index_i[r0 + r]); maxI = index_i[r0 + r] > maxI ? index_i[r0 + r] : maxI; }
        ensure_store(S, lmax, maxI);
        int *hi = RC.h_int;
        int *h_index = hi, *h_off = hi + RC_ROWS, *h_pack = hi + 3 * RC_ROWS + 2;
        const int pack_base = (int)(h_pack - hi);
        memcpy(h_index, index_i + r0, sizeof(int) * (size_t)nr);
        const size_t used = pack_rows(nr, lmax, lists + (size_t)r0 * lmax, h_pack, h_off, pack_base);
        CUDA_CHECK(cudaMemcpyAsync(RC.d_int, hi, sizeof(int) * ((size_t)pack_base + used), cudaMemcpyHostToDevice, S.stream));
        store_put_kernel<<<(nr * 32 + 127) / 128, 128, 0, S.stream>>>(nr, lmax, RC.d_int, RC.d_int, RC.d_int + RC_ROWS, RC.store, RC.store_stride);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(S.stream));
        S.counters[GPUNB_B200_CTR_H2D_BYTES] += sizeof(int) * ((double)pack_base + (double)used);
        S.counters[GPUNB_B200_CTR_LAUNCHES] += 1;
    }
}

void gpunb_b200_lists_get_(int *np, int index_i[], int *lmaxp, int lists[])
{
    const GpunbSnapshotView S = snapshot_or_die("gpunb_b200_lists_get_");
    const int n = *np, lmax = *lmaxp;
    if (!RC.store) FATAL("gpunb_b200_lists_get_: no resident list store");
    ensure_buffers(lmax);
    for (int r0 = 0; r0 < n; r0 += RC_ROWS) {
        const int nr = n - r0 < RC_ROWS ? n - r0 : RC_ROWS;
        int maxI = 0;
        for (int r = 0; r < nr; r++) { if (index_i[r0 + r] < 1) FATAL("gpunb_b200_lists_get_: particle number %d", index_i[r0 + r]); maxI = index_i[r0 + r] > maxI ? index_i[r0 + r] : maxI; }
        ensure_store(S, lmax, maxI);
        memcpy(RC.h_int, index_i + r0, sizeof(int) * (size_t)nr);
        CUDA_CHECK(cudaMemcpyAsync(RC.d_int, RC.h_int, sizeof(int) * (size_t)nr, cudaMemcpyHostToDevice, S.stream));
        store_get_kernel<<<(nr * 32 + 127) / 128, 128, 0, S.stream>>>(nr, lmax, RC.d_int, RC.store, RC.store_stride, RC.o_nlist_d);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(S.stream));
        for (int r = 0; r < nr; r++) {
            const int *row = RC.o_nlist + (size_t)r * lmax;
            memcpy(lists + (size_t)(r0 + r) * lmax, row, sizeof(int) * (size_t)(row[0] + 1));
        }
        S.counters[GPUNB_B200_CTR_LAUNCHES] += 1;
    }
}

// Resident STEP of the snapshot particles (index = J - IFIRST, like gpunb_b200_state_update_): all of them / the ones just changed.
void gpunb_b200_steps_all_(int *njp, double step[])
{
    const GpunbSnapshotView S = snapshot_or_die("gpunb_b200_steps_all_");
    ensure_step(S);
    if (*njp > RC.step_cap) FATAL("gpunb_b200_steps_all_: nj = %d exceeds nbmax", *njp);
    CUDA_CHECK(cudaMemcpyAsync(RC.d_step, step, sizeof(double) * (size_t)*njp, cudaMemcpyHostToDevice, S.stream));
    CUDA_CHECK(cudaStreamSynchronize(S.stream));
    RC.step_resident = true;
    S.counters[GPUNB_B200_CTR_H2D_BYTES] += sizeof(double) * (double)*njp;
}
void gpunb_b200_steps_update_(int *np, int idx[], double step[])
{
    const GpunbSnapshotView S = snapshot_or_die("gpunb_b200_steps_update_");
    if (!RC.step_resident) FATAL("gpunb_b200_steps_update_ before gpunb_b200_steps_all_");
    const int n = *np;
    ensure_buffers(RC.lmax ? RC.lmax : 64);
    for (int k0 = 0; k0 < n; k0 += RC_ROWS) {
        const int nk = n - k0 < RC_ROWS ? n - k0 : RC_ROWS;
        memcpy(RC.h_int, idx + k0, sizeof(int) * (size_t)nk);
        memcpy(RC.h_dbl, step + k0, sizeof(double) * (size_t)nk);
        CUDA_CHECK(cudaMemcpyAsync(RC.d_int, RC.h_int, sizeof(int) * (size_t)nk, cudaMemcpyHostToDevice, S.stream));
        CUDA_CHECK(cudaMemcpyAsync(RC.d_dbl, RC.h_dbl, sizeof(double) * (size_t)nk, cudaMemcpyHostToDevice, S.stream));
        step_scatter_kernel<<<(nk + 127) / 128, 128, 0, S.stream>>>(nk, RC.d_int, RC.d_dbl, RC.d_step, RC.step_cap);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(S.stream));
        S.counters[GPUNB_B200_CTR_LAUNCHES] += 1;
    }
}

}  // extern "C"
