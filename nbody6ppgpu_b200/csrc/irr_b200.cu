// irr_b200.cu -- B200-native irregular-force library behind the reference's irr_simd_* ABI
// (reference: src/Main/irr.avx.cpp:365-603, callers intgrt.F:199-207,545,1267-1273).
//
// The fp64 statement it implements (nbody6ppgpu_b200/irr.py: firr_f64) is pinned against the reference's own AVX library on
// the CPU (tests/test_irr_cpu.py); on a B200 the library reproduces that statement to 2e-14 with identical nearest-neighbour
// addresses (tests/test_irr_cpu.py::test_cuda_library_against_the_fp64_statement, tests/test_irr_gpu.py).
//
// What the reference does: one AVX thread per active particle; neighbour records (two-float position, FP32 velocity,
// F/2, FDOT/6, mass, fp64 time) are gathered through the list, predicted to the current time in FP32 and summed in FP32
// (irr.avx.cpp:143-157, 319-353).  Here: one WARP per active particle, lanes stride the neighbour list; one particle is
// one 128-byte record (a single cache line per gather); prediction, separation and sums are fp64 (the B200 has the
// fp64 rate for it and the sum is gather-bound), the result is the fp64 statement itself.  set_jp / set_list only
// append to pinned host buffers; they are flushed in two launches by the next firr_vec, which returns its results
// through mapped pinned memory.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <vector>
#include <sys/time.h>

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    fprintf(stderr, "irr_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, \
            cudaGetErrorString(e_)); abort(); } } while (0)
#define FATAL(...) do { fprintf(stderr, "irr_b200: " __VA_ARGS__); fprintf(stderr, "\n"); abort(); } while (0)

namespace {

constexpr int REC = 16;            // doubles per particle record: x0[3] m | v0[3] t0 | a2[3] - | j6[3] -
constexpr int PCAP = 1 << 16;      // pending particle records before an early flush
constexpr int LCAP = 1 << 14;      // pending neighbour lists before an early flush
constexpr int ICAP = 1 << 16;      // active particles per launch
constexpr int SMALL_FLUSH = 2048;  // pending records up to which the scatter kernels read the mapped host buffers directly

__global__ void scatter_particles_kernel(int n, const int *__restrict__ addr, const double *__restrict__ rec,
                                         double *__restrict__ ptcl)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * (REC / 2)) return;
    const int p = k / (REC / 2), q = k - p * (REC / 2);
    reinterpret_cast<double2 *>(ptcl + (size_t)addr[p] * REC)[q] = reinterpret_cast<const double2 *>(rec + (size_t)p * REC)[q];
}

// slot = [addr, nnb, entries (1-based) ...]; one warp per pending list
__global__ void scatter_lists_kernel(int n, int slot_ints, int lstride, const int *__restrict__ slots, int *__restrict__ list,
                                     int *__restrict__ nnb)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const int *s = slots + (size_t)w * slot_ints;
    const int a = s[0], c = s[1];
    if (lane == 0) nnb[a] = c;
    for (int k = lane; k < c; k += 32) list[(size_t)a * lstride + k] = s[2 + k] - 1;
}

__device__ __forceinline__ void predict(const double *__restrict__ r, double ti, double xp[3], double vp[3], double &m)
{   // irr.avx.cpp:143-157 in fp64: pos = x0 + s (v0 + s (a2 + s j6)),  vel = v0 + 2 s (a2 + 1.5 s j6)
    const double2 a = reinterpret_cast<const double2 *>(r)[0], b = reinterpret_cast<const double2 *>(r)[1];
    const double2 c = reinterpret_cast<const double2 *>(r)[2], d = reinterpret_cast<const double2 *>(r)[3];
    const double2 e = reinterpret_cast<const double2 *>(r)[4], f = reinterpret_cast<const double2 *>(r)[5];
    const double2 g = reinterpret_cast<const double2 *>(r)[6], h = reinterpret_cast<const double2 *>(r)[7];
    const double x0[3] = {a.x, a.y, b.x}, v0[3] = {c.x, c.y, d.x}, a2[3] = {e.x, e.y, f.x}, j6[3] = {g.x, g.y, h.x};
    m = b.y;
    const double s = ti - d.y, s2 = 2.0 * s, s15 = 1.5 * s;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        xp[q] = x0[q] + s * (v0[q] + s * (a2[q] + s * j6[q]));
        vp[q] = v0[q] + s2 * (a2[q] + s15 * j6[q]);
    }
}

// One warp per active particle.  res = acc[ni][3] | jrk[ni][3] (the caller's layouts: plain copies on the host side);
// nn[i] = 1-based address of the nearest neighbour (0: empty list); inter: pair interactions (profile line).
__global__ void __launch_bounds__(128) firr_kernel(int ni, double ti, const int *__restrict__ addr, const double *__restrict__ ptcl,
                                                    const int *__restrict__ list, const int *__restrict__ nnb, int lstride,
                                                    double *__restrict__ res, int *__restrict__ nn, unsigned long long *__restrict__ inter)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= ni) return;
    const int i = addr[w] - 1;
    double xi[3], vi[3], mi;
    predict(ptcl + (size_t)i * REC, ti, xi, vi, mi);
    double f[6] = {0, 0, 0, 0, 0, 0};
    double r2min = 1.0e300;
    int jmin = INT_MAX;
    const int c = nnb[i];
    const int *nb = list + (size_t)i * lstride;
    for (int k = lane; k < c; k += 32) {
        const int j = nb[k];
        double xj[3], vj[3], mj;
        predict(ptcl + (size_t)j * REC, ti, xj, vj, mj);
        const double dx = xj[0] - xi[0], dy = xj[1] - xi[1], dz = xj[2] - xi[2];
        const double dvx = vj[0] - vi[0], dvy = vj[1] - vi[1], dvz = vj[2] - vi[2];
        const double r2 = dx * dx + dy * dy + dz * dz;
        const double rv = dx * dvx + dy * dvy + dz * dvz;
        const double rinv = rsqrt(r2), rinv2 = rinv * rinv;
        const double mr3 = mj * rinv * rinv2, al = -3.0 * rv * rinv2;           // :329-333
        f[0] += mr3 * dx; f[1] += mr3 * dy; f[2] += mr3 * dz;
        f[3] += mr3 * (dvx + al * dx); f[4] += mr3 * (dvy + al * dy); f[5] += mr3 * (dvz + al * dz);
        if (r2 < r2min || (r2 == r2min && j < jmin)) { r2min = r2; jmin = j; }     // :339-342: min r2, then min index
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 6; q++) f[q] += __shfl_xor_sync(0xffffffffu, f[q], o);
        const double r2o = __shfl_xor_sync(0xffffffffu, r2min, o);
        const int jo = __shfl_xor_sync(0xffffffffu, jmin, o);
        if (r2o < r2min || (r2o == r2min && jo < jmin)) { r2min = r2o; jmin = jo; }
    }
    if (lane < 3) res[(size_t)w * 3 + lane] = f[lane];
    else if (lane < 6) res[3 * (size_t)ni + (size_t)w * 3 + lane - 3] = f[lane];
    if (lane == 0) { nn[w] = c > 0 ? jmin + 1 : 0; if (c > 0) atomicAdd(inter, (unsigned long long)c); }
}

struct Irr {
    bool is_open = false;
    int dev = 0, nmax = 0, lmax = 0, lstride = 0, slot_ints = 0;
    cudaStream_t st = nullptr;
    double *ptcl = nullptr; int *list = nullptr, *nnb = nullptr;          // device state
    // pending updates (pinned host) and their device copies
    double *h_rec = nullptr, *d_rec = nullptr; int *h_paddr = nullptr, *d_paddr = nullptr; int np = 0;
    int *h_slots = nullptr, *d_slots = nullptr; int nl = 0;
    double *h_rec_dev = nullptr; int *h_paddr_dev = nullptr, *h_slots_dev = nullptr;     // device aliases of the (mapped) pending buffers
    std::vector<int> pslot, lslot;                                        // address -> pending slot (-1: none)
    unsigned long long *d_inter = nullptr;                                // pair interactions since the last profile line (device counter)
    int *h_addr = nullptr, *d_addr = nullptr, *h_addr_dev = nullptr;      // active list: mapped pinned (small blocks read it over PCIe) / device copy
    bool timing = false;                                                  // IRR_B200_TIMING=1 or irr_b200_set_timing: events around the kernel
    cudaEvent_t ev0 = nullptr, ev1 = nullptr; double kernel_ms = 0;       // device time of firr_kernel (CUDA events on S.st)
    double *h_res = nullptr, *d_res = nullptr; int *h_nn = nullptr, *d_nn = nullptr;   // mapped pinned results
    double time_grav = 0; unsigned long long num_inter = 0, num_fcall = 0, num_steps = 0;
} S;

double wtime() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + 1e-6 * tv.tv_usec; }

// sync = false (from the force call): the copies, the scatter and the force kernel are ordered on S.st and the force call
// synchronises once at its end, before the pinned staging buffers can be refilled
void flush_particles(bool sync = true)
{
    if (!S.np) return;
    // a block step's worth of records (a few hundred) is read by the scatter kernel straight from the mapped pinned buffer:
    // one enqueue instead of three on a call that is all latency; large batches go through the copy engine first
    const double *rec = S.h_rec_dev; const int *paddr = S.h_paddr_dev;
    if (S.np > SMALL_FLUSH) {
        CUDA_CHECK(cudaMemcpyAsync(S.d_rec, S.h_rec, sizeof(double) * REC * S.np, cudaMemcpyHostToDevice, S.st));
        CUDA_CHECK(cudaMemcpyAsync(S.d_paddr, S.h_paddr, sizeof(int) * S.np, cudaMemcpyHostToDevice, S.st));
        rec = S.d_rec; paddr = S.d_paddr;
    }
    const int threads = S.np * (REC / 2);
    scatter_particles_kernel<<<(threads + 255) / 256, 256, 0, S.st>>>(S.np, paddr, rec, S.ptcl);
    CUDA_CHECK(cudaGetLastError());
    if (sync) CUDA_CHECK(cudaStreamSynchronize(S.st));        // the pinned buffers are refilled right away
    for (int k = 0; k < S.np; k++) S.pslot[S.h_paddr[k]] = -1;
    S.np = 0;
}

void flush_lists(bool sync = true)
{
    if (!S.nl) return;
    const int *slots = S.h_slots_dev;
    if (S.nl > SMALL_FLUSH / 4) {
        CUDA_CHECK(cudaMemcpyAsync(S.d_slots, S.h_slots, sizeof(int) * (size_t)S.slot_ints * S.nl, cudaMemcpyHostToDevice, S.st));
        slots = S.d_slots;
    }
    scatter_lists_kernel<<<(S.nl * 32 + 127) / 128, 128, 0, S.st>>>(S.nl, S.slot_ints, S.lstride, slots, S.list, S.nnb);
    CUDA_CHECK(cudaGetLastError());
    if (sync) CUDA_CHECK(cudaStreamSynchronize(S.st));
    for (int k = 0; k < S.nl; k++) S.lslot[S.h_slots[(size_t)k * S.slot_ints]] = -1;
    S.nl = 0;
}

}  // namespace

extern "C" {

// reference: irr.avx.cpp:365-402, :567-569
void irr_simd_open_(int *nmaxp, int *lmaxp, int *rank)
{
    if (S.is_open) { fprintf(stderr, "irr_simd: it is already open\n"); return; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FATAL("no CUDA device available; this library has no CPU fallback");
    const char *gl = getenv("GPU_LIST");
    S.dev = (gl && *gl) ? atoi(gl) : 0;
    if (S.dev < 0 || S.dev >= ndev) FATAL("GPU_LIST names device %d but only %d are visible", S.dev, ndev);
    CUDA_CHECK(cudaSetDevice(S.dev));
    S.nmax = *nmaxp; S.lmax = *lmaxp;
    S.lstride = (S.lmax + 7) & ~7;
    S.slot_ints = S.lstride + 2;
    CUDA_CHECK(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
    CUDA_CHECK(cudaMalloc((void **)&S.ptcl, sizeof(double) * REC * ((size_t)S.nmax + 1)));
    CUDA_CHECK(cudaMemsetAsync(S.ptcl, 0, sizeof(double) * REC * ((size_t)S.nmax + 1), S.st));
    CUDA_CHECK(cudaMalloc((void **)&S.list, sizeof(int) * (size_t)S.lstride * S.nmax));
    CUDA_CHECK(cudaMalloc((void **)&S.nnb, sizeof(int) * (size_t)S.nmax));
    CUDA_CHECK(cudaMemsetAsync(S.nnb, 0, sizeof(int) * (size_t)S.nmax, S.st));
    CUDA_CHECK(cudaHostAlloc((void **)&S.h_rec, sizeof(double) * REC * PCAP, cudaHostAllocMapped));
    CUDA_CHECK(cudaHostGetDevicePointer((void **)&S.h_rec_dev, S.h_rec, 0));
    CUDA_CHECK(cudaMalloc((void **)&S.d_rec, sizeof(double) * REC * PCAP));
    CUDA_CHECK(cudaHostAlloc((void **)&S.h_paddr, sizeof(int) * PCAP, cudaHostAllocMapped));
    CUDA_CHECK(cudaHostGetDevicePointer((void **)&S.h_paddr_dev, S.h_paddr, 0));
    CUDA_CHECK(cudaMalloc((void **)&S.d_paddr, sizeof(int) * PCAP));
    CUDA_CHECK(cudaHostAlloc((void **)&S.h_slots, sizeof(int) * (size_t)S.slot_ints * LCAP, cudaHostAllocMapped));
    CUDA_CHECK(cudaHostGetDevicePointer((void **)&S.h_slots_dev, S.h_slots, 0));
    CUDA_CHECK(cudaMalloc((void **)&S.d_slots, sizeof(int) * (size_t)S.slot_ints * LCAP));
    CUDA_CHECK(cudaHostAlloc((void **)&S.h_addr, sizeof(int) * ICAP, cudaHostAllocMapped));
    CUDA_CHECK(cudaHostGetDevicePointer((void **)&S.h_addr_dev, S.h_addr, 0));
    CUDA_CHECK(cudaEventCreate(&S.ev0)); CUDA_CHECK(cudaEventCreate(&S.ev1));
    CUDA_CHECK(cudaMalloc((void **)&S.d_addr, sizeof(int) * ICAP));
    CUDA_CHECK(cudaHostAlloc((void **)&S.h_res, sizeof(double) * 6 * ICAP, cudaHostAllocMapped));
    CUDA_CHECK(cudaHostGetDevicePointer((void **)&S.d_res, S.h_res, 0));
    CUDA_CHECK(cudaHostAlloc((void **)&S.h_nn, sizeof(int) * ICAP, cudaHostAllocMapped));
    CUDA_CHECK(cudaHostGetDevicePointer((void **)&S.d_nn, S.h_nn, 0));
    S.pslot.assign((size_t)S.nmax + 1, -1); S.lslot.assign((size_t)S.nmax + 1, -1);
    CUDA_CHECK(cudaMalloc((void **)&S.d_inter, sizeof(unsigned long long)));
    CUDA_CHECK(cudaMemsetAsync(S.d_inter, 0, sizeof(unsigned long long), S.st));
    S.np = S.nl = 0;
    { const char *e = getenv("IRR_B200_TIMING"); S.timing = e && atoi(e) > 0; }
    S.time_grav = 0; S.num_inter = S.num_fcall = S.num_steps = 0;
    CUDA_CHECK(cudaStreamSynchronize(S.st));
    fprintf(stderr, "# Opening IRR lib. B200 ver. - rank: %d; nmax: %d, lmax: %d\n", *rank, S.nmax, S.lmax);
    S.is_open = true;
}

// reference: irr.avx.cpp:421-435
void irr_simd_close_(int *rank)
{
    if (!S.is_open) { fprintf(stderr, "irr_simd: it is already close\n"); return; }
    CUDA_CHECK(cudaSetDevice(S.dev));
    CUDA_CHECK(cudaStreamSynchronize(S.st));
    cudaFree(S.d_inter); cudaFree(S.ptcl); cudaFree(S.list); cudaFree(S.nnb); cudaFree(S.d_rec); cudaFree(S.d_paddr); cudaFree(S.d_slots); cudaFree(S.d_addr);
    cudaFreeHost(S.h_rec); cudaFreeHost(S.h_paddr); cudaFreeHost(S.h_slots); cudaFreeHost(S.h_addr); cudaFreeHost(S.h_res); cudaFreeHost(S.h_nn);
    CUDA_CHECK(cudaStreamDestroy(S.st));
    cudaEventDestroy(S.ev0); cudaEventDestroy(S.ev1);
    S = Irr();
    fprintf(stderr, "Closing IRR lib. B200 ver. - rank: %d\n", *rank);
}

// reference: irr.avx.cpp:404-419 (same line format)
void irr_simd_profile_(int *rank)
{
    if (!S.is_open || !S.num_fcall) return;
    CUDA_CHECK(cudaSetDevice(S.dev));
    CUDA_CHECK(cudaMemcpy(&S.num_inter, S.d_inter, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemset(S.d_inter, 0, sizeof(unsigned long long)));
    fprintf(stderr, "[R.%d B200 Irr.F ] Ncall: %llu <NI>: %d <NB>: %f grav: %f s, %f Gflops, %f usec\n", *rank, S.num_fcall,
            (int)(S.num_inter / S.num_fcall), (double)S.num_inter / (double)S.num_steps, S.time_grav,
            60.0 * (double)S.num_inter * 1e-9 / S.time_grav, 1e6 * S.time_grav / S.num_fcall);
    S.time_grav = 0; S.num_inter = S.num_fcall = S.num_steps = 0;
}

// reference: irr.avx.cpp:437-447, :576-586 (addr is 1-based)
void irr_simd_set_jp_(int *addr, double pos[3], double vel[3], double acc2[3], double jrk6[3], double *mass, double *time)
{
    if (!S.is_open) FATAL("irr_simd_set_jp called while the library is closed");
    const int a = *addr - 1;
    if (a < 0 || a >= S.nmax) FATAL("irr_simd_set_jp: address %d outside 1..%d", *addr, S.nmax);
    int k = S.pslot[a];
    if (k < 0) {
        if (S.np == PCAP) { CUDA_CHECK(cudaSetDevice(S.dev)); flush_particles(); }
        k = S.np++;
        S.pslot[a] = k;
        S.h_paddr[k] = a;
    }
    double *r = S.h_rec + (size_t)k * REC;
    r[0] = pos[0]; r[1] = pos[1]; r[2] = pos[2]; r[3] = *mass;
    r[4] = vel[0]; r[5] = vel[1]; r[6] = vel[2]; r[7] = *time;
    r[8] = acc2[0]; r[9] = acc2[1]; r[10] = acc2[2]; r[11] = 0.0;
    r[12] = jrk6[0]; r[13] = jrk6[1]; r[14] = jrk6[2]; r[15] = 0.0;
}

// reference: irr.avx.cpp:449-494, :587-592: nblist = [nnb, j1, j2, ...], 1-based
void irr_simd_set_list_(int *addr, int *nblist)
{
    if (!S.is_open) FATAL("irr_simd_set_list called while the library is closed");
    const int a = *addr - 1, c = nblist[0];
    if (a < 0 || a >= S.nmax) FATAL("irr_simd_set_list: address %d outside 1..%d", *addr, S.nmax);
    if (c < 0 || c > S.lstride) FATAL("irr_simd_set_list: %d neighbours exceed lmax = %d", c, S.lmax);
    int k = S.lslot[a];
    if (k < 0) {
        if (S.nl == LCAP) { CUDA_CHECK(cudaSetDevice(S.dev)); flush_lists(); }
        k = S.nl++;
        S.lslot[a] = k;
    }
    int *s = S.h_slots + (size_t)k * S.slot_ints;
    s[0] = a; s[1] = c;
    memcpy(s + 2, nblist + 1, sizeof(int) * c);
}

// reference: irr.avx.cpp:539-563, :593-602
void irr_simd_firr_vec_(double *ti, int *nip, int addr[], double acc[][3], double jrk[][3], int nnbid[])
{
    if (!S.is_open) FATAL("irr_simd_firr_vec called while the library is closed");
    const double t0 = wtime();
    CUDA_CHECK(cudaSetDevice(S.dev));
    flush_particles(false);
    flush_lists(false);
    const int ni = *nip;
    if (ni <= 0) CUDA_CHECK(cudaStreamSynchronize(S.st));
    for (int i0 = 0; i0 < ni; i0 += ICAP) {
        const int n = ni - i0 < ICAP ? ni - i0 : ICAP;
        memcpy(S.h_addr, addr + i0, sizeof(int) * n);
        for (int k = 0; k < n; k++)
            if (S.h_addr[k] < 1 || S.h_addr[k] > S.nmax) FATAL("irr_simd_firr_vec: address %d outside 1..%d", S.h_addr[k], S.nmax);
        // small blocks (the rule: NXTLEN is a few tens, intgrt.F:545): the kernel reads the addresses straight from the mapped
        // pinned buffer -- one enqueue less on a call that is all latency; large blocks get a device copy first
        const int *addr_dev = S.h_addr_dev;
        if (n > 1024) { CUDA_CHECK(cudaMemcpyAsync(S.d_addr, S.h_addr, sizeof(int) * n, cudaMemcpyHostToDevice, S.st)); addr_dev = S.d_addr; }
        if (S.timing) CUDA_CHECK(cudaEventRecord(S.ev0, S.st));
        firr_kernel<<<(n + 3) / 4, 128, 0, S.st>>>(n, *ti, addr_dev, S.ptcl, S.list, S.nnb, S.lstride, S.d_res, S.d_nn, S.d_inter);
        CUDA_CHECK(cudaGetLastError());
        if (S.timing) CUDA_CHECK(cudaEventRecord(S.ev1, S.st));
        CUDA_CHECK(cudaStreamSynchronize(S.st));
        if (S.timing) { float ms = 0.f; CUDA_CHECK(cudaEventElapsedTime(&ms, S.ev0, S.ev1)); S.kernel_ms += ms; }
        memcpy(&acc[i0][0], S.h_res, sizeof(double) * 3 * (size_t)n);
        memcpy(&jrk[i0][0], S.h_res + 3 * (size_t)n, sizeof(double) * 3 * (size_t)n);
        memcpy(nnbid + i0, S.h_nn, sizeof(int) * (size_t)n);
    }
    S.time_grav += wtime() - t0;
    S.num_fcall++;
    S.num_steps += ni;
}

int irr_b200_version(void) { return 2; }

// Additive batch forms of set_jp / set_list (one call for the particles a block step has just advanced / the rows a
// regular block has just renewed) -- the same pending buffers, flushed by the next force call.
void irr_b200_set_jp_batch_(int *n, int addr[], double pos[][3], double vel[][3], double acc2[][3], double jrk6[][3],
                            double mass[], double time[])
{
    for (int k = 0; k < *n; k++) irr_simd_set_jp_(addr + k, pos[k], vel[k], acc2[k], jrk6[k], mass + k, time + k);
}
// lists[k] (stride *lstride ints) = [nnb, j1, j2, ...] of particle addr[k], 1-based
void irr_b200_set_list_batch_(int *n, int addr[], int *lstride, int lists[])
{
    for (int k = 0; k < *n; k++) irr_simd_set_list_(addr + k, lists + (size_t)k * *lstride);
}
// out[0] = device ms of the force kernel since open / the last call of this function, out[1] = force calls,
// out[2] = pair interactions (CUDA events on the library's stream)
void irr_b200_set_timing(int on) { S.timing = on != 0; }
// The particle table on the device (record of address 1; REC = 16 doubles per particle: x0[3] m | v0[3] t0 | a2[3] . | j6[3] .)
// for a consumer on the same device (gpunb_b200_predict_send_records_), and the call that makes it complete: pending
// set_jp / set_list reach the device and the stream is drained.
const double *irr_b200_particle_records_(int *stride)
{
    if (!S.is_open) FATAL("irr_b200_particle_records called while the library is closed");
    if (stride) *stride = REC;
    return S.ptcl;
}
void irr_b200_flush_(void)
{
    if (!S.is_open) FATAL("irr_b200_flush called while the library is closed");
    CUDA_CHECK(cudaSetDevice(S.dev));
    flush_particles(false);
    flush_lists(false);
    CUDA_CHECK(cudaStreamSynchronize(S.st));
}
void irr_b200_counters(double out[3])
{
    unsigned long long inter = 0;
    if (S.is_open) { CUDA_CHECK(cudaSetDevice(S.dev)); CUDA_CHECK(cudaMemcpy(&inter, S.d_inter, sizeof(inter), cudaMemcpyDeviceToHost)); }
    out[0] = S.kernel_ms; out[1] = (double)S.num_fcall; out[2] = (double)inter;
    S.kernel_ms = 0;
}

}  // extern "C"
