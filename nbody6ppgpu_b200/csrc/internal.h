// internal.h -- what the translation units of libgpunb_b200.so share besides the public C-ABI (not installed, not exported).
#pragma once
#include <cuda_runtime.h>

#define GPUNB_HIDDEN __attribute__((visibility("hidden")))

// The snapshot of the last gpunb_send_ / gpunb_b200_predict_send_ as it sits on the root device: the caller's fp64 arrays
// m[nj], x[nj][3], v[nj][3] (particle J of the Fortran program at index J - IFIRST), and the stream every call of the
// library is ordered on.
struct GpunbSnapshotView {
    int device = -1;
    cudaStream_t stream = nullptr;
    int nj = 0, nbmax = 0;
    const double *m = nullptr, *x = nullptr, *v = nullptr;
    double *counters = nullptr;          // GPUNB_B200_CTR_* array of the library
    // rows of the last gpunb_regf_ call as they were delivered ([ni][lmax], row i of the call), still on the device
    const int *last_rows = nullptr; int last_rows_ni = 0, last_rows_lmax = 0;
    void *(*pinned_alias)(const void *p, size_t bytes) = nullptr;     // device alias of caller-pinned host memory, or NULL
};
GPUNB_HIDDEN bool gpunb_b200_internal_snapshot(GpunbSnapshotView *out);      // false: library closed or nothing sent yet
GPUNB_HIDDEN void gpunb_b200_internal_regcor_close();                        // frees the buffers of regcor_b200.cu (gpunb_close_)
// Size of every host-side OpenMP team of the library (staging copies, row delivery, list packing, per-device enqueue): the
// CALLER's default team size unless GPUNB_B200_HOST_THREADS says otherwise.  libgomp re-docks its threads whenever
// consecutive parallel regions ask for different team sizes; inside an OpenMP host program (NBODY6++ is one, and so is the
// reference library: every region of gpunb.velocity.cu / reg.avx.cpp runs with the default team) teams of 4 between the
// caller's teams of 16 cost 0.2 s per N-body time unit at N = 16k (profiles/r2zk_host_team_ab.txt).
GPUNB_HIDDEN int gpunb_b200_internal_host_team();
