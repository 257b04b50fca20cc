/*
 * irr_b200.h -- C-ABI of libirr_b200.so: the reference's irregular-force interface (irr_simd_*).
 *
 * The fp64 statement it implements is pinned against the reference's AVX library on the CPU (tests/test_irr_cpu.py) and
 * the CUDA library reproduces it to 2e-14 on a B200 (tests/test_irr_gpu.py).  The regular-force library (gpunb_b200.h)
 * does not depend on it.  Semantics that differ from the AVX library, both invisible to a conforming caller: set_jp /
 * set_list are buffered on the host and reach the device with the next force call; an empty list returns nnbid = 0.
 *
 * Fortran-callable like the reference (trailing underscore, scalars by reference, 1-based particle addresses).
 */
#ifndef IRR_B200_H
#define IRR_B200_H
#ifdef __cplusplus
extern "C" {
#endif

/* reference: src/Main/irr.avx.cpp:365-402, :567-569.  nmax particles, lists of at most lmax entries. */
void irr_simd_open_(int *nmax, int *lmax, int *rank);
/* reference: irr.avx.cpp:421-435, :570-572 */
void irr_simd_close_(int *rank);
/* reference: irr.avx.cpp:404-419, :573-575 (stderr line, counters reset) */
void irr_simd_profile_(int *rank);
/* reference: irr.avx.cpp:437-447, :576-586.  X0, X0DOT, F/2, FDOT/6, BODY, T0 of particle addr. */
void irr_simd_set_jp_(int *addr, double pos[3], double vel[3], double acc2[3], double jrk6[3], double *mass, double *time);
/* reference: irr.avx.cpp:449-494, :587-592.  nblist = [nnb, j1, ..., j_nnb], 1-based addresses. */
void irr_simd_set_list_(int *addr, int *nblist);
/* reference: irr.avx.cpp:539-563, :593-602.  Force / derivative over each active particle's list at time ti (every
 * involved particle predicted to ti) and the address of its nearest neighbour. */
void irr_simd_firr_vec_(double *ti, int *ni, int addr[], double acc[][3], double jrk[][3], int nnbid[]);

int irr_b200_version(void);

/* Additive batch forms: the n particles a block step has just advanced / the n lists a regular block has just renewed in
 * ONE call (entry k belongs to particle addr[k]; lists[k] has stride *lstride ints and holds [nnb, j1, ...], 1-based). */
void irr_b200_set_jp_batch_(int *n, int addr[], double pos[][3], double vel[][3], double acc2[][3], double jrk6[][3],
                            double mass[], double time[]);
void irr_b200_set_list_batch_(int *n, int addr[], int *lstride, int lists[]);
/* out[0] = device ms of the force kernel since the last call of this function (CUDA events on the library's stream),
 * out[1] = force calls, out[2] = pair interactions since open / the last irr_simd_profile_. */
void irr_b200_counters(double out[3]);
/* The particle table on the device, for a consumer on the same device: returns a DEVICE pointer to the record of address 1;
 * *stride = doubles per record (16: x0[3] m | v0[3] t0 | F/2 [3] . | FDOT/6 [3] .).  irr_b200_flush_ sends the pending
 * set_jp / set_list to the device and drains the stream: afterwards the table is what the integrator last set.
 * gpunb_b200_predict_send_records_ (include/gpunb_b200.h) predicts the regular-force snapshot from it. */
const double *irr_b200_particle_records_(int *stride);
void irr_b200_flush_(void);
/* Kernel timing (two event records and one query per force call) is off by default; on: IRR_B200_TIMING=1 or this call. */
void irr_b200_set_timing(int on);

#ifdef __cplusplus
}
#endif
#endif
