/*
 * gpunb_b200.h -- C-ABI of libgpunb_b200.so, a B200-native (sm_100a) drop-in
 * for the regular-force library of NBODY6++GPU.
 *
 * Part 1 is EXACTLY the Fortran-callable ABI the reference exports (lower
 * case, trailing underscore, every scalar by reference, INTEGER*4 / REAL*8,
 * column-major X(3,N) == C x[N][3]).  Each prototype cites the reference
 * definition it replaces.  Part 2 are additive extension entry points
 * (prefix gpunb_b200_) used by the harness, bench.py and the multi-process
 * NCCL mode; a Fortran caller never needs them.
 *
 * All buffers are caller-owned, pageable, valid only during the call.
 */
#ifndef GPUNB_B200_H
#define GPUNB_B200_H
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- Part 1: the reference ABI (drop-in) ---------------- */

/* reference: src/Main/gpunb.velocity.cu:905 (GPUNB_devinit, :574-626).
 * Device discovery; honours GPU_LIST="0 1 ..." (:582-591); prints the banner. Idempotent. */
void gpunb_devinit_(int *irank);

/* reference: gpunb.velocity.cu:908 (GPUNB_open, :628-666). nbmax = largest nj of the session
 * (NTOT+10).  Re-open while open only warns (:636-639). */
void gpunb_open_(int *nbmax, int *irank);

/* reference: gpunb.velocity.cu:911 (GPUNB_close, :668-701). Close while closed only warns. */
void gpunb_close_(void);

/* reference: gpunb.velocity.cu:914 (GPUNB_send, :703-729). Uploads ALL nj j-particles
 * (mass, position, velocity; fp64) -- the snapshot every following regf call sums over. */
void gpunb_send_(int *nj, double mj[], double xj[][3], double vj[][3]);

/* reference: gpunb.velocity.cu:921-935 (GPUNB_regf, :731-880).
 * For i in [0,ni): acc/jrk/pot over all j outside the neighbour sphere; list[i*lmax+0] =
 * count (or -(count) when count > nnbmax, the reg.avx.cpp:320-321 encoding), list[i*lmax+1..] =
 * 0-based j indices, strictly ascending, self included.  m_flag=1: criterion r2min < mj*h2. */
void gpunb_regf_(int *ni, double h2[], double dtr[], double xi[][3], double vi[][3],
                 double acc[][3], double jrk[][3], double pot[],
                 int *lmax, int *nnbmax, int *list, int *m_flag);

/* reference: gpunb.velocity.cu:936 (GPUNB_profile, :882-902). stderr perf line, counters reset. */
void gpunb_profile_(int *irank);

/* reference: src/Main/gpupot.gpu.cu:117-126 (gpupot, :61-114). pot[ii] = sum_{j, r>0} m_j/r_ij
 * for i = istart-1+ii (istart is 1-based), ii in [0,ni), over all n particles. */
void gpupot_(int *irank, int *istart, int *ni, int *n, double m[], double x[][3], double pot[]);

/* ---------------- Part 2: additive extensions ---------------- */

/* Library/ABI version and a string naming the kernels compiled in. */
int         gpunb_b200_version(void);
const char *gpunb_b200_build_info(void);

/* Number of GPUs this process drives (after devinit). */
int gpunb_b200_num_devices(void);
/* Resident warps of the pair kernel on device 0 (SMs x CTAs per SM x 4): a block of 32 * (resident_warps / S) i-particles
 * with integer S fills the machine exactly (1024 at S = 74 on a B200; 2048 / 4736 / 9472 at S = 37 / 16 / 8). */
int gpunb_b200_resident_warps(void);

/* Counters since the last reset (doubles, see GPUNB_B200_CTR_*): device time of the pair kernel
 * measured with CUDA events on its launching stream, launches, bytes moved, interactions. */
enum {
    GPUNB_B200_CTR_GRAV_MS = 0,     /* sum of regf pair-kernel durations, ms (device 0)      */
    GPUNB_B200_CTR_GRAV_LAUNCHES,   /* pair-kernel launches (device 0)                        */
    GPUNB_B200_CTR_LAUNCHES,        /* all kernel launches by this library (all devices)      */
    GPUNB_B200_CTR_H2D_BYTES,
    GPUNB_B200_CTR_D2H_BYTES,
    GPUNB_B200_CTR_INTERACTIONS,    /* sum ni*nj as the reference counts (gpunb.velocity.cu:747) */
    GPUNB_B200_CTR_MERGE_MS,        /* sum of merge-kernel durations, ms (device 0)           */
    GPUNB_B200_CTR_POT_MS,          /* sum of gpupot kernel durations, ms                     */
    GPUNB_B200_CTR_NEAR_TILES,      /* (warp, j-tile) visits that ran the full NEAR body (GPUNB_B200_STATS=1) */
    GPUNB_B200_CTR_ALL_TILES,       /* all (warp, j-tile) visits (GPUNB_B200_STATS=1)          */
    /* per-kernel device timeline of resident sweeps (GPUNB_B200_TIMELINE=1: CUDA events around every kernel) */
    GPUNB_B200_CTR_TL_BLOCKS,       /* i-blocks timed                                          */
    GPUNB_B200_CTR_TL_ISORT_MS,     /* isort_kernel                                            */
    GPUNB_B200_CTR_TL_REGF_MS,      /* regf_kernel                                             */
    GPUNB_B200_CTR_TL_MERGE_MS,     /* merge_kernel                                            */
    GPUNB_B200_CTR_TL_EXCH_MS,      /* combine_kernel (multi-GPU; includes waiting for the slowest shard's flag) */
    /* host-side wall-clock buckets of gpunb_regf_ (ms): pack + NaN check of the i-block, enqueueing copies and
     * kernels, waiting for the device, copying result rows into the caller's arrays */
    GPUNB_B200_CTR_HOST_PACK_MS,
    GPUNB_B200_CTR_HOST_ENQUEUE_MS,
    GPUNB_B200_CTR_HOST_WAIT_MS,
    GPUNB_B200_CTR_HOST_SCATTER_MS,
    /* gpunb_send_: calls, wall-clock ms in total / in the chunked pinned staging + upload loop, and device ms of the
     * tile construction (keys, radix sort, tilepack) */
    GPUNB_B200_CTR_SENDS,
    GPUNB_B200_CTR_SEND_MS,
    GPUNB_B200_CTR_SEND_STAGE_MS,
    GPUNB_B200_CTR_SEND_TILES_MS,
    GPUNB_B200_CTR_TRANSPOSED_TILES, /* NEAR (warp, j-tile) visits handled by the transposed path (GPUNB_B200_STATS=1) */
    GPUNB_B200_CTR_SENDS_ORDER_KEPT,  /* snapshots whose tiles were re-packed in the kept Hilbert order (no sort) */
    GPUNB_B200_CTR_HOST_RENDEZVOUS_MS, /* i-slice mode: publishing this rank's slice and waiting for the other ranks' (ms) */
    GPUNB_B200_CTR_REGCOR_MS,         /* gpunb_b200_regcor_: wall-clock ms inside the calls (pack, upload, kernel, results) */
    GPUNB_B200_CTR_REGCOR_ROWS,       /* ... rows handled */
    GPUNB_B200_CTR_COUNT
};
void gpunb_b200_get_counters(double out[GPUNB_B200_CTR_COUNT]);
void gpunb_b200_reset_counters(void);

/* Device-resident sweep (bench "value" leg): i-particles are j-particles [i0, i0+ni) of the
 * snapshot already uploaded by gpunb_send_; h2/dtr (length nj, host) are uploaded once by
 * gpunb_b200_set_radii.  Blocks of `block` i-particles are launched back to back with no host
 * synchronisation; results stay on the device (last block readable with fetch).  Returns the
 * device time in ms between the first and the last kernel (CUDA events). */
void  gpunb_b200_set_radii(int *nj, double h2[], double dtr[]);
float gpunb_b200_sweep_resident(int *i0, int *ni, int *block, int *lmax, int *nnbmax, int *m_flag);
/* Copy the results of the LAST resident block (n_last rows) to host arrays laid out as regf's. */
void  gpunb_b200_fetch_last(int *n_last, double acc[][3], double jrk[][3], double pot[],
                            int *lmax, int *list);

/* Device-resident predictor: replaces the host predictor xbpredall (src/Main/xbpredall.f:15-26,
 * called at intgrt.F:516-520) AND the gpunb_send_ upload that follows it before every regular block (intgrt.F:912-918).
 * Fortran-callable like the reference entries (scalars by reference, REAL*8, X(3,N) == x[N][3]); strictly additive:
 * a caller that never uses them keeps the reference behaviour.
 *   state_all     BODY, X0, X0DOT, F, FDOT, T0 of all nj j-particles (arrays start at IFIRST) in the integrator's own
 *                 conventions: F is half the force, FDOT one sixth of its derivative (xbpredall.f:20-25).  Call once,
 *                 and again whenever the particle table is re-ordered (KS creation / termination, escapers).
 *   state_update  the same six quantities for the n particles idx[k] (0-based, relative to the j array, like the
 *                 returned neighbour indices) the integrator has just advanced; entry k belongs to particle idx[k].
 *   predict_send  X = ((FDOT*S + F)*S + X0DOT)*S + X0, XDOT = (FDOT*1.5S + F)*2S + X0DOT with S = time - T0 for the
 *                 first nj particles, on the device, in fp64 without FMA contraction (bit-for-bit an unfused host
 *                 build), followed by the tile construction of gpunb_send_.  gpunb_regf_ then works as usual.
 *   get_predicted predicted x / xdot of the particles idx[k] from that snapshot. */
void gpunb_b200_state_all_(int *nj, double body[], double x0[][3], double x0dot[][3], double f[][3], double fdot[][3],
                           double t0[]);
void gpunb_b200_state_update_(int *n, int idx[], double body[], double x0[][3], double x0dot[][3], double f[][3],
                              double fdot[][3], double t0[]);
void gpunb_b200_predict_send_(int *nj, double *time);
void gpunb_b200_get_predicted_(int *n, int idx[], double x[][3], double xdot[][3]);
/* predict_send from particle records that ANOTHER library keeps on the same device: records_dev is a DEVICE pointer to the
 * record of the first j-particle, *stride doubles apart (>= 16, even), laid out x0[3] m | v0[3] t0 | f[3] . | fdot[3] . with
 * F = force/2, FDOT = derivative/6 -- the table of libirr_b200.so (irr_b200_particle_records_, include/irr_b200.h), which the
 * integrator refreshes with irr_simd_set_jp_ after every corrector.  One copy of the state then serves the irregular force
 * AND the predictor of the regular force: no gpunb_b200_state_update_, no upload before a regular block.  The caller makes
 * the records complete first (irr_b200_flush_).  One process, one GPU.  Same arithmetic as predict_send. */
void gpunb_b200_predict_send_records_(int *nj, double *time, const double *records_dev, int *stride);

/* Pipeline depth: nslot = pipeline slots a resident sweep cycles through (1 = one block after the other on one
 * stream), nsub = sub-blocks one gpunb_regf_ call is split into (1 = the whole i-block in one pair-kernel launch).
 * A call is split only while every sub-block keeps >= 256 i-particles and ~1.5e8 pairs (the pair kernel must outlast
 * the fixed costs of a launch); nsub = -k forces k sub-blocks of any size (tests).  Other values leave the setting
 * unchanged.  Environment: GPUNB_B200_NSLOT / GPUNB_B200_NSUB. */
void  gpunb_b200_set_tuning(int nslot, int nsub);
/* Caller-pinned arrays (optional).  The reference ABI hands the library pageable arrays, so gpunb_send_ stages the
 * snapshot through pinned memory and gpunb_regf_ copies result rows into the caller's arrays on the host.  A caller
 * whose arrays live for the whole run (the Fortran COMMON blocks / static GPUACC, GPUJRK, GPUPHI, LISTGP of
 * util_gpu.F:14) can pin them ONCE: gpunb_send_ then DMAs straight from them and the kernels write acc / jrk / pot and
 * the list rows straight into them over PCIe.  Detection is per call and per array range; results are identical.
 * pin_host returns 0 on success (non-zero: the range stays pageable and the staged path is used).  Unpin before the
 * memory is freed. */
int   gpunb_b200_pin_host_(void *ptr, long long *bytes);
void  gpunb_b200_unpin_host_(void *ptr);

/* Hilbert order of the j-tiles across snapshots (gpunb_send_ / gpunb_b200_predict_send_).  The tiles are ALWAYS re-packed from
 * the current positions (boxes and offsets recomputed: lists bit-exact, forces within the same bars); what may be kept is
 * the permutation, which saves the eight launches of the sort (72 of the 150 us of a gpunb_send_ at N = 10^4, 0.19 of the
 * 0.29 ms of a predict_send at N = 10^6).  k = 0 (default): adaptive on one GPU -- the order is kept while the summed
 * half-extents of the tiles (an exact integer sum: the decision is reproducible) stay within 10 % of their value right
 * after the last sort, for at most 64 snapshots; sharded runs always sort.  k = 1: always sort (results are a function of
 * the snapshot alone).  k > 1: sort every k-th snapshot.  Environment: GPUNB_B200_RESORT_EVERY. */
void  gpunb_b200_set_resort_every(int k);

/* i-slice mode (one process per GPU, after gpunb_b200_nccl_init; environment GPUNB_B200_ISLICE=1).  Off (default): every
 * rank passes the SAME i-block to gpunb_regf_ and receives the complete result (replicated data).  On: gpunb_regf_ is a
 * collective call in which every rank passes ITS OWN i-slice -- ni may differ between ranks and may be 0, lmax / nnbmax /
 * m_flag must agree -- and receives the results of that slice only.  This is the calling pattern of NBODY6++'s MPI build,
 * where every rank integrates its share of the regular block (intgrt.F:982-1231): the ranks' slices meet in a shared-
 * memory segment of the node, every GPU runs ONE pair-kernel launch on the union of the slices against its j-shard, and
 * each rank's combine kernel pulls only its own rows from the peers over NVLink.  Every rank must make the same number
 * of gpunb_regf_ calls (pad with ni = 0). */
void  gpunb_b200_set_islice(int on);

/* Pairs (i x local j) a sub-block of gpunb_regf_ must keep for the call to be split (default 1.5e8; tests lower it). */
void  gpunb_b200_set_sub_pairs(double pairs);
/* gpunb_regf_ calls with fewer than `pairs` (ni x nj) pairs keep the caller's order of the i-block instead of Morton-sorting
 * it (default 2.5e7: the sort would add more latency than it saves; 0 = always sort).  Environment: GPUNB_B200_ISORT_PAIRS. */
void  gpunb_b200_set_isort_pairs(double pairs);

/* One process per GPU (after gpunb_b200_nccl_init): gpunb_send_ of at least min_nj particles uploads only this rank's 1/R slice
 * of the snapshot over its own PCIe link and completes it on every GPU with one all-gather over NVLink ("send
 * scatters instead of broadcasts"); smaller snapshots, where the all-gather's latency would exceed the saving, are uploaded
 * whole by every rank.  Default 40000; negative: never.  Every rank must use the same value (the all-gather is collective).
 * Environment: GPUNB_B200_SEND_SCATTER_MIN. */
void  gpunb_b200_set_send_scatter(int min_nj);

/* Work items per resident warp slot of a gpunb_regf_ call that is ONE pair-kernel launch (1 ... 4, default 4: four times
 * shorter work items against the tail of a launch that runs alone -- DESIGN.md section 3; calls split into sub-blocks
 * are not affected).  Lists are identical for every setting, sums differ in the last bits (more fp64 partials per i).
 * Environment: GPUNB_B200_REGF_OVERSUB. */
void  gpunb_b200_set_regf_oversub(int k);

/* Sub-block sizes of one gpunb_regf_ call: 0 = equal (default), 1 = tapering (weights 7:5:3:1 for four sub-blocks;
 * measured, no gain).  Environment: GPUNB_B200_TAPER. */
void  gpunb_b200_set_taper(int on);

/* A-B builds only (make EXTRA=-DNEAR_SCALAR_AB; gpunb_b200_has_near_scalar_ab() == 1): 1 = NEAR tiles through the
 * scalar pair body, 0 = through the packed (f32x2) pair body (bit-for-bit the same results), -1 = follow the
 * environment variable GPUNB_B200_NEAR_EXACT.  No effect in the default build (packed body only). */
void  gpunb_b200_set_near_exact(int on);
int   gpunb_b200_has_near_scalar_ab(void);

/* FP32 pipe microbenchmark: returns achieved scalar/packed FFMA TFLOP/s on device 0
 * (mode 0: FFMA, 1: FFMA2 (f32x2), 2: FADD2, 3: FMUL2, 4: MUFU.RSQ Gop/s, 5: FFMA2+ALU mix). */
double gpunb_b200_fp32_microbench(int mode, int iters);

/* Far-body microbenchmark (the 27-op pair body in isolation); returns Gint/s.  mode bit0: no MUFU, bit1: operands
 * from registers instead of LDS.128, mode>=4: two i-particles per lane. */
double gpunb_b200_farbody_microbench(int mode, int iters, int ctas_per_sm);

/* Tuning aid: per-work-item {start, end} %globaltimer of the last pair-kernel launch (GPUNB_B200_STATS=2). */
int gpunb_b200_debug_wtimes(unsigned long long *out, int max_items);

/* Multi-GPU j-sharding (the exchange step of the path; reference: the j split over GPUs inside
 * gpunb.velocity.cu:713-715 and its host-side fp64 combine :822-879).
 *
 *  (A) one process, several GPUs (the reference's own model): export GPUNB_B200_MULTI=1 (and optionally
 *      GPU_LIST); nothing else changes for the caller.  Shards are combined by a kernel on the first GPU
 *      that pulls partial sums and neighbour rows from the peers over NVLink P2P.
 *  (B) one process per GPU (MPI / torchrun): rank 0 creates a 128-byte ncclUniqueId with
 *      gpunb_b200_nccl_unique_id, the caller broadcasts it, every rank calls gpunb_b200_nccl_init after
 *      gpunb_devinit_ and before gpunb_open_.  From then on EVERY rank makes identical calls (same snapshot,
 *      same i-blocks) and every rank receives the complete result; rank r sums over every R-th tile of the
 *      Hilbert-sorted j-set.  NCCL only bootstraps (cudaIpc handles, barriers): per call, each rank's merge kernel
 *      raises a flag in every peer's exchange buffer over NVLink and the combine kernel pulls the 64 B per i of
 *      partial sums and the valid part of the neighbour rows from the peers' memory.  gpupot_ uses one
 *      ncclAllGather of the partial potentials.
 * Return 0 on success. */
int  gpunb_b200_nccl_unique_id(unsigned char id128[128]);
int  gpunb_b200_nccl_init(int rank, int nranks, const unsigned char id128[128]);
void gpunb_b200_nccl_finalize(void);

/* ---------------- Part 3: neighbour-list bookkeeping after gpunb_regf_ ----------------
 * Replaces, for the ni rows of one regular block in ONE call, what the Fortran caller does per particle on the host:
 *   src/Main/util_gpu.F:102-111     row of gpunb_regf_ -> NLIST: + IFIRST, self dropped
 *   src/Main/regcor_gpu.F:267-336   NBLOSS / NBGAIN / JJLIST: old list LIST(:,I) against NLIST
 *   src/Main/regcor_gpu.F:338-420   lost members with STEP(J) <= SMIN inside 2 RS are put back (NBSMIN, FREG / FDR corrected)
 *   src/Main/regcor_gpu.F:425-470   DFIRR / DFD: - pair terms of the lost members + pair terms of the gained ones, fp64
 * X, XDOT, BODY of those lines are the predicted values of the block = the snapshot the last gpunb_send_ /
 * gpunb_b200_predict_send_ left on the device: particle J sits at index J - IFIRST, nothing is uploaded again.
 * Fortran-callable (scalars by reference, INTEGER*4 / REAL*8).  All particle numbers are the Fortran program's (1-based).
 *   index_i[ni]        particle I of each row
 *   new_list[ni][lmax] in: rows exactly as gpunb_regf_ returned them ([count, 0-based j ascending, self included]; rows with
 *                      a negative count are passed through untouched -- the overflow retry of util_gpu.F:71-97 comes first);
 *                      out: NLIST = [NNB, members ascending] after self removal and retention
 *   old_list[ni][lmax] LIST(1:LMAX, I) = [NNB0, members ascending], or NULL: the rows of the device-resident list store
 *                      (gpunb_b200_lists_put_; once the store exists every call commits the final NLIST of its rows to it,
 *                      so a caller that changes lists nowhere else never uploads a list)
 *   rs2[ni]            RS(I)**2 as regcor sees it at entry (regcor_gpu.F:41)
 *   step               STEP(IFIRST:NTOT) on the host (uploaded by this call), or NULL: the resident copy kept by
 *                      gpunb_b200_steps_all_ / _update_; without either no lost member is ever retained (JMIN stays 0)
 *   freg, fdr          in/out [ni][3]: FREG, FDR (retention subtracts the retained members' terms)
 *   dfirr, dfd         in/out [ni][3]: the caller passes zeros (regcor_gpu.F:262-263) or its NNB = 0 values (:121-130)
 *   nbloss, nbgain     out [ni]
 *   jjlist             out [ni][2*lmax]: JJLIST(1..NBLOSS) lost, JJLIST(NNB0+1..NNB0+NBGAIN) gained (JJLIST(k) = jjlist[k-1])
 *   nbsmin             out: members retained in this call (the caller adds it to NBSMIN, regcor_gpu.F:363-365)
 * Every fp64 operation is the single IEEE operation the Fortran names, in its order: results are bit for bit those of an
 * unfused host build (tests/test_regcor_gpu.py against oracle/regcor_oracle.c). */
void gpunb_b200_regcor_(int *ni, int index_i[], int *ifirst, int *n, int *ntot, int *lmax, int new_list[], int old_list[],
                        double rs2[], double step[], double *smin, int *nnbmax, double freg[][3], double fdr[][3],
                        double dfirr[][3], double dfd[][3], int nbloss[], int nbgain[], int jjlist[], int *nbsmin);
/* The same bookkeeping for the rows of the LAST gpunb_regf_ call, taken from the copy that call left on the device (ni and
 * lmax must be that call's; index_i[k] names the particle of its row k; not available in i-slice mode): new_list is output
 * only.  With old_list = NULL (resident list store) and step = NULL (resident steps) no list crosses PCIe on the way in. */
void gpunb_b200_regcor_last_(int *ni, int index_i[], int *ifirst, int *n, int *ntot, int *lmax, int new_list[], int old_list[],
                             double rs2[], double step[], double *smin, int *nnbmax, double freg[][3], double fdr[][3],
                             double dfirr[][3], double dfd[][3], int nbloss[], int nbgain[], int jjlist[], int *nbsmin);
/* Device-resident list store (one row of lmax entries per particle number): lists[k] = LIST(1:LMAX, index_i[k]).
 * put after FPOLY0 and whenever the caller edits a list itself (KS, CHECKL, ...); get reads rows back. */
void gpunb_b200_lists_put_(int *n, int index_i[], int *lmax, int lists[]);
void gpunb_b200_lists_get_(int *n, int index_i[], int *lmax, int lists[]);
/* Resident STEP of the snapshot particles (idx 0-based relative to the j array, like gpunb_b200_state_update_). */
void gpunb_b200_steps_all_(int *nj, double step[]);
void gpunb_b200_steps_update_(int *n, int idx[], double step[]);

#ifdef __cplusplus
}
#endif
#endif
