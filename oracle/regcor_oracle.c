/* regcor_oracle.c -- CPU restatement of the neighbour-list bookkeeping the Fortran caller does after every gpunb_regf_
 * (SURVEY.md section 8f rank 4).  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load it; the product path (nbody6ppgpu_b200/csrc/regcor_b200.cu) never does.
 *
 * PARITY PINNED BY INTERPRETATION: the reference for this row is Fortran (src/Main/util_gpu.F:102-111,
 * src/Main/regcor_gpu.F:263-459), this image has no Fortran compiler and the reference has no tests or golden vectors for
 * it.  The reference's own TEXT is therefore executed statement by statement by oracle/f77_interp.py
 * (a fixed-form Fortran 77 subset interpreter: IEEE double operations in the order the statements name, no contraction)
 * on seeded rows -- oracle/make_regcor_golden.py, source read where it lies under /root/reference -- and its outputs are
 * the golden vectors tests/golden/regcor_f77_*.npz; tests/test_regcor_cpu.py holds this file to them (integers equal, fp64
 * bit for bit) and, where the reference is present, to the interpreter run live.  What this pin does NOT cover: a Fortran
 * compiler's freedom to contract a*b+c into an FMA (the vectors are the unfused evaluation, which is also what the CUDA
 * path computes).  Two further independent statements cross-check it: a hand-made GO TO transcription and plain set
 * differences / vectorised sums (tests/regcor_cases.py).  Built with -ffp-contract=off so that every fp64 operation is the
 * single IEEE operation the Fortran statement names, in the order it names them.
 *
 * Per row (one i-particle I of the regular block):
 *   1. util_gpu.F:102-111   the row gpunb_regf_ returned ([count, 0-based j ascending, self included]) becomes NLIST:
 *                           ITEMP = j + IFIRST, self dropped, count = L1 - 1.
 *   2. regcor_gpu.F:267-336 NBLOSS / NBGAIN by a two-pointer comparison of the old list LIST(:,I) with NLIST (both strictly
 *                           ascending): lost members -> JJLIST(1..NBLOSS), gained -> JJLIST(NNB0+1..NNB0+NBGAIN);
 *                           JMIN != 0 when a lost member has STEP(J) < SMIN.  NNB0 = 0: everything is gained (:271-283).
 *   3. regcor_gpu.F:338-420 when JMIN != 0: lost members with STEP(J) <= SMIN, IFIRST <= J <= N, inside 2 RS are put back
 *                           into NLIST (ordered insertion), NBLOSS--, NBSMIN++, and their pair force / derivative is
 *                           subtracted from FREG / FDR.
 *   4. regcor_gpu.F:425-470 DFIRR / DFD: minus the pair terms of the lost members, plus those of the gained members, fp64,
 *                           in list order.
 * X, XDOT, BODY are the predicted values of the current regular block, i.e. exactly the snapshot gpunb_send_ uploaded
 * (intgrt.F:918 passes X(1,IFIRST), XDOT(1,IFIRST), BODY(IFIRST)): particle J sits at index J - IFIRST of m / x / v.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>

typedef struct { double f[3], fd[3]; } pair_t;

/* regcor_gpu.F:393-404 (= :428-438 = :450-460): A = X(J) - XI, DV = XDOT(J) - XIDOT, RIJ2 = A1*A1 + A2*A2 + A3*A3,
 * DR2I = 1/RIJ2, DR3I = BODY(J)*DR2I*SQRT(DR2I), DRDV = A1*DV1 + A2*DV2 + A3*DV3, DRDP = 3*DRDV*DR2I;
 * force term A*DR3I, derivative term (DV - A*DRDP)*DR3I. */
static pair_t pair_terms(const double *xi, const double *vi, const double *xj, const double *vj, double mj)
{
    pair_t p;
    const double a1 = xj[0] - xi[0], a2 = xj[1] - xi[1], a3 = xj[2] - xi[2];
    const double d1 = vj[0] - vi[0], d2 = vj[1] - vi[1], d3 = vj[2] - vi[2];
    const double rij2 = a1 * a1 + a2 * a2 + a3 * a3;
    const double dr2i = 1.0 / rij2;
    const double dr3i = mj * dr2i * sqrt(dr2i);
    const double drdv = a1 * d1 + a2 * d2 + a3 * d3;
    const double drdp = 3.0 * drdv * dr2i;
    p.f[0] = a1 * dr3i; p.f[1] = a2 * dr3i; p.f[2] = a3 * dr3i;
    p.fd[0] = (d1 - a1 * drdp) * dr3i; p.fd[1] = (d2 - a2 * drdp) * dr3i; p.fd[2] = (d3 - a3 * drdp) * dr3i;
    return p;
}

/* One row.  Index conventions: I, list members and JJLIST are Fortran particle numbers (1-based); m/x/v are the sent
 * snapshot (particle J at J - ifirst); step is STEP(IFIRST..NTOT) laid out the same way (NULL: no member is ever retained).
 * row_gpu = [count, j...] as returned by gpunb_regf_; old = LIST(1:, I) = [NNB0, members...].
 * nlist (out, >= lmax ints) = [NNB, members...]; jjlist (out, 2*lmax ints): JJLIST(k) = jjlist[k-1].
 * freg / fdr / dfirr / dfd are updated in place (the caller passes DFIRR = DFD = 0, or the NNB = 0 values of :121-130).
 * Returns the number of retained members (the row's contribution to NBSMIN). */
int oracle_regcor_row(int I, int ifirst, int n, int ntot, int lmax, int nnbmax, const int *row_gpu, const int *old,
                      const double *m, const double *x, const double *v, const double *step, double smin, double rs2,
                      int *nlist, double *freg, double *fdr, double *dfirr, double *dfd, int *nbloss_out, int *nbgain_out,
                      int *jjlist)
{
    (void)ntot; (void)lmax;
    const double *xi = x + 3 * (size_t)(I - ifirst), *vi = v + 3 * (size_t)(I - ifirst);
    /* 1. util_gpu.F:102-111 */
    int nnb = 0;
    for (int ll = 1; ll <= row_gpu[0]; ll++) {
        const int itemp = row_gpu[ll] + ifirst;
        if (itemp != I) nlist[++nnb] = itemp;
    }
    nlist[0] = nnb;
    const int nnb0 = old[0];
    int nbloss = 0, nbgain = 0, jmin = 0, nbsmin = 0;
    if (nnb0 == 0) {
        /* regcor_gpu.F:271-283 */
        nbgain = nnb;
        for (int l = 1; l <= nnb; l++) jjlist[l - 1] = nlist[l];
    } else {
        /* 2. regcor_gpu.F:297-336: both lists strictly ascending; the Fortran walks them with the sentinel NTOT+1 behind
         * the new list and over the last old member -- the walk visits the members in ascending order and files every
         * old member absent from the new list as lost, every new member absent from the old list as gained. */
        int l = 1, lg = 1;
        while (l <= nnb0 || lg <= nnb) {
            const int jo = l <= nnb0 ? old[l] : ntot + 1, jn = lg <= nnb ? nlist[lg] : ntot + 1;
            if (jo == jn) { l++; lg++; }
            else if (jo > jn) { nbgain++; jjlist[nnb0 + nbgain - 1] = jn; lg++; }
            else {
                nbloss++; jjlist[nbloss - 1] = jo; l++;
                if (step && step[jo - ifirst] < smin) jmin = jo;           /* :317 */
            }
        }
        /* 3. regcor_gpu.F:338-420 */
        if (jmin != 0) {
            int k = 1;
            while (k <= nbloss) {
                if (nnb > nnbmax || I > n) break;                          /* :342 */
                const int j = jjlist[k - 1];
                int keep = !(step[j - ifirst] > smin || j < ifirst || j > n);      /* :345 */
                const double *xj = x + 3 * (size_t)(j - ifirst), *vj = v + 3 * (size_t)(j - ifirst);
                if (keep) {
                    const double e1 = xi[0] - xj[0], e2 = xi[1] - xj[1], e3 = xi[2] - xj[2];
                    const double rij2 = e1 * e1 + e2 * e2 + e3 * e3;       /* :347-348 */
                    if (rij2 > 4.0 * rs2) keep = 0;
                }
                if (!keep) { k++; continue; }
                /* :351-358 ordered insertion.  The Fortran starts its comparison at NLIST(NNB+1); for NNB >= 1 it never looks
                 * at NLIST(1), which holds scratch at this point (the last old member, saved at :304).  For NNB = 0 it DOES:
                 * "IF (NLIST(1).LT.J)" fails (J is an old member, so J <= the last old member), NLIST(2) = NLIST(1) and J
                 * goes to NLIST(1) -- the list receives the LAST OLD MEMBER instead of J while FREG / FDR are corrected for
                 * J.  A latent slip of the reference in a rare corner (first retention into an empty new list); restated as
                 * written, because the bar is the reference's results.  (tests/regcor_cases.py: fortran_walk shows it.) */
                if (nnb == 0) {
                    const int s = old[nnb0];
                    nlist[1] = s < j ? j : s;
                } else {
                    int l2 = nnb;
                    while (l2 >= 1 && !(nlist[l2] < j)) { nlist[l2 + 1] = nlist[l2]; l2--; }
                    nlist[l2 + 1] = j;
                }
                nnb++; nbloss--; nbsmin++;
                /* :367-392 */
                const pair_t p = pair_terms(xi, vi, xj, vj, m[j - ifirst]);
                for (int c = 0; c < 3; c++) { freg[c] = freg[c] - p.f[c]; fdr[c] = fdr[c] - p.fd[c]; }
                /* :408-417: drop J from JJLIST unless it was the last lost member; the same slot is looked at again */
                if (k > nbloss) break;
                for (int l3 = k; l3 <= nbloss; l3++) jjlist[l3 - 1] = jjlist[l3];
            }
            nlist[0] = nnb;
        }
    }
    /* 4. regcor_gpu.F:425-470 */
    for (int l = 1; l <= nbloss; l++) {
        const int j = jjlist[l - 1];
        const pair_t p = pair_terms(xi, vi, x + 3 * (size_t)(j - ifirst), v + 3 * (size_t)(j - ifirst), m[j - ifirst]);
        for (int c = 0; c < 3; c++) { dfirr[c] = dfirr[c] - p.f[c]; dfd[c] = dfd[c] - p.fd[c]; }
    }
    for (int l = 1; l <= nbgain; l++) {
        const int j = jjlist[nnb0 + l - 1];
        const pair_t p = pair_terms(xi, vi, x + 3 * (size_t)(j - ifirst), v + 3 * (size_t)(j - ifirst), m[j - ifirst]);
        for (int c = 0; c < 3; c++) { dfirr[c] = dfirr[c] + p.f[c]; dfd[c] = dfd[c] + p.fd[c]; }
    }
    *nbloss_out = nbloss; *nbgain_out = nbgain;
    return nbsmin;
}

/* The batch the product entry point gpunb_b200_regcor_ handles in one call; rows are independent (the Fortran runs them
 * inside an OpenMP loop with NBSMIN in a critical section, regcor_gpu.F:363-365).  Same argument meaning as
 * include/gpunb_b200.h: new_list in/out [ni][lmax], old_list [ni][lmax], jjlist [ni][2*lmax]. */
void oracle_regcor(int ni, const int *index_i, int ifirst, int n, int ntot, int lmax, int *new_list, const int *old_list,
                   const double *m, const double *x, const double *v, const double *rs2, const double *step, double smin,
                   int nnbmax, double *freg, double *fdr, double *dfirr, double *dfd, int *nbloss, int *nbgain, int *jjlist,
                   int *nbsmin)
{
    int total = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total)
    for (int r = 0; r < ni; r++) {
        int *nl = (int *)malloc(sizeof(int) * (size_t)(lmax + 2));
        int *row = new_list + (size_t)r * lmax;
        if (row[0] < 0) {                       /* overflow rows are the caller's retry business (util_gpu.F:71-97) */
            nbloss[r] = nbgain[r] = 0;
            free(nl);
            continue;
        }
        total += oracle_regcor_row(index_i[r], ifirst, n, ntot, lmax, nnbmax, row, old_list + (size_t)r * lmax, m, x, v, step,
                                   smin, rs2[r], nl, freg + 3 * (size_t)r, fdr + 3 * (size_t)r, dfirr + 3 * (size_t)r,
                                   dfd + 3 * (size_t)r, nbloss + r, nbgain + r, jjlist + 2 * (size_t)r * lmax);
        memcpy(row, nl, sizeof(int) * (size_t)(nl[0] + 1));
        free(nl);
    }
    *nbsmin = total;
}
