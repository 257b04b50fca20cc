/* fpoly0_sweep.c -- a compiled caller of the regular-force C-ABI, written the way the Fortran caller uses it.
 *
 * What FPOLY0 does at start-up (reference: src/Main/fpoly0.F:72-195): send all particles, then for every block of
 * NIMAX = 1024 i-particles call gpunb_regf_, shrink RS and retry the block while any row reports overflow
 * (fpoly0.F:136-151 / util_gpu.F:71-97), drop self from the returned rows and shift the indices (fpoly0.F:186-195);
 * then the potential energy through gpupot_ (energy.F:37-39).  The program links against ANY library exporting the
 * reference symbols -- this repo's libgpunb_b200.so or the reference's own objects -- without a line of difference:
 *
 *   gcc -O2 examples/fpoly0_sweep.c -o fpoly0_b200 -Lnbody6ppgpu_b200 -lgpunb_b200 -Wl,-rpath,$PWD/nbody6ppgpu_b200 -lm
 *   gcc -O2 -DNO_DEVINIT examples/fpoly0_sweep.c -o fpoly0_avx -Loracle/_ref -lgpunb_ref_avx -Wl,-rpath,$PWD/oracle/_ref -lm
 *
 * Usage: fpoly0_sweep [N=4096] [seed=1] [nnbopt=64] [lmax=400]
 * Prints one line: N, mean neighbour number, overflow retries, checksums of forces / lists, total energy, wall times.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

/* the reference ABI (include/gpunb_b200.h, part 1) */
void gpunb_devinit_(int *irank);
void gpunb_open_(int *nbmax, int *irank);
void gpunb_close_(void);
void gpunb_send_(int *nj, double mj[], double xj[][3], double vj[][3]);
void gpunb_regf_(int *ni, double h2[], double dtr[], double xi[][3], double vi[][3], double acc[][3], double jrk[][3],
                 double pot[], int *lmax, int *nnbmax, int *list, int *m_flag);
void gpunb_profile_(int *irank);
void gpupot_(int *irank, int *istart, int *ni, int *n, double m[], double x[][3], double pot[]);

#define NIMAX 1024
#define PAD 8 /* the reference AVX library reads / writes a few rows past ni */

static unsigned long long rng_state;
static double urand(void)
{   /* splitmix64 */
    unsigned long long z = (rng_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
static double wtime(void) { struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + 1e-6 * tv.tv_usec; }

int main(int argc, char **argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 4096;
    const int seed = argc > 2 ? atoi(argv[2]) : 1;
    const int nnbopt = argc > 3 ? atoi(argv[3]) : 64;
    int lmax = argc > 4 ? atoi(argv[4]) : 400;
    int nnbmax = lmax - 50, irank = 0, m_flag = 0;
    rng_state = 0x1234567ull + (unsigned long long)seed;

    double *m = calloc((size_t)n + PAD, sizeof(double));
    double(*x)[3] = calloc((size_t)n + PAD, sizeof(*x)), (*v)[3] = calloc((size_t)n + PAD, sizeof(*v));
    double *rs = calloc((size_t)n + PAD, sizeof(double)), *phi = calloc((size_t)n + PAD, sizeof(double));
    /* Plummer sphere, Aarseth-Henon-Wielen (setup.F:62-107), N-body units, equal masses */
    const double sx = 3.0 * M_PI / 16.0;
    double cm[6] = {0};
    for (int i = 0; i < n; i++) {
        double r;
        do { const double a = urand(); r = 1.0 / sqrt(pow(a > 1e-10 ? a : 1e-10, -2.0 / 3.0) - 1.0); } while (r > 10.0);
        const double cz = 1.0 - 2.0 * urand(), ph = 2.0 * M_PI * urand(), sz = sqrt(1.0 - cz * cz);
        x[i][0] = r * sz * cos(ph); x[i][1] = r * sz * sin(ph); x[i][2] = r * cz;
        double q, g;
        do { q = urand(); g = 0.1 * urand(); } while (g > q * q * pow(1.0 - q * q, 3.5));
        const double ve = sqrt(2.0) * pow(1.0 + r * r, -0.25) * q;
        const double cz2 = 1.0 - 2.0 * urand(), ph2 = 2.0 * M_PI * urand(), sz2 = sqrt(1.0 - cz2 * cz2);
        v[i][0] = ve * sz2 * cos(ph2); v[i][1] = ve * sz2 * sin(ph2); v[i][2] = ve * cz2;
        m[i] = 1.0 / n;
        for (int c = 0; c < 3; c++) { x[i][c] *= sx; v[i][c] /= sqrt(sx); cm[c] += x[i][c] / n; cm[3 + c] += v[i][c] / n; }
    }
    for (int i = 0; i < n; i++)
        for (int c = 0; c < 3; c++) { x[i][c] -= cm[c]; v[i][c] -= cm[3 + c]; }
    const double rs0 = sx * cbrt((double)nnbopt / n) * 1.6;
    for (int i = 0; i < n; i++) rs[i] = rs0 * sqrt(1.0 + (x[i][0] * x[i][0] + x[i][1] * x[i][1] + x[i][2] * x[i][2]));

    static double h2[NIMAX + PAD], dtr[NIMAX + PAD], acc[NIMAX + PAD][3], jrk[NIMAX + PAD][3], pot[NIMAX + PAD];
    int *list = calloc((size_t)(NIMAX + PAD) * lmax, sizeof(int));
    int nbmax = n + 10;
#ifndef NO_DEVINIT
    gpunb_devinit_(&irank);
#endif
    gpunb_open_(&nbmax, &irank);
    const double t0 = wtime();
    gpunb_send_(&n, m, x, v);
    const double t1 = wtime();
    long long nnb_sum = 0, retries = 0;
    unsigned long long list_xor = 0;
    double fsum = 0, jsum = 0, psum = 0;
    for (int i0 = 0; i0 < n; i0 += NIMAX) {
        int ni = n - i0 < NIMAX ? n - i0 : NIMAX;
        for (;;) {
            for (int k = 0; k < ni; k++) {
                const double r2 = x[i0 + k][0] * x[i0 + k][0] + x[i0 + k][1] * x[i0 + k][1] + x[i0 + k][2] * x[i0 + k][2];
                h2[k] = rs[i0 + k] * rs[i0 + k];
                dtr[k] = fmin(0.125 / 8.0 * sqrt(1.0 + r2), 0.125);       /* fpoly0.F:53-56 */
            }
            gpunb_regf_(&ni, h2, dtr, &x[i0], &v[i0], acc, jrk, pot, &lmax, &nnbmax, list, &m_flag);
            int over = 0;
            for (int k = 0; k < ni; k++) {
                const int nnb = list[(size_t)k * lmax];
                if (nnb < 0) {              /* util_gpu.F:83-87 */
                    rs[i0 + k] *= (-nnb > nnbopt) ? pow((double)nnbopt / -nnb, 0.333) : pow((double)nnbopt / nnbmax, 0.4);
                    over++;
                }
            }
            if (!over) break;
            retries++;
        }
        for (int k = 0; k < ni; k++) {
            int *row = list + (size_t)k * lmax, l1 = 0;
            for (int l = 1; l <= row[0]; l++)       /* fpoly0.F:186-195: drop self, 1-based caller indices */
                if (row[l] != i0 + k) { row[++l1] = row[l] + 1; list_xor ^= (unsigned long long)(row[l] + 1) * 0x9e3779b97f4a7c15ull + (unsigned long long)(i0 + k); }
            row[0] = l1;
            nnb_sum += l1;
            fsum += sqrt(acc[k][0] * acc[k][0] + acc[k][1] * acc[k][1] + acc[k][2] * acc[k][2]);
            jsum += sqrt(jrk[k][0] * jrk[k][0] + jrk[k][1] * jrk[k][1] + jrk[k][2] * jrk[k][2]);
            psum += pot[k];
        }
    }
    const double t2 = wtime();
    int one = 1;
    gpupot_(&irank, &one, &n, &n, m, x, phi);
    const double t3 = wtime();
    double ekin = 0, epot = 0;
    for (int i = 0; i < n; i++) {
        ekin += 0.5 * m[i] * (v[i][0] * v[i][0] + v[i][1] * v[i][1] + v[i][2] * v[i][2]);
        epot -= 0.5 * m[i] * phi[i];
    }
    gpunb_profile_(&irank);
    gpunb_close_();
    printf("FPOLY0 n %d mean_nnb %.6f retries %lld list_xor %016llx fsum %.12e jsum %.12e psum %.12e etot %.12e "
           "send_s %.6f regf_s %.6f pot_s %.6f\n",
           n, (double)nnb_sum / n, retries, list_xor, fsum, jsum, psum, ekin + epot, t1 - t0, t2 - t1, t3 - t2);
    return 0;
}
