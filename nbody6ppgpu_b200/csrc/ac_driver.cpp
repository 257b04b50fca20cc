// ac_driver.cpp -- native Ahmad-Cohen block-step driver around the Fortran-callable C-ABIs.
//
// The C++ twin of nbody6ppgpu_b200/hermite_ac.py (same statements, same order; see that file's docstring for the
// reference lines each step restates: intgrt.F:284-293,912-974, util_gpu.F:33-111, regcor_gpu.F:510-552,623-760,
// xbpredall.f:17-26, energy.F).  It exists so that "wall s per N-body time unit" measures the LIBRARIES and their call
// latencies, not a numpy harness: the host side of a block step is a few microseconds here.
// Any library exporting the reference's ABI can sit behind it (dlopen by path): this repo's libgpunb_b200.so /
// libirr_b200.so, or the reference's own libraries built under oracle/_ref.  Optional device paths, this repo's
// libraries only: device-resident predictor (gpunb_b200_state_* / predict_send), list bookkeeping
// (gpunb_b200_regcor_), batched set_jp / set_list.  Without an irregular-force library the irregular sums run on the
// host in fp64 (what nbint.f does).
// Left out, like the Python twin: KS / chain regularisation, stellar evolution, tides, retention of small-step neighbours.
#include <algorithm>
#include <functional>
#include <map>
#include <utility>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <dlfcn.h>
#include <sys/time.h>
#include <time.h>

namespace {

// AC_DRIVER_PHASES=1: where the driver's own time goes, printed to stderr at the end of the run (a dozen clock reads per block step)
static double PH[16];
static inline double pt() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
double wtime() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + 1e-6 * tv.tv_usec; }
constexpr int MAXTHR = 1024;      // i-particles per gpunb_regf_ call (util_gpu.F:7)
constexpr int OMP_MIN = 512;      // per-particle loops of a block go parallel above this many particles
constexpr int PAD = 8;            // rows behind every array: the reference AVX library reads / writes past ni (reg.avx.cpp:204-314)

typedef double d3[3];
struct Abi {
    void *h = nullptr, *hi = nullptr;
    void (*devinit)(int *) = nullptr;
    void (*open)(int *, int *) = nullptr;
    void (*close)() = nullptr;
    void (*send)(int *, double *, d3 *, d3 *) = nullptr;
    void (*regf)(int *, double *, double *, d3 *, d3 *, d3 *, d3 *, double *, int *, int *, int *, int *) = nullptr;
    void (*pot)(int *, int *, int *, int *, double *, d3 *, double *) = nullptr;
    // gpunb_b200 extensions
    void (*state_all)(int *, double *, d3 *, d3 *, d3 *, d3 *, double *) = nullptr;
    void (*state_update)(int *, int *, double *, d3 *, d3 *, d3 *, d3 *, double *) = nullptr;
    void (*predict_send)(int *, double *) = nullptr;
    void (*regcor)(int *, int *, int *, int *, int *, int *, int *, int *, double *, double *, double *, int *, d3 *, d3 *, d3 *,
                   d3 *, int *, int *, int *, int *) = nullptr;
    decltype(regcor) regcor_last = nullptr;
    void (*predict_send_records)(int *, double *, const double *, int *) = nullptr;
    void (*lists_put)(int *, int *, int *, int *) = nullptr;
    const double *(*irr_records)(int *) = nullptr;
    void (*irr_flush)() = nullptr;
    // irregular-force library
    void (*iopen)(int *, int *, int *) = nullptr;
    void (*iclose)(int *) = nullptr;
    void (*set_jp)(int *, double *, double *, double *, double *, double *, double *) = nullptr;
    void (*set_list)(int *, int *) = nullptr;
    void (*firr)(double *, int *, int *, d3 *, d3 *, int *) = nullptr;
    void (*set_jp_batch)(int *, int *, d3 *, d3 *, d3 *, d3 *, double *, double *) = nullptr;
    void (*set_list_batch)(int *, int *, int *, int *) = nullptr;
};
template <class F> bool sym(void *h, const char *name, F &f) { f = reinterpret_cast<F>(dlsym(h, name)); return f != nullptr; }

}  // namespace

extern "C" {

struct ACParams {
    int nnbopt, lmax, m_flag, use_predictor, use_regcor;
    double eta_i, eta_r, dtmax, dtmin, t_end, rs0;      // rs0 <= 0: default
};
struct ACStats {
    double t, wall_total, wall_send, wall_regf, wall_irr, wall_regcor, wall_energy, wall_init, e0, e1;
    long long irr_steps, reg_steps, block_steps, reg_blocks, regf_calls, overflow_retries;
    double mean_nnb;
};

// Returns 0 on success.  x_out / v_out (optional): the final X0 / X0DOT.
int ac_driver_run(const char *gpunb_so, const char *irr_so, int n, const double *m_in, const double *x_in, const double *v_in,
                  const ACParams *p, ACStats *st, double *x_out, double *v_out)
{
    Abi A;
    A.h = dlopen(gpunb_so, RTLD_NOW | RTLD_LOCAL);
    if (!A.h) { fprintf(stderr, "ac_driver: %s\n", dlerror()); return 1; }
    sym(A.h, "gpunb_devinit_", A.devinit);
    if (!sym(A.h, "gpunb_open_", A.open) || !sym(A.h, "gpunb_close_", A.close) || !sym(A.h, "gpunb_send_", A.send) ||
        !sym(A.h, "gpunb_regf_", A.regf) || !sym(A.h, "gpupot_", A.pot)) { fprintf(stderr, "ac_driver: %s lacks the reference ABI\n", gpunb_so); return 2; }
    const bool b200 = sym(A.h, "gpunb_b200_state_all_", A.state_all);
    if (b200) { sym(A.h, "gpunb_b200_state_update_", A.state_update); sym(A.h, "gpunb_b200_predict_send_", A.predict_send); sym(A.h, "gpunb_b200_regcor_", A.regcor); sym(A.h, "gpunb_b200_regcor_last_", A.regcor_last); sym(A.h, "gpunb_b200_predict_send_records_", A.predict_send_records); sym(A.h, "gpunb_b200_lists_put_", A.lists_put); }
    const bool predictor = p->use_predictor && b200, use_regcor = p->use_regcor && b200 && A.regcor;
    if ((p->use_predictor || p->use_regcor) && !b200) { fprintf(stderr, "ac_driver: device paths need libgpunb_b200.so\n"); return 3; }
    const bool use_irr = irr_so && *irr_so;
    if (use_irr) {
        A.hi = dlopen(irr_so, RTLD_NOW | RTLD_LOCAL);
        if (!A.hi) { fprintf(stderr, "ac_driver: %s\n", dlerror()); return 1; }
        if (!sym(A.hi, "irr_simd_open_", A.iopen) || !sym(A.hi, "irr_simd_close_", A.iclose) || !sym(A.hi, "irr_simd_set_jp_", A.set_jp) ||
            !sym(A.hi, "irr_simd_set_list_", A.set_list) || !sym(A.hi, "irr_simd_firr_vec_", A.firr)) { fprintf(stderr, "ac_driver: %s lacks irr_simd_*\n", irr_so); return 2; }
        sym(A.hi, "irr_b200_set_jp_batch_", A.set_jp_batch); sym(A.hi, "irr_b200_set_list_batch_", A.set_list_batch);
        sym(A.hi, "irr_b200_particle_records_", A.irr_records); sym(A.hi, "irr_b200_flush_", A.irr_flush);
    }
    // use_predictor = 2: ONE copy of the particle state on the device -- the irregular-force library's table, which set_jp keeps
    // current -- also feeds the regular-force predictor (gpunb_b200_predict_send_records_): no state upload of its own
    const bool shared_state = p->use_predictor == 2;
    // use_regcor = 2: the old lists come from the library's device-resident list store (put once after the initial forces,
    // committed by every regcor call since): no list is uploaded in steady state
    const bool resident_lists = p->use_regcor == 2 && use_regcor && A.lists_put;
    if (shared_state && !(A.predict_send_records && A.irr_records && A.irr_flush)) { fprintf(stderr, "ac_driver: shared predictor state needs libgpunb_b200.so and libirr_b200.so\n"); return 3; }
    memset(st, 0, sizeof(*st));
    const double w_init = wtime();
    const int nnbopt = p->nnbopt, lmax = p->lmax, m_flag = p->m_flag;
    const int nnbmax = lmax > 100 ? std::min(n / 2, lmax - 50) : lmax - 8;
    const double dtmax = p->dtmax, dtmin = p->dtmin;
    const int lstride = 1 + 8 * ((nnbmax + 8) / 8) + 8;              // irr library rows: [nnb, members (1-based) ...] + room for 8-wide reads
    const int N = n + PAD;
    std::vector<double> m(N, 0.0), t0(N, 0.0), t0r(N, 0.0), dt(N, dtmin), dtr(N, dtmax), rs(N, 0.0);
    std::vector<double> x0(3 * N, 0.0), v0(3 * N, 0.0), fi(3 * N, 0.0), fid(3 * N, 0.0), fr(3 * N, 0.0), frd(3 * N, 0.0), f(3 * N, 0.0), fd(3 * N, 0.0);
    std::vector<int> nb((size_t)n * (nnbmax + 1), -1), nnb(n, 0);
    std::vector<char> dirty(n, 0);
    memcpy(m.data(), m_in, sizeof(double) * n); memcpy(x0.data(), x_in, sizeof(double) * 3 * n); memcpy(v0.data(), v_in, sizeof(double) * 3 * n);
    double bodym = 0; for (int i = 0; i < n; i++) bodym += m[i]; bodym /= n;
    const double rs0 = p->rs0 > 0 ? p->rs0 : cbrt((double)nnbopt / (0.5 * n) * 0.8 * 0.8 * 0.8);
    for (int i = 0; i < n; i++) rs[i] = rs0 * sqrt(1.0 + x0[3 * i] * x0[3 * i] + x0[3 * i + 1] * x0[3 * i + 1] + x0[3 * i + 2] * x0[3 * i + 2]);   // fpoly0.F:53-56
    auto pow2_floor = [&](double d) { const double e = floor(log2(std::max(d, 1e-300) / dtmax)); return dtmax * exp2(std::min(e, 0.0)); };
    auto norm3 = [](const double *a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); };

    int irank = 0, nbmax = n + 10;
    if (A.devinit) A.devinit(&irank);
    A.open(&nbmax, &irank);
    if (use_irr) { int nm = n, lm = lstride, rk = 0; A.iopen(&nm, &lm, &rk); }

    // ---- scratch (block-sized or N-sized, allocated once) -------------------------------------------------------------
    std::vector<double> xp(3 * N), vp(3 * N);                       // predicted snapshot of all particles (host predictor)
    std::vector<int> act, reg, regpos;
    std::vector<double> xa, va, fia, fida, frn_, frdn_, fin, fidn, h2(MAXTHR + PAD), dtrb(MAXTHR + PAD), pot(MAXTHR + PAD);
    std::vector<double> xi(3 * (MAXTHR + PAD)), vi(3 * (MAXTHR + PAD)), acc(3 * (MAXTHR + PAD)), jrk(3 * (MAXTHR + PAD));
    std::vector<int> lst((size_t)(MAXTHR + PAD) * lmax), addr, irows, nnid;
    std::vector<int> rows_raw, lnew, cnew, old_rows, idx1, nbl, nbg, jj;
    std::vector<double> rs2, zf, zd, dfi, dfd, upd;

    auto predict_one = [&](int i, double t, double *xo, double *vo) {   // xbpredall.f:17-26 (F2 = F/2, FD6 = FDOT/6), unfused order
        const double s = t - t0[i];
        for (int c = 0; c < 3; c++) {
            const double f2 = 0.5 * f[3 * i + c], fd6 = fd[3 * i + c] * (1.0 / 6.0);
            xo[c] = ((fd6 * s + f2) * s + v0[3 * i + c]) * s + x0[3 * i + c];
            vo[c] = (fd6 * (1.5 * s) + f2) * (2.0 * s) + v0[3 * i + c];
        }
    };
    auto predict_all = [&](double t) {
#pragma omp parallel for schedule(static) if (n >= OMP_MIN)
        for (int i = 0; i < n; i++) predict_one(i, t, &xp[3 * i], &vp[3 * i]);
    };
    auto irr_push_particles = [&](const std::vector<int> &idx) {
        const int k = (int)idx.size();
        if (!k) return;
        addr.resize(k); upd.resize((size_t)14 * k);
        double *px = upd.data(), *pv = px + 3 * k, *pa = pv + 3 * k, *pj = pa + 3 * k, *pm = pj + 3 * k, *pt = pm + k;
#pragma omp parallel for schedule(static) if (k >= OMP_MIN)
        for (int q = 0; q < k; q++) {
            const int i = idx[q];
            addr[q] = i + 1; pm[q] = m[i]; pt[q] = t0[i];
            for (int c = 0; c < 3; c++) { px[3 * q + c] = x0[3 * i + c]; pv[3 * q + c] = v0[3 * i + c]; pa[3 * q + c] = 0.5 * f[3 * i + c]; pj[3 * q + c] = fd[3 * i + c] * (1.0 / 6.0); }
        }
        if (A.set_jp_batch) { int kk = k; A.set_jp_batch(&kk, addr.data(), (d3 *)px, (d3 *)pv, (d3 *)pa, (d3 *)pj, pm, pt); }
        else for (int q = 0; q < k; q++) A.set_jp(&addr[q], px + 3 * q, pv + 3 * q, pa + 3 * q, pj + 3 * q, pm + q, pt + q);
    };
    auto irr_push_lists = [&](const std::vector<int> &idx) {
        const int k = (int)idx.size();
        if (!k) return;
        addr.resize(k);
        if (irows.size() < (size_t)k * lstride) irows.resize((size_t)k * lstride);
#pragma omp parallel for schedule(static) if (k >= OMP_MIN)
        for (int q = 0; q < k; q++) {
            const int i = idx[q];
            addr[q] = i + 1;
            int *r = &irows[(size_t)q * lstride];
            r[0] = nnb[i];
            for (int l = 0; l < nnb[i]; l++) r[1 + l] = nb[(size_t)i * (nnbmax + 1) + l] + 1;
            for (int l = nnb[i]; l < std::min(lstride - 1, nnb[i] + 8); l++) r[1 + l] = 0;     // the AVX library reads the list 8 at a time
        }
        if (A.set_list_batch) { int kk = k, ls = lstride; A.set_list_batch(&kk, addr.data(), &ls, irows.data()); }
        else for (int q = 0; q < k; q++) A.set_list(&addr[q], &irows[(size_t)q * lstride]);
    };
    // irregular force / derivative of the particles idx over the lists they hold, at time t
    auto irregular = [&](const std::vector<int> &idx, double t, std::vector<double> &ofi, std::vector<double> &ofd, bool have_snapshot) {
        const int k = (int)idx.size();
        ofi.assign((size_t)3 * k + 3 * PAD, 0.0); ofd.assign((size_t)3 * k + 3 * PAD, 0.0);
        if (!k) return;
        const double w0 = wtime();
        if (use_irr) {
            addr.resize(k); nnid.resize(k + PAD);
            for (int q = 0; q < k; q++) addr[q] = idx[q] + 1;
            int kk = k; double tt = t;
            A.firr(&tt, &kk, addr.data(), (d3 *)ofi.data(), (d3 *)ofd.data(), nnid.data());
        } else {
            if (!have_snapshot) predict_all(t);
            for (int q = 0; q < k; q++) {
                const int i = idx[q];
                double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
                for (int l = 0; l < nnb[i]; l++) {
                    const int j = nb[(size_t)i * (nnbmax + 1) + l];
                    const double dx[3] = {xp[3 * j] - xp[3 * i], xp[3 * j + 1] - xp[3 * i + 1], xp[3 * j + 2] - xp[3 * i + 2]};
                    const double dv[3] = {vp[3 * j] - vp[3 * i], vp[3 * j + 1] - vp[3 * i + 1], vp[3 * j + 2] - vp[3 * i + 2]};
                    const double r2 = dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2], rinv2 = 1.0 / r2;
                    const double mr3 = m[j] * rinv2 * sqrt(rinv2), rv = 3.0 * (dx[0] * dv[0] + dx[1] * dv[1] + dx[2] * dv[2]) * rinv2;
                    for (int c = 0; c < 3; c++) { a[c] += mr3 * dx[c]; b[c] += mr3 * (dv[c] - rv * dx[c]); }
                }
                for (int c = 0; c < 3; c++) { ofi[3 * q + c] = a[c]; ofd[3 * q + c] = b[c]; }
            }
        }
        st->wall_irr += wtime() - w0;
    };
    // gpunb_send_ (or predict_send) + gpunb_regf_ over the block idx with its predicted xi / vi (3 k doubles each).
    // raw: the untouched rows go to rows_raw (for regcor); else lnew / cnew get the caller-side lists (self removed).
    auto regular = [&](const std::vector<int> &idx, const double *bx, const double *bv, double t, bool snapshot_is_x0, bool raw) {
        const int k = (int)idx.size();
        double w0 = wtime();
        int nj = n;
        if (shared_state && !snapshot_is_x0) {
            A.irr_flush();
            int stride = 0; const double *rec = A.irr_records(&stride);
            double tt = t;
            A.predict_send_records(&nj, &tt, rec, &stride);
        } else if (predictor && !snapshot_is_x0) {
            std::vector<int> d;
            for (int i = 0; i < n; i++) if (dirty[i]) { d.push_back(i); dirty[i] = 0; }
            if (!d.empty()) {
                const int kd = (int)d.size();
                upd.resize((size_t)14 * kd);
                double *pb = upd.data(), *px = pb + kd, *pv = px + 3 * kd, *pf = pv + 3 * kd, *pfd = pf + 3 * kd, *pt = pfd + 3 * kd;
                for (int q = 0; q < kd; q++) {
                    const int i = d[q];
                    pb[q] = m[i]; pt[q] = t0[i];
                    for (int c = 0; c < 3; c++) { px[3 * q + c] = x0[3 * i + c]; pv[3 * q + c] = v0[3 * i + c]; pf[3 * q + c] = 0.5 * f[3 * i + c]; pfd[3 * q + c] = fd[3 * i + c] * (1.0 / 6.0); }
                }
                int kk = kd;
                A.state_update(&kk, d.data(), pb, (d3 *)px, (d3 *)pv, (d3 *)pf, (d3 *)pfd, pt);
            }
            double tt = t;
            A.predict_send(&nj, &tt);
        } else if (snapshot_is_x0) A.send(&nj, m.data(), (d3 *)x0.data(), (d3 *)v0.data());
        else A.send(&nj, m.data(), (d3 *)xp.data(), (d3 *)vp.data());
        st->wall_send += wtime() - w0;
        frn_.assign((size_t)3 * k, 0.0); frdn_.assign((size_t)3 * k, 0.0);
        if (raw) { if (rows_raw.size() < (size_t)k * lmax) rows_raw.resize((size_t)k * lmax); }
        else { if (lnew.size() < (size_t)k * (nnbmax + 1)) lnew.resize((size_t)k * (nnbmax + 1)); cnew.assign(k, 0); }
        for (int c0 = 0; c0 < k; c0 += MAXTHR) {
            int ni = std::min(MAXTHR, k - c0);
            memcpy(xi.data(), bx + 3 * (size_t)c0, sizeof(double) * 3 * ni); memcpy(vi.data(), bv + 3 * (size_t)c0, sizeof(double) * 3 * ni);
            for (;;) {
                for (int q = 0; q < ni; q++) { const int i = idx[c0 + q]; h2[q] = rs[i] * rs[i] / (m_flag ? bodym : 1.0); dtrb[q] = dtr[i]; }
                w0 = wtime();
                int lm = lmax, nm = nnbmax, mf = m_flag;
                A.regf(&ni, h2.data(), dtrb.data(), (d3 *)xi.data(), (d3 *)vi.data(), (d3 *)acc.data(), (d3 *)jrk.data(), pot.data(), &lm, &nm, lst.data(), &mf);
                st->wall_regf += wtime() - w0;
                st->regf_calls++;
                bool over = false;
                for (int q = 0; q < ni; q++) {
                    const int c = lst[(size_t)q * lmax];
                    if (c < 0) {           // util_gpu.F:83-90 (NB_FLAG = 1): RS towards NNBOPT members
                        over = true;
                        const double cnt = -(double)c;
                        rs[idx[c0 + q]] *= cnt > nnbopt ? pow(nnbopt / cnt, 0.333) : pow((double)nnbopt / nnbmax, 0.4);
                    }
                }
                if (!over) break;
                st->overflow_retries++;
            }
            memcpy(&frn_[3 * (size_t)c0], acc.data(), sizeof(double) * 3 * ni); memcpy(&frdn_[3 * (size_t)c0], jrk.data(), sizeof(double) * 3 * ni);
#pragma omp parallel for schedule(static) if (ni >= OMP_MIN)
            for (int q = 0; q < ni; q++) {
                const int *row = &lst[(size_t)q * lmax];
                if (raw) { memcpy(&rows_raw[(size_t)(c0 + q) * lmax], row, sizeof(int) * (row[0] + 1)); continue; }
                int *o = &lnew[(size_t)(c0 + q) * (nnbmax + 1)], cnt = 0;
                for (int l = 1; l <= row[0]; l++) if (row[l] != idx[c0 + q]) o[cnt++] = row[l];      // util_gpu.F:102-111
                cnew[c0 + q] = cnt;
            }
        }
    };
    auto adjust_rs = [&](const std::vector<int> &idx, const std::vector<int> &counts) {     // regcor_gpu.F:623-760, volume rule
        for (size_t q = 0; q < idx.size(); q++) {
            const double ratio = std::min(std::max(cbrt((double)nnbopt / std::max(counts[q], 1)), 0.9), 1.1);
            rs[idx[q]] *= ratio;
        }
    };
    auto hermite = [&](const double *f0, const double *fd0, const double *f1, const double *fd1, double d, double *a2, double *a3) {
        for (int c = 0; c < 3; c++) {
            a2[c] = (-6.0 * (f0[c] - f1[c]) - d * (4.0 * fd0[c] + 2.0 * fd1[c])) / (d * d);
            a3[c] = (12.0 * (f0[c] - f1[c]) + 6.0 * d * (fd0[c] + fd1[c])) / (d * d * d);
        }
    };
    auto aarseth = [&](double eta, const double *f1, const double *fd1, const double *a2, const double *a3, double d) {
        const double a2n[3] = {a2[0] + d * a3[0], a2[1] + d * a3[1], a2[2] + d * a3[2]};
        const double num = norm3(f1) * norm3(a2n) + norm3(fd1) * norm3(fd1), den = norm3(fd1) * norm3(a3) + norm3(a2n) * norm3(a2n) + 1e-300;
        return sqrt(eta * num / den);
    };
    auto energy = [&]() {
        const double w0 = wtime();
        int ir = 0, is = 1, ni = n, nn = n;
        std::vector<double> phi(N, 0.0);
        A.pot(&ir, &is, &ni, &nn, m.data(), (d3 *)x0.data(), phi.data());
        double e = 0;
        for (int i = 0; i < n; i++) e += 0.5 * m[i] * (v0[3 * i] * v0[3 * i] + v0[3 * i + 1] * v0[3 * i + 1] + v0[3 * i + 2] * v0[3 * i + 2]) - 0.5 * m[i] * phi[i];
        st->wall_energy += wtime() - w0;
        return e;
    };

    // ---- initial force polynomials --------------------------------------------------------------------------------------
    std::vector<int> all(n);
    for (int i = 0; i < n; i++) all[i] = i;
    regular(all, x0.data(), v0.data(), 0.0, true, false);
    for (int i = 0; i < n; i++) { nnb[i] = cnew[i]; memcpy(&nb[(size_t)i * (nnbmax + 1)], &lnew[(size_t)i * (nnbmax + 1)], sizeof(int) * cnew[i]); }
    if (use_irr) { irr_push_particles(all); irr_push_lists(all); }
    memcpy(xp.data(), x0.data(), sizeof(double) * 3 * n); memcpy(vp.data(), v0.data(), sizeof(double) * 3 * n);
    irregular(all, 0.0, fia, fida, true);
    for (int i = 0; i < n; i++) {
        for (int c = 0; c < 3; c++) {
            fr[3 * i + c] = frn_[3 * i + c]; frd[3 * i + c] = frdn_[3 * i + c]; fi[3 * i + c] = fia[3 * i + c]; fid[3 * i + c] = fida[3 * i + c];
            f[3 * i + c] = fi[3 * i + c] + fr[3 * i + c]; fd[3 * i + c] = fid[3 * i + c] + frd[3 * i + c];
        }
        const double fa = norm3(&f[3 * i]), fda = norm3(&fd[3 * i]) + 1e-300, fia_ = norm3(&fi[3 * i]) + 1e-300, fida_ = norm3(&fid[3 * i]) + 1e-300;
        const double fra = norm3(&fr[3 * i]) + 1e-300, frda = norm3(&frd[3 * i]) + 1e-300;
        const double dt_i = 0.1 * sqrt(p->eta_i) * std::min(fa / fda, fia_ / fida_), dt_r = 0.1 * sqrt(p->eta_r) * fra / frda;
        dt[i] = std::max(pow2_floor(dt_i), dtmin);
        dtr[i] = std::max(pow2_floor(std::max(dt_r, dt[i])), dt[i]);
    }
    adjust_rs(all, cnew);
    if (use_irr) irr_push_particles(all);
    if (resident_lists) {
        std::vector<int> rows((size_t)MAXTHR * lmax), idx(MAXTHR);
        for (int i0 = 0; i0 < n; i0 += MAXTHR) {
            int k = std::min(MAXTHR, n - i0), lm = lmax;
            for (int q = 0; q < k; q++) {
                const int i = i0 + q;
                idx[q] = i + 1;
                int *o = &rows[(size_t)q * lmax];
                o[0] = nnb[i];
                for (int l = 0; l < nnb[i]; l++) o[1 + l] = nb[(size_t)i * (nnbmax + 1) + l] + 1;
            }
            A.lists_put(&k, idx.data(), &lm, rows.data());
        }
    }
    if (predictor && !shared_state) {
        std::vector<double> f2(3 * N), fd6(3 * N);
        for (int k = 0; k < 3 * n; k++) { f2[k] = 0.5 * f[k]; fd6[k] = fd[k] * (1.0 / 6.0); }
        int nj = n;
        A.state_all(&nj, m.data(), (d3 *)x0.data(), (d3 *)v0.data(), (d3 *)f2.data(), (d3 *)fd6.data(), t0.data());
    }
    // the raw rows of the largest possible regular block (every particle): allocated and touched here, not inside the timed run
    // (at N = 262 144 the first regular block of the run otherwise pays 0.5 s for 629 MB of fresh pages)
    if (use_regcor) rows_raw.assign((size_t)n * lmax, 0);
    st->wall_init = wtime() - w_init;
    st->wall_send = st->wall_regf = st->wall_irr = st->wall_regcor = 0.0;      // the buckets cover the run, like wall_total (the initial
                                                                               // force polynomials are wall_init)

    // ---- run --------------------------------------------------------------------------------------------------------------
    double t = 0.0;
    memset(PH, 0, sizeof(PH));
    st->e0 = energy();
    const double w_run = wtime();
    std::vector<double> frnew, frdnew, fr_old, frd_old, dtr_new;
    // next block time and its particles: block steps are powers of two, so the particles share a few dozen distinct next
    // step times -- one bucket of particle numbers per next time (ordered map), O(1) per particle and block step instead of
    // the O(N) scan of the Python twin (a binary heap over all N was 1/3 of the driver's own time: 14 levels of cache
    // misses per pop).  The block is put in ascending particle order, which is the scan's order (the order of the i-block
    // decides which particles share a warp of the pair kernel, hence the last bits of its sums).
    std::map<double, std::vector<int>> due;
    std::vector<std::vector<int>> spare;                 // emptied buckets, kept for their capacity
    auto due_at = [&](double tnext) -> std::vector<int> & {
        auto it = due.find(tnext);
        if (it != due.end()) return it->second;
        std::vector<int> b;
        if (!spare.empty()) { b.swap(spare.back()); spare.pop_back(); }
        return due.emplace(tnext, std::move(b)).first->second;
    };
    for (int i = 0; i < n; i++) due_at(t0[i] + dt[i]).push_back(i);
    while (t < p->t_end) {
        double P0 = pt();
        reg.clear(); regpos.clear();
        const double tn = due.begin()->first;
        act.swap(due.begin()->second);
        due.begin()->second.clear();
        spare.emplace_back(std::move(due.begin()->second));
        due.erase(due.begin());
        std::sort(act.begin(), act.end());
        const int na = (int)act.size();
        PH[0] += pt() - P0; P0 = pt();
        st->block_steps++;
        xa.resize((size_t)3 * na); va.resize((size_t)3 * na);
#pragma omp parallel for schedule(static) if (na >= OMP_MIN)
        for (int q = 0; q < na; q++) predict_one(act[q], tn, &xa[3 * q], &va[3 * q]);
        for (int q = 0; q < na; q++) if (t0r[act[q]] + dtr[act[q]] <= tn) { reg.push_back(act[q]); regpos.push_back(q); }
        const int nr = (int)reg.size();
        frnew.resize((size_t)3 * na); frdnew.resize((size_t)3 * na);
#pragma omp parallel for schedule(static) if (na >= OMP_MIN)
        for (int q = 0; q < na; q++) {
            const int i = act[q];
            for (int c = 0; c < 3; c++) { frnew[3 * q + c] = fr[3 * i + c] + frd[3 * i + c] * (tn - t0r[i]); frdnew[3 * q + c] = frd[3 * i + c]; }   // intgrt.F:284-293
        }
        bool have_snapshot = false;
        PH[1] += pt() - P0; P0 = pt();
        if (!use_irr || (nr && !predictor)) { predict_all(tn); have_snapshot = true; }
        PH[2] += pt() - P0; P0 = pt();
        irregular(act, tn, fia, fida, have_snapshot);
        PH[3] += pt() - P0; P0 = pt();          // over the lists held now (the old lists of the regular particles)
        if (nr) {
            st->reg_blocks++; st->reg_steps += nr;
            std::vector<double> bx((size_t)3 * nr + 3 * PAD), bv((size_t)3 * nr + 3 * PAD);
            fr_old.resize((size_t)3 * nr); frd_old.resize((size_t)3 * nr); dtr_new.resize(nr);
            std::vector<double> fio((size_t)3 * nr), fido((size_t)3 * nr);
#pragma omp parallel for schedule(static) if (nr >= OMP_MIN)
            for (int q = 0; q < nr; q++) for (int c = 0; c < 3; c++) {
                bx[3 * q + c] = xa[3 * regpos[q] + c]; bv[3 * q + c] = va[3 * regpos[q] + c];
                fio[3 * q + c] = fia[3 * regpos[q] + c]; fido[3 * q + c] = fida[3 * regpos[q] + c];
            }
            PH[9] += pt() - P0; P0 = pt();
            regular(reg, bx.data(), bv.data(), tn, false, use_regcor);
            PH[10] += pt() - P0; P0 = pt();
            if (use_regcor) {
                // the device diffs the lists and returns the force swap: F_irr(new list) = F_irr(old list) + DFIRR.  In chunks of
                // 2048 rows (what one launch of the library handles): scratch stays a few MB however large the block is
                const double w0 = wtime();
                constexpr int RCH = 2048;
                if (lnew.size() < (size_t)nr * (nnbmax + 1)) lnew.resize((size_t)nr * (nnbmax + 1));
                cnew.assign(nr, 0);
                fin.resize((size_t)3 * nr); fidn.resize((size_t)3 * nr);
                if (!resident_lists && old_rows.size() < (size_t)RCH * lmax) old_rows.resize((size_t)RCH * lmax);
                if (jj.size() < (size_t)2 * RCH * lmax) jj.resize((size_t)2 * RCH * lmax);
                idx1.resize(RCH); rs2.resize(RCH); nbl.resize(RCH); nbg.resize(RCH);
                for (int c0 = 0; c0 < nr; c0 += RCH) {
                    const int nc = std::min(RCH, nr - c0);
                    zf.assign((size_t)3 * nc, 0.0); zd.assign((size_t)3 * nc, 0.0); dfi.assign((size_t)3 * nc, 0.0); dfd.assign((size_t)3 * nc, 0.0);
#pragma omp parallel for schedule(static) if (nc >= OMP_MIN)
                    for (int q = 0; q < nc; q++) {
                        const int i = reg[c0 + q];
                        idx1[q] = i + 1; rs2[q] = rs[i] * rs[i];
                        if (resident_lists) continue;
                        int *o = &old_rows[(size_t)q * lmax];
                        o[0] = nnb[i];
                        for (int l = 0; l < nnb[i]; l++) o[1 + l] = nb[(size_t)i * (nnbmax + 1) + l] + 1;
                    }
                    int kk = nc, ifirst = 1, nn = n, lm = lmax, nm = nnbmax, nbsmin = 0; double smin = 0.0;
                    int *rows_c = &rows_raw[(size_t)c0 * lmax];
                    // a block that went through ONE gpunb_regf_ call still has its rows on the device: nothing is uploaded but the old lists
                    (nr <= MAXTHR && A.regcor_last ? A.regcor_last : A.regcor)(
                        &kk, idx1.data(), &ifirst, &nn, &nn, &lm, rows_c, resident_lists ? nullptr : old_rows.data(), rs2.data(), nullptr, &smin, &nm,
                        (d3 *)zf.data(), (d3 *)zd.data(), (d3 *)dfi.data(), (d3 *)dfd.data(), nbl.data(), nbg.data(), jj.data(), &nbsmin);
#pragma omp parallel for schedule(static) if (nc >= OMP_MIN)
                    for (int q = 0; q < nc; q++) {
                        const int *row = rows_c + (size_t)q * lmax;
                        cnew[c0 + q] = row[0];
                        for (int l = 0; l < row[0] && l <= nnbmax; l++) lnew[(size_t)(c0 + q) * (nnbmax + 1) + l] = row[1 + l] - 1;
                        for (int c = 0; c < 3; c++) {
                            fin[3 * (c0 + q) + c] = fio[3 * (c0 + q) + c] + dfi[3 * q + c]; fidn[3 * (c0 + q) + c] = fido[3 * (c0 + q) + c] + dfd[3 * q + c];
                        }
                    }
                }
                st->wall_regcor += wtime() - w0;
            } else {
#pragma omp parallel for schedule(static) if (nr >= OMP_MIN)
                for (int q = 0; q < nr; q++) { const int i = reg[q]; nnb[i] = cnew[q]; memcpy(&nb[(size_t)i * (nnbmax + 1)], &lnew[(size_t)q * (nnbmax + 1)], sizeof(int) * cnew[q]); }
                if (use_irr) irr_push_lists(reg);
                irregular(reg, tn, fin, fidn, have_snapshot);
            }
            PH[12] += pt() - P0; P0 = pt();
#pragma omp parallel for schedule(static) if (nr >= OMP_MIN)
            for (int q = 0; q < nr; q++) {
                const int i = reg[q];
                double fol[3], fdol[3], a2r[3], a3r[3];
                for (int c = 0; c < 3; c++) {      // regular polynomial over the OLD list at both ends (regcor_gpu.F:510-552)
                    fol[c] = (fin[3 * q + c] + frn_[3 * q + c]) - fio[3 * q + c]; fdol[c] = (fidn[3 * q + c] + frdn_[3 * q + c]) - fido[3 * q + c];
                }
                const double d = tn - t0r[i];
                hermite(&fr[3 * i], &frd[3 * i], fol, fdol, d, a2r, a3r);
                dtr_new[q] = aarseth(p->eta_r, fol, fdol, a2r, a3r, d);
                for (int c = 0; c < 3; c++) {
                    frnew[3 * regpos[q] + c] = frn_[3 * q + c]; frdnew[3 * regpos[q] + c] = frdn_[3 * q + c];
                    fia[3 * regpos[q] + c] = fin[3 * q + c]; fida[3 * regpos[q] + c] = fidn[3 * q + c];
                }
            }
        }
        PH[4] += pt() - P0; P0 = pt();
        // corrector (4th-order Hermite on the total force) and the new irregular steps (particles are independent: the host
        // side of the reference integrator is OpenMP as well)
#pragma omp parallel for schedule(static) if (na >= OMP_MIN)
        for (int q = 0; q < na; q++) {
            const int i = act[q];
            const double d = tn - t0[i];
            double f1[3], fd1[3], a2[3], a3[3];
            for (int c = 0; c < 3; c++) { f1[c] = fia[3 * q + c] + frnew[3 * q + c]; fd1[c] = fida[3 * q + c] + frdnew[3 * q + c]; }
            hermite(&f[3 * i], &fd[3 * i], f1, fd1, d, a2, a3);
            const double d2 = d * d, d3_ = d2 * d, d4 = d3_ * d, d5 = d4 * d;
            for (int c = 0; c < 3; c++) {
                x0[3 * i + c] = xa[3 * q + c] + d4 / 24.0 * a2[c] + d5 / 120.0 * a3[c];
                v0[3 * i + c] = va[3 * q + c] + d3_ / 6.0 * a2[c] + d4 / 24.0 * a3[c];
                fi[3 * i + c] = fia[3 * q + c]; fid[3 * i + c] = fida[3 * q + c]; f[3 * i + c] = f1[c]; fd[3 * i + c] = fd1[c];
            }
            t0[i] = tn;
            if (!shared_state) dirty[i] = 1;            // one byte per particle, distinct particles
            const double dt_new = aarseth(p->eta_i, f1, fd1, a2, a3, d), old = dt[i];
            double qd = dt_new < old ? std::max(pow2_floor(dt_new), dtmin) : old;
            if (dt_new >= 2.0 * old && fmod(tn, 2.0 * old) == 0.0 && 2.0 * old <= dtmax) qd = 2.0 * old;
            dt[i] = qd;
        }
        st->irr_steps += na;
        PH[5] += pt() - P0; P0 = pt();
        if (nr) {
#pragma omp parallel for schedule(static) if (nr >= OMP_MIN)
            for (int q = 0; q < nr; q++) {
                const int i = reg[q];
                for (int c = 0; c < 3; c++) { fr[3 * i + c] = frn_[3 * q + c]; frd[3 * i + c] = frdn_[3 * q + c]; }
                t0r[i] = tn;
                nnb[i] = cnew[q]; memcpy(&nb[(size_t)i * (nnbmax + 1)], &lnew[(size_t)q * (nnbmax + 1)], sizeof(int) * cnew[q]);
                const double oldr = dtr[i];
                double qr = dtr_new[q] < oldr ? pow2_floor(dtr_new[q]) : oldr;
                if (dtr_new[q] >= 2.0 * oldr && fmod(tn, 2.0 * oldr) == 0.0 && 2.0 * oldr <= dtmax) qr = 2.0 * oldr;
                dtr[i] = std::max(qr, dt[i]);
            }
            PH[11] += pt() - P0; P0 = pt();
            adjust_rs(reg, cnew);
            if (use_irr && use_regcor) irr_push_lists(reg);       // the host-list branch above has pushed the new lists already
        }
        PH[6] += pt() - P0; P0 = pt();
        for (int q = 0; q < na; q++) {          // an irregular step never exceeds the distance to the particle's next regular time
            const int i = act[q];
            const double nxt = t0r[i] + dtr[i] - tn;
            // pow2_floor(nxt) > nxt / 2 (or it is the cap dtmax >= dt): a step of at most nxt / 2 is never cut -- the rule for
            // most particles, and log2 / exp2 are the expensive part of this loop
            if (nxt > 0 && dt[i] > 0.5 * nxt) dt[i] = std::min(dt[i], pow2_floor(nxt));
        }
        {
            double key = -1.0; std::vector<int> *bucket = nullptr;          // most particles of a block keep their step
            for (int q = 0; q < na; q++) {
                const double tnext = t0[act[q]] + dt[act[q]];
                if (tnext != key) { key = tnext; bucket = &due_at(tnext); }
                bucket->push_back(act[q]);
            }
        }
        PH[7] += pt() - P0; P0 = pt();
        if (use_irr) irr_push_particles(act);
        PH[8] += pt() - P0;
        t = tn;
    }
    st->wall_total = wtime() - w_run;
    if (getenv("AC_DRIVER_PHASES")) fprintf(stderr, "PHASES pop+sort %.3f predict+frnew %.3f predict_all %.3f irregular(call) %.3f | reg: gather %.3f regular() %.3f lists+new-irregular %.3f regular-polynomial %.3f | corrector %.3f reg-update-loop %.3f adjust+push_lists %.3f cap+bucket %.3f push_particles %.3f\n", PH[0], PH[1], PH[2], PH[3], PH[9], PH[10], PH[12], PH[4], PH[5], PH[11], PH[6], PH[7], PH[8]);
    st->t = t;
    st->e1 = energy();
    double s = 0; for (int i = 0; i < n; i++) s += nnb[i];
    st->mean_nnb = s / n;
    if (x_out) memcpy(x_out, x0.data(), sizeof(double) * 3 * n);
    if (v_out) memcpy(v_out, v0.data(), sizeof(double) * 3 * n);
    if (use_irr) { int rk = 0; A.iclose(&rk); }
    A.close();
    return 0;
}

}  // extern "C"
