"""Device-resident predictor (gpunb_b200_state_all_/_state_update_/_predict_send_): the snapshot predicted
on the device must be BIT-FOR-BIT the one an (unfused fp64) host predictor restating xbpredall.f:17-26 uploads with
gpunb_send_, so every downstream result is identical."""
import numpy as np
import pytest

import oracle_lib
from nbody6ppgpu_b200 import hermite_ac as H
from nbody6ppgpu_b200 import snapshots as S

pytestmark = pytest.mark.gpu


def host_predict(x0, v0, f2, fd6, t0, time):
    s = (time - t0)[:, None]
    return ((fd6 * s + f2) * s + v0) * s + x0, (fd6 * (1.5 * s) + f2) * (2.0 * s) + v0


def test_predict_send_equals_host_predict_plus_send(b200):
    n = 5003
    rng = np.random.default_rng(11)
    m, x0, v0 = S.plummer(n, 7, "kroupa")
    f2 = 0.5 * rng.normal(size=(n, 3)); fd6 = rng.normal(size=(n, 3)) * 3.0
    t0 = rng.integers(0, 64, size=n) * 2.0 ** -10
    time = 0.0703125
    h2, dtr = S.radii(x0, m, S.rs0_for_nnb(n, 60.0))
    isel = slice(100, 100 + 777)
    b200.open(n + 10, 0)
    try:
        def both(nj, t):
            xp, vp = host_predict(x0[:nj], v0[:nj], f2[:nj], fd6[:nj], t0[:nj], t)
            b200.send(m[:nj], xp, vp)
            a = [q.copy() for q in b200.regf(h2[isel], dtr[isel], xp[isel], vp[isel], 400, 350, 0)]
            b200.predict_send(nj, t)
            gx, gv = b200.get_predicted(np.arange(isel.start, isel.stop))
            assert np.array_equal(gx, xp[isel]) and np.array_equal(gv, vp[isel])
            b = b200.regf(h2[isel], dtr[isel], xp[isel], vp[isel], 400, 350, 0)
            for q in range(3):
                assert np.array_equal(a[q], b[q])
            assert not oracle_lib.list_rows_equal(a[3], b[3])

        b200.state_all(m, x0, v0, f2, fd6, t0)
        both(n, time)
        # the integrator advances a few particles: only they are pushed
        idx = rng.choice(n, size=300, replace=False).astype(np.int32)
        x0[idx] += 1e-3 * rng.normal(size=(300, 3)); v0[idx] += 1e-3 * rng.normal(size=(300, 3))
        f2[idx] *= 1.01; fd6[idx] *= 0.99; t0[idx] = time; m[idx] *= 1.0 + 1e-6
        b200.state_update(idx, m[idx], x0[idx], v0[idx], f2[idx], fd6[idx], t0[idx])
        both(n, time + 2.0 ** -7)
        both(n - 131, time + 2.0 ** -6)           # NTOT shrinks: the first nj particles of the state
    finally:
        b200.close()


def test_device_predictor_against_the_golden_vectors_of_the_interpreted_fortran(b200):
    """gpunb_b200_state_all_ + _predict_send_ against tests/golden/xbpredall_f77.npz: the reference's own predictor statements
    (xbpredall.f:18-26) executed by oracle/f77_interp.py.  Bit for bit."""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "xbpredall_f77.npz")
    n = g["m"].shape[0]
    b200.open(n + 10, 0)
    try:
        b200.state_all(g["m"], g["x0"], g["x0dot"], g["f"], g["fdot"], g["t0"])
        b200.predict_send(n, float(g["time"]))
        gx, gv = b200.get_predicted(np.arange(n))
        assert np.array_equal(gx, g["f77_x"]) and np.array_equal(gv, g["f77_xdot"])
    finally:
        b200.close()


def test_predict_send_from_the_irregular_force_librarys_table(b200):
    """One copy of the particle state on the device: gpunb_b200_predict_send_records_ predicts the regular-force snapshot
    from the table irr_simd_set_jp_ keeps current in libirr_b200.so (X0, X0DOT, F/2, FDOT/6, BODY, T0 per particle).  The
    snapshot is bit for bit the one predict_send builds from its own state, for particles IFIRST..NTOT of the table."""
    from nbody6ppgpu_b200 import irr
    ntot, ifirst = 4100, 41                           # the j-set starts behind 40 KS components
    nj = ntot - ifirst + 1
    rng = np.random.default_rng(21)
    m, x0, v0 = S.plummer(ntot, 9, "kroupa")
    f2 = 0.5 * rng.normal(size=(ntot, 3)); fd6 = rng.normal(size=(ntot, 3)) * 3.0
    t0 = rng.integers(0, 64, size=ntot) * 2.0 ** -10
    il = irr.IrrLib(irr.lib_path())
    il.open(ntot, 64, 0)
    b200.open(nj + 10, 0)
    try:
        addr = np.arange(1, ntot + 1, dtype=np.int32)
        il.set_jp_batch(addr, x0, v0, f2, fd6, m, t0)
        for time in (0.0703125, 0.078125):
            il.flush()
            base, stride = il.particle_records()
            b200.predict_send_records(nj, time, base + 8 * stride * (ifirst - 1), stride)
            idx = np.arange(0, nj, 7, dtype=np.int32)
            gx, gv = b200.get_predicted(idx)
            j = slice(ifirst - 1, ntot)
            xp, vp = host_predict(x0[j], v0[j], f2[j], fd6[j], t0[j], time)
            assert np.array_equal(gx, xp[idx]) and np.array_equal(gv, vp[idx])
            # a block of particles advances: set_jp is all the integrator does
            adv = rng.choice(ntot, size=200, replace=False)
            x0[adv] += 1e-3 * rng.normal(size=(200, 3)); f2[adv] *= 1.01; t0[adv] = time
            il.set_jp_batch(adv.astype(np.int32) + 1, x0[adv], v0[adv], f2[adv], fd6[adv], m[adv], t0[adv])
    finally:
        b200.close()
        il.close(0)


def test_ac_driver_with_device_predictor_is_bitwise_the_host_path(b200):
    m, x, v = S.plummer(512, 3, "equal")
    res = {}
    for dev in (False, True):
        ac = H.AhmadCohen(b200, m, x, v, nnbopt=30, device_predictor=dev)
        try:
            st = ac.run(0.25)
        finally:
            ac.close()
        res[dev] = (ac.x0.copy(), ac.v0.copy(), st.energies[-1][1], st.regf_calls)
    assert np.array_equal(res[False][0], res[True][0]) and np.array_equal(res[False][1], res[True][1])
    assert res[False][2] == res[True][2] and res[False][3] == res[True][3]


def test_reused_tile_order_stays_exact(b200, oracle):
    """GPUNB_B200_RESORT_EVERY = 4: three of four snapshots keep the previous Hilbert permutation and only re-pack the
    tiles.  Boxes and offsets come from the current positions, so lists stay bit-exact and forces within 1e-6 while the
    particles drift (here far more than between two regular blocks)."""
    n = 12000
    m, x, v = S.plummer(n, 17, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 80.0))
    isel = slice(2000, 2000 + 1024)
    b200.open(n + 10, 0)
    try:
        b200.set_resort_every(4)
        z3 = np.zeros((n, 3))
        for step in range(7):
            xs = x + v * (0.03 * step)
            if step % 2:
                b200.state_all(m, xs, v, z3, z3, np.zeros(n))
                b200.predict_send(n, 0.0)
            else:
                b200.send(m, xs, v)
            acc, jrk, pot, lst = b200.regf(h2[isel], dtr[isel], xs[isel], v[isel], 400, 350, 0)
            a64, j64, p64, l64, band, _ = oracle.regf_f64(m, xs, v, h2[isel], dtr[isel], xs[isel], v[isel], 400, 350, 0, 4.0)
            bad = [i for i in oracle_lib.list_rows_equal(lst, l64) if band[i] > 4.0]
            assert not bad, (step, bad[:5])
            assert oracle_lib.relerr(acc, a64) <= 1e-6 and oracle_lib.relerr(pot, p64) <= 1e-6
            assert oracle_lib.relerr_scaled(jrk, j64, oracle.scale[:, 1]) <= 1e-6
            phi = b200.gpupot(1, n, m, xs) if step == 3 else None      # gpupot in between invalidates the kept order
    finally:
        b200.set_resort_every(0)
        b200.close()


def test_adaptive_tile_order(b200, oracle):
    """Default mode (GPUNB_B200_RESORT_EVERY = 0): the Hilbert order of the previous snapshot is kept while the tiles stay
    compact (summed half-extents within 10 % of their value after the last sort) and refreshed once the particles have
    drifted.  Small drifts keep the order, a large one triggers a sort; lists stay bit-exact and forces within 1e-6 on
    every snapshot, and the decision sequence is reproducible."""
    n = 12000
    m, x, v = S.plummer(n, 19, "kroupa")
    h2, dtr = S.radii(x, m, S.rs0_for_nnb(n, 80.0))
    isel = slice(3000, 3000 + 512)
    drifts = [0.0, 1e-4, 2e-4, 3e-4, 0.3, 0.3001, 0.3002]
    kept_runs = []
    for rep in range(2):
        b200.open(n + 10, 0)
        try:
            b200.set_resort_every(0)
            kept = []
            for dt in drifts:
                xs = x + v * dt
                b200.reset_counters()
                b200.send(m, xs, v)
                kept.append(int(b200.counters()["sends_order_kept"]))
                acc, jrk, pot, lst = b200.regf(h2[isel], dtr[isel], xs[isel], v[isel], 400, 350, 0)
                a64, j64, p64, l64, band, _ = oracle.regf_f64(m, xs, v, h2[isel], dtr[isel], xs[isel], v[isel], 400, 350, 0, 4.0)
                assert not [i for i in oracle_lib.list_rows_equal(lst, l64) if band[i] > 4.0], dt
                assert oracle_lib.relerr(acc, a64) <= 1e-6 and oracle_lib.relerr(pot, p64) <= 1e-6
                assert oracle_lib.relerr_scaled(jrk, j64, oracle.scale[:, 1]) <= 1e-6
            kept_runs.append(kept)
        finally:
            b200.close()
    assert kept_runs[0] == kept_runs[1]                    # reproducible decisions
    assert kept_runs[0][0] == 0                            # the first snapshot after open is always sorted
    assert kept_runs[0][1:4] == [1, 1, 1]                  # tiny drifts keep the order
    # a large drift is seen when its tiles are packed (in the kept order, still exact) and triggers the sort of the NEXT snapshot
    assert kept_runs[0][4] == 1 and kept_runs[0][5] == 0 and kept_runs[0][6] == 1, kept_runs[0]
