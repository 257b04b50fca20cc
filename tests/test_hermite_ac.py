"""Ahmad-Cohen driver (nbody6ppgpu_b200/hermite_ac.py): energy conservation over the regular-force ABI.

CPU: the driver itself is exercised with the reference's AVX library behind the ABI (the checker; tests only).
GPU: the north-star's drift criterion -- the energy error over one N-body time unit with libgpunb_b200.so is laid
beside the reference library's on the identical snapshot and integrator settings."""
import numpy as np
import pytest

from nbody6ppgpu_b200 import hermite_ac as H
from nbody6ppgpu_b200 import snapshots as S


def test_ac_driver_conserves_energy_with_reference_library(ref_avx):
    if ref_avx is None:
        pytest.skip("oracle/_ref not built")
    m, x, v = S.plummer(256, 5, "equal")
    ac = H.AhmadCohen(ref_avx, m, x, v, nnbopt=30)
    try:
        st = ac.run(0.25)
    finally:
        ac.close()
    e0, e1 = st.energies[0][1], st.energies[-1][1]
    assert abs(e0 + 0.25) < 0.05                      # Plummer sphere in N-body units: E = -1/4
    assert abs((e1 - e0) / e0) < 1e-4, (e0, e1)
    assert st.t == 0.25 and st.reg_steps > 256 and st.irr_steps > st.reg_steps
    assert np.all(ac.t0 == 0.25)                      # block steps re-synchronise at multiples of dtmax
    assert 0.3 * 30 < ac.nnb.mean() < 3.0 * 30        # RS control keeps the lists near NNBOPT
    # lists are ascending, self-free and irregular steps never exceed regular ones
    for i in range(0, 256, 17):
        row = ac.nb[i, :ac.nnb[i]]
        assert np.all(np.diff(row) > 0) and i not in row
    assert np.all(ac.dt <= ac.dtr)


@pytest.mark.gpu
def test_energy_drift_matches_reference_library(b200, ref_avx):
    """Energy error over one N-body time unit: this library vs the reference's behind the same driver."""
    out, stats = {}, {}
    for name, lib in (("b200", b200), ("avx", ref_avx)):
        de, st = H.energy_drift(lib, n=1024, seed=5, t_end=1.0, nnbopt=40)
        out[name] = de
        stats[name] = {"irr_steps": st.irr_steps, "reg_steps": st.reg_steps, "regf_calls": st.regf_calls,
                       "overflow_retries": st.overflow_retries, "wall_s": st.wall_total, "wall_regf_s": st.wall_regf,
                       "wall_send_s": st.wall_send}
        print(f"{name}: dE/E = {de:+.3e} over t=1 ({st.irr_steps} irregular, {st.reg_steps} regular steps, "
              f"{st.regf_calls} gpunb_regf_ calls, {st.overflow_retries} overflow retries, {st.wall_total:.1f} s)")
    import json, os
    if os.environ.get("GPUNB_DRIFT_OUT"):         # GPU sessions keep the two numbers (profiles/)
        with open(os.environ["GPUNB_DRIFT_OUT"], "w") as fh:
            json.dump({"n": 1024, "t_end": 1.0, "nnbopt": 40, "dE_over_E": out, "stats": stats}, fh)
    assert abs(out["b200"]) <= 3.0 * abs(out["avx"]) + 1e-5, out


def test_ac_driver_with_imf_and_mass_weighted_criterion(ref_avx):
    """Kroupa IMF + m_flag = 1 (neighbour criterion r^2 < h2 * m_j, KZ(39) = 2 as in samples/N16k.input): the RS control,
    the overflow retry loop and the energy bookkeeping of the driver hold with the reference library behind it."""
    if ref_avx is None:
        pytest.skip("oracle/_ref not built")
    m, x, v = S.plummer(384, 8, "kroupa")
    ac = H.AhmadCohen(ref_avx, m, x, v, nnbopt=24, lmax=128, m_flag=1)
    try:
        st = ac.run(0.125)
    finally:
        ac.close()
    e0, e1 = st.energies[0][1], st.energies[-1][1]
    assert abs((e1 - e0) / e0) < 5e-4, (e0, e1)
    assert st.reg_steps > 384 and np.all(ac.t0 == 0.125)
    assert ac.nnb.max() <= ac.nnbmax


def test_ac_driver_with_the_irregular_force_library(ref_avx):
    """The driver's library path for the irregular force (set_jp / set_list / firr_vec, intgrt.F:199-207,545) with the
    reference's own AVX libraries on both sides: same block structure as the numpy path, energy conserved."""
    from pathlib import Path
    from nbody6ppgpu_b200 import irr
    so = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libirr_ref_avx.so"
    if ref_avx is None or not so.exists():
        pytest.skip("oracle/_ref not built")
    m, x, v = S.plummer(256, 5, "equal")
    ac = H.AhmadCohen(ref_avx, m, x, v, nnbopt=30, irr_lib=irr.IrrLib(so))
    try:
        st = ac.run(0.25)
    finally:
        ac.close()
    e0, e1 = st.energies[0][1], st.energies[-1][1]
    assert abs((e1 - e0) / e0) < 2e-4, (e0, e1)          # the AVX library predicts and sums in FP32
    assert st.t == 0.25 and st.irr_steps > st.reg_steps > 256 and st.wall_irr > 0
    for i in range(0, 256, 17):
        row = ac.nb[i, :ac.nnb[i]]
        assert np.all(np.diff(row) > 0) and i not in row


@pytest.mark.gpu
def test_device_paths_of_the_driver(b200):
    """Irregular force through libirr_b200.so, list bookkeeping through gpunb_b200_regcor_, snapshot from the device-resident
    predictor: the same integration as the host statements of those steps (numpy), to rounding."""
    from nbody6ppgpu_b200 import irr
    m, x, v = S.plummer(1024, 5, "kroupa")
    out = {}
    for name, kw in (("host", {}),
                     ("irr", dict(irr_lib=irr.IrrLib(irr.lib_path()))),
                     ("device", dict(irr_lib=irr.IrrLib(irr.lib_path()), use_regcor=True, device_predictor=True))):
        ac = H.AhmadCohen(b200, m, x, v, nnbopt=40, m_flag=1, **kw)
        try:
            st = ac.run(0.25)
        finally:
            ac.close()
        e0, e1 = st.energies[0][1], st.energies[-1][1]
        out[name] = dict(de=(e1 - e0) / abs(e0), irr=st.irr_steps, reg=st.reg_steps, blocks=st.block_steps, wall=st.wall_total,
                         x=ac.x0.copy(), nnb=float(ac.nnb.mean()))
        print(name, {k: out[name][k] for k in ("de", "irr", "reg", "blocks", "wall", "nnb")})
    for name in ("irr", "device"):
        assert abs(out[name]["de"]) < 1e-4
        # same block structure up to the few steps a last-bit difference in a force can shift
        assert abs(out[name]["irr"] - out["host"]["irr"]) < 0.02 * out["host"]["irr"]
        assert abs(out[name]["reg"] - out["host"]["reg"]) < 0.02 * out["host"]["reg"]
        assert np.median(np.linalg.norm(out[name]["x"] - out["host"]["x"], axis=1)) < 1e-6


def test_native_driver_is_the_python_driver_bit_for_bit(ref_avx):
    """csrc/ac_driver.cpp (libac_driver.so) restates hermite_ac.py statement by statement: behind the same (reference)
    libraries the two integrations agree in every step count and in every bit of the final positions."""
    from pathlib import Path
    from nbody6ppgpu_b200 import ac_native, irr
    ref = Path(__file__).resolve().parent.parent / "oracle" / "_ref"
    if ref_avx is None or not (ref / "libirr_ref_avx.so").exists():
        pytest.skip("oracle/_ref not built")
    assert ac_native.lib_path().exists(), "libac_driver.so not built: run __graft_entry__.build()"
    m, x, v = S.plummer(512, 7, "kroupa")
    for irr_so, m_flag in ((None, 0), (ref / "libirr_ref_avx.so", 1)):
        kw = dict(nnbopt=30, lmax=128, m_flag=m_flag)
        ac = H.AhmadCohen(ref_avx, m, x, v, irr_lib=irr.IrrLib(irr_so) if irr_so else None, **kw)
        try:
            st = ac.run(0.25)
        finally:
            ac.close()
        out, xo, vo = ac_native.run(ref / "libgpunb_ref_avx.so", irr_so, m, x, v, 0.25, **kw)
        assert (out["block_steps"], out["irr_steps"], out["reg_steps"], out["regf_calls"], out["overflow_retries"]) == \
               (st.block_steps, st.irr_steps, st.reg_steps, st.regf_calls, st.overflow_retries)
        assert np.array_equal(xo, ac.x0) and np.array_equal(vo, ac.v0)
        assert abs(out["e0"] - st.energies[0][1]) < 1e-13 and abs(out["e1"] - st.energies[-1][1]) < 1e-13    # summation order of E only
        assert abs(out["dE_over_E"]) < 2e-4


@pytest.mark.gpu
def test_native_driver_on_the_device_paths(b200):
    """The native driver behind this repo's three device paths (libgpunb_b200.so with the device-resident predictor and
    gpunb_b200_regcor_, libirr_b200.so): the same integration as the Python driver on the same paths."""
    from nbody6ppgpu_b200 import ac_native, irr, lib_path
    m, x, v = S.plummer(2048, 5, "kroupa")
    kw = dict(nnbopt=40, lmax=128, m_flag=1)
    ac = H.AhmadCohen(b200, m, x, v, irr_lib=irr.IrrLib(irr.lib_path()), use_regcor=True, device_predictor=True, **kw)
    try:
        st = ac.run(0.25)
    finally:
        ac.close()
    out, xo, vo = ac_native.run(lib_path(), irr.lib_path(), m, x, v, 0.25, use_predictor=True, use_regcor=True, **kw)
    print({k: out[k] for k in ("wall_total", "wall_send", "wall_regf", "wall_irr", "wall_regcor", "block_steps", "irr_steps", "dE_over_E")},
          "python wall", st.wall_total)
    assert (out["block_steps"], out["irr_steps"], out["reg_steps"]) == (st.block_steps, st.irr_steps, st.reg_steps)
    assert np.array_equal(xo, ac.x0) and np.array_equal(vo, ac.v0)
    assert abs(out["dE_over_E"]) < 1e-4
    # one copy of the state on the device: the predictor reads the irr library's particle table -- the same snapshots, bit for bit
    shared, xs, vs = ac_native.run(lib_path(), irr.lib_path(), m, x, v, 0.25, use_predictor=2, use_regcor=2, **kw)
    assert np.array_equal(xs, xo) and np.array_equal(vs, vo)
    print("shared state:", {k: shared[k] for k in ("wall_total", "wall_send", "wall_regf", "wall_irr", "wall_regcor")})
    host, xh, _ = ac_native.run(lib_path(), irr.lib_path(), m, x, v, 0.25, **kw)         # reference ABI only: host predictor, host lists
    assert abs(host["irr_steps"] - out["irr_steps"]) < 0.02 * out["irr_steps"]
    assert np.median(np.linalg.norm(xh - xo, axis=1)) < 1e-6
