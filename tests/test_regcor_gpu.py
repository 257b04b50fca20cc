"""gpunb_b200_regcor_ (device) against oracle/regcor_oracle.c on the same rows: every integer result equal, every fp64
result BIT FOR BIT (both sides perform the single IEEE operations the Fortran names, in its order)."""
import json
import os
import time

import numpy as np
import pytest

import sys
from pathlib import Path

import regcor_cases as RC

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import make_regcor_golden as MG  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = sorted((ROOT / "tests" / "golden").glob("regcor_f77_*.npz"))
KEYS = ("nlist_used", "nbloss", "nbgain", "jj_used", "freg", "fdr", "dfirr", "dfd")


def used(out, c_old):
    """Entries in use of NLIST / JJLIST (what lies behind them is unspecified)."""
    nl, jj = out["nlist"], out["jjlist"]
    a = [list(nl[r, :max(nl[r, 0], 0) + 1]) for r in range(nl.shape[0])]
    b = [(list(jj[r, :out["nbloss"][r]]), list(jj[r, c_old[r, 0]:c_old[r, 0] + out["nbgain"][r]])) for r in range(nl.shape[0])]
    return a, b


def same(dev, ora, old):
    da, db = used(dev, old); oa, ob = used(ora, old)
    assert da == oa, "NLIST differs"
    assert db == ob, "JJLIST differs"
    for k in ("nbloss", "nbgain", "freg", "fdr", "dfirr", "dfd"):
        assert np.array_equal(dev[k], ora[k]), k
    assert dev["nbsmin"] == ora["nbsmin"]


def both(b200, oracle, c, old="host", step="host"):
    args = (c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"])
    tail = (c["rs2"], c["step"] if step == "host" else None, c["smin"], c["nnbmax"], c["freg"], c["fdr"])
    dev = b200.regcor(*args, c["old"] if old == "host" else None, *tail)
    ora = oracle.regcor(*args, c["old"], c["m"], c["x"], c["v"], c["rs2"], c["step"] if step else None, c["smin"], c["nnbmax"],
                        c["freg"], c["fdr"])
    return dev, ora


def test_device_rows_equal_the_oracle_bit_for_bit(b200, oracle):
    c = RC.make_case(overflow_rows=(5,), empty_old_rows=(0, 17))
    nj = c["m"].shape[0]
    b200.open(nj + 10, 0)
    try:
        b200.send(c["m"], c["x"], c["v"])
        dev, ora = both(b200, oracle, c)
        same(dev, ora, c["old"])
        assert ora["nbsmin"] > 0 and ora["nbloss"].sum() > 100 and ora["nbgain"].sum() > 100
        # no steps at all: nothing is retained
        dev0 = b200.regcor(c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"], c["old"], c["rs2"], None, c["smin"],
                           c["nnbmax"], c["freg"], c["fdr"])
        ora0 = oracle.regcor(c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"], c["old"], c["m"], c["x"], c["v"],
                             c["rs2"], None, c["smin"], c["nnbmax"], c["freg"], c["fdr"])
        same(dev0, ora0, c["old"])
        assert dev0["nbsmin"] == 0
        # c.m. rows (I > N) and a single row
        c1 = {**c}
        for k in ("index_i", "new", "old", "rs2", "freg", "fdr"):
            c1[k] = c[k][33:34]
        d1, o1 = both(b200, oracle, c1)
        same(d1, o1, c1["old"])
    finally:
        b200.close()


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: p.stem)
def test_device_rows_against_the_golden_vectors_of_the_interpreted_fortran(b200, path):
    """The CUDA path against what the reference's OWN Fortran text (util_gpu.F:102-111 + regcor_gpu.F:263-459, executed by
    oracle/f77_interp.py; fixtures by oracle/make_regcor_golden.py) leaves behind: integers equal, fp64 bit for bit."""
    c, g = MG.load_case(path)
    b200.open(c["m"].shape[0] + 10, 0)
    try:
        b200.send(c["m"], c["x"], c["v"])
        dev = b200.regcor(c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"], c["old"], c["rs2"], c["step"], c["smin"],
                          c["nnbmax"], c["freg"], c["fdr"])
        rows = [r for r in range(c["index_i"].shape[0]) if g["f77_valid"][r]]
        retained = RC.compare_rows(dev, c, rows, MG.golden_walk(g))
        assert retained == dev["nbsmin"] == int(g["f77_nbsmin"].sum())
    finally:
        b200.close()


def test_resident_lists_and_steps(b200, oracle):
    c = RC.make_case(seed=11, n_tot=5000, ni=700, lmax=160, nnb_mean=45.0)
    b200.open(c["m"].shape[0] + 10, 0)
    try:
        b200.send(c["m"], c["x"], c["v"])
        b200.lists_put(c["index_i"], c["old"])
        back = b200.lists_get(c["index_i"], c["lmax"])
        for r in range(back.shape[0]):
            assert np.array_equal(back[r, :back[r, 0] + 1], c["old"][r, :c["old"][r, 0] + 1])
        b200.steps_all(c["step"])
        # a few steps change afterwards (the integrator has advanced those particles)
        idx = np.arange(0, c["m"].shape[0], 37, dtype=np.int32)
        c["step"][idx] *= 0.5
        b200.steps_update(idx, c["step"][idx])
        dev, ora = both(b200, oracle, c, old="store", step=None)
        ora = oracle.regcor(c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"], c["old"], c["m"], c["x"], c["v"], c["rs2"],
                            c["step"], c["smin"], c["nnbmax"], c["freg"], c["fdr"])
        same(dev, ora, c["old"])
        assert ora["nbsmin"] > 0
        # the kernel committed the final NLIST of every row: the next block diffs against it without any upload
        held = b200.lists_get(c["index_i"], c["lmax"])
        for r in range(held.shape[0]):
            assert np.array_equal(held[r, :held[r, 0] + 1], dev["nlist"][r, :dev["nlist"][r, 0] + 1])
        c2 = {**c, "old": dev["nlist"].copy()}
        dev2, _ = both(b200, oracle, c2, old="store", step=None)
        ora2 = oracle.regcor(c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"], c2["old"], c["m"], c["x"], c["v"],
                             c["rs2"], c["step"], c["smin"], c["nnbmax"], c["freg"], c["fdr"])
        same(dev2, ora2, c2["old"])
        assert dev2["nbgain"].sum() == 0          # same new rows again: nothing is gained (retained members are lost and put back)
    finally:
        b200.close()


def test_rows_of_gpunb_regf_at_lmax_600_in_batches(b200, oracle):
    """The real pipeline: gpunb_send_, gpunb_regf_ for 2500 particles in calls of 1024 (old lists: the same spheres 6 % larger at
    a drifted snapshot), then ONE gpunb_b200_regcor_ over all rows (the library cuts it at 2048)."""
    from nbody6ppgpu_b200 import snapshots as S
    n, lmax, nnbmax, ifirst = 40000, 600, 550, 21
    m, x, v = S.plummer(n, 4, "kroupa")
    h2, dtr = S.radii_nnb(x, m, 120.0)
    rng = np.random.default_rng(2)
    rows_j = np.sort(rng.choice(n, 2500, replace=False))
    b200.open(n + 10, 0)
    try:
        lists = {}
        for tag, (xs, fac) in {"old": (x + 0.002 * rng.normal(size=x.shape), 1.06), "new": (x, 1.0)}.items():
            b200.send(m, xs, v)
            out = []
            for k0 in range(0, rows_j.size, 1024):
                sel = rows_j[k0:k0 + 1024]
                out.append(b200.regf(h2[sel] * fac, dtr[sel], xs[sel], v[sel], lmax, nnbmax, 0)[3].copy())
            lists[tag] = np.concatenate(out)
        assert (lists["new"][:, 0] > 0).all() and (lists["old"][:, 0] > 0).all()
        old = np.zeros_like(lists["old"])
        for r in range(old.shape[0]):                                        # the caller's LIST: Fortran numbers, self removed
            row = lists["old"][r, 1:1 + lists["old"][r, 0]]
            row = row[row != rows_j[r]] + ifirst
            old[r, 0] = row.size; old[r, 1:1 + row.size] = row
        step = 2.0 ** -rng.integers(3, 12, size=n).astype(np.float64)
        c = dict(m=m, x=x, v=v, index_i=(rows_j + ifirst).astype(np.int32), ifirst=ifirst, n=n + ifirst - 1 - 100,
                 ntot=n + ifirst - 1, lmax=lmax, nnbmax=nnbmax, new=lists["new"], old=old, rs2=h2[rows_j], step=step,
                 smin=float(np.quantile(step, 0.2)), freg=rng.normal(size=(rows_j.size, 3)), fdr=rng.normal(size=(rows_j.size, 3)))
        t0 = time.perf_counter()
        dev = b200.regcor(c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"], c["old"], c["rs2"], c["step"], c["smin"],
                          c["nnbmax"], c["freg"], c["fdr"])
        t_dev = time.perf_counter() - t0
        t0 = time.perf_counter()
        ora = oracle.regcor(c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"], c["old"], c["m"], c["x"], c["v"], c["rs2"],
                            c["step"], c["smin"], c["nnbmax"], c["freg"], c["fdr"])
        t_ora = time.perf_counter() - t0
        same(dev, ora, c["old"])
        assert ora["nbloss"].sum() > 1000 and ora["nbsmin"] > 0
        # latency of one block of 1024 rows: device call (host lists in, results out) vs the CPU restatement on all host threads
        c1 = {**c}
        for k in ("index_i", "new", "old", "rs2", "freg", "fdr"):
            c1[k] = c[k][:1024]
        b200.lists_put(c1["index_i"], c1["old"])
        b200.steps_all(step)
        t = {}
        for name, fn in (("device_host_lists", lambda: b200.regcor(c1["index_i"], ifirst, c["n"], c["ntot"], c1["new"], c1["old"], c1["rs2"], None, c["smin"], nnbmax, c1["freg"], c1["fdr"])),
                         ("device_resident_lists", lambda: (b200.lists_put(c1["index_i"][:1], c1["old"][:1]), b200.regcor(c1["index_i"], ifirst, c["n"], c["ntot"], c1["new"], None, c1["rs2"], None, c["smin"], nnbmax, c1["freg"], c1["fdr"]))),
                         ("oracle_all_threads", lambda: oracle.regcor(c1["index_i"], ifirst, c["n"], c["ntot"], c1["new"], c1["old"], m, x, v, c1["rs2"], step, c["smin"], nnbmax, c1["freg"], c1["fdr"]))):
            fn()
            b200.reset_counters()
            t0 = time.perf_counter()
            inner = 0.0
            for _ in range(20):
                fn()
                inner += getattr(oracle, "last_regcor_s", 0.0) if name.startswith("oracle") else 0.0
            t[name] = (time.perf_counter() - t0) / 20 * 1e6
            # the C call alone (without the numpy marshalling of the Python mirrors)
            t[name + "_c_call"] = inner / 20 * 1e6 if name.startswith("oracle") else b200.counters()["regcor_ms"] / 20 * 1e3
        print("regcor, 1024 rows, lmax 600, <nnb> ~ %.0f: us per call through the Python mirror" % lists["new"][:, 0].mean(), json.dumps(t))
        if os.environ.get("GPUNB_REGCOR_OUT"):
            with open(os.environ["GPUNB_REGCOR_OUT"], "w") as f:
                json.dump({"rows": 1024, "lmax": lmax, "mean_nnb": float(lists["new"][:, 0].mean()), "us_per_call": t,
                           "first_batch_2500_rows_s": {"device": t_dev, "oracle": t_ora}}, f, indent=1)
    finally:
        b200.close()


def test_rows_still_on_the_device(b200, oracle):
    """gpunb_b200_regcor_last_: the rows of the last gpunb_regf_ call are taken from the copy that call left on the device
    (nothing is uploaded but index_i, rs2 and the four force vectors); with the resident list store and resident steps no
    list crosses PCIe on the way in.  Same results as the oracle on the rows the call returned to the host."""
    from nbody6ppgpu_b200 import snapshots as S
    n, lmax, nnbmax, ifirst = 30000, 400, 350, 11
    m, x, v = S.plummer(n, 6, "kroupa")
    h2, dtr = S.radii_nnb(x, m, 90.0)
    rng = np.random.default_rng(4)
    step = 2.0 ** -rng.integers(3, 12, size=n).astype(np.float64)
    smin = float(np.quantile(step, 0.2))
    b200.open(n + 10, 0)
    try:
        for ni, sub in ((1024, 1), (777, -3), (40, 1)):                        # one launch, forced sub-blocks, a small unsorted block
            rows_j = np.sort(rng.choice(n, ni, replace=False))
            index_i = (rows_j + ifirst).astype(np.int32)
            b200.set_tuning(0, sub)
            b200.send(m, x + 0.002 * rng.normal(size=x.shape), v)
            old_raw = b200.regf(h2[rows_j] * 1.05, dtr[rows_j], x[rows_j], v[rows_j], lmax, nnbmax, 0)[3].copy()
            old = np.zeros_like(old_raw)
            for r in range(ni):
                row = old_raw[r, 1:1 + old_raw[r, 0]]
                row = row[row != rows_j[r]] + ifirst
                old[r, 0] = row.size; old[r, 1:1 + row.size] = row
            b200.send(m, x, v)
            new = b200.regf(h2[rows_j], dtr[rows_j], x[rows_j], v[rows_j], lmax, nnbmax, 0)[3].copy()
            freg = rng.normal(size=(ni, 3)); fdr = rng.normal(size=(ni, 3))
            b200.lists_put(index_i, old); b200.steps_all(step)
            b200.reset_counters()
            dev = b200.regcor(index_i, ifirst, n + ifirst - 1, n + ifirst - 1, None, None, h2[rows_j], None, smin, nnbmax, freg, fdr, last_lmax=lmax)
            c = b200.counters()
            ora = oracle.regcor(index_i, ifirst, n + ifirst - 1, n + ifirst - 1, new, old, m, x, v, h2[rows_j], step, smin, nnbmax, freg, fdr)
            same(dev, ora, old)
            assert ora["nbloss"].sum() > 0 and ora["nbgain"].sum() > 0
            assert c["h2d_bytes"] < 200.0 * ni + 64, c["h2d_bytes"]             # index, rs2, four vectors per row: no list went up
            print(f"regcor_last ni {ni}: {c['regcor_ms'] * 1e3:.1f} us inside the call, oracle {oracle.last_regcor_s * 1e6:.1f} us")
        b200.set_tuning(0, 4)
    finally:
        b200.close()


@pytest.mark.parametrize("smin_mode,nnbmax", [("case", None), ("all_small", None), ("case", 20)])
def test_edge_rows(b200, oracle, smin_mode, nnbmax):
    """The hand-built rows of tests/test_regcor_cpu.py::test_edge_rows on the device: only self, both lists empty, identical
    lists, no self, the longest rows lmax allows; every lost member retained / none retained."""
    c = RC.make_edge_case(seed=21, ni=120, n_tot=1500, lmax=96, nnb_mean=24.0)
    if smin_mode == "all_small":
        c["smin"] = 10.0
    if nnbmax is not None:
        c["nnbmax"] = nnbmax
    b200.open(c["m"].shape[0] + 10, 0)
    try:
        b200.send(c["m"], c["x"], c["v"])
        dev, ora = both(b200, oracle, c)
        same(dev, ora, c["old"])
        # the same through the resident store
        b200.lists_put(c["index_i"], c["old"]); b200.steps_all(c["step"])
        dev2 = b200.regcor(c["index_i"], c["ifirst"], c["n"], c["ntot"], c["new"], None, c["rs2"], None, c["smin"], c["nnbmax"],
                           c["freg"], c["fdr"])
        same(dev2, ora, c["old"])
    finally:
        b200.close()


def test_random_small_lists(b200, oracle):
    """The 400 random short rows of tests/test_regcor_cpu.py on the device."""
    c = RC.make_random_case()
    b200.open(c["m"].shape[0] + 10, 0)
    try:
        b200.send(c["m"], c["x"], c["v"])
        dev, ora = both(b200, oracle, c)
        same(dev, ora, c["old"])
        assert ora["nbsmin"] > 30
    finally:
        b200.close()
