#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own CPU library
(oracle/_ref/libgpunb_ref_avx.so, compiled from /root/reference/src/Main/{reg,pot}.avx.cpp by
oracle/Makefile) on seeded snapshots.  Test infrastructure only.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Fixtures store only the generator parameters and the reference outputs (a few hundred KB).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib  # noqa: E402
from nbody6ppgpu_b200 import snapshots as S  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)
ref = oracle_lib.ref_avx()

CASES = [  # n, seed, imf, m_flag, nnb target, i-selection, lmax, nnbmax
    (2048, 1, "equal", 0, 100.0, np.arange(0, 128), 400, 350),
    (2048, 2, "kroupa", 1, 100.0, np.arange(1900, 2048), 400, 350),
    (16384, 1, "kroupa", 0, 100.0, np.arange(0, 16384, 128), 400, 350),
    (16384, 3, "kroupa", 1, 150.0, np.arange(5, 16384, 97), 400, 350),
    (4099, 4, "kroupa", 0, 300.0, np.arange(0, 61), 128, 78),        # odd nj, overflow rows
]
for n, seed, imf, m_flag, nnb, isel, lmax, nnbmax in CASES:
    m, x, v = S.plummer(n, seed, imf)
    rs0 = S.rs0_for_nnb(n, nnb)
    h2, dtr = S.radii(x, m, rs0, 0.125, m_flag)
    ref.open(n + 10, 0)
    ref.send(m, x, v)
    acc, jrk, pot, lst = ref.regf(h2[isel], dtr[isel], x[isel], v[isel], lmax, nnbmax, m_flag)
    ref.close()
    # entries past the count are unspecified: zero them so the fixture is deterministic
    for i in range(lst.shape[0]):
        lst[i, 1 + max(int(lst[i, 0]), 0):] = 0
    f = OUT / f"regf_n{n}_s{seed}_{imf}_m{m_flag}.npz"
    np.savez_compressed(f, n=n, seed=seed, imf=imf, m_flag=m_flag, rs0=rs0, isel=isel, lmax=lmax, nnbmax=nnbmax,
                        acc=acc, jrk=jrk, pot=pot, list=lst)
    print(f.name, "mean |nnb|", np.abs(lst[:, 0]).mean(), "overflow rows", int((lst[:, 0] < 0).sum()))

for n, seed, imf, istart, ni in [(3001, 11, "kroupa", 1, 3001), (3001, 12, "equal", 100, 800)]:
    m, x, v = S.plummer(n, seed, imf)
    pot = ref.gpupot(istart, ni, m, x)
    f = OUT / f"pot_n{n}_s{seed}_{imf}_i{istart}.npz"
    np.savez_compressed(f, n=n, seed=seed, imf=imf, istart=istart, ni=ni, pot=pot)
    print(f.name)
