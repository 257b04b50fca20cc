#!/usr/bin/env python
"""Generate tests/golden/xbpredall_f77.npz: the reference's OWN predictor statements (src/Main/xbpredall.f lines 18-26, the
body of the all-particle prediction that precedes every regular block, intgrt.F:516-520) executed by oracle/f77_interp.py
for every particle of a seeded state.  Test infrastructure: the pin of the device-resident predictor and
of its numpy restatement in the tests.  The source is read where it lies (/root/reference); nothing is copied.

Run in the build container:   python oracle/make_xbpredall_golden.py
"""
import hashlib
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
from f77_interp import Machine, farray_numpy, read_statements  # noqa: E402

REFERENCE = os.environ.get("NBODY6_REFERENCE", "/root/reference")
SPEC = ("src/Main/xbpredall.f", 18, 26)
OUT = ROOT / "tests" / "golden" / "xbpredall_f77.npz"


def statements():
    return read_statements(os.path.join(REFERENCE, SPEC[0]), SPEC[1], SPEC[2])


def interpreted_predict(x0, x0dot, f, fdot, t0, time, ifirst=1):
    """X, XDOT of particles IFIRST..IFIRST+n-1 as the reference's statements compute them (F = force/2, FDOT = derivative/6)."""
    n = x0.shape[0]
    x, xdot = np.full((n, 3), np.nan), np.full((n, 3), np.nan)
    arrays = {"X": farray_numpy(x, ifirst), "XDOT": farray_numpy(xdot, ifirst), "X0": farray_numpy(x0, ifirst),
              "X0DOT": farray_numpy(x0dot, ifirst), "F": farray_numpy(f, ifirst), "FDOT": farray_numpy(fdot, ifirst),
              "T0": farray_numpy(t0, ifirst)}
    st = statements()
    for j in range(ifirst, ifirst + n):
        Machine(st, {"J": j, "TIME": float(time)}, arrays).run()
    return x, xdot


def make_state(n=600, seed=5):
    from nbody6ppgpu_b200 import snapshots as S
    rng = np.random.default_rng(seed)
    m, x0, v0 = S.plummer(n, seed, "kroupa")
    f2 = 0.5 * rng.normal(size=(n, 3)) * 3.0
    fd6 = rng.normal(size=(n, 3)) * 7.0
    t0 = rng.integers(0, 64, size=n) * 2.0 ** -10
    x0[:40] += 1000.0                                   # a far-away clump: large |x| against small steps
    return m, x0, v0, f2, fd6, t0, 0.0703125


def main():
    if not os.path.isfile(os.path.join(REFERENCE, SPEC[0])):
        raise SystemExit("needs the reference sources under %s" % REFERENCE)
    m, x0, v0, f2, fd6, t0, time = make_state()
    x, xdot = interpreted_predict(x0, v0, f2, fd6, t0, time)
    fp = hashlib.sha256("\n".join("%s|%s|%d" % s for s in statements()).encode()).hexdigest()
    np.savez_compressed(OUT, m=m, x0=x0, x0dot=v0, f=f2, fdot=fd6, t0=t0, time=np.array(time), f77_x=x, f77_xdot=xdot,
                        source=np.array("%s:%d-%d" % SPEC), source_sha256=np.array(fp))
    print("xbpredall: %d particles -> %s (%d KB)" % (x.shape[0], OUT.name, OUT.stat().st_size // 1024))


if __name__ == "__main__":
    main()
