"""The predictor statements of the reference (xbpredall.f:18-26) executed by oracle/f77_interp.py -- the golden vectors
tests/golden/xbpredall_f77.npz (oracle/make_xbpredall_golden.py) -- against the unfused numpy restatement the GPU tests and
the Ahmad-Cohen driver use.  Bit for bit: the device-resident predictor is then held to the same vectors on the GPU."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import make_xbpredall_golden as MX  # noqa: E402

GOLDEN = ROOT / "tests" / "golden" / "xbpredall_f77.npz"


def host_predict(x0, v0, f2, fd6, t0, time):
    """The restatement of tests/test_predictor_gpu.py and hermite_ac.Integrator._predict."""
    s = (time - t0)[:, None]
    return ((fd6 * s + f2) * s + v0) * s + x0, (fd6 * (1.5 * s) + f2) * (2.0 * s) + v0


def test_numpy_restatement_equals_the_interpreted_fortran_bit_for_bit():
    g = np.load(GOLDEN)
    xp, vp = host_predict(g["x0"], g["x0dot"], g["f"], g["fdot"], g["t0"], float(g["time"]))
    assert np.array_equal(xp, g["f77_x"]) and np.array_equal(vp, g["f77_xdot"])
    assert str(g["source"]) == "src/Main/xbpredall.f:18-26"
    assert np.abs(g["f77_x"] - g["x0"]).max() > 1e-3             # the prediction moved the particles


@pytest.mark.skipif(not Path(MX.REFERENCE, MX.SPEC[0]).is_file(), reason="the reference sources are not on this machine")
def test_interpreter_live_on_the_reference_text():
    g = np.load(GOLDEN)
    x, xdot = MX.interpreted_predict(g["x0"][:64], g["x0dot"][:64], g["f"][:64], g["fdot"][:64], g["t0"][:64], float(g["time"]))
    assert np.array_equal(x, g["f77_x"][:64]) and np.array_equal(xdot, g["f77_xdot"][:64])
    rng = np.random.default_rng(99)
    a = [rng.normal(size=(50, 3)) for _ in range(4)]
    t0 = rng.integers(0, 32, size=50) * 2.0 ** -9
    x, xdot = MX.interpreted_predict(*a, t0, 0.125)
    xp, vp = host_predict(*a, t0, 0.125)
    assert np.array_equal(x, xp) and np.array_equal(xdot, vp)
