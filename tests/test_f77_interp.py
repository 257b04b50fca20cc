"""oracle/f77_interp.py -- the small fixed-form Fortran 77 interpreter that executes the reference's own list-bookkeeping text
(oracle/regcor_fortran.py): statement semantics on hand-written snippets whose results are known from the standard."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
import f77_interp as F  # noqa: E402


def run(tmp_path, src, scalars=None, arrays=None):
    p = tmp_path / "snippet.f"
    p.write_text(src)
    st = F.read_statements(str(p), 1, src.count("\n") + 1)
    return F.Machine(st, scalars or {}, arrays or {}).run()


def test_do_loops_trip_counts_and_final_values(tmp_path):
    env = run(tmp_path, """
      K = 0
      DO 10 L = 1,5
         K = K + L
   10 CONTINUE
      M = 0
      DO L2 = 3,2
         M = M + 1
      END DO
      N3 = 0
      DO J = 1,3
         DO 20 I = 1,J
            N3 = N3 + 1
   20    CONTINUE
      END DO
""")
    assert env["K"] == 15 and env["L"] == 6            # the DO variable ends one step past the last trip
    assert env["M"] == 0 and env["L2"] == 3            # zero-trip loop: variable initialised, body skipped
    assert env["N3"] == 6 and env["J"] == 4


def test_block_if_chain_and_goto_out_of_a_block(tmp_path):
    src = """
      K = 0
    5 K = K + 1
      IF (K.LT.3) THEN
         M = 1
         GO TO 5
      ELSE IF (K.EQ.3) THEN
         M = 2
         GO TO 5
      ELSE
         M = 3
      END IF
      IF (M.EQ.3.AND.K.GE.4) J = 7
"""
    env = run(tmp_path, src)
    assert (env["K"], env["M"], env["J"]) == (4, 3, 7)


def test_arithmetic_follows_fortran(tmp_path):
    a = np.zeros(4)
    env = run(tmp_path, """
      I1 = 7/2
      I2 = -7/2
      I3 = 2**3**2
      X1 = -2.0**2
      X2 = 1.0/3.0D0
      X3 = (A0 - B0)**2
      Y(2) = A0*B0 + A0*B0*C0
     &       - C0
      Y(3:4) = 0.5D0
      IF (1.LT.2.AND..NOT.(I1.EQ.4)) L9 = 1
""", {"A0": 0.1, "B0": 0.7, "C0": 3.0}, {"Y": F.farray_numpy(a)})
    assert (env["I1"], env["I2"], env["I3"]) == (3, -3, 512)          # truncation toward zero, ** right associative
    assert env["X1"] == -4.0 and env["X2"] == 1.0 / 3.0
    assert env["X3"] == (0.1 - 0.7) * (0.1 - 0.7)
    assert a[1] == 0.1 * 0.7 + 0.1 * 0.7 * 3.0 - 3.0 and a[2] == 0.5 and a[3] == 0.5       # left to right, continuation line
    assert env["L9"] == 1


def test_comments_directives_and_unsupported_statements(tmp_path):
    env = run(tmp_path, """
* a comment
C$$$      K = 99
!$omp critical
      K = 1
!$omp end critical
c     another
""")
    assert env["K"] == 1
    with pytest.raises(F.F77Error):
        run(tmp_path, "      CALL CHECKL(I,NNB)\n")
    with pytest.raises(F.F77Error):
        run(tmp_path, "      X = 0.1\n")                 # a REAL*4 literal whose promotion is not the decimal value
    with pytest.raises(F.F77Error):
        run(tmp_path, "      K = J + 1\n")               # J undefined
