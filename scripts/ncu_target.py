"""Target of the `ncu --set full` capture: the launch roofline.frac is about -- ONE gpunb_regf_ call of 1024 i-particles against
N = 1M j, not split into sub-blocks (one regf_kernel launch per call).  Usage under ncu: -k regex:regf_kernel -s 1 -c 1."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nbody6ppgpu_b200 import load, snapshots as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
lib = load(); lib.devinit(0)
m, x, v = S.plummer(n, 1, "kroupa"); h2, dtr = S.radii_nnb(x, m, 200.0)
lib.open(n + 10, 0); lib.send(m, x, v)
lib.set_tuning(0, 1)
for b in range(3):
    s = slice(b * 1024, (b + 1) * 1024)
    lib.regf(h2[s], dtr[s], x[s], v[s], 600, 550, 0)
lib.close()
