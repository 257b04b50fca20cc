#!/usr/bin/env python
"""rfbank_analyze.py <binary> [timings.txt]: register triples of the FFMA2s in each rf_kernel<S1,S2,MODE> loop (from cuobjdump -sass)
and, when the timings printed by the binary on a GPU are given, a least-squares fit of the cost per conflict class.
Classes tried: pair-bank p = (reg >> 1) & 1 of the three source pairs (A, B, C); cost(#pairs in the fuller bank)."""
import re, subprocess, sys, collections
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
kern = {}
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    m = re.match(r"_Z9rf_kernelILi(\d+)ELi(\d+)ELi(\d+)E", name)
    if not m: continue
    key = tuple(int(x) for x in m.groups())
    ins = [(int(a, 16), t.strip()) for a, t in re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)]
    # innermost loop: last backward branch
    loop = None
    for a, t in ins:
        mm = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if mm and int(mm.group(1), 16) < a: loop = (int(mm.group(1), 16), a)
    body = [t for a, t in ins if loop and loop[0] <= a <= loop[1]]
    trip = []
    for t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0]
        if op.startswith(("FFMA2", "FADD2", "FFMA")):
            regs = [int(x) for x in re.findall(r"(?<![U\w])R(\d+)", t)]
            trip.append((op, regs[0], regs[1:]))
    kern[key] = trip
times = {}
if len(sys.argv) > 2:
    for ln in open(sys.argv[2]):
        m = re.match(r"rf mode (\d+) s1\s+(\d+) s2\s+(\d+)\s+[\d.]+ ms\s+([\d.]+) cycles", ln)
        if m: times[(int(m.group(2)), int(m.group(3)), int(m.group(1)))] = float(m.group(4))
def classes(srcs):
    ds = sorted(set(srcs))
    pb = collections.Counter((r >> 1) & 1 for r in ds)          # pair-bank: bit 1 of the register number
    qb = collections.Counter((r >> 1) & 3 for r in ds)          # 4 pair-banks: bits 1-2
    return len(ds), max(pb.values()), max(qb.values())
rows = []
for key in sorted(kern, key=lambda k: (k[2], k[0], k[1])):
    trip = kern[key]
    cls = collections.Counter(classes(s) for _, _, s in trip)
    t = times.get(key)
    print(key, "n=%d" % len(trip), dict(cls), ("%.3f cyc" % t) if t else "", "e.g.", trip[0][2] if trip else None)
    rows.append((key, cls, len(trip), t))
if times:
    import numpy as np
    keys = sorted({c for _, cls, _, t in rows if t for c in cls})
    A = np.array([[cls.get(c, 0) / n for c in keys] for _, cls, n, t in rows if t and n])
    y = np.array([t for _, cls, n, t in rows if t and n])
    sol, res, rk, sv = np.linalg.lstsq(A, y, rcond=None)
    print("fit: cycles per instruction by class (distinct pairs, max in one of 2 pair-banks, max in one of 4):")
    for c, v in zip(keys, sol): print("   ", c, "%.2f" % v)
    print("   rms residual %.3f" % float(np.sqrt(np.mean((A @ sol - y) ** 2))))
