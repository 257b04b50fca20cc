import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nbody6ppgpu_b200 import load
lib = load(); lib.devinit(0)
names = {0: "FFMA (shared b,c)", 1: "FFMA2 (shared b,c)", 2: "FADD2", 3: "FMUL2", 4: "MUFU.RSQ", 5: "FFMA2+ALU mix",
         6: "FFMA 3 distinct regs", 7: "FFMA2 3 distinct pairs", 8: "FFMA2 1 shared operand", 9: "FADD2 2 distinct pairs",
         10: "FFMA2 square (2 distinct)", 11: "FFMA2 distinct + MUFU/4",
         12: "FADD scalar (op/s)", 13: "FMUL scalar (op/s)", 14: "FADD:FMUL:FFMA 1:1:2 (inst/s x2)", 15: "FFMA:FADD 1:1"}
for mode in range(16):
    best = max(lib.fp32_microbench(mode, 8192) for _ in range(3))
    print(f"mode {mode:2d} {names[mode]:28s} {best:8.2f} T(fl)op/s", flush=True)
import ctypes as C
lib.lib.gpunb_b200_farbody_microbench.argtypes = [C.c_int, C.c_int, C.c_int]
lib.lib.gpunb_b200_farbody_microbench.restype = C.c_double
fn = {0: "IT1 full", 1: "IT1 no MUFU", 2: "IT1 no LDS", 3: "IT1 no MUFU no LDS", 4: "IT2 full", 5: "IT2 no MUFU", 6: "IT2 no LDS"}
for mode in range(7):
    for ctas in (2, 3, 4):
        print(f"farbody {fn[mode]:20s} {ctas} CTAs/SM x4 warps: {lib.lib.gpunb_b200_farbody_microbench(mode, 2000, ctas):8.1f} Gint/s", flush=True)
