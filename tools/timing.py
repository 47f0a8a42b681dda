#!/usr/bin/env python
"""Dev tool (GPU box): cycles per section of the cached-engine loop (k_lvg_small), from a build with -DV2S_TIMING.
  python tools/timing.py build            # here: compiles ab/timing.so
  python tools/timing.py run [log2n]      # on the GPU box: one scheduled sweep with ab/timing.so, prints the table
Sections per engine (lead levels 12..28): load, patch (radiative rates -> lead block), pivots, back-substitution,
relax (normalise, floor, under-relax), lines (Tex, tau, escape probability), rest of the loop body; cycles are per
WARP-iteration (two models per warp), averaged over warps; clock64() fences the sections, so overlap between them is lost."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SO = os.path.join(ROOT, "ab", "timing.so")

if sys.argv[1] == "build":
    csrc = os.path.join(ROOT, "radex_emcee_b200", "csrc")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-DV2S_TIMING",
                           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", csrc, "-o", SO,
                           os.path.join(csrc, "radex_b200.cu"), os.path.join(csrc, "moldata.cpp")] + sys.argv[2:])
    print(SO)
    sys.exit(0)

from radex_emcee_b200 import _lib
_lib.LIB_PATH = SO
from conftest import MOLFILE, draw_params
from test_gpu_solve import gpu_solve
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
ctx = _lib.Context(_lib.MolData(MOLFILE), 0)
lib = _lib.load()
P = draw_params(np.random.default_rng(1), 1 << log2n, 2.7315)
gpu_solve(ctx, P[:, 0], P[:, 1], P[:, 2], 2.7315, 2)            # warm-up
lib.rb_debug_timing.argtypes = [C.c_void_p, C.c_int]
lib.rb_debug_timing(None, 1)
gpu_solve(ctx, P[:, 0], P[:, 1], P[:, 2], 2.7315, 2)
out = (C.c_uint64 * 124)()
lib.rb_debug_timing(out, 0)
tm = np.array(out[:64], dtype=np.float64).reshape(8, 8)
tv = np.array(out[64:], dtype=np.float64).reshape(5, 12)
names = ["load", "patch", "pivots", "backsub", "relax", "lines", "rest"]
print("| lead levels | warp-iterations | cycles per warp-iteration | " + " | ".join(names) + " |")
print("|---|---|---|" + "---|" * len(names))
for kp in range(3, 8):
    it = tm[kp, 7]
    if it == 0:
        continue
    tot = tm[kp, :7].sum()
    print("| %d | %.3g | %.0f | " % (4 * kp, it, tot / it) + " | ".join("%.0f (%.0f %%)" % (tm[kp, i] / it, 100 * tm[kp, i] / tot) for i in range(7)) + " |")

vn = ["rates", "detailed balance", "line constants", "patch", "full elimination", "full back-sub", "cached lead solve", "capture", "relax", "lines", "park/resume/top"]
print()
print("v2::solve by launch kind (cycles summed over warps / 1e9; share of the launch):")
print("| launch | total Gcycles | calls | cycles per call | " + " | ".join(vn) + " |")
print("|---|---|---|---|" + "---|" * len(vn))
for sc, nm in ((0, "single"), (1, "A"), (2, "B"), (4, "C")):
    tot = tv[sc, :11].sum()
    if tot == 0:
        continue
    print("| %s | %.2f | %.3g | %.0f | " % (nm, tot / 1e9, tv[sc, 11], tot / max(tv[sc, 11], 1)) + " | ".join("%.1f %%" % (100 * tv[sc, i] / tot) for i in range(11)) + " |")
