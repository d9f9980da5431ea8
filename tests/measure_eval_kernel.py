#!/usr/bin/env python3
"""Time the fused residual + Jacobian kernel alone on BASELINE config [3] (1k cameras / 200k lines / 2M observations):
`lba_residuals_jacobian` with no host outputs, profiling mode on, N launches.  Used for A/B runs of kernel variants
(LBA_LIB=<path to a liblba.so> selects the library).  Not a test; prints one JSON line."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from amuya_sfm_lines_beta_b200 import _lib as _L
if os.environ.get("LBA_LIB"):
    _L.LIB_PATH = os.path.abspath(os.environ["LBA_LIB"])
from amuya_sfm_lines_beta_b200 import LineBA
from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cams = int(os.environ.get("CAMS", 1000))
s = make_synthetic_scene(cams, 200_000, 10, seed=0)
ba = LineBA(0)
ba.set_problem(s.n_cameras, s.n_lines, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1)
cp = torch.from_numpy(s.camera_params).cuda(); pp = torch.from_numpy(s.points_3d).cuda()
torch.cuda.synchronize()
from amuya_sfm_lines_beta_b200._lib import check
import ctypes as C
def call():
    check(ba._lib.lba_residuals_jacobian(ba._h, C.c_void_p(cp.data_ptr()), C.c_void_p(pp.data_ptr()), None, None, None))
for _ in range(3): call()
ba.set_profiling(True)
for _ in range(n): call()
p = ba.get_profile()
print(json.dumps({"lib": os.environ.get("LBA_LIB", "default"), "jacobian_us": round(1e3 * p["jacobian"]["ms"] / p["jacobian"]["launches"], 2), "launches": p["jacobian"]["launches"]}))
