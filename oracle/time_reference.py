#!/usr/bin/env python3
"""B1 baseline: time the REFERENCE ITSELF (imported from /root/reference, build container
only) on synthetic scenes, and fit t_fun(m) = alpha*m + beta*m^2 to extrapolate to the headline sizes the
reference cannot reach.  Writes profiles/reference_cpu_timing.json.  TEST/BENCH INFRASTRUCTURE ONLY.

    python oracle/time_reference.py
"""
import contextlib
import io
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from make_golden import import_reference  # noqa: E402
from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene  # noqa: E402


class _P:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def scene_dict(s):
    lines_2d = [_P(n=s.n[i], points=[_P(p=s.p0[i]), _P(p=s.p1[i])]) for i in range(s.n_obs)]
    lines_3d = [_P(points=[_P(coord=s.points_3d[2 * L]), _P(coord=s.points_3d[2 * L + 1])]) for L in range(s.n_lines)]
    return {"camera_params": s.camera_params, "Lines_3d": lines_3d, "camera_indices": s.camera_indices,
            "lines_indices": s.lines_indices, "lines_2d": lines_2d}


def main():
    ref = import_reference()
    out = {"host": {"cpu_count": os.cpu_count()}, "fun": [], "run_bundle_adjustment": []}
    for (nc, nl, k) in ((10, 50, 10), (20, 200, 10), (40, 800, 10), (80, 3200, 10)):
        s = make_synthetic_scene(nc, nl, k, seed=0)
        d = scene_dict(s)
        with contextlib.redirect_stdout(io.StringIO()):
            pts, p2d, cam_pt, pt_idx = ref.ba.point_wise_input(d["Lines_3d"], d["lines_2d"], d["camera_indices"], d["lines_indices"])
        x0 = np.hstack((s.camera_params.ravel(), pts.ravel()))
        args = (nc, pts.shape[0], cam_pt, pt_idx, p2d, d["lines_2d"])
        ref.ba.fun(x0, *args)
        t0 = time.perf_counter(); ref.ba.fun(x0, *args); dt = time.perf_counter() - t0
        out["fun"].append({"m": s.n_obs, "seconds": dt, "obs_per_s": s.n_obs / dt})
        print(f"fun  m={s.n_obs:6d}  {dt:8.3f} s  ({s.n_obs / dt:8.0f} obs/s)", flush=True)
    for (nc, nl, k) in ((10, 50, 10), (20, 200, 10)):
        s = make_synthetic_scene(nc, nl, k, seed=0)
        d = scene_dict(s)
        buf = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(buf):
            ref.ba.run_bundle_adjustment(d, "linear")
        dt = time.perf_counter() - t0
        lines = [l for l in buf.getvalue().splitlines() if l.strip() and l.split()[0].isdigit()]
        nfev = int(lines[-1].split()[1]) if lines else None
        iters = int(lines[-1].split()[0]) if lines else None
        out["run_bundle_adjustment"].append({"m": s.n_obs, "n_cameras": nc, "seconds": dt, "iterations": iters, "nfev": nfev,
                                             "s_per_iteration": dt / max(iters or 1, 1)})
        print(f"run_bundle_adjustment m={s.n_obs}: {dt:.1f} s, {iters} TRF iterations, nfev {nfev}", flush=True)
    m = np.array([r["m"] for r in out["fun"]], float); t = np.array([r["seconds"] for r in out["fun"]])
    A = np.stack([m, m * m], axis=1)
    alpha, beta = np.linalg.lstsq(A, t, rcond=None)[0]
    out["fit"] = {"alpha_s_per_obs": alpha, "beta_s_per_obs2": beta,
                  "fun_seconds_at_2M": alpha * 2e6 + beta * 4e12, "fun_seconds_at_20M": alpha * 2e7 + beta * 4e14}
    print("fit:", out["fit"])
    json.dump(out, open(os.path.join(ROOT, "profiles", "reference_cpu_timing.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
