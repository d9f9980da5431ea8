#!/usr/bin/env python3
"""Measurement for the N3 row: batched two-view endpoint triangulation on the GPU against the reference's
per-line cv2.triangulatePoints loop (src/seq_utils.py:113-121) on a bounded sample.  Prints one JSON line.
Lives under tests/ because it times the oracle as its CPU arm (only test infrastructure may execute oracle/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from amuya_sfm_lines_beta_b200 import LineBA
    from test_triangulation import synthetic_tracks
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    P, view0, p0, p1 = synthetic_tracks(n, n_views=8, seed=5)
    ba = LineBA(0)
    dP, dv, d0, d1 = (torch.from_numpy(a).cuda() for a in (P, view0, p0, p1))
    out = torch.empty(n * 6, dtype=torch.float32, device="cuda")
    times = []
    for _ in range(6):
        ba.triangulate_lines(dP, dv, d0, d1, out=out)
        times.append(ba.last_triangulate_seconds)
    t_kernel = float(np.median(times[2:]))
    t0 = time.perf_counter()
    X = ba.triangulate_lines(P, view0, p0, p1)                  # host arrays in, host array out
    t_e2e = time.perf_counter() - t0
    # CPU arm: the reference's loop body on a bounded sample
    cpu = None
    try:
        import cv2
        k = 20000
        t0 = time.perf_counter()
        for t in range(k):
            Xc = cv2.triangulatePoints(P[view0[t]], P[view0[t] + 1], p0[t].T.copy(), p1[t].T.copy())
            Xc = Xc[:3] / Xc[3]
        cpu = k / (time.perf_counter() - t0)
    except ImportError:
        pass
    fp64_tf = ba.measure_fp64_peak()
    ba.close()
    bytes_per_track = 2 * (16 + 12) + 8                         # two endpoints (2 x float2 in, 3 floats out) + view index
    print(json.dumps({"workload": f"{n} line tracks = {2 * n} endpoints, 8 views", "kernel_ms": t_kernel * 1e3,
                      "tracks_per_s_kernel": n / t_kernel, "tracks_per_s_e2e_host_arrays": n / t_e2e,
                      "hbm_GBps_algorithmic": bytes_per_track * n / t_kernel / 1e9,
                      "cpu_reference_loop_tracks_per_s": cpu, "fp64_peak_tflops": fp64_tf,
                      "nan_or_inf": int(np.count_nonzero(~np.isfinite(X)))}))


if __name__ == "__main__":
    main()
