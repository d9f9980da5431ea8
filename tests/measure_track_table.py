#!/usr/bin/env python3
"""Measurement for the N2 row: track-table construction on the GPU (epipolar verification of every match + table
merge, view pair by view pair) against the sequential oracle restatement of the reference on a bounded sample.
Prints one JSON line.
Lives under tests/ because it times the oracle as its CPU arm (only test infrastructure may execute oracle/)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def scene(V, n_per_view, seed=0):
    """Consecutive views of one random 3D segment cloud: line j of view v has global id v * n + j and is matched to
    line j of view v+1 (a share of the matches is corrupted so that the epipolar test rejects them)."""
    from oracle import tracker_oracle as K
    rng = np.random.default_rng(seed)
    Kc = np.array([[1500.0, 0, 1024], [0, 1500.0, 1024], [0, 0, 1]])
    X = rng.uniform(-2, 2, (n_per_view, 2, 3))
    R, t, E = [], [], []
    for v in range(V):
        ang = 0.15 * v
        c = np.array([8 * np.sin(ang), 0.3 * v, -8 * np.cos(ang)])
        z = -c / np.linalg.norm(c); x = np.cross([0, 1.0, 0], z); x /= np.linalg.norm(x); y = np.cross(z, x)
        Rv = np.stack([x, y, z]); tv = -Rv @ c
        p = np.einsum("ij,nej->nei", Kc @ Rv, X) + Kc @ tv
        E.append(p[..., :2] / p[..., 2:3] + rng.normal(0, 0.5, (n_per_view, 2, 2)))
        R.append(Rv); t.append(tv)
    E = np.concatenate(E)
    pairs, Fs = [], []
    for v in range(V - 1):
        j = np.arange(n_per_view)
        keep = rng.random(n_per_view) < 0.9
        j1 = np.where(rng.random(n_per_view) < 0.1, rng.integers(0, n_per_view, n_per_view), j)   # 10 % wrong partners
        pairs.append(np.stack([v * n_per_view + j, (v + 1) * n_per_view + j1], axis=1)[keep].astype(np.int64))
        Fs.append(K.fundamental_from_cameras(Kc, R[v], t[v], R[v + 1], t[v + 1]))
    return E, pairs, np.array(Fs)


def main():
    from amuya_sfm_lines_beta_b200 import LineBA
    from oracle import tracker_oracle as K
    V = 8
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
    E, pairs, Fs = scene(V, n)
    M = sum(len(p) for p in pairs)
    ba = LineBA(0)
    ba.track_table(E, pairs, Fs)                                  # warm
    t0 = time.perf_counter()
    ids, bi, n_new, fl = ba.track_table(E, pairs, Fs, want_flags=True)
    t_gpu = time.perf_counter() - t0
    # device-resident variant (row N1): endpoints uploaded once, tables left on the device
    import torch
    zeros = np.zeros((E.shape[0], 3))
    ba.set_lines2d(zeros, zeros, zeros, E)
    ba.track_table_device(pairs, Fs)                              # warm
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_ids, d_bin, d_new = ba.track_table_device(pairs, Fs)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    assert np.array_equal(d_ids.cpu().numpy(), ids) and np.array_equal(d_new, n_new)
    ba.close()
    # CPU arm on a bounded sample of the same construction
    Es, ps, Fss = scene(V, 2000, seed=1)
    Ms = sum(len(p) for p in ps)
    t0 = time.perf_counter()
    o = K.comp_tracker(Es, ps, Fss)
    t_cpu = time.perf_counter() - t0
    print(json.dumps({"workload": f"{V} views x {n} 2D lines, {M} matches", "gpu_call_ms_host_arrays": t_gpu * 1e3,
                      "matches_per_s_gpu_call": M / t_gpu,
                      "gpu_call_ms_resident_endpoints_device_tables": t_dev * 1e3, "matches_per_s_device_resident": M / t_dev, "tracks": int(ids.shape[0]), "inlier_share": float(fl.mean()),
                      "n_new_tracks": n_new.tolist(), "cpu_oracle_matches_per_s": Ms / t_cpu,
                      "cpu_sample": f"{Ms} matches, sequential numpy restatement of tracker.py + verify_duplet, 1 core"}))


if __name__ == "__main__":
    main()
