#!/usr/bin/env python3
"""Golden vectors for the track-table construction.

Runs the reference's ``Tracker.comp_tracker`` (src/tracker.py:11-96, which calls ``relativeCameraMotion``,
``F_from_KRt2`` and ``verify_duplet`` of src/utils.py) on the toy domains of make_golden.py -- as shipped, and on
copies whose 2D endpoints of the odd views are perturbed by seeded pixel noise so that ``verify_duplet`` rejects a
share of the matches -- and records inputs and outputs:

    endpoints (n_lines2d,2,2) f64   Point.coord of the two endpoints of domain.lines2d[id]
    K, R (V,3,3), t (V,3)            intrinsics and poses the fundamental matrices are built from
    F (V-1,3,3)                      reference F of every consecutive pair
    match_ptr (V,), matches (M,2)    domain.line_match["<v><v+1>"][1], concatenated
    flags (M,) bool                  inlier flags: membership of each match in verify_duplet's returned list
    ids, bin (T,V) f64, n_new_tracks the reference's tables

Needs /root/reference (build container only); the fixtures are committed.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def record(ref, toy_dir, json_name, last_view, noise=0.0, seed=0):
    import utils as ref_utils
    dom = mg.build_toy_domain(ref, toy_dir, json_name, last_view + 1)
    if noise > 0:
        rng = np.random.default_rng(seed)
        for v in dom.views[1::2]:
            for L in v.lines:
                for p in L.points:
                    p.coord = list(np.asarray(p.coord, dtype=np.float64) + rng.normal(0, noise, 2) * rng.integers(0, 2))
    with contextlib.redirect_stdout(io.StringIO()):
        Tr = ref.tracker.Tracker()
        Tr.comp_tracker(dom, last_view=last_view)
    endpoints = np.array([[np.asarray(p.coord, dtype=np.float64) for p in L.points] for L in dom.lines2d])
    assert all(dom.lines2d[i].id == i for i in range(len(dom.lines2d)))   # verify_duplet indexes the list by id
    Fs, ptr, matches, flags = [], [0], [], []
    for v in range(last_view):
        P0, P1 = dom.views[v].P, dom.views[v + 1].P
        rR, rt = ref_utils.relativeCameraMotion(P0, P1)
        F = ref_utils.F_from_KRt2(dom.K, rR, rt)
        m = np.array(dom.line_match[str(v) + str(v + 1)][1], dtype=np.int64)
        with contextlib.redirect_stdout(io.StringIO()):
            _, inl = ref_utils.verify_duplet(dom, v, v + 1, F, 0.1, tri_match=None)
        inl = {tuple(r) for r in np.asarray(inl, dtype=np.int64).reshape(-1, 2)}
        Fs.append(F); matches.append(m); ptr.append(ptr[-1] + len(m))
        flags.append(np.array([tuple(r) in inl for r in m], dtype=bool))
    return dict(endpoints=endpoints, K=np.asarray(dom.K, dtype=np.float64),
                R=np.array([np.asarray(v.P.R, dtype=np.float64) for v in dom.views]),
                t=np.array([np.asarray(v.P.t, dtype=np.float64).ravel() for v in dom.views]),
                F=np.array(Fs), match_ptr=np.asarray(ptr, dtype=np.int64), matches=np.concatenate(matches),
                flags=np.concatenate(flags), ids=np.asarray(Tr.ids), bin=np.asarray(Tr.bin),
                n_new_tracks=np.asarray(Tr.n_new_tracks, dtype=np.int64))


def main():
    ref = mg.import_reference()
    for name, tdir, jname, lv, noise, seed in (
            ("trk_toy1_lv5", "toy_example1", "toy_example1_3Dvision.json", 5, 0.0, 0),
            ("trk_toy1_lv9_noisy", "toy_example1", "toy_example1_3Dvision.json", 9, 40.0, 1),
            ("trk_toy2_lv2_noisy", "toy_example2", "toy_example2_3Dvision.json", 2, 25.0, 2)):
        q = record(ref, tdir, jname, lv, noise, seed)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **q)
        print(name, "views", q["R"].shape[0], "matches", len(q["matches"]), "inliers", int(q["flags"].sum()),
              "tracks", q["ids"].shape, "n_new", q["n_new_tracks"].tolist())


if __name__ == "__main__":
    main()
