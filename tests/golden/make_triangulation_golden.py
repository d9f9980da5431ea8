#!/usr/bin/env python3
"""Golden vectors for the two-view endpoint triangulation.

Runs the reference's own ``Tracker.comp_tracker`` -> ``triangulate_model`` (src/seq_utils.py:75-135, one
``cv2.triangulatePoints`` call per line) on the toy domains of ``make_golden.py`` and records, per triangulated
track, exactly what the reference fed to OpenCV and what it stored:

    P        (V,3,4) f64   camera matrices domain.views[i].P.P
    view0    (n,)   int64  first view of the pair (the second is view0 + 1)
    pts0/1   (n,2,2) f32   [track][endpoint][x,y] -- the reference's float32 arrays (:114-117), transposed
    X_ref    (n,2,3) f32   Point3D.coord of the two endpoints (:119-127): X[:3] / X[3] in float32
    ids, bin, n_new_tracks  the tracker tables the loop walks (:108-112)

Needs /root/reference and cv2 (build container only); the fixtures are committed.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def record(ref, toy_dir, json_name, last_view):
    dom = mg.build_toy_domain(ref, toy_dir, json_name, last_view + 1)
    with contextlib.redirect_stdout(io.StringIO()):
        Tr = ref.tracker.Tracker()
        Tr.comp_tracker(dom, last_view=last_view)
        dom.tracker = Tr
        ref.seq.triangulate_model(dom, last_view=last_view)
    P = np.array([np.asarray(v.P.P, dtype=np.float64) for v in dom.views[:last_view + 1]])
    ids = np.asarray(Tr.ids)
    view0, pts0, pts1 = [], [], []
    tr = 0
    for i in range(last_view):                                   # seq_utils.py:105-112
        k = Tr.n_new_tracks[i]
        for l in ids[tr:tr + k, i:i + 2]:
            view0.append(i)
            pts0.append(np.array([p.coord for p in dom.lines2d[int(l[0])].points], dtype=np.float32))
            pts1.append(np.array([p.coord for p in dom.lines2d[int(l[1])].points], dtype=np.float32))
        tr += k
    X = np.array([[np.asarray(L.points[0].coord), np.asarray(L.points[1].coord)] for L in dom.lines3d])
    assert X.dtype == np.float32 and X.shape == (len(view0), 2, 3)
    return dict(P=P, view0=np.asarray(view0, dtype=np.int64), pts0=np.asarray(pts0), pts1=np.asarray(pts1), X_ref=X,
                ids=ids, bin=np.asarray(Tr.bin), n_new_tracks=np.asarray(Tr.n_new_tracks, dtype=np.int64))


def main():
    ref = mg.import_reference()
    for name, tdir, jname, lv in (("tri_toy1_lv5", "toy_example1", "toy_example1_3Dvision.json", 5),
                                  ("tri_toy2_lv2", "toy_example2", "toy_example2_3Dvision.json", 2)):
        q = record(ref, tdir, jname, lv)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **q)
        print(name, "tracks", q["view0"].shape[0], "views", q["P"].shape[0], "pts dtype", q["pts0"].dtype,
              "X range", float(np.abs(q["X_ref"]).max()))


if __name__ == "__main__":
    main()
