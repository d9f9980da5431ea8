"""CPU, build container only: the array-native glue (row N1) against the reference's own input_bundle_adj /
update_structure_motion on the toy fixtures.  Skipped where /root/reference is absent (the GPU box)."""
import contextlib
import copy
import io
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference sources not present")


@pytest.fixture(scope="module")
def ref_env():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as mg
    ref = mg.import_reference()
    return mg, ref


@pytest.mark.parametrize("toy,json_name,last_view", [("toy_example1", "toy_example1_3Dvision.json", 2),
                                                     ("toy_example1", "toy_example1_3Dvision.json", 5),
                                                     ("toy_example2", "toy_example2_3Dvision.json", 2)])
def test_input_arrays_equal_reference_input_bundle_adj(ref_env, toy, json_name, last_view):
    from amuya_sfm_lines_beta_b200.glue import input_arrays_from_domain
    mg, ref = ref_env
    dom = mg.build_toy_domain(ref, toy, json_name, last_view + 1)
    inp = mg.ba_input_from_domain(ref, dom, last_view)          # reference: tracker -> triangulate -> input_bundle_adj
    flat = mg.flatten_input(inp)
    a = input_arrays_from_domain(dom, last_view)
    assert np.array_equal(a["camera_params"], flat["camera_params"])
    assert np.array_equal(a["camera_indices"], flat["camera_indices"]) and a["camera_indices"].dtype == np.int64
    assert np.array_equal(a["lines_indices"], flat["lines_indices"]) and a["lines_indices"].dtype == np.int64
    assert np.array_equal(a["points_3d"], flat["points_3d_raw"]) and a["points_3d"].dtype == flat["points_3d_raw"].dtype
    for k in ("n", "p0", "p1"):
        assert np.array_equal(a[k], flat[k])


def test_update_domain_equals_reference_update_structure_motion(ref_env):
    from amuya_sfm_lines_beta_b200.glue import input_arrays_from_domain, update_domain_from_arrays
    mg, ref = ref_env
    dom_a = mg.build_toy_domain(ref, "toy_example1", "toy_example1_3Dvision.json", 3)
    mg.ba_input_from_domain(ref, dom_a, 2)
    dom_b = copy.deepcopy(dom_a)
    a = input_arrays_from_domain(dom_a, 2)
    rng = np.random.default_rng(0)
    cams = a["camera_params"] + rng.normal(0, 1e-3, a["camera_params"].shape)
    pts = np.asarray(a["points_3d"], dtype=np.float64) + rng.normal(0, 1e-3, a["points_3d"].shape)
    with contextlib.redirect_stdout(io.StringIO()):
        ref.seq.update_structure_motion(dom_a, cams, pts, last_view=2)
    update_domain_from_arrays(dom_b, cams, pts, last_view=2)
    for va, vb in zip(dom_a.views, dom_b.views):
        assert np.array_equal(va.P.P, vb.P.P) and np.array_equal(va.P.R, vb.P.R) and np.array_equal(va.P.c, vb.P.c)
    for La, Lb in zip(dom_a.lines3d, dom_b.lines3d):
        assert np.array_equal(La.points[0].coord, Lb.points[0].coord) and np.array_equal(La.points[1].coord, Lb.points[1].coord)


def test_lines2d_tables_gathered_by_id_equal_reference_observation_constants(ref_env):
    """The once-per-domain flattening used by the device-resident incremental step (IncrementalBA.step): rows of the
    per-2D-line tables, gathered by the 2D line id of every observation, are the reference's per-observation arrays."""
    from amuya_sfm_lines_beta_b200.glue import input_arrays_from_domain, lines2d_arrays_from_domain
    mg, ref = ref_env
    dom = mg.build_toy_domain(ref, "toy_example1", "toy_example1_3Dvision.json", 6)
    inp = mg.ba_input_from_domain(ref, dom, 5)
    flat = mg.flatten_input(inp)
    n, p0, p1, E = lines2d_arrays_from_domain(dom)
    head = lines2d_arrays_from_domain(dom, 0)[0][:10]
    tail = lines2d_arrays_from_domain(dom, 10)[0]
    assert np.array_equal(np.concatenate((head, tail)), n)          # a grown tail can be flattened on its own
    a = input_arrays_from_domain(dom, 5)
    ids = a["line_ids"]
    assert np.array_equal(n[ids], flat["n"]) and np.array_equal(p0[ids], flat["p0"]) and np.array_equal(p1[ids], flat["p1"])
    assert E.shape == (len(dom.lines2d), 2, 2)
    assert np.array_equal(E[3], np.array([np.asarray(p.coord, dtype=np.float64)[:2] for p in dom.lines2d[3].points]))
