"""Track-table construction: the oracle against the reference's own Tracker outputs, the
data-parallel formulation against the sequential one (including the reference's quirks), the device arithmetic of
the epipolar test checked on the host, and (``-m gpu``) the CUDA path through the C ABI.

Parity bar: everything bit-exact with the flags computed HERE (nothing of the reference fed in): inlier flags, ids, bin
and n_new_tracks equal the reference's own Tracker on all three toy domains -- including the 88 matches whose
segment is parallel to an epipolar line within 1e-9, where the reference divides by rounding noise and its flag is
set by the rounding of every operation (the epipolar arithmetic reproduces those roundings, see csrc/lba_tracks.cuh)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import tracker_oracle as K

GOLD = os.path.join(ROOT, "tests", "golden")
FIXTURES = ["trk_toy1_lv5", "trk_toy1_lv9_noisy", "trk_toy2_lv2_noisy"]


def load(name):
    q = np.load(os.path.join(GOLD, name + ".npz"))
    q = {k: q[k] for k in q.files}
    V = q["R"].shape[0]
    q["pairs"] = [q["matches"][q["match_ptr"][v]:q["match_ptr"][v + 1]] for v in range(V - 1)]
    return q


def degenerate(q, tol=1e-9):
    """Matches whose segment is parallel to one of its epipolar lines to within `tol` (|sin| of the angle)."""
    out = np.zeros(len(q["matches"]), dtype=bool)
    E = q["endpoints"]
    for v, m in enumerate(q["pairs"]):
        F = q["F"][v]
        for j, (l0, l1) in enumerate(m):
            b0, b1 = E[l1]
            A, B = b0[1] - b1[1], b1[0] - b0[0]
            for a in E[l0]:
                l = F @ np.array([a[0], a[1], 1.0])
                if abs(A * l[1] - B * l[0]) <= tol * np.hypot(A, B) * np.hypot(l[0], l[1]):
                    out[q["match_ptr"][v] + j] = True
    return out


def random_case(seed, V=5, n_per_view=40, k=60):
    """Random match lists with repeated first / second lines, line id 0 beyond view 0 and flags drawn at random:
    exercises the membership-with-zeros and last-writer behaviour of tracker.py:60-70."""
    rng = np.random.default_rng(seed)
    n_lines = V * n_per_view
    pairs, flags = [], []
    for v in range(V - 1):
        lo0, lo1 = v * n_per_view, (v + 1) * n_per_view
        m = np.stack([rng.integers(lo0, lo0 + n_per_view, k), rng.integers(lo1, lo1 + n_per_view, k)], axis=1)
        if v > 0:
            m[rng.integers(0, k, 3), 0] = 0                      # the id-0 quirk
        pairs.append(m.astype(np.int64))
        flags.append(rng.random(k) < 0.7)
    return n_lines, pairs, flags


def tables_from_flags(pairs, flags, V, formulation):
    """Run an oracle formulation with the epipolar test replaced by given flags."""
    it = iter(flags)
    orig = K.verify_duplets
    K.verify_duplets = lambda F, E, m, thr=0.1: next(it)
    try:
        E = np.zeros((max((int(p.max()) for p in pairs if len(p)), default=0) + 1, 2, 2))
        return formulation(E, pairs, [np.eye(3)] * (V - 1))[:3]
    finally:
        K.verify_duplets = orig


# ---- oracle vs the reference ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_fundamental_matrices_bit_equal_reference(name):
    q = load(name)
    for v in range(q["R"].shape[0] - 1):
        F = K.fundamental_from_cameras(q["K"], q["R"][v], q["t"][v], q["R"][v + 1], q["t"][v + 1])
        assert np.array_equal(F, q["F"][v])


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_flags_and_tables_match_reference(name):
    q = load(name)
    ids, bi, n_new, fl = K.comp_tracker(q["endpoints"], q["pairs"], q["F"])
    fl = np.concatenate(fl)
    assert np.array_equal(fl, q["flags"])                        # every match, the near-parallel ones included
    if name == "trk_toy1_lv5":
        assert degenerate(q).sum() >= 40 and q["flags"][degenerate(q)].sum() >= 1   # the fixture does exercise them
    assert np.array_equal(ids, q["ids"].astype(np.int64)) and np.array_equal(bi, q["bin"].astype(np.int64))
    assert np.array_equal(n_new, q["n_new_tracks"])
    # the opt-in tolerance drops the matches that are parallel within rounding
    assert not any(K.duplet_inlier(q["F"][v], *q["endpoints"][l0], *q["endpoints"][l1], 0.1, parallel_tol=1e-12)
                   for v, m in enumerate(q["pairs"]) for j, (l0, l1) in enumerate(m)
                   if degenerate(q, 1e-13)[q["match_ptr"][v] + j])
    # with the reference's own flags the tables are the reference's, exactly
    pf = [q["flags"][q["match_ptr"][v]:q["match_ptr"][v + 1]] for v in range(len(q["pairs"]))]
    ids2, bi2, n2 = tables_from_flags(q["pairs"], pf, q["R"].shape[0], K.comp_tracker)
    assert np.array_equal(ids2, q["ids"].astype(np.int64)) and np.array_equal(bi2, q["bin"].astype(np.int64))
    assert np.array_equal(n2, q["n_new_tracks"])
    assert ids2.dtype == np.int64 and np.array_equal(q["ids"], q["ids"].astype(np.int64))   # reference floats are integers


@pytest.mark.parametrize("seed", range(8))
def test_parallel_formulation_equals_sequential_on_quirky_input(seed):
    n_lines, pairs, flags = random_case(seed)
    a = tables_from_flags(pairs, flags, len(pairs) + 1, K.comp_tracker)
    b = tables_from_flags(pairs, flags, len(pairs) + 1, K.comp_tracker_tables)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert a[0].shape[0] == int(a[2].sum())


def test_formulations_agree_on_empty_pairs():
    n_lines, pairs, flags = random_case(0)
    pairs2 = [pairs[0], pairs[1][:0], pairs[2]]
    fl2 = [flags[0], flags[1][:0], flags[2]]
    a = tables_from_flags(pairs2, fl2, 4, K.comp_tracker)
    b = tables_from_flags(pairs2, fl2, 4, K.comp_tracker_tables)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert a[2][1] == 0
    e = tables_from_flags([pairs[0][:0]], [flags[0][:0]], 2, K.comp_tracker_tables)
    assert e[0].shape == (0, 2) and e[2].tolist() == [0]


@pytest.mark.parametrize("name", FIXTURES)
def test_parallel_formulation_equals_sequential_on_fixtures(name):
    q = load(name)
    a = K.comp_tracker(q["endpoints"], q["pairs"], q["F"])
    b = K.comp_tracker_tables(q["endpoints"], q["pairs"], q["F"])
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)


def test_observation_list_matches_reference_input_builder():
    """The observation order derived from the tables is the one the BA fixtures record (seq_utils.py:186-195)."""
    q = load("trk_toy1_lv5")
    g = np.load(os.path.join(GOLD, "toy1_lv5.npz"))
    pf = [q["flags"][q["match_ptr"][v]:q["match_ptr"][v + 1]] for v in range(len(q["pairs"]))]
    ids, bi, _ = tables_from_flags(q["pairs"], pf, q["R"].shape[0], K.comp_tracker)
    cam, line, _ = K.observations_from_tables(ids, bi)
    assert np.array_equal(cam, g["camera_indices"]) and np.array_equal(line, g["lines_indices"])


# ---- the device arithmetic of the epipolar test, compiled for the host ----------------------------------------------
HOST_SRC = r"""
#include "lba_tracks.cuh"
extern "C" void duplets_host(long k, const long* m, const double* E, const double* F, double thr, unsigned char* out) {
  for (long j = 0; j < k; ++j) {
    const double* a = E + 4 * m[2 * j]; const double* b = E + 4 * m[2 * j + 1];
    out[j] = lba::duplet_inlier(F, a, a + 2, b, b + 2, thr) ? 1 : 0;
  }
}
"""


@pytest.fixture(scope="module")
def host_duplets(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    d = tmp_path_factory.mktemp("trk_host")
    (d / "host.cpp").write_text(HOST_SRC)
    so = d / "libtrk_host.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "amuya_sfm_lines_beta_b200", "csrc"),
                    "-o", str(so), str(d / "host.cpp")], check=True)
    lib = C.CDLL(str(so))

    def run(F, E, m, thr=0.1):
        m = np.ascontiguousarray(m, dtype=np.int64); E = np.ascontiguousarray(E, dtype=np.float64)
        F = np.ascontiguousarray(F, dtype=np.float64); out = np.zeros(len(m), dtype=np.uint8)
        vp = C.c_void_p
        lib.duplets_host(C.c_long(len(m)), vp(m.ctypes.data), vp(E.ctypes.data), vp(F.ctypes.data), C.c_double(thr),
                         vp(out.ctypes.data))
        return out.astype(bool)
    return run


@pytest.mark.parametrize("name", FIXTURES)
def test_device_epipolar_test_on_host_matches_reference(host_duplets, name):
    q = load(name)
    fl = np.concatenate([host_duplets(q["F"][v], q["endpoints"], m) for v, m in enumerate(q["pairs"])])
    assert np.array_equal(fl, q["flags"])                        # the reference's flag for every match
    assert np.array_equal(fl, np.concatenate(K.comp_tracker(q["endpoints"], q["pairs"], q["F"])[3]))


# ---- glue -------------------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources not present")
def test_glue_arrays_equal_reference_inputs():
    import sys
    sys.path.insert(0, GOLD)
    import make_golden as mg
    from amuya_sfm_lines_beta_b200.glue import tracker_arrays_from_domain
    ref = mg.import_reference()
    dom = mg.build_toy_domain(ref, "toy_example1", "toy_example1_3Dvision.json", 6)
    a = tracker_arrays_from_domain(dom, 5)
    q = load("trk_toy1_lv5")
    assert np.array_equal(a["endpoints"], q["endpoints"]) and np.array_equal(a["F"], q["F"])
    assert all(np.array_equal(x, y) for x, y in zip(a["matches"], q["pairs"]))


# ---- CUDA path ----------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu_ba():
    from amuya_sfm_lines_beta_b200 import LineBA
    ba = LineBA(0)
    yield ba
    ba.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_gpu_track_table_matches_oracle_and_reference(gpu_ba, name):
    q = load(name)
    ids, bi, n_new, fl = gpu_ba.track_table(q["endpoints"], q["pairs"], q["F"], want_flags=True)
    o_ids, o_bin, o_new, o_fl = K.comp_tracker(q["endpoints"], q["pairs"], q["F"])
    assert np.array_equal(fl, np.concatenate(o_fl))
    assert np.array_equal(ids, o_ids) and np.array_equal(bi, o_bin) and np.array_equal(n_new, o_new)
    # own flags, own tables == the reference's Tracker, every match and every entry
    assert np.array_equal(fl, q["flags"])
    assert np.array_equal(ids, q["ids"].astype(np.int64)) and np.array_equal(bi, q["bin"].astype(np.int64))
    assert np.array_equal(n_new, q["n_new_tracks"])
    # (and with the reference's flags fed in, too)
    ids2, bi2, n2 = gpu_ba.track_table(q["endpoints"], q["pairs"], q["F"], flags_in=q["flags"])
    assert np.array_equal(ids2, q["ids"].astype(np.int64)) and np.array_equal(bi2, q["bin"].astype(np.int64))
    assert np.array_equal(n2, q["n_new_tracks"])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_track_table_quirks_and_edges(gpu_ba, seed):
    n_lines, pairs, flags = random_case(seed, V=6, n_per_view=300, k=500)
    want = tables_from_flags(pairs, flags, len(pairs) + 1, K.comp_tracker)
    E = np.zeros((n_lines, 2, 2))
    got = gpu_ba.track_table(E, pairs, [np.eye(3)] * len(pairs), flags_in=np.concatenate(flags))
    for x, y in zip(want, got):
        assert np.array_equal(x, y)
    if seed == 0:
        # a pair without matches, and no matches at all
        pairs2 = [pairs[0], pairs[1][:0], pairs[2]]
        fl2 = [flags[0], flags[1][:0], flags[2]]
        want2 = tables_from_flags(pairs2, fl2, 4, K.comp_tracker)
        got2 = gpu_ba.track_table(E, pairs2, [np.eye(3)] * 3, flags_in=np.concatenate(fl2))
        for x, y in zip(want2, got2):
            assert np.array_equal(x, y)
        ids, bi, n_new = gpu_ba.track_table(E, [pairs[0][:0]], [np.eye(3)])
        assert ids.shape == (0, 2) and n_new.tolist() == [0]
        from amuya_sfm_lines_beta_b200._lib import LbaError
        bad = pairs[0].copy(); bad[5, 1] = n_lines
        with pytest.raises(LbaError):
            gpu_ba.track_table(E, [bad], [np.eye(3)])
