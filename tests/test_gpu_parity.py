"""GPU parity tests (B200): the CUDA path, called through the C ABI, against the CPU oracle and the
golden fixtures generated from the reference itself.  Parity ladder of SURVEY.md 8c:

P0 indexing bit-exact | P1 residuals <= 1e-9 rel | P2 Jacobian vs oracle / vs FD of the reference fun
P3 normal equations, Schur solve | P4 LM trajectory vs the oracle | P5 final cost vs the reference's run
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_names

pytestmark = pytest.mark.gpu

EPS = np.finfo(np.float64).eps
RTOL_RESIDUAL = 1e-9       # north_star: per-residual values within 1e-9 relative in fp64
RTOL_FINAL = 1e-6          # north_star: final cost and parameters within 1e-6 relative
RTOL_TRAJ = 1e-9           # per-iteration cost against the oracle running the same algorithm


def residual_tol(r_ref):
    # 1e-9 relative plus the fp64 floor of the formula itself: r = (2a)^2 + b^2 + ... with a, b dot
    # products of unit vectors (|da| ~ eps), so |dr| <~ 8 eps sqrt(r).
    return RTOL_RESIDUAL * np.abs(r_ref) + 16 * EPS * np.sqrt(np.abs(r_ref)) + 64 * EPS * EPS


@pytest.fixture(scope="module")
def orc():
    from oracle import line_ba_oracle
    return line_ba_oracle


@pytest.fixture(scope="module")
def LineBA():
    from amuya_sfm_lines_beta_b200 import LineBA as cls
    return cls


def make_ctx(LineBA, s):
    ba = LineBA(0)
    ba.set_problem(s.n_cameras, s.n_lines, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1)
    return ba


def load(name):
    from amuya_sfm_lines_beta_b200.scenes import scene_from_npz
    return scene_from_npz(os.path.join(GOLDEN, name + ".npz")), np.load(os.path.join(GOLDEN, name + ".npz"))


def synth(*a, **k):
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    return make_synthetic_scene(*a, **k)


# ---------------------------------------------------------------- P0
@pytest.mark.parametrize("name", golden_names())
def test_p0_indexing_bit_exact(LineBA, name):
    s, q = load(name)
    ba = make_ctx(LineBA, s)
    cam_pt, pt_idx = ba.point_wise_indices()
    assert np.array_equal(cam_pt, q["ref_camera_indices_pt"])
    assert np.array_equal(pt_idx, q["ref_pt_indices"])
    ba.close()


# ---------------------------------------------------------------- P1
@pytest.mark.parametrize("name", golden_names())
def test_p1_residuals_vs_reference_fun(LineBA, name):
    s, q = load(name)
    ba = make_ctx(LineBA, s)
    nc = s.n_cameras
    for xk, fk in (("ref_x0", "ref_fun_x0"), ("ref_x1", "ref_fun_x1")):
        x = q[xk]
        r, cost = ba.residuals(x[:6 * nc].reshape(nc, 6), x[6 * nc:].reshape(-1, 3))
        assert np.all(np.abs(r - q[fk]) <= residual_tol(q[fk]))
        assert abs(cost - 0.5 * np.sum(q[fk] ** 2)) <= 1e-12 * cost
    ba.close()


@pytest.mark.parametrize("name", ["toy1_lv2", "toy2_lv2"])
def test_p1_project_vs_reference(LineBA, name):
    s, q = load(name)
    ba = make_ctx(LineBA, s)
    ph = ba.project(s.camera_params, s.points_3d)
    assert np.max(np.abs(ph - q["ref_project_x0"])) < 1e-14
    ba.close()


def test_p1_dropin_fun_and_project(LineBA):
    """The module-level mirrors of the reference helpers."""
    from amuya_sfm_lines_beta_b200 import bundle_adjustment as mod
    from test_cabi_and_host import _objects_from_golden
    s, q = load("toy1_lv2")
    lines_3d, lines_2d = _objects_from_golden(q)
    pts, p2d, cam_pt, pt_idx = mod.point_wise_input(lines_3d, lines_2d, q["camera_indices"], q["lines_indices"])
    r = mod.fun(q["ref_x0"], s.n_cameras, pts.shape[0], cam_pt, pt_idx, p2d, lines_2d)
    assert np.all(np.abs(r - q["ref_fun_x0"]) <= residual_tol(q["ref_fun_x0"]))
    x0 = q["ref_x0"]; nc = s.n_cameras
    ph = mod.project(x0[6 * nc:].reshape(-1, 3)[pt_idx], x0[:6 * nc].reshape(nc, 3, 2)[cam_pt])
    assert np.max(np.abs(ph - q["ref_project_x0"])) < 1e-14


def test_p1_residuals_synthetic_32k(LineBA, orc):
    s = synth(64, 3200, 10, seed=0)
    ba = make_ctx(LineBA, s)
    r, cost = ba.residuals(s.camera_params, s.points_3d)
    ref = orc.residuals(*s.arrays())
    assert np.all(np.abs(r - ref) <= residual_tol(ref))
    assert abs(cost - 0.5 * ref @ ref) <= 1e-12 * cost
    ba.close()


# ---------------------------------------------------------------- P2
def _oracle_jac(orc, s, cams=None, pts=None):
    cams = s.camera_params if cams is None else cams
    pts = s.points_3d if pts is None else pts
    om, ce = orc.split_camera_params(cams)
    return orc.residuals_jacobian(om, ce, pts.reshape(-1, 6), s.camera_indices, s.lines_indices, s.n,
                                  np.cross(s.n, s.p0), np.cross(s.n, s.p1))


def _interleave(jc):
    out = np.empty_like(jc)
    out[:, 0::2] = jc[:, 0:3]
    out[:, 1::2] = jc[:, 3:6]
    return out


@pytest.mark.parametrize("name", ["toy1_lv2", "toy1_lv5", "toy2_lv2"])
def test_p2_jacobian_vs_reference_fd_and_oracle(LineBA, orc, name):
    s, q = load(name)
    nc = s.n_cameras
    x1 = q["ref_x1"]
    cams, pts = x1[:6 * nc].reshape(nc, 6), x1[6 * nc:].reshape(-1, 3)
    ba = make_ctx(LineBA, s)
    r, jcam, jline = ba.residuals_jacobian(cams, pts)
    r_o, jc_o, jl_o = _oracle_jac(orc, s, cams, pts)
    scale = max(np.abs(jc_o).max(), np.abs(jl_o).max())
    assert np.max(np.abs(jcam - _interleave(jc_o))) < 1e-12 * scale
    assert np.max(np.abs(jline - jl_o)) < 1e-12 * scale
    # against central finite differences of the REFERENCE fun
    J = np.zeros((s.n_obs, 6 * nc + 6 * s.n_lines))
    rows = np.arange(s.n_obs)
    for k in range(6):
        J[rows, 6 * s.camera_indices + k] = jcam[:, k]
        J[rows, 6 * nc + 6 * s.lines_indices + k] = jline[:, k]
    Ja = J[np.ix_(q["ref_fd_rows"], q["ref_fd_cols"])]
    assert np.max(np.abs(Ja - q["ref_fd_jac_x1"])) < 1e-7 * max(1.0, np.abs(Ja).max())
    ba.close()


# ---------------------------------------------------------------- P3
def _unpack21(P):
    M = np.empty((P.shape[0], 6, 6))
    k = 0
    for a in range(6):
        for b in range(a, 6):
            M[:, a, b] = P[:, k]; M[:, b, a] = P[:, k]; k += 1
    return M


@pytest.mark.parametrize("scene", ["toy1_lv9", "synth"])
def test_p3_normal_equations_and_schur_solve(LineBA, orc, scene):
    s = load(scene)[0] if scene != "synth" else synth(40, 1500, 8, seed=2)
    ba = make_ctx(LineBA, s)
    r, jcam, jline = ba.residuals_jacobian(s.camera_params, s.points_3d)
    U, gc, V, gl = ba.normal_equations()
    r_o, jc_o, jl_o = _oracle_jac(orc, s)
    Uo, gco, Vo, glo = orc.normal_equations(r_o, jc_o, jl_o, s.camera_indices, s.lines_indices, s.n_cameras, s.n_lines)
    for got, ref in ((_unpack21(U), Uo), (gc, gco), (_unpack21(V), Vo), (gl, glo)):
        assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))
    idx = np.arange(6)
    Dc = orc.update_scaling(None, Uo[:, idx, idx]); Dl = orc.update_scaling(None, Vo[:, idx, idx])
    for lam in (1e-2, 1e-4):
        sysm = orc.SchurSystem(r_o, jc_o, jl_o, s.camera_indices, s.lines_indices, Uo, gco, Vo, glo, Dc, Dl, lam)
        dco, qo = sysm.pcg(1e-10, 500)
        dlo, _ = sysm.back_substitute(dco)
        dc, dl, qd = ba.schur_solve(lam, pcg_tol=1e-10, pcg_max_iter=500)
        assert np.max(np.abs(dc - dco)) <= 1e-7 * np.max(np.abs(dco))
        assert np.max(np.abs(dl - dlo)) <= 1e-7 * np.max(np.abs(dlo))
        assert abs(qd - qo) <= max(2, 0.1 * qo)   # at eta=1e-10 the tail of CG is rounding-driven
    ba.close()


# ---------------------------------------------------------------- P4
def _traj(res):
    return np.array([t["cost_new"] for t in res.trace]), [t["accepted"] for t in res.trace], [t["pcg"] for t in res.trace]


def _conditioning_floor(orc, s, kw, ref_costs, ref_out):
    """How far the ORACLE's own trajectory moves when its inputs are perturbed by one ulp: the
    rounding-sensitivity of the problem itself (no camera is held fixed, so 7 gauge directions are
    only damped).  Device-vs-oracle differences below a small multiple of this floor
    are not attributable to the implementation."""
    rng = np.random.default_rng(12345)
    worst = np.zeros_like(ref_costs)
    fin = dict(cost=0.0, cams=0.0, pts=0.0)
    for _ in range(2):
        n2 = s.n * (1.0 + rng.uniform(-1, 1, s.n.shape) * EPS)
        c2, p2, r2, rep2 = orc.lm_solve(s.camera_params, s.points_3d, s.camera_indices, s.lines_indices, n2, s.p0, s.p1,
                                        opts=orc.LMOptions(**kw))
        c = np.array([t["cost_new"] for t in rep2.trace])
        k = min(len(c), len(ref_costs))
        worst[:k] = np.maximum(worst[:k], np.abs(c[:k] - ref_costs[:k]) / np.abs(ref_costs[:k]))
        fin["cost"] = max(fin["cost"], abs(rep2.cost - ref_out[3].cost) / ref_out[3].cost)
        fin["cams"] = max(fin["cams"], np.max(np.abs(c2 - ref_out[0])) / np.max(np.abs(ref_out[0])))
        fin["pts"] = max(fin["pts"], np.max(np.abs(p2 - ref_out[1])) / np.max(np.abs(ref_out[1])))
    return np.maximum.accumulate(worst), fin


FLOOR_MULT = 50.0


@pytest.mark.parametrize("scene,iters", [("toy1_lv2", 12), ("toy1_lv9", 12), ("toy2_lv2", 12), ("manual3", 10),
                                         ("synth_small", 12), ("synth_mid", 8)])
def test_p4_lm_trajectory_matches_oracle(LineBA, orc, scene, iters):
    """Same algorithm, same iteration count, device vs oracle: per-iteration cost within 1e-9 relative,
    final cost and parameters within 1e-6 relative -- or within FLOOR_MULT x the problem's own
    rounding sensitivity (see _conditioning_floor) where that is larger."""
    if scene == "synth_small":
        s = synth(12, 200, 5, seed=1)
    elif scene == "synth_mid":
        s = synth(50, 2000, 10, seed=0)
    else:
        s = load(scene)[0]
    kw = dict(max_iter=iters, ftol=0.0, xtol=0.0, gtol=0.0)
    ba = make_ctx(LineBA, s)
    res = ba.lm_solve(s.camera_params, s.points_3d, **kw)
    ref = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(**kw))
    cams_o, pts_o, r_o, rep = ref
    assert res.iterations == rep.iterations == iters
    cost_d, acc_d, pcg_d = _traj(res)
    cost_o = np.array([t["cost_new"] for t in rep.trace])
    floor, fin = _conditioning_floor(orc, s, kw, cost_o, ref)
    print(f"\n[{scene}] max per-iteration rel diff {np.max(np.abs(cost_d - cost_o) / cost_o):.2e} "
          f"(conditioning floor {floor.max():.2e}); final cost rel {abs(res.cost - rep.cost) / rep.cost:.2e} "
          f"(floor {fin['cost']:.2e})")
    assert acc_d == [t["accepted"] for t in rep.trace]
    assert pcg_d == [t["pcg"] for t in rep.trace]
    assert [t["it"] for t in res.trace] == [t["it"] for t in rep.trace]          # same retries
    assert np.allclose([t["alpha"] for t in res.trace], [t["alpha"] for t in rep.trace], rtol=max(1e-6, FLOOR_MULT * floor.max()))
    assert res.n_accepted == rep.n_accepted
    assert np.all(np.abs(cost_d - cost_o) <= np.maximum(RTOL_TRAJ, FLOOR_MULT * floor) * np.abs(cost_o))
    assert abs(res.cost - rep.cost) <= max(RTOL_FINAL, FLOOR_MULT * fin["cost"]) * rep.cost
    assert np.max(np.abs(res.camera_params - cams_o)) <= max(RTOL_FINAL, FLOOR_MULT * fin["cams"]) * np.max(np.abs(cams_o))
    assert np.max(np.abs(res.points_3d - pts_o)) <= max(RTOL_FINAL, FLOOR_MULT * fin["pts"]) * np.max(np.abs(pts_o))
    # the returned residual is fun at the returned parameters (bundle_adjustment.py:332)
    r_chk = orc.residuals(res.camera_params, res.points_3d, *s.arrays()[2:])
    assert np.all(np.abs(res.residuals - r_chk) <= residual_tol(r_chk))
    assert res.chol_clamped == 0
    ba.close()


@pytest.mark.parametrize("scene", ["synth_small", "synth_mid", "toy1_lv9"])
def test_p4_strict_on_well_conditioned_scenes(LineBA, orc, scene):
    """No floor allowance: 1e-9 per iteration, 1e-6 final, on scenes with many cameras per gauge.  (toy1_lv9 is
    noise-free: from the sixth iteration on its damping has fallen below 1e-4 and the seven gauge directions, which
    only the damping holds, set a rounding floor above 1e-9 -- the floor-aware test above covers those iterations.)"""
    s = {"synth_small": lambda: synth(12, 200, 5, seed=1), "synth_mid": lambda: synth(50, 2000, 10, seed=0),
         "toy1_lv9": lambda: load("toy1_lv9")[0]}[scene]()
    kw = dict(max_iter=5 if scene == "toy1_lv9" else 8, ftol=0.0, xtol=0.0, gtol=0.0)
    ba = make_ctx(LineBA, s)
    res = ba.lm_solve(s.camera_params, s.points_3d, **kw)
    cams_o, pts_o, r_o, rep = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(**kw))
    cost_d = np.array([t["cost_new"] for t in res.trace]); cost_o = np.array([t["cost_new"] for t in rep.trace])
    assert np.all(np.abs(cost_d - cost_o) <= RTOL_TRAJ * cost_o)
    assert abs(res.cost - rep.cost) <= RTOL_FINAL * rep.cost
    assert np.max(np.abs(res.camera_params - cams_o)) <= RTOL_FINAL * np.max(np.abs(cams_o))
    assert np.max(np.abs(res.points_3d - pts_o)) <= RTOL_FINAL * np.max(np.abs(pts_o))
    ba.close()


def test_p4_termination_rules_match_oracle(LineBA, orc):
    """Default tolerances (ftol=1e-4 as the reference call site, xtol=gtol=1e-8)."""
    s = synth(12, 200, 5, seed=1)
    ba = make_ctx(LineBA, s)
    res = ba.lm_solve(s.camera_params, s.points_3d, max_iter=60)
    rep = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(max_iter=60))[3]
    assert (res.iterations, res.status) == (rep.iterations, rep.status)
    assert abs(res.cost - rep.cost) <= RTOL_FINAL * rep.cost
    ba.close()


# ---------------------------------------------------------------- P5
def test_p5_dropin_beats_reference_run(LineBA, capsys):
    """run_bundle_adjustment on the scene whose reference output was recorded: same returned shapes
    and layouts; final cost below the reference's (its as-shipped pattern stalls)."""
    from amuya_sfm_lines_beta_b200.bundle_adjustment import run_bundle_adjustment
    from test_cabi_and_host import _objects_from_golden
    s, q = load("toy1_lv2")
    lines_3d, lines_2d = _objects_from_golden(q)
    inp = {"camera_params": q["camera_params"], "Lines_3d": lines_3d, "camera_indices": q["camera_indices"],
           "lines_indices": q["lines_indices"], "lines_2d": lines_2d}
    cams_before = q["camera_params"].copy()
    opc, opp, res = run_bundle_adjustment(inp, "linear")
    out = capsys.readouterr().out
    assert out.splitlines()[0] == str(q["ref_ba_stdout_head"]).splitlines()[0]      # same banner
    assert out.splitlines()[1:6] == str(q["ref_ba_stdout_head"]).splitlines()[1:6]  # same size lines
    assert "Optimization took" in out
    assert opc.shape == q["ref_ba_op_camera_params"].shape and opc.dtype == np.float64
    assert opp.shape == q["ref_ba_op_points_3d"].shape and opp.dtype == np.float64
    assert res.shape == q["ref_ba_residual"].shape
    assert np.array_equal(inp["camera_params"], cams_before)                        # inputs not mutated
    ref_cost = 0.5 * float(np.sum(q["ref_ba_residual"] ** 2))
    assert 0.5 * float(res @ res) < ref_cost


def test_p5_final_cost_vs_reference_on_every_fixture(LineBA, capsys):
    """SURVEY 8(c) P5 on every fixture, at the SAME number of evaluations as the reference's scipy call was given
    (max_nfev = 15: the initial point + 14 trial points; every one of ours is a residual+Jacobian evaluation):
    the CUDA solver ends at or below BOTH the reference as shipped (finite differences with the shipped sparsity
    pattern, ftol=1e-4) and scipy with the analytic Jacobian in the right columns (tests/golden/p5_costs.json).
    The default run (ftol=1e-4 as the reference call site) is printed beside it with its accept rate."""
    import json
    p5j = json.load(open(os.path.join(GOLDEN, "p5_costs.json")))
    budget = p5j["njev_budget"]
    lines = []
    for name, ref in p5j["fixtures"].items():
        s = synth(12, 200, 5, seed=1) if name.startswith("synth") else load(name)[0]
        ba = make_ctx(LineBA, s)
        res = ba.lm_solve(s.camera_params, s.points_3d, max_iter=50, max_evals=budget)
        dflt = ba.lm_solve(s.camera_params, s.points_3d)
        ba.close()
        lines.append(f"[P5] {name:16s} init {res.cost_init:.3e}  ours @{res.n_jac} evals {res.cost:.3e} (accepted "
                     f"{res.n_accepted}/{len(res.trace)} trials)  reference as shipped {ref['as_shipped_cost']:.3e}  "
                     f"scipy+analytic J @{budget} {ref['analytic_cost']:.3e} | default run: {dflt.iterations} it, "
                     f"{dflt.n_jac} evals, status {dflt.status}, cost {dflt.cost:.3e}, accepted {dflt.n_accepted}/{len(dflt.trace)}")
        assert abs(res.cost_init - ref["cost_init"]) <= 1e-9 * ref["cost_init"]
        assert res.n_jac <= budget
        assert res.cost <= min(ref["as_shipped_cost"], ref["analytic_cost"]), lines[-1]
        assert dflt.cost <= ref["as_shipped_cost"], lines[-1]
    with capsys.disabled():
        print("\n" + "\n".join(lines))


# ---------------------------------------------------------------- edge cases
def test_unsorted_observations_give_identical_results(LineBA):
    s = synth(10, 150, 5, seed=7)
    perm = np.random.default_rng(0).permutation(s.n_obs)
    ba = make_ctx(LineBA, s)
    r0, c0 = ba.residuals(s.camera_params, s.points_3d)
    res0 = ba.lm_solve(s.camera_params, s.points_3d, max_iter=5, ftol=0.0)
    ba2 = LineBA(0)
    ba2.set_problem(s.n_cameras, s.n_lines, s.camera_indices[perm], s.lines_indices[perm], s.n[perm], s.p0[perm], s.p1[perm])
    r1, c1 = ba2.residuals(s.camera_params, s.points_3d)
    assert np.array_equal(r1, r0[perm])
    cam_pt, pt_idx = ba2.point_wise_indices()
    assert np.array_equal(cam_pt, np.repeat(s.camera_indices[perm], 2))
    assert np.array_equal(pt_idx[0::2], 2 * s.lines_indices[perm])
    # a stable sort by line keeps the per-line order only if the permutation did; costs agree to rounding
    res1 = ba2.lm_solve(s.camera_params, s.points_3d, max_iter=5, ftol=0.0)
    assert abs(res1.cost - res0.cost) <= 1e-9 * res0.cost
    assert np.allclose(res1.residuals, res0.residuals[perm], rtol=1e-7, atol=1e-12)
    ba.close(); ba2.close()


def test_bitwise_reproducible(LineBA):
    s = synth(30, 800, 8, seed=3)
    ba = make_ctx(LineBA, s)
    a = ba.lm_solve(s.camera_params, s.points_3d, max_iter=6, ftol=0.0)
    b = ba.lm_solve(s.camera_params, s.points_3d, max_iter=6, ftol=0.0)
    assert np.array_equal(a.camera_params, b.camera_params) and np.array_equal(a.points_3d, b.points_3d)
    assert a.cost == b.cost
    ba.close()


def test_lines_and_cameras_without_observations(LineBA, orc):
    s = synth(8, 60, 4, seed=11)
    # append two 3D lines and one camera that nothing observes (Lines_3d holds ALL lines, seq_utils.py:198)
    pts = np.vstack([s.points_3d, np.array([[1., 2, 3], [4, 5, 6], [0, 0, 1], [1, 1, 1]])])
    cams = np.vstack([s.camera_params, np.array([[0.1, 5, 0.2, 6, 0.3, 7]])])
    ba = LineBA(0)
    ba.set_problem(9, 62, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1)
    res = ba.lm_solve(cams, pts, max_iter=6, ftol=0.0)
    ref = orc.lm_solve(cams, pts, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1,
                       opts=orc.LMOptions(max_iter=6, ftol=0.0))
    assert abs(res.cost - ref[3].cost) <= RTOL_FINAL * ref[3].cost
    assert np.array_equal(res.points_3d[-4:], pts[-4:])      # untouched
    assert np.array_equal(res.camera_params[-1], cams[-1])
    ba.close()


def test_empty_problem(LineBA):
    ba = LineBA(0)
    z3 = np.zeros((0, 3))
    ba.set_problem(2, 3, np.zeros(0, np.int64), np.zeros(0, np.int64), z3, z3, z3)
    r, cost = ba.residuals(np.zeros((2, 6)), np.ones((6, 3)))
    assert r.shape == (0,) and cost == 0.0
    res = ba.lm_solve(np.zeros((2, 6)), np.ones((6, 3)), max_iter=3)
    assert res.cost == 0.0 and res.status == 1          # gradient is zero: gtol
    ba.close()


def test_error_paths(LineBA):
    from amuya_sfm_lines_beta_b200._lib import LbaError, LBA_ERR_ARG, LBA_ERR_NONFINITE, LBA_ERR_STATE
    ba = LineBA(0)
    with pytest.raises(LbaError) as e:
        ba.point_wise_indices()                            # no problem set yet
    assert e.value.code == LBA_ERR_STATE
    one = np.array([[0.0, 0.0, 1.0]])
    with pytest.raises(LbaError) as e:
        ba.set_problem(1, 1, np.array([1]), np.array([0]), one, one, one)
    assert e.value.code == LBA_ERR_ARG
    s = synth(4, 10, 3, seed=0)
    ba.set_problem(s.n_cameras, s.n_lines, s.camera_indices, s.lines_indices, s.n, s.p0, s.p1)
    bad = s.points_3d.copy(); bad[0, 0] = np.nan
    with pytest.raises(LbaError, match="not finite in the initial point") as e:
        ba.lm_solve(s.camera_params, bad)
    assert e.value.code == LBA_ERR_NONFINITE
    ba.close()


def test_device_resident_inputs(LineBA):
    """torch CUDA tensors are consumed in place (no host round trip of the observation arrays)."""
    import torch
    s = synth(10, 150, 5, seed=7)
    ba = make_ctx(LineBA, s)
    r0, _ = ba.residuals(s.camera_params, s.points_3d)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ba2 = LineBA(0)
    ba2.set_problem(s.n_cameras, s.n_lines, t(s.camera_indices), t(s.lines_indices), t(s.n), t(s.p0), t(s.p1))
    r1, _ = ba2.residuals(t(s.camera_params), t(s.points_3d))
    assert np.array_equal(r0, r1)
    ba.close(); ba2.close()


@pytest.mark.parametrize("loss,f_scale", [("soft_l1", 1e-3), ("huber", 1e-3), ("cauchy", 1e-3), ("arctan", 5e-3), ("cauchy", 1.0)])
def test_robust_losses_follow_the_oracle(LineBA, orc, loss, f_scale):
    """N4: scipy's robust losses (loss_f is forwarded to least_squares at bundle_adjustment.py:319).  The
    oracle's rho / row scaling are pinned to scipy's own functions (CPU test); here the CUDA LM must follow
    the oracle LM with the same loss on a scene with gross outliers."""
    s = synth(12, 300, 6, seed=2)
    rng = np.random.default_rng(5)
    bad = rng.choice(s.n_obs, 60, replace=False)
    s.n[bad] = np.roll(s.n[bad], 1, axis=1); s.p0[bad] = np.roll(s.p0[bad], 1, axis=1); s.p1[bad] = np.roll(s.p1[bad], 1, axis=1)
    kw = dict(max_iter=8, ftol=0.0, xtol=0.0, gtol=0.0)
    ba = make_ctx(LineBA, s)
    res = ba.lm_solve(s.camera_params, s.points_3d, loss=loss, f_scale=f_scale, **kw)
    cams_o, pts_o, r_o, rep = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(loss=loss, f_scale=f_scale, **kw))
    assert abs(res.cost_init - rep.cost_init) <= 1e-12 * rep.cost_init
    cost_d = np.array([t["cost_new"] for t in res.trace]); cost_o = np.array([t["cost_new"] for t in rep.trace])
    assert [t["accepted"] for t in res.trace] == [t["accepted"] for t in rep.trace]
    assert np.all(np.abs(cost_d - cost_o) <= RTOL_TRAJ * cost_o)
    assert abs(res.cost - rep.cost) <= RTOL_FINAL * rep.cost
    assert np.max(np.abs(res.points_3d - pts_o)) <= RTOL_FINAL * np.max(np.abs(pts_o))
    # the returned residual vector is fun itself (scipy's res.fun), not the rescaled one
    r_chk = orc.residuals(res.camera_params, res.points_3d, *s.arrays()[2:])
    assert np.all(np.abs(res.residuals - r_chk) <= residual_tol(r_chk))
    assert abs(res.cost - orc.robust_cost(r_chk, loss, f_scale)) <= 1e-10 * res.cost
    ba.close()


def test_dropin_forwards_loss_f(LineBA, capsys):
    from amuya_sfm_lines_beta_b200.bundle_adjustment import run_bundle_adjustment
    from test_cabi_and_host import _objects_from_golden
    s, q = load("toy1_lv2")
    lines_3d, lines_2d = _objects_from_golden(q)
    inp = {"camera_params": q["camera_params"], "Lines_3d": lines_3d, "camera_indices": q["camera_indices"],
           "lines_indices": q["lines_indices"], "lines_2d": lines_2d}
    _, _, _, rep_lin = run_bundle_adjustment(inp, "linear", return_report=True, max_iter=5)
    _, _, _, rep_hub = run_bundle_adjustment(inp, "soft_l1", return_report=True, max_iter=5, f_scale=1e-4)
    capsys.readouterr()
    assert rep_hub.cost_init < rep_lin.cost_init * 1.0000001 and rep_hub.cost_init != rep_lin.cost_init
    with pytest.raises(ValueError, match="`loss` must be one of"):
        run_bundle_adjustment(inp, "bogus")


def test_many_cameras_take_the_global_gather_variant(LineBA, orc):
    """More cameras than the shared-memory camera table holds (~1.3k): the evaluation kernel falls back to
    the global-memory gather variant; results must not change."""
    s = synth(3000, 6000, 6, seed=4)
    ba = make_ctx(LineBA, s)
    r, jcam, jline = ba.residuals_jacobian(s.camera_params, s.points_3d)
    r_o, jc_o, jl_o = _oracle_jac(orc, s)
    assert np.all(np.abs(r - r_o) <= residual_tol(r_o))
    scale = max(np.abs(jc_o).max(), np.abs(jl_o).max())
    assert np.max(np.abs(jcam - _interleave(jc_o))) < 1e-12 * scale
    assert np.max(np.abs(jline - jl_o)) < 1e-12 * scale
    kw = dict(max_iter=5, ftol=0.0, xtol=0.0, gtol=0.0)
    res = ba.lm_solve(s.camera_params, s.points_3d, **kw)
    rep = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(**kw))[3]
    cost_d = np.array([t["cost_new"] for t in res.trace]); cost_o = np.array([t["cost_new"] for t in rep.trace])
    assert np.all(np.abs(cost_d - cost_o) <= RTOL_TRAJ * cost_o)
    ba.close()


@pytest.mark.parametrize("shape", [(40, 3000, 9), (12, 4000, 3), (1500, 3000, 6)])
def test_jacobian_kernels_agree(LineBA, monkeypatch, shape):
    """The Jacobian kernel of round 2 (k_eval_z: decoupled load pipeline, sums by two warps, bulk-copied V_l / g_l, C
    recomputed) against the round-1 kernel it replaces (k_eval<true>, LBA_EVAL_Z=0), with and without the
    shared-memory camera table and with several batches of lines per tile (3 observations per line).  Same formulas and
    the same order of every sum; the two instantiations are compiled separately (FMA contraction may differ), so the
    comparison is to a few ulps of each array's scale, not bitwise."""
    s = synth(*shape, seed=17)
    out = {}
    for z in ("0", "1"):
        monkeypatch.setenv("LBA_EVAL_Z", z)
        ba = make_ctx(LineBA, s)
        r, jc, jl = ba.residuals_jacobian(s.camera_params, s.points_3d)
        U, gc, V, gl = ba.normal_equations()
        out[z] = (r, jc, jl, V, gl, U, gc)
        ba.close()
    names = ("r", "jc", "jl", "V", "gl", "U", "gc")
    for name, a0, a1 in zip(names, out["0"], out["1"]):
        scale = max(np.abs(a0).max(), 1e-300)
        assert np.max(np.abs(a0 - a1)) <= 1e-13 * scale, name


@pytest.mark.parametrize("scene", ["manual3", "toy1_lv2", "synth_noisy"])
def test_retry_and_evaluation_budget_match_oracle(LineBA, orc, scene):
    """Scenes whose first Gauss-Newton steps overshoot: the rejected trial is retried at the measured length along
    the same direction (no new solve) and `max_evals` stops the loop like scipy's max_nfev.  Same trials, same
    lengths, same costs as the oracle."""
    s = synth(12, 200, 5, seed=1) if scene == "synth_noisy" else load(scene)[0]
    kw = dict(max_iter=40, max_evals=22, ftol=0.0, xtol=0.0, gtol=0.0)
    ba = make_ctx(LineBA, s)
    res = ba.lm_solve(s.camera_params, s.points_3d, **kw)
    rep = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(**kw))[3]
    assert res.n_jac == 22 == 1 + len(res.trace)
    assert [t["it"] for t in res.trace] == [t["it"] for t in rep.trace]
    assert [t["accepted"] for t in res.trace] == [t["accepted"] for t in rep.trace]
    if scene != "toy1_lv2":
        assert len(res.trace) > res.iterations          # at least one retry happened
    assert np.allclose([t["alpha"] for t in res.trace], [t["alpha"] for t in rep.trace], rtol=1e-6)
    cost_d = np.array([t["cost_new"] for t in res.trace]); cost_o = np.array([t["cost_new"] for t in rep.trace])
    assert np.all(np.abs(cost_d - cost_o) <= 1e-6 * cost_o)
    assert abs(res.cost - rep.cost) <= RTOL_FINAL * rep.cost
    ba.close()


def test_config4_trajectory_matches_recorded_oracle(LineBA):
    """BASELINE config [3] at full size (1k cameras / 200k lines / 2M observations): the device trajectory against
    the oracle's, recorded by tests/golden/make_config4_trajectory.py (the numpy oracle needs minutes for it)."""
    import json
    fx = json.load(open(os.path.join(GOLDEN, "config4_trajectory.json")))
    s = synth(1000, 200_000, 10, seed=0, variant="A")
    ba = make_ctx(LineBA, s)
    res = ba.lm_solve(s.camera_params, s.points_3d, **fx["options"])
    ba.close()
    assert abs(res.cost_init - fx["cost_init"]) <= 1e-12 * fx["cost_init"]
    assert [t["it"] for t in res.trace] == fx["it"]
    assert [bool(t["accepted"]) for t in res.trace] == fx["accepted"]
    assert [t["pcg"] for t in res.trace] == fx["pcg"]
    cost_d = np.array([t["cost_new"] for t in res.trace]); cost_o = np.array(fx["cost_new"])
    print(f"\n[config 4] per-trial rel diff vs recorded oracle: {np.max(np.abs(cost_d - cost_o) / cost_o):.2e}")
    assert np.all(np.abs(cost_d - cost_o) <= RTOL_TRAJ * cost_o)
    assert abs(res.cost - fx["cost"]) <= RTOL_FINAL * fx["cost"]
    assert abs(np.linalg.norm(res.camera_params) - fx["cams_norm"]) <= RTOL_FINAL * fx["cams_norm"]
    assert abs(np.linalg.norm(res.points_3d) - fx["pts_norm"]) <= RTOL_FINAL * fx["pts_norm"]


# ---------------------------------------------------------------- full size (config 4) properties
@pytest.mark.parametrize("variant", ["A"])
def test_config4_full_size_properties(LineBA, orc, variant):
    """1k cameras / 200k lines / 2M observations: residuals checked against the oracle on the whole
    scene (vectorised numpy), cost monotone over accepted steps, bit-reproducible, launches counted."""
    s = synth(1000, 200_000, 10, seed=0, variant=variant)
    ba = make_ctx(LineBA, s)
    r, cost = ba.residuals(s.camera_params, s.points_3d)
    ref = orc.residuals(*s.arrays())
    assert np.all(np.abs(r - ref) <= residual_tol(ref))
    n0 = ba.launch_count()
    a = ba.lm_solve(s.camera_params, s.points_3d, max_iter=6, ftol=0.0)
    assert ba.launch_count() > n0
    b = ba.lm_solve(s.camera_params, s.points_3d, max_iter=6, ftol=0.0)
    assert a.cost == b.cost and np.array_equal(a.points_3d, b.points_3d)
    costs = [t["cost_new"] for t in a.trace if t["accepted"]]
    assert all(x > y for x, y in zip([a.cost_init] + costs[:-1], costs))
    assert a.cost < 0.5 * a.cost_init
    r_fin = orc.residuals(a.camera_params, a.points_3d, *s.arrays()[2:])
    assert np.all(np.abs(a.residuals - r_fin) <= residual_tol(r_fin))
    assert abs(a.cost - 0.5 * r_fin @ r_fin) <= 1e-10 * a.cost
    ba.close()


@pytest.mark.parametrize("transport", ["peer", "peer_two_phase", "nccl"])
def test_multi_gpu_sharded_solve_matches_single_gpu(transport):
    """N > 1: run tests/run_multigpu_parity.py under torchrun when the box has at least two GPUs
    (observations sharded at line boundaries, per-camera partial sums all-reduced; the sharded trajectory
    must equal the single-GPU one).  Transports: NVLink peer memory (one-phase reduce), the same with the
    reduce-scatter + all-gather kernel forced for every per-camera exchange, and the NCCL fallback."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    env.pop("LBA_NO_PEER", None)
    env.pop("LBA_PEER_TWO_PHASE_MIN", None)
    if transport == "peer_two_phase":
        env["LBA_PEER_TWO_PHASE_MIN"] = "0"
    elif transport == "nccl":
        env["LBA_NO_PEER"] = "1"
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_multigpu_parity.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_multi_gpu_stalled_peer_is_an_error_not_a_hang():
    """A rank whose peers never join the exchange gets LBA_ERR_COMM from the bounded device-side waits."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "run_multigpu_parity.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29537", script, "--stall"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "LBA_ERR_COMM" in out.stdout or "error -6" in out.stdout, out.stdout[-2000:]


def test_cuda_tensor_inputs_are_ordered_behind_the_torch_stream(LineBA, orc):
    """Inputs that are CUDA tensors still being written by kernels on a torch side stream: the context's stream
    waits for that stream (lba_wait_stream) before reading them, so the results equal the host-array path."""
    import torch
    s = synth(20, 4000, 8, seed=5)
    ba = make_ctx(LineBA, s)
    r_ref, cost_ref = ba.residuals(s.camera_params, s.points_3d)
    side = torch.cuda.Stream()
    big = torch.randn(6144, 6144, device="cuda")
    hc = torch.from_numpy(s.camera_params).pin_memory(); hp = torch.from_numpy(s.points_3d).pin_memory()
    for _ in range(3):
        with torch.cuda.stream(side):
            for _ in range(6):
                big = big @ big * 1e-3                     # keeps the side stream busy for milliseconds
            cams = hc.to("cuda", non_blocking=True) * 1.0    # written behind that work, on the side stream
            pts = hp.to("cuda", non_blocking=True) * 1.0
            r, cost = ba.residuals(cams, pts)
        assert cost == cost_ref and np.array_equal(r, r_ref)
    ba.close()


def _long_line_scene():
    """Three 3D lines seen from 700, 300 and 257 cameras (more than one 256-slot tile each) among 300 ordinary
    lines; the reference has no limit on the observations of a line (bundle_adjustment.py:194-250)."""
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    ks = np.full(303, 5)
    ks[[7, 150, 302]] = [700, 300, 257]
    return make_synthetic_scene(720, 303, ks, seed=11)


def test_lines_with_more_observations_than_a_tile(LineBA, orc):
    """Lines split over several fragment tiles: blocks, Schur solve and the LM trajectory against the oracle."""
    s = _long_line_scene()
    assert np.bincount(s.lines_indices).max() == 700
    ba = make_ctx(LineBA, s)
    r, jcam, jline = ba.residuals_jacobian(s.camera_params, s.points_3d)
    r_o, jc_o, jl_o = _oracle_jac(orc, s)
    assert np.all(np.abs(r - r_o) <= residual_tol(r_o))
    U, gc, V, gl = ba.normal_equations()
    Uo, gco, Vo, glo = orc.normal_equations(r_o, jc_o, jl_o, s.camera_indices, s.lines_indices, s.n_cameras, s.n_lines)
    iu = np.triu_indices(6)
    assert np.max(np.abs(V - Vo[:, iu[0], iu[1]])) <= 1e-10 * np.abs(Vo).max()
    assert np.max(np.abs(gl - glo)) <= 1e-10 * np.abs(glo).max()
    assert np.max(np.abs(gc - gco)) <= 1e-10 * np.abs(gco).max()
    kw = dict(max_iter=8, ftol=0.0, xtol=0.0, gtol=0.0)
    res = ba.lm_solve(s.camera_params, s.points_3d, **kw)
    cams_o, pts_o, _, rep = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(**kw))
    assert [t["accepted"] for t in res.trace] == [t["accepted"] for t in rep.trace]
    assert [t["pcg"] for t in res.trace] == [t["pcg"] for t in rep.trace]
    cost_d = np.array([t["cost_new"] for t in res.trace]); cost_o = np.array([t["cost_new"] for t in rep.trace])
    print(f"\n[long lines] per-trial rel diff {np.max(np.abs(cost_d - cost_o) / cost_o):.2e}")
    assert np.all(np.abs(cost_d - cost_o) <= RTOL_TRAJ * cost_o)
    assert abs(res.cost - rep.cost) <= RTOL_FINAL * rep.cost
    assert np.max(np.abs(res.camera_params - cams_o)) <= RTOL_FINAL * np.max(np.abs(cams_o))
    assert np.max(np.abs(res.points_3d - pts_o)) <= RTOL_FINAL * np.max(np.abs(pts_o))
    assert res.cost < 0.2 * res.cost_init
    # bit-reproducible run to run; unsorted input (another order inside the lines) agrees to rounding
    res1 = ba.lm_solve(s.camera_params, s.points_3d, **kw)
    assert res1.cost == res.cost and np.array_equal(res1.points_3d, res.points_3d)
    perm = np.random.default_rng(1).permutation(s.n_obs)
    ba2 = LineBA(0)
    ba2.set_problem(s.n_cameras, s.n_lines, s.camera_indices[perm], s.lines_indices[perm], s.n[perm], s.p0[perm], s.p1[perm])
    res2 = ba2.lm_solve(s.camera_params, s.points_3d, **kw)
    assert abs(res2.cost - res.cost) <= 1e-9 * res.cost
    assert np.array_equal(res2.residuals[np.argsort(perm)].shape, res.residuals.shape)
    ba.close(); ba2.close()


# ---------------------------------------------------------------- 12-parameter [R|c] camera rows
def _p12():
    q = np.load(os.path.join(GOLDEN, "toy1_lv2_p12.npz"))
    return {k: q[k] for k in q.files}


def test_p12_residuals_and_project_match_reference(LineBA):
    """(Nc,12) camera rows: device ``fun`` against the reference's own outputs at two points, ``project`` against its
    [R|c] branch (bundle_adjustment.py:43-63)."""
    q = _p12()
    nc = q["camera_params"].shape[0]
    ba = LineBA(0)
    ba.set_problem(nc, q["points_3d_raw"].shape[0] // 2, q["camera_indices"], q["lines_indices"], q["n"], q["p0"], q["p1"])
    for key in ("x0", "x1"):
        x = q["ref_" + key]
        r, cost = ba.residuals(x[:12 * nc].reshape(nc, 12), x[12 * nc:].reshape(-1, 3))
        ref = q["ref_fun_" + key]
        assert np.all(np.abs(r - ref) <= residual_tol(ref))
        assert abs(cost - 0.5 * ref @ ref) <= 1e-12 * cost
    x0 = q["ref_x0"]
    pts = x0[12 * nc:].reshape(-1, 3)
    cam_pt = np.repeat(q["camera_indices"], 2)
    pt_idx = np.empty(2 * len(q["lines_indices"]), dtype=np.int64)
    pt_idx[0::2] = 2 * q["lines_indices"]; pt_idx[1::2] = 2 * q["lines_indices"] + 1
    proj = ba.project12(pts[pt_idx], x0[:12 * nc].reshape(nc, 12)[cam_pt])
    assert np.max(np.abs(proj - q["ref_project_x0"])) < 1e-14
    ba.close()
    from amuya_sfm_lines_beta_b200.bundle_adjustment import project
    proj2 = project(pts[pt_idx], x0[:12 * nc].reshape(nc, 3, 4)[cam_pt])
    assert np.array_equal(proj2, proj)


def test_p12_solver_follows_oracle_and_beats_reference_run(LineBA, orc):
    """The dense 12-parameter LM: same trials as the oracle's dense LM (same policy), and a final cost below what the
    reference's own run_bundle_adjustment reached on this input; the drop-in returns (Nc,12) rows."""
    q = _p12()
    nc = q["camera_params"].shape[0]
    cams, pts = q["camera_params"], q["points_3d_raw"].astype(np.float64)
    a = (q["camera_indices"], q["lines_indices"], q["n"], q["p0"], q["p1"])
    ba = LineBA(0)
    ba.set_problem(nc, pts.shape[0] // 2, *a)
    kw = dict(max_iter=12, ftol=0.0, xtol=0.0, gtol=0.0)
    res = ba.lm_solve(cams, pts, **kw)
    cams_o, pts_o, _, rep = orc.lm_solve12(cams, pts, *a, opts=orc.LMOptions(**kw))
    assert res.camera_params.shape == (nc, 12)
    assert [t["it"] for t in res.trace] == [t["it"] for t in rep.trace]
    assert [t["accepted"] for t in res.trace] == [t["accepted"] for t in rep.trace]
    cost_d = np.array([t["cost_new"] for t in res.trace]); cost_o = np.array([t["cost_new"] for t in rep.trace])
    print(f"\n[12-parameter rows] per-trial rel diff vs oracle {np.max(np.abs(cost_d - cost_o) / cost_o):.2e}")
    # 12 + Nc gauge directions (a 3D affinity and one scale per camera) are held by the damping alone: rounding
    # differences between the two dense Cholesky solvers grow along them, hence 1e-6 per trial instead of 1e-9
    assert np.all(np.abs(cost_d - cost_o) <= 1e-6 * cost_o)
    assert abs(res.cost - rep.cost) <= 1e-5 * rep.cost
    ref_cost = 0.5 * float(np.sum(q["ref_ba_residual"] ** 2))
    dflt = ba.lm_solve(cams, pts)
    assert dflt.cost < ref_cost
    r_chk = orc.residuals12(dflt.camera_params, dflt.points_3d, *a)
    assert np.all(np.abs(dflt.residuals - r_chk) <= residual_tol(r_chk))
    ba.close()
    from amuya_sfm_lines_beta_b200.bundle_adjustment import run_bundle_adjustment
    from test_cabi_and_host import _objects_from_golden
    lines_3d, lines_2d = _objects_from_golden(q)
    inp = {"camera_params": cams, "Lines_3d": lines_3d, "camera_indices": a[0], "lines_indices": a[1], "lines_2d": lines_2d}
    opc, opp, rr = run_bundle_adjustment(inp, "linear", verbose=0)
    assert opc.shape == q["ref_ba_op_camera_params"].shape == (nc, 12) and opp.shape == q["ref_ba_op_points_3d"].shape
    assert 0.5 * float(rr @ rr) < ref_cost
