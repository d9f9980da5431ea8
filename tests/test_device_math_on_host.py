"""CPU: the device-side arithmetic of the hot path (csrc/lba_geometry.cuh: Rodrigues coefficients, the residual and
analytic Jacobian of one observation, the robust-loss functions, the packed 6x6 Cholesky inverse) compiled for the
HOST with g++ and checked against the reference-generated fixtures and the oracle.  The library never runs this
code on the CPU; this is test infrastructure that lets a change to the kernels' arithmetic be caught without a GPU.
Tolerances are those of the GPU parity ladder (P1: 1e-9 relative per residual; P2: 1e-12 vs the oracle)."""
import ctypes as C
import math
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import line_ba_oracle as orc

GOLD = os.path.join(ROOT, "tests", "golden")
CUDA_INC = "/usr/local/cuda/include"

HOST_SRC = r"""
#include "lba_geometry.cuh"
using namespace lba;
extern "C" {
// one observation per row: camera (w, c), endpoints X0, X1, constants n, m0 = n x p0, m1 = n x p1
void geo_eval(long m, const double* w, const double* c, const double* X0, const double* X1, const double* n,
              const double* m0, const double* m1, double* r, double* jl, double* jc) {
  for (long i = 0; i < m; ++i) {
    double A, B, Cc, R[9], G[9];
    rodrigues_coeffs(w + 3 * i, A, B, Cc);
    // as k_eval_z does: C is not gathered with the camera record but recomputed from |w|^2 and A
    Cc = coeff_C(w[3 * i] * w[3 * i] + w[3 * i + 1] * w[3 * i + 1] + w[3 * i + 2] * w[3 * i + 2], A);
    build_R(w + 3 * i, A, B, R);
    build_G(w + 3 * i, B, Cc, G);
    double tq[3] = {0, 0, 0};
    const double e0 = endpoint_term<true>(R, c + 3 * i, X0 + 3 * i, n + 3 * i, m0 + 3 * i, jl + 6 * i, tq);
    const double e1 = endpoint_term<true>(R, c + 3 * i, X1 + 3 * i, n + 3 * i, m1 + 3 * i, jl + 6 * i + 3, tq);
    r[i] = e0 + e1;
    for (int k = 0; k < 3; ++k) {   // the kernel's d r / d w = G^T tq, d r / d c = -(jl0 + jl1)
      jc[6 * i + k] = G[k] * tq[0] + G[3 + k] * tq[1] + G[6 + k] * tq[2];
      jc[6 * i + 3 + k] = -(jl[6 * i + k] + jl[6 * i + 3 + k]);
    }
  }
}
void geo_rotation(long n, const double* w, double* R, double* G) {
  for (long i = 0; i < n; ++i) {
    double A, B, Cc;
    rodrigues_coeffs(w + 3 * i, A, B, Cc);
    build_R(w + 3 * i, A, B, R + 9 * i);
    build_G(w + 3 * i, B, Cc, G + 9 * i);
  }
}
// C of the camera record (rodrigues_coeffs) and C as the Jacobian kernel recomputes it
void geo_coeff_c(long n, const double* w, double* c_rec, double* c_kernel) {
  for (long i = 0; i < n; ++i) {
    double A, B;
    rodrigues_coeffs(w + 3 * i, A, B, c_rec[i]);
    c_kernel[i] = coeff_C(w[3 * i] * w[3 * i] + w[3 * i + 1] * w[3 * i + 1] + w[3 * i + 2] * w[3 * i + 2], A);
  }
}
void geo_rho(long n, int loss, const double* z, double* r0, double* r1, double* r2) {
  for (long i = 0; i < n; ++i) robust_rho(loss, z[i], r0[i], r1[i], r2[i]);
}
void geo_spd6(long n, double* S, int* clamped) {
  for (long i = 0; i < n; ++i) clamped[i] = spd6_inverse(S + 21 * i);
}
}
"""


@pytest.fixture(scope="module")
def geo(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("g++ or the CUDA headers are not available")
    d = tmp_path_factory.mktemp("geo_host")
    (d / "host.cpp").write_text(HOST_SRC)
    so = d / "libgeo_host.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", CUDA_INC, "-I", os.path.join(ROOT, "amuya_sfm_lines_beta_b200", "csrc"),
                    "-o", str(so), str(d / "host.cpp")], check=True)
    return C.CDLL(str(so))


def ptr(a):
    return C.c_void_p(a.ctypes.data)


def gathered(q, cams, pts):
    """Per-observation rows of the fixture at parameters (cams (Nc,6) interleaved, pts (2Nl,3))."""
    om, ce = orc.split_camera_params(cams)
    ci, li = q["camera_indices"].astype(np.int64), q["lines_indices"].astype(np.int64)
    X = np.asarray(pts, dtype=np.float64).reshape(-1, 2, 3)
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    return dict(w=c(om[ci]), c=c(ce[ci]), X0=c(X[li, 0]), X1=c(X[li, 1]), n=c(q["n"]), m0=c(np.cross(q["n"], q["p0"])),
                m1=c(np.cross(q["n"], q["p1"])), om=om, ce=ce, ci=ci, li=li)


def run_eval(geo, g):
    m = g["w"].shape[0]
    r = np.empty(m); jl = np.empty((m, 6)); jc = np.empty((m, 6))
    geo.geo_eval(C.c_long(m), ptr(g["w"]), ptr(g["c"]), ptr(g["X0"]), ptr(g["X1"]), ptr(g["n"]), ptr(g["m0"]), ptr(g["m1"]),
                 ptr(r), ptr(jl), ptr(jc))
    return r, jc, jl


@pytest.mark.parametrize("name", ["toy1_lv2", "toy1_lv5", "toy1_lv9", "toy2_lv2"])
def test_residuals_match_reference_fun(geo, name):
    """P1 on the host: the kernel's per-observation arithmetic against the reference's recorded fun(x0), fun(x1)."""
    q = dict(np.load(os.path.join(GOLD, name + ".npz")))
    nc = q["camera_params"].shape[0]
    for x, want in ((q["ref_x0"], q["ref_fun_x0"]), (q["ref_x1"], q["ref_fun_x1"])):
        cams, pts = x[:6 * nc].reshape(nc, 6), x[6 * nc:].reshape(-1, 3)
        r, _, _ = run_eval(geo, gathered(q, cams, pts))
        assert np.all(np.abs(r - want) <= 1e-9 * np.abs(want) + 16 * np.finfo(float).eps * np.sqrt(np.abs(want)))


@pytest.mark.parametrize("name", ["toy1_lv5", "toy2_lv2", "manual3"])
def test_jacobian_matches_oracle(geo, name):
    """P2 on the host: analytic Jacobian rows against the oracle (itself pinned to finite differences of the
    reference's fun)."""
    q = dict(np.load(os.path.join(GOLD, name + ".npz")))
    rng = np.random.default_rng(3)
    cams = q["camera_params"] + rng.normal(0, 1e-3, q["camera_params"].shape)
    pts = np.asarray(q["points_3d_raw"], dtype=np.float64) + rng.normal(0, 1e-3, q["points_3d_raw"].shape)
    g = gathered(q, cams, pts)
    r, jc, jl = run_eval(geo, g)
    ro, jco, jlo = orc.residuals_jacobian(g["om"], g["ce"], pts.reshape(-1, 6), g["ci"], g["li"], g["n"], g["m0"], g["m1"])
    scale = max(np.abs(jco).max(), np.abs(jlo).max())
    assert np.max(np.abs(r - ro)) <= 1e-12 * max(1.0, np.abs(ro).max())
    assert np.max(np.abs(jc - jco)) <= 1e-12 * scale and np.max(np.abs(jl - jlo)) <= 1e-12 * scale


def test_rotation_and_so3_jacobian_match_oracle_including_small_angles(geo):
    rng = np.random.default_rng(0)
    w = rng.normal(0, 1.0, (400, 3))
    w[:100] *= 1e-3; w[100:150] *= 1e-7; w[150] = 0.0; w[151] = [0.1 / np.sqrt(3)] * 3   # series branch, zero, the switch
    w = np.ascontiguousarray(w)
    R = np.empty((len(w), 9)); G = np.empty((len(w), 9))
    geo.geo_rotation(C.c_long(len(w)), ptr(w), ptr(R), ptr(G))
    assert np.max(np.abs(R.reshape(-1, 3, 3) - orc.rodrigues(w))) <= 4e-15
    assert np.max(np.abs(G.reshape(-1, 3, 3) - orc.so3_right_jacobian(w))) <= 4e-15


def test_recomputed_so3_coefficient_matches_the_camera_record(geo):
    """k_eval_z keeps 64 B camera records (w, c, A, B) in shared memory and recomputes C = (t - sin t) / t^3 from
    t^2 and A.  Against the C of the full record and against the Taylor series / closed form evaluated in
    longdouble, over rotation angles from 1e-9 to beyond pi, both sides of the series switch at t = 0.1 included."""
    t = np.concatenate([np.logspace(-9, np.log10(3.5), 600), [0.0, 0.0999999, 0.1, 0.1000001, np.pi]])
    rng = np.random.default_rng(5)
    d = rng.normal(size=(len(t), 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    w = np.ascontiguousarray(d * t[:, None])
    c_rec = np.empty(len(t)); c_ker = np.empty(len(t))
    geo.geo_coeff_c(C.c_long(len(t)), ptr(w), ptr(c_rec), ptr(c_ker))
    tl = np.linalg.norm(w.astype(np.longdouble), axis=1)
    exact = np.where(tl < 0.5, sum((-1) ** k * tl ** (2 * k) / np.longdouble(math.factorial(2 * k + 3)) for k in range(12)),
                     (tl - np.sin(tl)) / np.where(tl > 0, tl, 1) ** 3)
    # P2's tolerance against the oracle is 1e-12 of the Jacobian's scale; C enters it multiplied by |w|^2 <= 10
    assert np.max(np.abs(c_ker - c_rec) / c_rec) <= 2e-13
    assert np.max(np.abs(c_ker - exact.astype(np.float64)) / c_rec) <= 2e-13


@pytest.mark.parametrize("loss,code", [("soft_l1", 1), ("huber", 2), ("cauchy", 3), ("arctan", 4)])
def test_robust_rho_matches_oracle(geo, loss, code):
    z = np.ascontiguousarray(np.concatenate([np.logspace(-12, 4, 200), [0.0, 1.0, 1.0 + 1e-12]]))
    r0 = np.empty_like(z); r1 = np.empty_like(z); r2 = np.empty_like(z)
    geo.geo_rho(C.c_long(len(z)), C.c_int(code), ptr(z), ptr(r0), ptr(r1), ptr(r2))
    o0, o1, o2 = orc.robust_rho(loss, z)
    for a, b in ((r0, o0), (r1, o1), (r2, o2)):
        assert np.allclose(a, b, rtol=1e-14, atol=1e-300)


def test_spd6_inverse_matches_numpy(geo):
    rng = np.random.default_rng(1)
    n = 300
    A = rng.normal(size=(n, 6, 8))
    M = A @ np.swapaxes(A, 1, 2) + 1e-3 * np.eye(6)
    idx = [(a, b) for a in range(6) for b in range(a, 6)]
    S = np.ascontiguousarray(np.stack([M[:, a, b] for a, b in idx], axis=1))
    cl = np.empty(n, dtype=np.int32)
    geo.geo_spd6(C.c_long(n), ptr(S), ptr(cl))
    inv = np.linalg.inv(M)
    got = np.zeros_like(M)
    for k, (a, b) in enumerate(idx):
        got[:, a, b] = got[:, b, a] = S[:, k]
    assert not cl.any()
    assert np.max(np.abs(got - inv) / np.abs(inv).max(axis=(1, 2), keepdims=True)) <= 1e-10
    # a singular block is clamped and reported, not NaN
    Z = np.zeros((1, 21)); c1 = np.empty(1, dtype=np.int32)
    geo.geo_spd6(C.c_long(1), ptr(Z), ptr(c1))
    assert c1[0] > 0 and np.all(np.isfinite(Z))


# ---------------------------------------------------------------- the LM step policy (csrc/lba_policy.cuh)
POLICY_SRC = r"""
#include <string.h>
#include "lba_policy.cuh"
using namespace lba;
extern "C" {
// replays a sequence of trials through lm_decide; st_out rows: lam, sigma, alpha(next), cur, need_solve, done, status, it, evals, n_acc
int policy_replay(int n, const double* cost_new, const double* rjd, const double* jd2, const double* d2, const double* x2,
                  const double* g_norm, const int* pcg, double cost0, double lam0, int max_iter, int max_evals,
                  int max_backtrack, double ftol, double xtol, double gtol, double* st_out) {
  LmCtl L; memset(&L, 0, sizeof L);
  L.ftol = ftol; L.xtol = xtol; L.gtol = gtol; L.lambda_min = 1e-6; L.lambda_max = 1e12; L.min_ratio = 1e-4;
  L.max_iter = max_iter; L.max_evals = max_evals; L.pcg_max_iter = 200; L.max_backtrack = max_backtrack;
  L.lam = lam0; L.nu = 2.0; L.sigma = 1.0; L.alpha = 1.0; L.cost = cost0; L.need_solve = 1; L.evals = 1; L.scaling_mode = 2;
  int k = 0;
  for (; k < n && !L.done; ++k) {
    LmTrial t; t.cost_new = cost_new[k]; t.rjd = rjd[k]; t.jd2 = jd2[k]; t.d2 = d2[k]; t.x2 = x2[k]; t.g_norm = g_norm[k];
    t.pcg_iters = pcg[k]; t.pcg_done = 1;
    lm_decide(L, t, nullptr);
    double* o = st_out + 10 * k;
    o[0] = L.lam; o[1] = L.sigma; o[2] = L.alpha; o[3] = L.cur; o[4] = L.need_solve; o[5] = L.done; o[6] = L.status;
    o[7] = L.it; o[8] = L.evals; o[9] = L.n_acc;
  }
  return k;
}
}
"""


@pytest.fixture(scope="module")
def policy(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ is not available")
    d = tmp_path_factory.mktemp("policy_host")
    (d / "policy.cpp").write_text(POLICY_SRC)
    so = d / "libpolicy_host.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "amuya_sfm_lines_beta_b200", "csrc"),
                    "-o", str(so), str(d / "policy.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.policy_replay.restype = C.c_int
    return lib


@pytest.mark.parametrize("scene,kw", [("manual3", dict(max_iter=40, max_evals=22)), ("toy1_lv9", dict(max_iter=12)),
                                      ("synth", dict(max_iter=60, ftol=1e-4))])
def test_step_policy_header_follows_the_oracle(policy, scene, kw):
    """The decision kernel's function, compiled for the host and fed the oracle's own per-trial scalars, takes the
    oracle's decisions: same accepts / retries / termination and the same damping and step lengths (bit-equal)."""
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene, scene_from_npz
    s = make_synthetic_scene(12, 200, 5, seed=1) if scene == "synth" else scene_from_npz(os.path.join(GOLD, scene + ".npz"))
    kw = dict(dict(ftol=0.0, xtol=0.0, gtol=0.0), **kw)
    rep = orc.lm_solve(*s.arrays(), opts=orc.LMOptions(**kw))[3]
    tr = rep.trace
    n = len(tr)
    arr = lambda key: np.ascontiguousarray([t[key] for t in tr], dtype=np.float64)   # noqa: E731
    pcg = np.ascontiguousarray([t["pcg"] for t in tr], dtype=np.int32)
    out = np.zeros((n, 10))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))   # noqa: E731
    k = policy.policy_replay(n, dp(arr("cost_new")), dp(arr("rjd")), dp(arr("jd2")), dp(arr("d2")), dp(arr("x2")), dp(arr("g_norm")),
                             pcg.ctypes.data_as(C.POINTER(C.c_int)), C.c_double(rep.cost_init), C.c_double(1e-2),
                             kw["max_iter"], kw.get("max_evals", 0), 1, C.c_double(kw["ftol"]), C.c_double(kw["xtol"]),
                             C.c_double(kw["gtol"]), dp(out))
    assert k == n
    # state after trial j is what the oracle used for trial j+1
    for j in range(n - 1):
        assert out[j, 0] == tr[j + 1]["lam"], (j, out[j, 0], tr[j + 1]["lam"])
        assert out[j, 2] == tr[j + 1]["alpha"], (j, out[j, 2], tr[j + 1]["alpha"])
        assert int(out[j, 4]) == (1 if tr[j + 1]["it"] != tr[j]["it"] else 0)
        assert int(out[j, 5]) == 0
    assert int(out[n - 1, 9]) == rep.n_accepted and int(out[n - 1, 7]) == rep.iterations
    assert int(out[n - 1, 5]) == 1 and int(out[n - 1, 6]) == rep.status
    assert int(out[n - 1, 3]) == rep.n_accepted % 2
