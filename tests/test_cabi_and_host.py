"""CPU: the C-ABI library loads and exports every symbol include/lba.h declares; host-side logic
(index expansion, sharding, scene generator, drop-in argument handling).  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_names


def _ensure_built():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "amuya_sfm_lines_beta_b200", "liblba.so")):
        g.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    from amuya_sfm_lines_beta_b200 import _lib
    header = open(os.path.join(ROOT, "include", "lba.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(lba_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/lba.h but not exported"
    assert declared == set(_lib.EXPORTS)
    assert _lib.load().lba_version() >= 100


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors of lba_options / lba_report have the layout a C compiler gives include/lba.h."""
    _ensure_built()
    import shutil
    import subprocess
    from amuya_sfm_lines_beta_b200 import _lib
    lib = _lib.load()
    o = _lib.LbaOptions()
    lib.lba_default_options(ctypes.byref(o))
    assert (o.max_iter, o.pcg_max_iter) == (50, 200)
    assert (o.ftol, o.xtol, o.gtol) == (1e-4, 1e-8, 1e-8)   # reference call site :319 + scipy defaults
    assert o.lambda0 == 1e-2 and o.lambda_min == 1e-6 and o.pcg_tol == 1e-2 and o.min_ratio == 1e-4
    assert (o.loss, o.flags, o.f_scale, o.max_evals, o.max_backtrack) == (0, 0, 1.0, 0, 1)
    if shutil.which("gcc") is None:
        pytest.skip("gcc is not available")
    fields_o = [f[0] for f in _lib.LbaOptions._fields_]
    fields_r = [f[0] for f in _lib.LbaReport._fields_]
    src = "#include <stdio.h>\n#include <stddef.h>\n#include \"lba.h\"\nint main(void){\n"
    src += 'printf("%zu %zu\\n", sizeof(lba_options), sizeof(lba_report));\n'
    for f in fields_o:
        src += f'printf("o {f} %zu\\n", offsetof(lba_options, {f}));\n'
    for f in fields_r:
        src += f'printf("r {f} %zu\\n", offsetof(lba_report, {f}));\n'
    src += "return 0;}\n"
    (tmp_path / "lay.c").write_text(src)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "lay"), str(tmp_path / "lay.c")], check=True)
    out = subprocess.run([str(tmp_path / "lay")], capture_output=True, text=True, check=True).stdout.splitlines()
    so, sr = (int(x) for x in out[0].split())
    assert ctypes.sizeof(_lib.LbaOptions) == so and ctypes.sizeof(_lib.LbaReport) == sr
    for line in out[1:]:
        which, name, off = line.split()
        cls = _lib.LbaOptions if which == "o" else _lib.LbaReport
        assert getattr(cls, name).offset == int(off), line


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _ensure_built()
    from amuya_sfm_lines_beta_b200 import LineBA
    from amuya_sfm_lines_beta_b200._lib import LbaError
    with pytest.raises(LbaError, match="no CPU fallback"):
        LineBA(0)


def test_product_never_imports_the_oracle():
    """The product package neither imports oracle/ nor touches /root/reference: the path may appear in comments and
    docstrings (citations) only -- never in code or in a string the code uses."""
    import ast
    pkg = os.path.join(ROOT, "amuya_sfm_lines_beta_b200")
    checked = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                src = open(path).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                tree = ast.parse(src)
                doc_nodes = set()
                for node in ast.walk(tree):
                    if isinstance(node, (ast.Module, ast.ClassDef, ast.FunctionDef, ast.AsyncFunctionDef)) and node.body and \
                            isinstance(node.body[0], ast.Expr) and isinstance(node.body[0].value, ast.Constant):
                        doc_nodes.add(id(node.body[0].value))
                for node in ast.walk(tree):
                    if isinstance(node, ast.Constant) and isinstance(node.value, str) and id(node) not in doc_nodes:
                        assert "/root/reference" not in node.value and "oracle" not in node.value.split("/"), (f, node.value[:80])
                checked += 1
            elif f.endswith((".cu", ".cuh", ".h")):
                src = open(path).read()
                code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
                code = re.sub(r"//[^\n]*", "", code)
                assert "/root/reference" not in code and "oracle" not in code, f
                checked += 1
    assert checked >= 8


class _P:  # minimal stand-ins for the reference's Point/Line/Line3D attribute bags (src/point.py, src/line.py)
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _objects_from_golden(q):
    lines_2d = [_P(n=q["n"][i], points=[_P(p=q["p0"][i]), _P(p=q["p1"][i])]) for i in range(q["n"].shape[0])]
    pts = q["points_3d_raw"]
    lines_3d = [_P(points=[_P(coord=pts[2 * L]), _P(coord=pts[2 * L + 1])]) for L in range(pts.shape[0] // 2)]
    return lines_3d, lines_2d


@pytest.mark.parametrize("name", golden_names())
def test_point_wise_input_bit_exact(name):
    from amuya_sfm_lines_beta_b200.bundle_adjustment import point_wise_input
    q = np.load(os.path.join(GOLDEN, name + ".npz"))
    lines_3d, lines_2d = _objects_from_golden(q)
    pts, p2d, cam_pt, pt_idx = point_wise_input(lines_3d, lines_2d, q["camera_indices"], q["lines_indices"])
    assert pts.dtype == np.dtype(str(q["ref_points_3d_dtype"])) and np.array_equal(pts, q["ref_points_3d"])
    assert np.array_equal(cam_pt, q["ref_camera_indices_pt"]) and cam_pt.dtype == q["ref_camera_indices_pt"].dtype
    assert np.array_equal(pt_idx, q["ref_pt_indices"]) and pt_idx.dtype == q["ref_pt_indices"].dtype
    assert p2d.shape == (2 * len(lines_2d), 3)
    assert np.array_equal(p2d[0::2], q["p0"]) and np.array_equal(p2d[1::2], q["p1"])


def test_sparsity_as_shipped():
    from amuya_sfm_lines_beta_b200.bundle_adjustment import bundle_adjustment_sparsity
    q = np.load(os.path.join(GOLDEN, "toy1_lv2.npz"))
    nc = q["camera_params"].shape[0]; npts = q["points_3d_raw"].shape[0]; m = q["n"].shape[0]
    A = bundle_adjustment_sparsity(nc, npts, 6 * nc + 3 * npts, m, q["camera_indices"], q["lines_indices"]).tocsr()
    A.sort_indices()
    assert np.array_equal(A.indptr, q["ref_sparsity_indptr"]) and np.array_equal(A.indices, q["ref_sparsity_indices"])


def test_loss_argument_errors():
    from amuya_sfm_lines_beta_b200.bundle_adjustment import run_bundle_adjustment
    with pytest.raises(ValueError, match="`loss` must be one of"):
        run_bundle_adjustment({}, "bogus")


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_plan_partitions_at_line_boundaries(world):
    from amuya_sfm_lines_beta_b200.sharding import plan_shards
    rng = np.random.default_rng(world)
    nl = 500
    k = rng.integers(0, 12, nl)            # some lines have no observation at all
    li = np.repeat(np.arange(nl), k)
    shards = plan_shards(li, nl, world)
    assert [s.line_start for s in shards][0] == 0 and shards[-1].line_stop == nl
    seen = np.concatenate([s.obs_index for s in shards])
    assert np.array_equal(np.sort(seen), np.arange(li.size))          # every observation exactly once
    for a, b in zip(shards[:-1], shards[1:]):
        assert a.line_stop == b.line_start
    for s in shards:
        ll = li[s.obs_index]
        assert ll.size == 0 or (ll.min() >= s.line_start and ll.max() < s.line_stop)
    if world > 1:
        sizes = np.array([s.obs_index.size for s in shards])
        assert sizes.max() - sizes.min() <= 2 * k.max() + 1


def test_shard_plan_unsorted_input():
    from amuya_sfm_lines_beta_b200.sharding import plan_shards
    rng = np.random.default_rng(0)
    li = rng.integers(0, 40, 300)
    shards = plan_shards(li, 40, 4)
    for s in shards:
        ll = li[s.obs_index]
        assert np.all(np.diff(ll) >= 0)
        assert ll.size == 0 or (ll.min() >= s.line_start and ll.max() < s.line_stop)


def test_synthetic_scene_shapes_and_order():
    from amuya_sfm_lines_beta_b200.scenes import make_synthetic_scene
    for variant in ("A", "B"):
        s = make_synthetic_scene(20, 100, 5, seed=1, variant=variant)
        assert s.n_obs == 500 and s.camera_params.shape == (20, 6) and s.points_3d.shape == (200, 3)
        assert s.camera_indices.dtype == np.int64 and s.lines_indices.dtype == np.int64
        assert np.all(np.diff(s.lines_indices) >= 0)
        same = np.diff(s.lines_indices) == 0
        assert np.all(np.diff(s.camera_indices)[same] > 0)     # distinct cameras, ascending within a line
        for v in (s.n, s.p0, s.p1):
            assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-12)
        assert np.allclose(np.einsum('ni,ni->n', s.n, s.p0), 0, atol=1e-12)
    s2 = make_synthetic_scene(20, 100, 5, seed=1, variant="B")
    assert np.array_equal(s.n, s2.n) and np.array_equal(s.camera_params, s2.camera_params)   # seeded
