"""The tcgen05 split-integer composition path (AM_B200_GEMM_VARIANT=2, csrc/split.cuh).

(1) plane rows equal the exact-integer CPU restatement tests/split_emul.py bit for bit;
(2) whole marches give the region set / edge loops of the oracle and of the reference's golden vectors.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.golden.cases import build_case
from tests import parity, split_emul

pytestmark = pytest.mark.gpu
GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(autouse=True)
def _split_variant(monkeypatch):
    monkeypatch.setenv("AM_B200_GEMM_VARIANT", "2")
    yield
    from analyticmesh_b200 import cuam
    cuam.Destroy()          # the next test module must not inherit a split-variant environment


def _planes(case, n, sd, monkeypatch):
    from analyticmesh_b200 import cuam
    monkeypatch.setenv("AM_B200_SPLIT_DIGITS", str(sd))
    info = case["info"]
    cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
    cuam.load_weights(info.weights, info.biases, info.arc_tm)
    st = case["states"][:n]
    planes, equ = cuam.debug_planes(st, iso=0.125)
    return st, planes, equ


@pytest.mark.parametrize("name,sd", [("mlp4x128s", 7), ("skipnet", 7), ("chair_cube", 7), ("sphere", 7),
                                     ("mlp4x128s", 6), ("mlp4x128s", 8), ("mlp8x512s_cube", 7), ("mlp3x256s_cube", 7)])
def test_planes_equal_integer_restatement(name, sd, monkeypatch, oracle_lib):
    case = build_case(name)
    n = 37 if case["info"].state_len < 2000 else 19
    st, planes, equ = _planes(case, n, sd, monkeypatch)
    for i in range(st.shape[0]):
        want = split_emul.compose(case["info"], st[i], sd)
        bad = np.argwhere(planes[i] != want)
        assert bad.size == 0, (name, sd, i, len(bad), bad[:8].tolist(), planes[i][tuple(bad[0])], want[tuple(bad[0])])
    # the level plane is the FMA chain over these rows: close to the oracle's
    p, e = oracle_lib.compose(case["info"], st[0], iso=0.125)
    assert np.allclose(equ[0], e, rtol=0, atol=1e-11 * max(1.0, np.abs(e).max()))
    # ... and the rows themselves stay within the split's error bound of the oracle's FP64 FMA chain: operands are
    # truncated 8 SD - 2 bits below their row / column maximum and the digit products with i + j >= SD are dropped
    # (1.0e-15 / 5.7e-16 / 2.3e-13 of the layer's largest entry for SD = 7 / 8 / 6, compounded over depth)
    bound = {6: 5e-12, 7: 5e-14, 8: 5e-14}[sd]
    for i in range(min(4, st.shape[0])):
        p, _ = oracle_lib.compose(case["info"], st[i], iso=0.125)
        assert np.abs(planes[i] - p).max() <= bound * max(1.0, np.abs(p).max()), (name, sd, i, np.abs(planes[i] - p).max())


@pytest.mark.parametrize("name,sd", [("mlp3x256s_cube", 7), ("mlp4x128s", 6), ("skipnet", 7)])
def test_pipeline_variants_of_the_split_gemm_are_bit_identical(name, sd, monkeypatch):
    """Stage depth (64-byte K slabs with SWIZZLE_64B boxes, the default, vs the 32-byte slabs of rounds 1-2) and the number
    of epilogue warps per TMEM lane quarter only change how the operands travel and who drains the accumulators: every
    output is produced by the same instruction sequence, so the rows must agree bit for bit (tools/epi_check.py does the
    same on the full 8x512 march; profiles/r02_split_bk64.md, r02_split_epilogue_warps.md)."""
    case = build_case(name)
    rows = {}
    for label, env in (("bk64", {"AM_B200_SPLIT_BK": "64"}), ("bk32", {"AM_B200_SPLIT_BK": "32"}),
                       ("bk32_epi2", {"AM_B200_SPLIT_BK": "32", "AM_B200_SPLIT_EPI": "2"}),
                       ("bk32_epi4", {"AM_B200_SPLIT_BK": "32", "AM_B200_SPLIT_EPI": "4"})):
        for k in ("AM_B200_SPLIT_BK", "AM_B200_SPLIT_EPI"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        _, planes, equ = _planes(case, 37, sd, monkeypatch)
        rows[label] = (np.ascontiguousarray(planes).copy(), np.ascontiguousarray(equ).copy())
    ref_p, ref_e = rows["bk32"]
    assert np.isfinite(ref_p).all() and np.abs(ref_p).max() > 0
    for label, (p, e) in rows.items():
        assert np.array_equal(p.view(np.uint64), ref_p.view(np.uint64)), label
        assert np.array_equal(e.view(np.uint64), ref_e.view(np.uint64)), label


@pytest.mark.parametrize("name", ["skipnet", "chair", "mlp4x128s", "mlp3x256s_cube"])
def test_region_set_and_loops_match_oracle(oracle_lib, name):
    case = build_case(name)
    eng = parity.run_engine(case)
    orc = oracle_lib.march(case["info"], case["states"], case["points"], case["w_extra"], case["b_extra"])
    rep = parity.compare_with_oracle(eng, orc, case["info"].state_len)
    assert rep["keys_equal"] and rep["loops_equal"], rep
    assert rep["max_vertex_err"] < 1e-9, rep          # north_star tolerance: 1e-5 relative
    st = eng["stats"]
    assert st["n_overflow"] == 0 and st["n_inconsistent"] == 0 and st["n_stitch_miss"] == 0, st


@pytest.mark.parametrize("fname", ["ref_sphere.json", "ref_mlp8x512s_cube.json", "ref_chair.json"])
def test_against_reference_golden(fname):
    g = json.load(open(os.path.join(GOLD_DIR, fname)))
    case = build_case(g["case"])
    eng = parity.run_engine(case)
    ef = {k: v for k, v in parity.engine_faces(eng, case["info"].state_len).items() if v is not None}
    assert len(ef) == g["n_faces"]
    assert hashlib.sha256(b"".join(sorted(ef))).hexdigest() == g["keys_sha256"]
    v, _, _ = eng["mesh"]
    assert np.abs(case["info"].forward(v)[0]).max() <= max(g["max_abs_f"], 1e-12) + 1e-12


def test_read_through_of_parent_rows_equals_copying(monkeypatch):
    """Children read the rows they inherit straight from the parent's level buffer (clip_kernel materialises
    them on the way); AM_B200_READ_THROUGH=0 restores copy_parent_rows_kernel: same bytes out."""
    case = build_case("mlp8x512s_cube")
    a = parity.run_engine(case, combine=False)
    monkeypatch.setenv("AM_B200_READ_THROUGH", "0")
    b = parity.run_engine(case, combine=False)
    assert np.array_equal(a["keys"], b["keys"]) and np.array_equal(a["edges"], b["edges"])
    assert np.array_equal(a["xyz"], b["xyz"]) and np.array_equal(a["face_off"], b["face_off"])
    monkeypatch.setenv("AM_B200_INCREMENTAL", "0")
    c = parity.run_engine(case, combine=False)
    assert np.array_equal(a["keys"], c["keys"]) and np.array_equal(a["xyz"], c["xyz"])
