"""CPU-only tests: ONNX codec, model container, initialiser, the oracle against known answers and the
reference-generated golden vectors, and the C-ABI library's exported symbols (no compute without a GPU)."""
import ctypes
import hashlib
import json
import os
import random
import re

import numpy as np
import pytest
import torch

from analyticmesh_b200 import MLP, load_model, save_model, zoo
from analyticmesh_b200.initializers import dichotomy, states_of
from analyticmesh_b200.netinfo import NetInfo
from tests.golden.cases import build_case, CASES

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
GOLD = os.path.join(ROOT, "tests", "golden")


# ---------------------------------------------------------------- ONNX codec / model -----------
def test_load_chair_onnx():
    m = load_model(os.path.join(GOLD, "chair.onnx"))
    assert m.nodes == [3, 60, 60, 60, 60, 1] and m.arc_table == [[0]] * 4 and m.arc_tm_shape == []
    assert load_model(open(os.path.join(GOLD, "chair.onnx"), "rb").read()).nodes == m.nodes
    with pytest.raises(Exception):
        load_model("/nonexistent.onnx")


def test_onnx_roundtrip_all_skip_kinds(tmp_path):
    """architecture of reference backend/test/test_onnx_io.py:148-154 (test_mlp_2)"""
    torch.manual_seed(1)
    m = MLP([3] + [80] * 5 + [1], [[1, 0, 0], [1, 1, 1], [0], [1, 3, 2], [1, 0, 3]],
            [[80, 3], [0, 0], [80, 80], [1, 3]], enable_print=False)
    p = str(tmp_path / "m.onnx")
    save_model(m, p)
    q = load_model(p)
    assert q.nodes == m.nodes and q.arc_table == m.arc_table and q.arc_tm_shape == m.arc_tm_shape
    x = torch.randn(17, 3)
    assert torch.equal(m(x), q(x))


def test_onnx_from_torch_exporter(tmp_path):
    """custom nn.Module like reference backend/test/test_onnx_io.py:99-141 (identity + linear skips)"""
    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.l0, self.l1 = torch.nn.Linear(3, 20), torch.nn.Linear(20, 50)
            self.l2, self.l3 = torch.nn.Linear(50, 20), torch.nn.Linear(20, 1)
            self.sk = torch.nn.Linear(20, 20, bias=False)

        def forward(self, x):
            y1 = torch.relu(self.l0(x))
            y2 = torch.relu(self.l1(y1))
            y3 = torch.relu(self.l2(y2) + y1 + self.sk(y1))
            return self.l3(y3)

    torch.manual_seed(0)
    net = Net()
    p = str(tmp_path / "n.onnx")
    save_model(net, p)
    q = load_model(p)
    assert q.nodes == [3, 20, 50, 20, 1]
    assert q.arc_table == [[0], [2, 1, 0, 1, 1], [0]] and q.arc_tm_shape == [[0, 0], [20, 20]]
    x = torch.randn(9, 3)
    assert torch.allclose(net(x), q(x), atol=1e-6)


def test_netinfo_forward_matches_torch():
    c = build_case("skipnet")
    x = np.random.RandomState(0).randn(50, 3) * 0.4
    f, bits = c["info"].forward(x)
    m = c["model"].double()
    ft = m(torch.from_numpy(x)).reshape(-1).detach().numpy()
    st = states_of(m, torch.from_numpy(x)).numpy()
    c["model"].float()
    assert np.allclose(f, ft, atol=1e-12) and np.array_equal(bits, st)


def test_dichotomy_reproducible_and_on_surface():
    m = zoo.sphere(width=64)
    a = dichotomy(m, 0.0, 200, generator=torch.Generator().manual_seed(3), rng=random.Random(3))
    b = dichotomy(m, 0.0, 200, generator=torch.Generator().manual_seed(3), rng=random.Random(3))
    assert torch.equal(a, b) and a.shape == (200, 3)
    assert float(m(a).abs().mean()) < 1e-3


def test_other_initialisers_find_the_surface():
    """`sphere_tracing`, `gradient_descent` and `dichotomy(provided_surfpts=...)` (reference
    backend/main.py:112-249, 263-264) with the reference's default argument dicts: points on the iso-surface
    (mean |f - iso| below avg_eps), inside the ball, honouring extra constraints; gradient_descent returns exactly
    init_num points."""
    from analyticmesh_b200.initializers import gradient_descent, sphere_tracing
    from analyticmesh_b200.main import GRADIENT_DESCENT_DICT, SPHERE_TRACING_DICT
    from analyticmesh_b200.utils import get_boundary
    m = zoo.sphere(width=64, depth=2)              # SDF-like: f ~ |x| - 0.5
    for p in m.parameters():
        p.requires_grad_(False)
    w_e, b_e = get_boundary('cube', min_vert=(-1.0, -1.0, 0.0), max_vert=(1.0, 1.0, 1.0))     # upper half space
    g = torch.Generator().manual_seed(3)
    a = dict(SPHERE_TRACING_DICT['args'], init_num=256, w_extra_constraints=w_e, b_extra_constraints=b_e, generator=g)
    pts = sphere_tracing(m, 0.0, **a)
    assert 0 < pts.shape[0] <= 256 and pts.shape[1] == 3
    assert float(m(pts).abs().mean()) < 2e-3 and float(pts[:, 2].min()) > 0.0 and float((pts ** 2).sum(1).max()) < 1.0
    a = dict(GRADIENT_DESCENT_DICT['args'], init_num=128, w_extra_constraints=w_e, b_extra_constraints=b_e, generator=g)
    pts = gradient_descent(m, 0.05, **a)
    assert pts.shape == (128, 3)
    assert float((m(pts) - 0.05).abs().mean()) < 5e-3 and float(pts[:, 2].min()) > 0.0
    # dichotomy around provided surface points: the trial points are the given points + N(0, std)
    seeds = dichotomy(m, 0.0, 64, generator=torch.Generator().manual_seed(1), rng=random.Random(1))
    near = dichotomy(m, 0.0, 200, provided_surfpts=seeds, provided_surfstd=0.02,
                     generator=torch.Generator().manual_seed(2), rng=random.Random(2))
    assert near.shape == (200, 3) and float(m(near).abs().mean()) < 1e-3
    d = torch.cdist(near, seeds).min(dim=1).values
    assert float(d.max()) < 0.2                                # they stay in the neighbourhood of the given points
    # a level set that lies outside the ball: every traced point is filtered out (the reference does the same)
    far = sphere_tracing(m, 5.0, **dict(SPHERE_TRACING_DICT['args'], init_num=32, time_out=20, w_extra_constraints=None,
                                        b_extra_constraints=None, generator=g))
    assert far.shape[0] == 0


def test_seed_fixtures_are_current():
    """the committed seeds are what make_seeds.py generates (guards against silent drift)"""
    for name in ("polytope", "chair_cube"):
        c = build_case(name)
        f, bits = c["info"].forward(c["points"])
        assert np.abs(f).mean() < 2e-3
        assert c["states"].shape == (CASES[name][1], c["info"].state_len)


# ---------------------------------------------------------------- oracle ----------------------
def _canon(orc, res):
    return orc.canonical_faces(res)


def test_oracle_chair_known_answer(oracle_lib):
    """SURVEY App. D (two independent float64 brute-force traversals): 248 228 faces, this histogram."""
    c = build_case("chair")
    r = oracle_lib.march(c["info"], c["states"], c["points"])
    assert r["n_faces"] == 248228
    cf = _canon(oracle_lib, r)
    sizes = {}
    for v in cf.values():
        if v is not None:
            sizes[len(v[0])] = sizes.get(len(v[0]), 0) + 1
    assert sizes == {3: 86083, 4: 97749, 5: 46744, 6: 14193, 7: 2963, 8: 427, 9: 59, 10: 9, 11: 1}
    keys = sorted(k for k, v in cf.items() if v is not None)
    assert len(set(keys)) == 248228


def test_oracle_self_consistency(oracle_lib):
    """closed surface: every neuron edge shared by exactly two faces, |f(v)| tiny, centroid pattern = key"""
    c = build_case("skipnet")
    info = c["info"]
    r = oracle_lib.march(info, c["states"], c["points"])
    cf = _canon(oracle_lib, r)
    L = info.state_len
    inc = {}
    verts, cents, keys = [], [], []
    for k, v in cf.items():
        if v is None:
            continue
        for e in v[0]:
            kb = bytearray(k)
            kb[e >> 3] &= ~(1 << (e & 7)) & 0xFF
            inc[(bytes(kb), e)] = inc.get((bytes(kb), e), 0) + 1
        verts.append(v[1])
        cents.append(v[1].mean(0))
        keys.append(k)
    assert set(inc.values()) == {2}
    assert np.abs(info.forward(np.concatenate(verts))[0]).max() < 1e-10
    _, bits = info.forward(np.stack(cents))
    got = [b.tobytes() for b in np.packbits(bits, axis=1, bitorder="little")]
    assert got == keys


def test_oracle_extra_constraints_and_seed_dedup(oracle_lib):
    c = build_case("chair_cube")
    r = oracle_lib.march(c["info"], c["states"], c["points"], c["w_extra"], c["b_extra"])
    cf = _canon(oracle_lib, r)
    L = c["info"].state_len
    lo, hi = np.array(c["cube"][0]), np.array(c["cube"][1])
    n_cut = 0
    for v in cf.values():
        if v is None:
            continue
        assert (v[1] >= lo - 1e-6).all() and (v[1] <= hi + 1e-6).all()
        n_cut += any(e >= L for e in v[0])
    assert n_cut > 10
    twice = oracle_lib.march(c["info"], np.concatenate([c["states"]] * 2), np.concatenate([c["points"]] * 2),
                             c["w_extra"], c["b_extra"])
    assert twice["n_states"] == r["n_states"]


def test_oracle_per_state_entry_reproduces_the_march(oracle_lib):
    """process_states() (used to check a sample of a march that is too large for the CPU) gives, state by state,
    what the whole march gave: start from any point of the face (its centroid) without a start edge, or from the
    midpoint of one of its edges with that edge as the start edge."""
    c = build_case("chair_cube")
    info = c["info"]
    r = oracle_lib.march(info, c["states"], c["points"], c["w_extra"], c["b_extra"])
    cf = _canon(oracle_lib, r)
    L = info.state_len
    keys, mids, starts, want = [], [], [], []
    for i, (k, v) in enumerate(cf.items()):
        if v is None:
            continue
        kw = np.frombuffer(k + b"\0" * (4 * ((L + 31) // 32) - len(k)), dtype="<u4")
        j = i % len(v[0])
        keys += [kw, kw]
        mids += [v[1].mean(0), 0.5 * (v[1][j] + v[1][(j + 1) % len(v[0])])]
        starts += [-1, v[0][j]]
        want += [v, v]
    got = oracle_lib.process_states(info, np.stack(keys), np.stack(mids), starts, c["w_extra"], c["b_extra"])
    gf = [None] * len(want)
    assert got["n_states"] == len(want) > 1000
    for o in range(len(want)):
        one = {kk: got[kk][o:o + 1] for kk in ("keys", "tab", "verts")}
        one.update(n_states=1, state_len=L)
        (val,) = oracle_lib.canonical_faces(one).values()
        assert val is not None and val[0] == want[o][0], (o, val, want[o][0])
        assert np.abs(val[1] - want[o][1]).max() < 1e-9
    # the order-independent checksum the engine computes on the device (cuam.digest) is order independent
    items = list(cf.items())
    assert oracle_lib.topology_sum(dict(items), L) == oracle_lib.topology_sum(dict(reversed(items)), L)
    items[0] = (items[0][0], None)
    assert oracle_lib.topology_sum(dict(items), L) != oracle_lib.topology_sum(cf, L)


def test_oracle_float32_variant_runs(oracle_lib):
    c = build_case("chair_cube")
    info32 = NetInfo.from_model(c["model"], dtype=np.float32)
    r = oracle_lib.march(info32, c["states"], c["points"], c["w_extra"], c["b_extra"])
    assert r["n_faces"] > 500


GOLDEN_REF = sorted(f for f in os.listdir(GOLD) if f.startswith("ref_") and f.endswith(".json"))


@pytest.mark.parametrize("fname", GOLDEN_REF or ["<none>"])
def test_oracle_against_reference_golden(oracle_lib, fname):
    """Pins the oracle to outputs of the reference's own CUDA build (tests/golden/make_golden_ref.py)."""
    if fname == "<none>":
        pytest.skip("no reference-generated golden vectors committed (see DESIGN.md: parity pinning)")
    g = json.load(open(os.path.join(GOLD, fname)))
    c = build_case(g["case"])
    info = c["info"]
    wsha = hashlib.sha256(b"".join(np.ascontiguousarray(w).tobytes() for w in info.weights)).hexdigest()
    assert wsha == g["weights_sha256"], "weights differ between this machine and the golden run"
    r = oracle_lib.march(info, c["states"], c["points"], c["w_extra"], c["b_extra"])
    cf = {k: v for k, v in _canon(oracle_lib, r).items() if v is not None}
    assert len(cf) == g["n_faces"] == g["n_unique_keys"]
    digest = hashlib.sha256(b"".join(sorted(cf))).hexdigest()
    assert digest == g["keys_sha256"]
    for f in g["faces"]:
        k = bytes.fromhex(f["key"])
        ref_v = np.asarray(f["verts"])
        ours = cf[k][1]
        assert len(ours) == len(ref_v)
        # same cyclic loop (any rotation), vertices within 1e-5 relative (north_star tolerance)
        d = np.abs(ours[None, :, :] - ref_v[:, None, :]).max(-1)
        start = int(np.argmin(d[0]))
        rolled = np.roll(ours, -start, axis=0)
        scale = max(1.0, float(np.abs(ref_v).max()))
        assert np.abs(rolled - ref_v).max() / scale < 1e-5


# ---------------------------------------------------------------- C ABI -----------------------
def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "am_b200.h")).read()
    declared = set(re.findall(r"\b(am_[a-z0-9_]+)\s*\(", header))
    assert {"am_create", "am_march", "am_combine", "am_export", "am_destroy"} <= declared
    lib_path = os.path.join(ROOT, "analyticmesh_b200", "libam_b200.so")
    assert os.path.exists(lib_path), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(lib_path)
    for sym in declared:
        assert hasattr(lib, sym), sym
    from analyticmesh_b200 import cuam
    assert set(cuam.EXPORTS) == declared


def test_cabi_rejects_malformed_architecture_without_gpu():
    lib = ctypes.CDLL(os.path.join(ROOT, "analyticmesh_b200", "libam_b200.so"))
    lib.am_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    nodes = (ctypes.c_int * 3)(2, 14, 1)          # input width must be 3
    arc = (ctypes.c_int * 1)(0)
    assert lib.am_create(ctypes.byref(h), 1, nodes, 3, arc, 1, 1, 0) == -1
    assert b"malformed" in lib.am_last_error(None)
    nodes = (ctypes.c_int * 3)(3, 14, 1)
    assert lib.am_create(ctypes.byref(h), 1, nodes, 3, arc, 2, 1, 0) == -1     # wrong number of arc rows


def test_product_path_does_not_touch_the_oracle():
    """the product package must never import / call oracle/ (it is the checker, not a fallback)"""
    pkg = os.path.join(ROOT, "analyticmesh_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "am_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


# ---------------------------------------------------------------- multi-process host logic ------
def _gloo_worker(rank, world, port, q):
    import ctypes as ct
    import torch.distributed as dist
    from analyticmesh_b200.parallel import make_allreduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the level's polygon scratch: every slot is filled by exactly one rank (owner = slot % world)
    n = 4096
    buf = np.zeros(n, dtype=np.int32)
    mine = np.arange(n) % world == rank
    buf[mine] = (np.arange(n, dtype=np.int64)[mine] * 2654435761 % 2**31).astype(np.int32)
    make_allreduce(device="cpu")(buf.ctypes.data, n)
    expect = (np.arange(n, dtype=np.int64) * 2654435761 % 2**31).astype(np.int32)
    # the handle swap of the peer-memory exchange (cuam.set_shard_p2p): 64-byte payloads, rank order on every rank
    from analyticmesh_b200.parallel import make_allgather
    parts = make_allgather(device="cpu")(bytes([rank + 1]) * 64)
    gathered = parts == [bytes([r + 1]) * 64 for r in range(world)]
    # hash ownership of the sharded visited set (csrc/frontier.cuh key_hash_owner): a partition of the key space
    hs = np.random.RandomState(7).randint(0, 2**63, 10000).astype(np.uint64) * np.uint64(2) + np.uint64(rank)
    own = ((hs >> np.uint64(32)) * np.uint64(world)) >> np.uint64(32)
    balanced = bool(own.min() == 0 and own.max() == world - 1 and abs(float((own == 0).mean()) - 1.0 / world) < 0.05)
    q.put((rank, bool(np.array_equal(buf, expect)) and gathered and balanced))
    dist.destroy_process_group()


def test_allreduce_union_world2_gloo():
    """N>1 host path (analyticmesh_b200/parallel.py) with world_size 2 on CPU: the integer all-reduce of
    disjointly filled buffers is the exact union on every rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def test_native_host_selftest(tmp_path):
    """The arithmetic helpers of the CUDA sources that run on the host as well (nvcc-compiled, executed on the CPU):
    pairing bijection + Philox of the device initialiser, digit split of the tcgen05 path, exchange-block layout, hash
    ownership and the chain deal of the sharded march (tests/native/host_selftest.cu)."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    exe = str(tmp_path / "host_selftest")
    src = os.path.join(ROOT, "tests", "native", "host_selftest.cu")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src], check=True,
                   capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host selftest ok" in r.stdout, r.stdout + r.stderr


def test_seed_owner_rule():
    from analyticmesh_b200.parallel import owner_of_seeds
    assert owner_of_seeds(5, 2).tolist() == [0, 1, 0, 1, 0]


# ---------------------------------------------------------------- polygon-mesh files ------------
def test_polymesh_roundtrip_and_poly2tri(tmp_path):
    """behaviour of reference backend/libpolytools (polylib.cpp:134-268, 349-393, 503-569)"""
    from analyticmesh_b200.polymesh import PolyMesh, get_faces_num, load_ply_header, poly2tri
    verts = [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 1]]
    faces = [[0, 1, 2, 3], [0, 1, 4], [1, 2, 4, 4]]
    m = PolyMesh(vertices=verts, faces=faces, colors=[[255, 0, 0], [0, 255, 0], [0, 0, 255]])
    ply, off, tri = str(tmp_path / "a.ply"), str(tmp_path / "a.off"), str(tmp_path / "t.ply")
    m.save(ply)
    n = PolyMesh(ply)
    assert n.faces() == faces and n.colors() == [[255, 0, 0], [0, 255, 0], [0, 0, 255]]
    assert np.allclose(n.vertices(), verts)
    assert n.is_polymesh() and n.num_polyfaces() == 3 and n.num_trifaces() == 5
    assert load_ply_header(ply) == {'storing_type': 'binary_little_endian', 'storing_version': 1.0, 'vertex_num': 5,
                                    'face_num': 3}
    assert get_faces_num(ply) == {'poly': 3, 'tri': 5}
    poly2tri(ply, tri)
    t = PolyMesh(tri)
    assert t.faces() == [[0, 1, 2], [0, 2, 3], [0, 1, 4], [1, 2, 4]]       # degenerate (1,4,4) dropped
    assert t.colors() == [[255, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255]] and not t.is_polymesh()
    m.save(off)
    o = PolyMesh(off)
    assert o.faces() == faces and np.allclose(o.vertices(), verts)
    with pytest.raises(RuntimeError):
        PolyMesh(str(tmp_path / "x.stl"))


# ---------------------------------------------------------------- split-integer restatement ------
def test_split_digits_reconstruct_exactly():
    """tests/split_emul.py (the CPU restatement of csrc/split.cuh): digits are int8 and rebuild the scaled integer."""
    from tests import split_emul
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, 64)) * np.exp(rng.uniform(-20, 20, (5, 1)))
    x[1, 3] = 0.0
    for sd in (6, 7, 8):
        e = split_emul._exponent(np.abs(x).max(axis=1))[:, None]
        d = split_emul.digits(x, e, sd)
        assert d.min() >= -128 and d.max() <= 127
        X = sum(d[t].astype(object) * (256 ** (sd - 1 - t)) for t in range(sd))
        want = np.rint(np.ldexp(x, (8 * sd - 2) - e)).astype(np.int64).astype(object)
        assert (X == want).all()
        back = np.ldexp(np.array(X, dtype=np.float64), e - (8 * sd - 2))
        assert np.abs(back - x).max() <= np.ldexp(np.abs(x).max(axis=1), -(8 * sd - 3)).max()


def test_split_gemm_restatement_matches_fp64_reference(oracle_lib):
    """The split-integer composition agrees with the FMA-chain oracle to a few ulp of the column maximum."""
    from tests import split_emul
    from tests.golden.cases import build_case
    case = build_case("skipnet")
    info = case["info"]
    for i in range(3):
        p, _ = oracle_lib.compose(info, case["states"][i], iso=0.0)
        q = split_emul.compose(info, case["states"][i], 7)
        scale = np.abs(p).max(axis=0)
        scale[scale == 0] = 1.0
        assert (np.abs(p - q) / scale).max() < 1e-14


def test_bench_roofline_objects_carry_the_contract_keys():
    """bench.py's roofline object for both composition paths: bound / achieved / peak / unit / frac / traffic, the
    fractions consistent with the peaks they name (no GPU needed: pure arithmetic on the committed peak files)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    r = bench.roofline(2, 7, 90.0, 37.0, 1928, 1300.0, 90.0e12 * 1.3)
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "launches", "avg_launch_ms"):
        assert k in r, k
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["digit_products"] == 28
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["int8_tops_achieved"] - 28 * r["achieved"]) < 1e-9
    if r["instruction_peak"] is not None:       # profiles/r02_int8_peak.json is committed
        assert abs(r["instruction_peak"]["frac_of_instruction_peak"] -
                   r["int8_tops_achieved"] / r["instruction_peak"]["int8_tops_instruction_peak"]) < 1e-12
        assert r["instruction_peak"]["frac_of_instruction_peak"] < r["frac"]
    d = bench.roofline(0, 7, 30.0, 37.0, 100, 50.0, 30.0e12 * 0.05)
    assert d["bound"] == "fp64" and abs(d["frac"] - 30.0 / 37.0) < 1e-12
    assert bench.flops_per_face([3] + [512] * 8 + [1]) == 14684160
