"""GPU tests of the public API: AnalyticMarching() (reference backend/main.py:335-559), voxel mode, and the
pybind11 `cuam` module with the reference's call sequence (backend/main.py:441-471)."""
import os

import numpy as np
import pytest
import torch

from tests.golden.cases import build_case
from analyticmesh_b200 import zoo
from analyticmesh_b200.polymesh import PolyMesh, load_ply_header

pytestmark = pytest.mark.gpu


def test_analytic_marching_api_chair(tmp_path):
    from analyticmesh_b200 import AnalyticMarching
    ply = str(tmp_path / "chair.ply")
    cfg = {'method': 'dichotomy', 'args': {'init_num': 512, 'try_pts_num': 4096, 'init_ball_radius': 1.0,
                                           'iter_max': 100, 'avg_eps': 1e-3, 'time_out': 60,
                                           'provided_surfpts': None, 'provided_surfstd': None}}
    ret = AnalyticMarching(zoo.chair(), ply, init_configs=cfg, seed=0)
    assert {'init_point_time', 'init_cuda_time', 'am_time', 'export_time'} <= set(ret)
    assert 'w_extra_constraints' not in cfg['args']            # the caller's dict is not mutated (App. B-14)
    head = load_ply_header(ply)
    assert head['face_num'] == 248228 and head['vertex_num'] == 248228      # SURVEY App. D known answer
    # second call re-uses the environment (reference main.py:437-449)
    ret2 = AnalyticMarching(zoo.chair(), ply, init_configs=cfg, seed=1, save_polymesh=False, scale=2.0,
                            center=[1.0, 0.0, 0.0])
    assert ret2['init_cuda_time'] == 0.0
    m = PolyMesh(ply)
    assert not m.is_polymesh() and m.num_polyfaces() == 2 * 248228
    v = np.asarray(m.vertices())
    assert v[:, 0].mean() > 0.9                                 # scaled and translated


def test_voxel_mode_covers_the_surface(tmp_path):
    from analyticmesh_b200 import AnalyticMarching
    ply = str(tmp_path / "vox.ply")
    ret = AnalyticMarching(zoo.chair(), ply, voxel_configs={'voxel_size': 0.5}, seed=0)
    assert ret['am_time'] > 0
    m = PolyMesh(ply)
    assert m.num_polyfaces() >= 248228                          # faces cut by voxel walls are split
    v = np.asarray(m.vertices(), dtype=np.float64)
    info = build_case("chair")["info"]
    assert np.abs(info.forward(v)[0]).max() < 1e-5              # float32 vertices on the surface


def test_pybind_module_matches_ctypes_binding(tmp_path):
    """the reference's five-call sequence through analyticmesh_b200/build/cuam*.so with torch CUDA tensors"""
    from analyticmesh_b200.build import cuam as pyb
    from tests import parity
    case = build_case("chair_cube")
    info = case["info"]
    dev = torch.device("cuda")
    t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev).contiguous()  # noqa: E731
    pyb.Init(float_type="float64", nodesnum=info.nodes, arc_table=torch.from_numpy(info.arc_table),
             num_extra_constraints=len(case["b_extra"]))
    pyb.AnalyticMarching(weights=[t(w) for w in info.weights], biases=[t(b) for b in info.biases],
                         states=torch.from_numpy(case["states"]).to(dev), points=t(case["points"]), arc_tm=[],
                         w_extra_constraints=t(case["w_extra"]).reshape(-1, 3), b_extra_constraints=t(case["b_extra"]),
                         iso=0.0, flip_insideout=False)
    pyb.CombineMesh(scale=1.0, center=[0.0, 0.0, 0.0])
    a = str(tmp_path / "a.ply")
    pyb.ExportMesh(file_path=a, is_polymesh=True, is_float32=False)
    st = pyb.Stats()
    pyb.Destroy()
    from analyticmesh_b200 import cuam
    parity.run_engine(case)
    b = str(tmp_path / "b.ply")
    cuam.ExportMesh(file_path=b, is_polymesh=True, is_float32=False)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert st["n_faces"] == 685
    with pytest.raises(RuntimeError):
        pyb.Init(float_type="float64", nodesnum=[2, 5, 1], arc_table=torch.zeros(1, 1, dtype=torch.int32),
                 num_extra_constraints=0)


def test_batch_of_latent_shapes_reuses_the_environment(oracle_lib):
    """BASELINE.json config 5 in small: shapes that share every weight but biases[0] are marched one after the
    other in ONE environment; each equals the oracle on its own network."""
    import random
    import torch
    from analyticmesh_b200 import zoo, cuam
    from analyticmesh_b200.netinfo import NetInfo
    from analyticmesh_b200.initializers import dichotomy, states_of
    from tests import parity
    model = zoo.sal(depth=3, width=32, skip=False, seed=1)
    biases0 = zoo.latent_shapes(model, 3, latent_dim=16, sigma=0.05)
    assert zoo.shapes_of_rank(7, 1, 3) == [1, 4] and sorted(sum((zoo.shapes_of_rank(7, r, 3) for r in range(3)), [])) == list(range(7))
    counts = []
    for k, b0 in enumerate(biases0):
        with torch.no_grad():
            model.linears[0].bias.copy_(b0)
        pts = dichotomy(model, 0.0, 64, generator=torch.Generator().manual_seed(k), rng=random.Random(k))
        info = NetInfo.from_model(model)
        case = dict(info=info, states=states_of(model, pts).numpy(), points=pts.double().numpy(),
                    w_extra=np.zeros((0, 3)), b_extra=np.zeros(0))
        if k == 0:
            eng = parity.run_engine(case, combine=False)          # Init happens here, once
        else:
            cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=case["states"], points=case["points"],
                                  arc_tm=info.arc_tm, w_extra_constraints=np.zeros((0, 3)),
                                  b_extra_constraints=np.zeros(0), iso=0.0, flip_insideout=False)
            keys, face_off, parent, via = cuam.states()
            edges, xyz = cuam.faces()
            eng = dict(keys=keys, face_off=face_off, parent=parent, via=via, edges=edges, xyz=xyz, stats=cuam.stats())
        orc = oracle_lib.march(info, case["states"], case["points"], None, None)
        rep = parity.compare_with_oracle(eng, orc, info.state_len)
        assert rep["keys_equal"] and rep["loops_equal"] and rep["max_vertex_err"] < 1e-9, (k, rep)
        counts.append(eng["stats"]["n_faces"])
    assert len(set(counts)) > 1 or counts[0] > 0


def test_native_dichotomy_seeds_on_the_device(oracle_lib):
    """the surface-point initialiser as device code (csrc/seeds.cuh, am_seed_dichotomy) against the
    behaviour of reference backend/main.py:252-326 -- init_num points inside the ball and the extra constraints, mean
    |f - iso| below avg_eps, activation patterns equal to a float64 forward pass, deterministic in the seed -- and the
    march from the stored seeds equals the march from the same seeds handed over as (states, points)."""
    from analyticmesh_b200 import cuam
    for name in ("chair", "skipnet", "chair_cube", "mlp3x256s_cube"):
        case = build_case(name)
        info = case["info"]
        we, be = np.ascontiguousarray(case["w_extra"]).reshape(-1, 3), np.ascontiguousarray(case["b_extra"]).reshape(-1)
        cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=len(be))
        rep = cuam.seed_dichotomy(info.weights, info.biases, info.arc_tm, we, be, 0.0, init_num=300, try_pts_num=4096,
                                  init_ball_radius=1.0, iter_max=100, avg_eps=1e-3, seed=5)
        pts, st = cuam.seeds()
        assert pts.shape == (300, 3) and st.shape == (300, info.state_len) and rep["n_points"] == 300
        f, bits = info.forward(pts)
        assert np.abs(f).mean() < 1e-3 and abs(np.abs(f).mean() - rep["avg_abs_error"]) < 1e-9, (name, rep)
        assert int((bits != st).sum()) == 0, name
        assert (np.square(pts).sum(1) < 1.0).all()
        if len(be):
            assert ((pts @ we.T + be) < 0).all()
        cuam.seed_dichotomy(info.weights, info.biases, info.arc_tm, we, be, 0.0, init_num=300, seed=5)
        assert np.array_equal(cuam.seeds()[0], pts)                       # same seed, same points
        kw = dict(weights=info.weights, biases=info.biases, arc_tm=info.arc_tm, w_extra_constraints=we,
                  b_extra_constraints=be, iso=0.0, flip_insideout=False)
        cuam.AnalyticMarching(states=None, points=None, **kw)             # from the seeds stored on the device
        d_stored, n_stored = cuam.digest(), cuam.stats()["n_faces"]
        cuam.AnalyticMarching(states=st, points=pts, **kw)
        assert cuam.digest()["raw"] == d_stored["raw"] and n_stored > 0
        cuam.seed_dichotomy(info.weights, info.biases, info.arc_tm, we, be, 0.0, init_num=300, seed=6)
        assert not np.array_equal(cuam.seeds()[0], pts)
        if name in ("chair", "skipnet"):        # the march from device-made seeds against the oracle from the same seeds
            orc = oracle_lib.march(info, st, pts)
            assert d_stored["topology_sum"] == oracle_lib.topology_sum(oracle_lib.canonical_faces(orc), info.state_len)
            assert n_stored == orc["n_faces"]
    cuam.Destroy()


def test_weight_cache_with_device_tensors():
    """Device-resident weights are compared by a checksum computed on the device (no copy to the host per march): the
    same tensors re-stage nothing, a changed bias re-stages exactly that tensor and changes the mesh, changing it back
    restores the first mesh bit for bit."""
    from analyticmesh_b200 import cuam
    case = build_case("mlp3x256s_cube")
    info = case["info"]
    dev = torch.device("cuda")
    W = [torch.from_numpy(w).to(dev) for w in info.weights]
    B = [torch.from_numpy(b).to(dev) for b in info.biases]
    TM = [torch.from_numpy(t).to(dev) for t in info.arc_tm]
    we = torch.from_numpy(np.ascontiguousarray(case["w_extra"])).reshape(-1, 3).to(dev)
    be = torch.from_numpy(np.ascontiguousarray(case["b_extra"])).reshape(-1).to(dev)
    kw = dict(weights=W, biases=B, arc_tm=TM, states=torch.from_numpy(case["states"]).to(dev),
              points=torch.from_numpy(case["points"]).to(dev), w_extra_constraints=we, b_extra_constraints=be, iso=0.0,
              flip_insideout=False)
    cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=len(be))
    cuam.AnalyticMarching(**kw)
    d0 = cuam.digest()["raw"]
    cuam.AnalyticMarching(**kw)
    assert cuam.stats()["n_tensors_reloaded"] == 0 and cuam.digest()["raw"] == d0
    B[1][3] += 1e-3                                        # one bias of the second layer, in place on the device
    cuam.AnalyticMarching(**kw)
    assert cuam.digest()["raw"] != d0                      # a changed bias is seen (biases are not counted as matrices)
    B[1][3] -= 1e-3
    W[2][0, 0] *= 1.0000001
    cuam.AnalyticMarching(**kw)
    assert cuam.stats()["n_tensors_reloaded"] == 1
    cuam.Destroy()
