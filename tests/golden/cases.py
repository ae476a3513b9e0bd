"""Parity cases: (network, seed points, seed states, extra constraints) with everything fixed.

Seed points/states are stored under tests/golden/seeds_<case>.npz (generated once by
make_seeds.py with fixed RNG seeds) so that every engine -- the reference CUDA build, the CPU
oracle and the B200 engine -- starts from byte-identical inputs on every machine.
"""
import os
import random

import numpy as np
import torch

from analyticmesh_b200 import zoo
from analyticmesh_b200.model import MLP
from analyticmesh_b200.netinfo import NetInfo
from analyticmesh_b200.initializers import dichotomy, states_of
from analyticmesh_b200.utils import get_boundary

HERE = os.path.dirname(os.path.abspath(__file__))
CONSTRAINTS_JITTER = 1e-8  # reference backend/main.py:381,427


def _skipnet(seed=0):
    """The four skip flavours of reference backend/test/test_onnx_io.py:148-154 except the skip
    into the output layer (which the reference silently drops)."""
    st = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        m = MLP(nodes=[3, 40, 40, 40, 40, 40, 1], arc_table=[[1, 0, 0], [1, 1, 1], [0], [1, 3, 2], [0]],
                arc_tm_shape=[[40, 3], [0, 0], [40, 40]], initialization='geometric', geometric_radius=0.5,
                enable_print=False)
    finally:
        torch.random.set_rng_state(st)
    return m


# name -> (builder, n_seeds, cube or None)
CASES = {
    "polytope": (zoo.polytope, 64, None),
    "chair": (zoo.chair, 256, None),
    "chair_cube": (zoo.chair, 128, ((-0.05, -0.3, -0.2), (0.25, 0.1, 0.2))),
    "skipnet": (_skipnet, 128, None),
    "sphere": (zoo.sphere, 256, None),
    "mlp4x128s": (lambda: zoo.sal(depth=4, width=128), 256, None),
    "mlp8x512s_cube": (lambda: zoo.sal(depth=8, width=512), 256, "auto"),
    "mlp3x256s_cube": (lambda: zoo.sal(depth=3, width=256), 64, "auto"),    # small, wide enough for the tcgen05 path
}


def extra_constraints(cube):
    if cube is None:
        return np.zeros((0, 3)), np.zeros((0,))
    w, b = get_boundary('cube', min_vert=cube[0], max_vert=cube[1])
    return w.double().numpy() + CONSTRAINTS_JITTER, b.double().numpy()


def _auto_cube(model, half=0.04):
    """A small cube centred on a surface point found by bisection along +x."""
    lo, hi = np.zeros(3), np.array([1.0, 0.0, 0.0])
    info = NetInfo.from_model(model)
    flo = info.forward(lo[None])[0][0]
    for _ in range(60):
        mid = (lo + hi) / 2
        if (info.forward(mid[None])[0][0] < 0) == (flo < 0):
            lo = mid
        else:
            hi = mid
    c = np.round((lo + hi) / 2, 3)
    return (tuple(c - half), tuple(c + half))


def build_case(name, regenerate=False):
    builder, n_seeds, cube = CASES[name]
    model = builder()
    if cube == "auto":
        cube = _auto_cube(model)
    w_extra, b_extra = extra_constraints(cube)
    path = os.path.join(HERE, f"seeds_{name}.npz")
    L = sum(model.nodes[1:-1])
    if regenerate or not os.path.exists(path):
        g = torch.Generator().manual_seed(0)
        rng = random.Random(0)
        we = torch.from_numpy(w_extra).float() if len(b_extra) else None
        be = torch.from_numpy(b_extra).float() if len(b_extra) else None
        pts = dichotomy(model, 0.0, n_seeds, w_extra_constraints=we, b_extra_constraints=be, generator=g, rng=rng)
        states = states_of(model, pts).numpy()
        np.savez_compressed(path, points=pts.double().numpy(), states=np.packbits(states, axis=1, bitorder="little"))
    z = np.load(path)
    points = z["points"]
    states = np.unpackbits(z["states"], axis=1, bitorder="little")[:, :L].astype(bool)
    return dict(name=name, model=model, info=NetInfo.from_model(model), points=points, states=states,
                w_extra=w_extra, b_extra=b_extra, cube=cube)
