"""Builders for the BASELINE.json configs (deterministic: fixed torch seeds, CPU RNG).

polytope  reference examples/2_polytope.py:14-36 (3-14-1, hand weights + 1e-7 noise)
sphere    reference examples/1_sphere.py:8-9     (3-500-500-1, geometric init r = 0.5)
chair     reference examples/chair.onnx          (3-60-60-60-60-1)
sal       SURVEY 8(d) config 4: geometric-init depth x width MLP with a linear skip from the
          input into the middle hidden layer (8 x 512 by default)
"""
import math
import os

import torch

from .model import MLP
from .onnx_io import load_model

_POLY_W0 = [[1, 1, 1], [-1, -1, -1], [0, 1, 1], [0, -1, -1], [1, 0, 1], [-1, 0, -1], [1, 1, 0],
            [-1, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]


def polytope(seed=0, noise=1e-7):
    g = torch.Generator().manual_seed(seed)
    m = MLP(nodes=[3, 14, 1], arc_table=[[0]], arc_tm_shape=[], initialization=None, enable_print=False)

    def jitter(x):
        return x + torch.randn(x.shape, generator=g) * noise

    with torch.no_grad():
        m.linears[0].weight.copy_(jitter(torch.tensor(_POLY_W0, dtype=torch.float32)))
        m.linears[0].bias.copy_(jitter(torch.zeros(14)))
        m.linears[1].weight.copy_(jitter(torch.ones(1, 14)))
        m.linears[1].bias.copy_(jitter(torch.tensor([-2.0])))
    return m


def _geometric(nodes, arc_table, arc_tm_shape, radius, seed):
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        m = MLP(nodes=nodes, arc_table=arc_table, arc_tm_shape=arc_tm_shape, initialization='geometric',
                geometric_radius=radius, enable_print=False)
        with torch.no_grad():
            for tm, shape in zip(m.tms, arc_tm_shape):
                if shape[0]:
                    torch.nn.init.normal_(tm.weight, 0.0, math.sqrt(2) / math.sqrt(shape[0]))
    finally:
        torch.random.set_rng_state(state)
    return m


def sphere(width=500, depth=2, radius=0.5, seed=0):
    nodes = [3] + [width] * depth + [1]
    return _geometric(nodes, [[0]] * depth, [], radius, seed)


def sal(depth=8, width=512, radius=0.5, seed=0, skip=True):
    arc = [[0] for _ in range(depth)]
    tms = []
    if skip:
        arc[depth // 2 - 1] = [1, 0, 0]
        tms = [[width, 3]]
    return _geometric([3] + [width] * depth + [1], arc, tms, radius, seed)


def chair(path=None):
    if path is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "..", "tests", "golden", "chair.onnx")
    return load_model(path)


def by_name(name):
    """'polytope' | 'sphere' | 'chair' | 'mlp<depth>x<width>[s]' (trailing s = input skip)."""
    if name == "polytope":
        return polytope()
    if name == "sphere":
        return sphere()
    if name == "chair":
        return chair()
    if name.startswith("mlp"):
        spec = name[3:]
        skip = spec.endswith("s")
        d, w = (int(x) for x in spec.rstrip("s").split("x"))
        return sal(depth=d, width=w, skip=skip)
    raise ValueError(name)


def latent_shapes(model, n_shapes, latent_dim=256, sigma=0.01, seed=0):
    """BASELINE.json config 5: a DeepSDF-style decoder whose first layer sees (x, latent code).  For a fixed code
    c_k the latent columns W_c fold into the bias, b'_0 = W_c c_k + b_0 (reference README.md:181), so every
    shape is a plain 3-input MLP that shares all weights with `model` except biases[0].
    Returns the list of first-layer biases (float32 tensors), shape k drawn with torch.manual_seed(k)."""
    lin0 = model.linears[0]
    n1 = lin0.weight.shape[0]
    g = torch.Generator().manual_seed(seed)
    w_c = torch.randn(n1, latent_dim, generator=g) * (math.sqrt(2) / math.sqrt(n1))
    out = []
    for k in range(n_shapes):
        c = torch.randn(latent_dim, generator=torch.Generator().manual_seed(k)) * sigma
        out.append((w_c @ c + lin0.bias.detach().cpu().float()).contiguous())
    return out


def shapes_of_rank(n_shapes, rank, world):
    """Replica sharding of a batch of shapes: shape k runs on GPU k mod world, no collective."""
    return [k for k in range(n_shapes) if k % world == rank]
