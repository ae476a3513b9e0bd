"""Re-fit of `estimate_am_time` (reference backend/utils.py:76-90) for THIS engine on the GPU at hand.

The reference's model   t(l, n) = (a n)^(b l + c) * n^d * l^e * f   (l hidden layers of mean width n) was fitted
to its own engine.  This script marches a grid of SAL geometric-init MLPs (the family of BASELINE config 4, with
the input skip for l >= 4), measures am_time, fits the same six constants in log space and prints them with the
measurements (JSON); analyticmesh_b200/utils.py carries the result.
    python tools/fit_am_time.py > gpurun_out/fit_am_time.json
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from scipy.optimize import least_squares

from analyticmesh_b200 import cuam, zoo
from analyticmesh_b200.initializers import dichotomy, states_of
from analyticmesh_b200.netinfo import NetInfo

GRID = [(2, 64), (2, 128), (2, 256), (2, 512), (3, 64), (3, 128), (3, 256), (3, 512), (4, 64), (4, 128), (4, 256),
        (4, 512), (6, 64), (6, 128), (6, 256), (6, 384), (8, 64), (8, 128), (8, 256), (8, 384), (8, 512)]
rows = []
for depth, width in GRID:
    model = zoo.sal(depth=depth, width=width, skip=depth >= 4, seed=0)
    pts = dichotomy(model, 0.0, 1024, generator=torch.Generator().manual_seed(0), rng=random.Random(0))
    states = states_of(model, pts).numpy()
    info = NetInfo.from_model(model)
    cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
    best = None
    for _ in range(3):        # first march of a network is cold (arenas grow, weights staged)
        cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=states, points=pts.double().numpy(),
                              arc_tm=info.arc_tm, w_extra_constraints=np.zeros((0, 3)), b_extra_constraints=np.zeros(0),
                              iso=0.0, flip_insideout=False)
        st = cuam.stats()
        best = st["seconds_march"] if best is None else min(best, st["seconds_march"])
    rows.append({"depth": depth, "width": width, "faces": st["n_faces"], "levels": st["n_levels"], "am_time": best})
    print(rows[-1], file=sys.stderr, flush=True)
cuam.Destroy()

L = np.array([r["depth"] for r in rows], float)
N = np.array([r["width"] for r in rows], float)
T = np.array([r["am_time"] for r in rows], float)


def model_log(p, l, n):
    a, b, c, d, e, f = p
    return (b * l + c) * np.log(np.abs(a) * n) + d * np.log(n) + e * np.log(l) + np.log(np.abs(f))


x0 = np.array([1.94452188, 0.13816182, -0.14536181, 0.59338494, -1.20459825, 1e-6])
fit = least_squares(lambda p: model_log(p, L, N) - np.log(T), x0, max_nfev=20000)
pred = np.exp(model_log(fit.x, L, N))
print(json.dumps({"constants": [float(v) for v in fit.x], "gpu": torch.cuda.get_device_name(0),
                  "max_rel_err": float(np.max(np.abs(pred / T - 1))), "rows": rows,
                  "pred": [float(v) for v in pred]}))
