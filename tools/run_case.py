"""Run one config through the engine and print its counters (development helper).
usage: python tools/run_case.py <zoo name> [n_seeds] [reps]"""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from analyticmesh_b200 import zoo, cuam
from analyticmesh_b200.netinfo import NetInfo
from analyticmesh_b200.initializers import dichotomy, states_of

name = sys.argv[1]
n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
m = zoo.by_name(name)
pts = dichotomy(m, 0.0, n_seeds, generator=torch.Generator().manual_seed(0), rng=random.Random(0))
st = states_of(m, pts).numpy()
info = NetInfo.from_model(m)
cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
for r in range(reps):
    t0 = time.time()
    cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=st, points=pts.double().numpy(),
                          arc_tm=info.arc_tm, w_extra_constraints=np.zeros((0, 3)), b_extra_constraints=np.zeros(0),
                          iso=0.0, flip_insideout=False)
    t1 = time.time()
    s = cuam.stats()
    p = cuam.compose_profile()
    cuam.CombineMesh(1.0, [0, 0, 0])
    t2 = time.time()
    s2 = cuam.stats()
    print(json.dumps(dict(case=name, wall_march=t1 - t0, wall_combine=t2 - t1, faces_per_s=s["n_faces"] / s["seconds_march"],
                          gemm_tflops=p["flops"] / max(p["ms_total"], 1e-9) / 1e9, gemm_ms=p["ms_total"],
                          gemm_launches=p["launches"], **{**s, "n_vertices": s2["n_vertices"], "n_stitch_miss": s2["n_stitch_miss"]})), flush=True)
