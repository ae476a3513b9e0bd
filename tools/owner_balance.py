"""Predicted load balance of the sharded march (owner of a child = owner of its parent): runs one
single-GPU march, replays the ownership rule for N = 2, 4, 8 and prints sum_levels(max load) vs the
ideal sum_levels(mean load).  usage: python tools/owner_balance.py <zoo name> [n_seeds]"""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from analyticmesh_b200 import zoo, cuam
from analyticmesh_b200.netinfo import NetInfo
from analyticmesh_b200.initializers import dichotomy, states_of

name = sys.argv[1]
n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
m = zoo.by_name(name)
pts = dichotomy(m, 0.0, n_seeds, generator=torch.Generator().manual_seed(0), rng=random.Random(0))
st = states_of(m, pts).numpy()
info = NetInfo.from_model(m)
cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=st, points=pts.double().numpy(),
                      arc_tm=info.arc_tm, w_extra_constraints=np.zeros((0, 3)), b_extra_constraints=np.zeros(0),
                      iso=0.0, flip_insideout=False)
keys, fo, parent, via = cuam.states()
n = len(parent)
level = np.zeros(n, dtype=np.int32)
seeds = parent < 0
# ids are in BFS order: a parent always precedes its children
for i in np.nonzero(~seeds)[0]:
    pass
level_of = np.zeros(n, dtype=np.int32)
idx = np.nonzero(~seeds)[0]
# vectorised by levels: children of level l are contiguous after level l
lb, le, lvl = 0, int(seeds.sum()), 0
bounds = [(lb, le)]
while le < n:
    nxt = le
    # children of [lb, le) are the maximal run starting at le whose parent lies in [lb, le)
    hi = np.searchsorted(parent[le:], le, side="left") + le   # parents are non-decreasing in id order
    lb, le = le, int(hi)
    bounds.append((lb, le))
for N in (2, 4, 8):
    owner = np.zeros(n, dtype=np.int16)
    s0, s1 = bounds[0]
    owner[s0:s1] = np.arange(s1 - s0) % N
    tot_max = tot_mean = 0.0
    worst = 0.0
    for (a, b) in bounds[1:]:
        owner[a:b] = owner[parent[a:b]]
    for (a, b) in bounds:
        c = np.bincount(owner[a:b], minlength=N)
        tot_max += c.max()
        tot_mean += (b - a) / N
        if b - a > 1000:
            worst = max(worst, c.max() / ((b - a) / N))
    print(f"{name}: N={N} levels={len(bounds)} predicted compute efficiency={tot_mean / tot_max:.3f} worst level max/mean={worst:.2f}", flush=True)
