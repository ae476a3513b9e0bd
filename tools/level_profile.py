"""Level structure of a march and the balance of the owner-inherits-parent sharding, from the parent array of
one single-GPU march:  python tools/level_profile.py [workload] [seeds]
Prints, per world size, sum_levels max_rank(work) / sum_levels mean_rank(work) for work = states and for
work = recomputed state-layers (the composition cost), plus the level-size histogram."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np

import bench
from analyticmesh_b200 import cuam

name = sys.argv[1] if len(sys.argv) > 1 else "mlp8x512s"
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
info, points, states, _ = bench.build_workload(name, 0, seeds)
cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=states, points=points, arc_tm=info.arc_tm,
                      w_extra_constraints=np.zeros((0, 3)), b_extra_constraints=np.zeros(0), iso=0.0, flip_insideout=False)
st = cuam.stats()
n = st["n_states"]
parent = np.zeros(n, np.int32)
via = np.zeros(n, np.int32)
p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
assert cuam.lib().am_copy_states(cuam._handle, None, None, p(parent), p(via)) == 0
level = np.zeros(n, np.int32)
n_seed = int((parent < 0).sum())
# states are numbered level by level, so parents precede children: one vectorised pass per level
lb, le, lv, bounds = 0, n_seed, 0, [0]
while le > lb:
    bounds.append(le)
    nxt = le
    # children of [lb, le) are contiguous after le
    hi = np.searchsorted(parent[le:], le, side="left") + le if le < n else n
    level[le:hi] = lv + 1
    lb, le, lv = le, hi, lv + 1
sizes = np.diff(np.array(bounds))
off = np.cumsum([0] + info.nodes[1:-1])
D = len(info.nodes) - 2
bucket = np.where(via < 0, 1, np.clip(np.searchsorted(off, via, side="right"), 1, D))    # hidden layer of the flipped neuron
layers_recomputed = D - bucket            # fc layers h = bucket .. D-1 are recomputed
out = {"workload": name, "seeds": seeds, "n_states": int(n), "levels": int(len(sizes)), "max_level": int(sizes.max()),
       "levels_below_1024": int((sizes < 1024).sum()), "levels_below_8192": int((sizes < 8192).sum()),
       "levels_below_32768": int((sizes < 32768).sum()),
       "states_in_levels_below_8192": int(sizes[sizes < 8192].sum()), "balance": {}}
for world in (2, 4, 8):
    owner = np.zeros(n, np.int8)
    owner[:n_seed] = np.arange(n_seed) % world
    for i in range(1, len(bounds) - 1):
        a, b = bounds[i], bounds[i + 1]
        owner[a:b] = owner[parent[a:b]]
    tot_s = tot_l = max_s = max_l = 0.0
    for i in range(len(bounds) - 1):
        a, b = bounds[i], bounds[i + 1]
        cs = np.bincount(owner[a:b], minlength=world)
        cl = np.bincount(owner[a:b], weights=layers_recomputed[a:b] + 2.0, minlength=world)   # + clip ~ 2 layer units
        tot_s += cs.sum() / world
        max_s += cs.max()
        tot_l += cl.sum() / world
        max_l += cl.max()
    out["balance"][world] = {"states_max_over_mean": max_s / tot_s, "work_max_over_mean": max_l / tot_l}
out["level_sizes"] = [int(x) for x in sizes]
print(json.dumps(out))
