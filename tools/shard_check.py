"""Sharded march == single-GPU march, bit for bit.  Launch with
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py [case ...]"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist

from analyticmesh_b200 import cuam
from analyticmesh_b200.parallel import make_allreduce
from tests.golden.cases import build_case
from tests import parity

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for name in (sys.argv[1:] or ["chair_cube", "skipnet", "chair", "mlp4x128s"]):
    case = build_case(name)

    def digest(shard):
        info = case["info"]
        cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table,
                  num_extra_constraints=len(case["b_extra"]))
        if shard == 1:
            cuam.set_shard(rank, world, make_allreduce())
        elif shard == 2:
            from analyticmesh_b200.parallel import broadcast_bytes
            cuam.set_shard_nccl(rank, world, broadcast_bytes)
        cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=case["states"], points=case["points"],
                              arc_tm=info.arc_tm, w_extra_constraints=case["w_extra"].reshape(-1, 3),
                              b_extra_constraints=case["b_extra"].reshape(-1), iso=0.0, flip_insideout=False)
        keys, fo, par, via = cuam.states()
        e, v = cuam.faces()
        st = cuam.stats()
        cuam.CombineMesh(1.0, [0, 0, 0])
        mv, fs, fi = cuam.mesh()
        h = hashlib.sha256()
        for a in (keys, fo, par, via, e, v, mv, fs, fi):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest(), st

    d1, s1 = digest(0)
    d2, s2 = digest(2)
    d3, s3 = digest(1)
    ok = d1 == d2 == d3
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name}: world={world} identical={bool(t.item())} faces={s2['n_faces']} "
              f"t_single={s1['seconds_march']:.4f}s t_sharded={s2['seconds_march']:.4f}s", flush=True)
    assert ok, (name, rank)
dist.destroy_process_group()
