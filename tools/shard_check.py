"""Sharded march == single-GPU march, bit for bit (keys, numbering, polygons, stitched mesh).  Launch with
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
       tools/shard_check.py [case ...]
Modes compared with the unsharded march: the exchange over NVLink peer memory (am_set_shard_p2p, the default
of bench.py) and -- on distinct GPUs -- the two all-reduce schemes of round 1 (am_set_shard_nccl / am_set_shard).
AM_SHARD_SAME_GPU=1: every rank uses cuda:0 (two processes time-slicing one GPU, CUDA IPC between them, gloo for
the handle swap) so that the multi-rank code path is exercised on a single-GPU box as well."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist

from analyticmesh_b200 import cuam
from analyticmesh_b200.parallel import broadcast_bytes, make_allgather, make_allreduce
from tests.golden.cases import build_case

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
same_gpu = os.environ.get("AM_SHARD_SAME_GPU", "0") == "1"
torch.cuda.set_device(0 if same_gpu else local)
if same_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
modes = [("p2p", 3)] if same_gpu else [("p2p", 3), ("nccl", 2), ("callback", 1)]
for name in (sys.argv[1:] or ["chair_cube", "skipnet", "chair", "mlp4x128s"]):
    case = build_case(name)

    def digest(shard):
        info = case["info"]
        cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table,
                  num_extra_constraints=len(case["b_extra"]))
        if shard == 1:
            cuam.set_shard(rank, world, make_allreduce())
        elif shard == 2:
            cuam.set_shard_nccl(rank, world, broadcast_bytes)
        elif shard == 3:
            cuam.set_shard_p2p(rank, world, make_allgather())
        cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=case["states"], points=case["points"],
                              arc_tm=info.arc_tm, w_extra_constraints=case["w_extra"].reshape(-1, 3),
                              b_extra_constraints=case["b_extra"].reshape(-1), iso=0.0, flip_insideout=False)
        keys, fo, par, via = cuam.states()
        e, v = cuam.faces()
        st = cuam.stats()
        dev_digest = cuam.digest()["ordered"]
        cuam.CombineMesh(1.0, [0, 0, 0])
        mv, fs, fi = cuam.mesh()
        inc = cuam.edge_incidence()
        h = hashlib.sha256()
        for a in (keys, fo, par, via, e, v, mv, fs, fi):
            h.update(np.ascontiguousarray(a).tobytes())
        h.update(repr(sorted(inc.items())).encode())
        return h.hexdigest() + dev_digest, st

    d0, s0 = digest(0)
    ok, times = True, []
    for label, m in modes:
        d, s = digest(m)
        ok = ok and (d == d0)
        times.append(f"t_{label}={s['seconds_march']:.4f}s")
    flag = torch.tensor([int(ok)])
    if not same_gpu:
        flag = flag.cuda()
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name}: world={world} identical={bool(flag.item())} faces={s0['n_faces']} "
              f"t_single={s0['seconds_march']:.4f}s " + " ".join(times), flush=True)
    assert ok, (name, rank)
cuam.Destroy()
dist.destroy_process_group()
