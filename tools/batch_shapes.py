"""BASELINE.json config 5: a batch of latent-conditioned decoders (shared weights, per-shape first-layer bias),
sharded per shape across the GPUs (replicas: shape k -> GPU k mod world, no data-path collective).

    python tools/batch_shapes.py [--net mlp8x512s] [--shapes 64] [--seeds 1024]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/batch_shapes.py ...

One Init per process, one march per shape (the environment is reused, reference backend/main.py:437-449).
Prints total faces / max-over-ranks wall time of the marches."""
import argparse
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="mlp8x512s")
    ap.add_argument("--shapes", type=int, default=64)
    ap.add_argument("--seeds", type=int, default=1024)
    ap.add_argument("--sigma", type=float, default=0.01)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = zoo.by_name(args.net)
    biases0 = zoo.latent_shapes(model, args.shapes, sigma=args.sigma)
    mine = zoo.shapes_of_rank(args.shapes, rank, world)
    jobs = []
    for k in mine:                       # seeds of every shape first: the timed region is the marching
        with torch.no_grad():
            model.linears[0].bias.copy_(biases0[k])
        pts = dichotomy(model, 0.0, args.seeds, generator=torch.Generator().manual_seed(k), rng=random.Random(k))
        jobs.append((k, NetInfo.from_model(model), pts.double().numpy(), states_of(model, pts).numpy()))
    info0 = jobs[0][1] if jobs else NetInfo.from_model(model)
    cuam.Init(float_type="float64", nodesnum=info0.nodes, arc_table=info0.arc_table, num_extra_constraints=0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    faces, per_shape = 0, []
    for k, info, pts, st in jobs:
        cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=st, points=pts, arc_tm=info.arc_tm,
                              w_extra_constraints=np.zeros((0, 3)), b_extra_constraints=np.zeros(0), iso=0.0,
                              flip_insideout=False)
        s = cuam.stats()
        faces += s["n_faces"]
        per_shape.append((k, s["n_faces"], round(s["seconds_march"], 3)))
    torch.cuda.synchronize()
    dt = time.time() - t0
    if world > 1:
        t = torch.tensor([dt, float(faces)], dtype=torch.float64, device="cuda")
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, faces = float(mx[0]), int(t[1])
    if rank == 0:
        print(json.dumps({"config": f"{args.shapes} shapes of {args.net}, latent sigma {args.sigma}, {args.seeds} seeds each",
                          "n_gpus": world, "faces": faces, "seconds": dt, "faces_per_s": faces / dt,
                          "rank0_shapes": per_shape}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
