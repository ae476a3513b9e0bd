"""bench.py -- faces/sec of one full Analytic Marching pass over the 8x512 SAL MLP (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one complete march of the workload: seed states -> region BFS -> all polygons (the
reference's `am_time` phase, backend/main.py:455-465).  Prints ONE JSON line on rank 0.

  value      faces/s, whole job, inputs (weights, seed states/points) already resident in HBM
  e2e        same metric through the C ABI with HOST buffers: host->device copies of weights and
             seeds and the device->host read of the stitched mesh (am_combine) inside the timed region
  roofline   the dominant kernel (compose_gemm_kernel): its algorithmic flops / its CUDA-event time,
             measured live on the engine's stream, against the FP64 DFMA peak measured in the same
             process (am_fp64_peak_tflops; tools/fp64_peak.cu measured 37.0 TFLOP/s for DFMA and DMMA)
  cpu_baseline  the CPU oracle (oracle/am_oracle.c, OpenMP) on a bounded sample of the same workload

N > 1: ONE march of the same network is spread over the N GPUs (am_set_shard): every rank composes and
clips the states it owns, the per-level polygons are combined with one NCCL all-reduce, the frontier
and visited set are replicated; value = faces of the mesh / max-over-ranks time ("strong").
--impl reference: the reference algorithm's CPU restatement on the host cores (rank 0 only).
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "faces_per_sec"
UNIT = "faces/s"


def build_workload(name, seed, n_seeds):
    from analyticmesh_b200 import zoo
    from analyticmesh_b200.netinfo import NetInfo
    from analyticmesh_b200.initializers import dichotomy, states_of
    if name.startswith("mlp"):
        spec = name[3:]
        skip = spec.endswith("s")
        d, w = (int(x) for x in spec.rstrip("s").split("x"))
        model = zoo.sal(depth=d, width=w, skip=skip, seed=seed)
    else:
        model = zoo.by_name(name)
    t0 = time.time()
    pts = dichotomy(model, 0.0, n_seeds, generator=torch.Generator().manual_seed(seed), rng=random.Random(seed))
    init_point_time = time.time() - t0
    states = states_of(model, pts).numpy()
    info = NetInfo.from_model(model)
    return info, np.ascontiguousarray(pts.double().numpy()), np.ascontiguousarray(states), init_point_time


def flops_per_face(nodes):
    """8 * sum_{l>=2} n_l n_{l-1} + 8 n_D (linear skips from hidden layers not present here)."""
    h = nodes[1:-1]
    return 8 * sum(h[i] * h[i - 1] for i in range(1, len(h))) + 8 * h[-1]


def roofline(variant, split_digits, achieved, fp64_peak, launches, ms, flops):
    """Dominant kernel against its bound.  FP64 DMMA path: the measured DFMA/DMMA pipe peak.  tcgen05 path
    (SURVEY 8d): the int8 tensor peak divided by the number of digit products, with int8 = 2 x the bf16 rate
    measured by the driver (MEASURED_PEAKS.json, sustained figure: the kernel is timed inside a long step)."""
    avg_ms = ms / max(launches, 1)
    if variant != 2:
        traffic = ncu_traffic("compose_gemm_kernel", flops / max(launches, 1))
        return {"bound": "fp64", "kernel": "compose_gemm_kernel", "achieved": achieved, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak > 0 else None, "traffic": traffic,
                "traffic_source": "scaled from the committed ncu --set full capture (profiles/ncu_traffic.json), not measured in this run",
                "launches": launches, "avg_launch_ms": avg_ms,
                "peak_source": "DFMA loop measured in this process (am_fp64_peak_tflops); FP64 is not in "
                               "MEASURED_PEAKS.json; tools/fp64_peak.cu: DFMA 37.0, DMMA 37.0, cuBLAS DGEMM 35.8"}
    products = split_digits * (split_digits + 1) // 2
    bf16, src = 1400.0, "fallback of /opt/skills/guides/B200_PROFILING.md (sustained bf16 1.4 PFLOP/s)"
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        bf16, src = float(mp["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained"
    except Exception:
        pass
    peak = 2.0 * bf16 / products
    return {"bound": "tensor", "kernel": "split_gemm_kernel", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak, "traffic": ncu_traffic("split_gemm_kernel", flops / max(launches, 1)),
            "traffic_source": "scaled from the committed ncu --set full capture (profiles/ncu_traffic.json), not measured in this run",
            "launches": launches, "avg_launch_ms": avg_ms,
            "int8_tops_achieved": achieved * products, "int8_tops_peak": 2.0 * bf16, "digit_products": products,
            "peak_source": f"FP64-equivalent TFLOP/s: algorithmic 2*M*K*4*S flops of the launch; peak = int8 tensor peak / "
                           f"{products} digit products, int8 peak = 2 x {src} ({bf16:.0f})"}


def ncu_traffic(kernel, flops_per_launch):
    """DRAM bytes per launch of the dominant kernel, scaled from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json: dram bytes and algorithmic flops of the captured launches)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]
        return t["dram_bytes"] / t["flops"] * flops_per_launch
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        busy = [v for v in sm if v > 500] or sm
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_sample(info, points, states, max_states, threads=0):
    """The restated reference algorithm on the host cores, first `max_states` processed states."""
    from oracle import am_oracle
    r = am_oracle.march(info, states, points, max_states=max_states, threads=threads)
    return r["n_faces"], r["seconds"], r["n_states"]


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    info, points, states, _ = build_workload(args.workload, 0, args.seeds)
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1: ask for every host core explicitly
    for _ in range(args.warmup):
        cpu_reference_sample(info, points, states, args.ref_states, threads=cores)
    faces, secs = 0, 0.0
    for _ in range(args.steps):
        f, s, _n = cpu_reference_sample(info, points, states, args.ref_states, threads=cores)
        faces += f
        secs += s
    v = faces / secs
    sample = f"first {args.ref_states} states of the march (LIFO batches of 1024), same network and seeds"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "CPU restatement of the reference algorithm "
                   "(oracle/am_oracle.c, OpenMP); the reference itself has no CPU path"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args):
    d, w = args.workload[3:].rstrip("s").split("x")
    return (f"{args.workload}: SAL geometric-init ReLU MLP 3-[{w}]x{d}-1" +
            (" with a linear skip from the input into the middle layer" if args.workload.endswith("s") else "") +
            f", r=0.5, torch.manual_seed(0), {args.seeds} dichotomy seeds, iso 0, float64")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="mlp8x512s")
    ap.add_argument("--seeds", type=int, default=1024)
    ap.add_argument("--ref-states", type=int, default=16384)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from analyticmesh_b200 import cuam
    info, points, states, init_point_time = build_workload(args.workload, 0, args.seeds)
    E0w, E0b = np.zeros((0, 3)), np.zeros(0)
    host = dict(weights=info.weights, biases=info.biases, states=states, points=points, arc_tm=info.arc_tm,
                w_extra_constraints=E0w, b_extra_constraints=E0b)
    devi = dict(weights=[torch.from_numpy(w).to(dev) for w in info.weights],
                biases=[torch.from_numpy(b).to(dev) for b in info.biases],
                states=torch.from_numpy(states).to(dev), points=torch.from_numpy(points).to(dev),
                arc_tm=[torch.from_numpy(t).to(dev) for t in info.arc_tm],
                w_extra_constraints=torch.zeros((0, 3), dtype=torch.float64, device=dev),
                b_extra_constraints=torch.zeros((0,), dtype=torch.float64, device=dev))
    t0 = time.time()
    cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
    torch.cuda.synchronize()
    init_cuda_time = time.time() - t0
    shard_mode = os.environ.get("AM_B200_SHARD", "p2p")
    if world > 1:
        from analyticmesh_b200.parallel import broadcast_bytes, make_allgather
        if shard_mode == "nccl":      # round-1 scheme: ncclAllReduce of the level's polygons, replicated visited set
            cuam.set_shard_nccl(rank, world, lambda b: broadcast_bytes(b, device=dev))
        else:                         # the engine's own exchange over NVLink peer memory (csrc/xchg.cuh)
            cuam.set_shard_p2p(rank, world, make_allgather(device=dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def march(bufs):
        cuam.AnalyticMarching(iso=0.0, flip_insideout=False, **bufs)
        return cuam.stats()

    for i in range(args.warmup):
        march(devi)
        if i == args.warmup - 1 and rank == 0:      # warm the stitching / pinned read-back path of the e2e leg too
            cuam.CombineMesh(scale=1.0, center=[0.0, 0.0, 0.0])

    peak = cuam.fp64_peak_tflops()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- timed: device-resident inputs ------------------------------------------------------
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    faces = launches = 0
    gemm_ms = gemm_flops = 0.0
    gemm_launches = 0
    variant, split_digits = cuam.gemm_variant()
    prof_kind = 4 if variant == 2 else None       # the tcgen05 GEMM kernel alone, else the FP64 DMMA launches
    engine_s = 0.0
    last = None
    for _ in range(args.steps):
        last = march(devi)
        faces += last["n_faces"]
        launches += last["n_launches"]
        engine_s += last["seconds_march"]
        phases = {k: last[k] for k in ("seconds_compose", "seconds_clip", "seconds_frontier", "seconds_host_wait",
                                       "seconds_host_total")}
        names = {4: "split_gemm", 5: "digits", 6: "xchg_barrier_wait", 7: "xchg_push", 8: "xchg_unpack_scan_csr",
                 9: "expand_insert", 11: "finalize"}
        phases["kernel_seconds"] = {v: cuam.kernel_profile(k)["ms_total"] * 1e-3 for k, v in names.items()}
        p = cuam.kernel_profile(prof_kind) if prof_kind is not None else cuam.compose_profile()
        if prof_kind is not None and p["launches"] == 0:      # multi-chain mode records no per-kernel events
            p = cuam.compose_profile()
        gemm_ms += p["ms_total"]
        gemm_flops += p["flops"]
        gemm_launches += p["launches"]
    ev1.record()
    barrier()
    dt = ev0.elapsed_time(ev1) * 1e-3

    # ---- timed: end to end through the C ABI with host buffers --------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e_faces = 0
    d2h = 0
    e2e_parts = {"march_host_buffers": 0.0, "combine_and_read_back": 0.0, "export_ply": 0.0}
    ply = f"/tmp/am_b200_bench_rank{rank}.ply"
    ply_bytes = 0
    for _ in range(args.steps):
        t0 = time.time()
        st = march(host)
        t1 = time.time()
        t2 = t1
        if rank == 0:   # the mesh is replicated: one rank stitches it, reads it back and writes the PLY
            cuam.CombineMesh(scale=1.0, center=[0.0, 0.0, 0.0])
            st = cuam.stats()
            d2h = st["n_vertices"] * 24 + st["n_corners"] * 4 + (st["n_states"] + 1) * 8
            t2 = time.time()
            cuam.ExportMesh(file_path=ply, is_polymesh=True, is_float32=True)     # the reference's export_time phase
        e2e_parts["march_host_buffers"] += (t1 - t0) / args.steps
        e2e_parts["combine_and_read_back"] += (t2 - t1) / args.steps
        e2e_parts["export_ply"] += (time.time() - t2) / args.steps
        e_faces += st["n_faces"]
    e1.record()
    barrier()
    e_dt = e0.elapsed_time(e1) * 1e-3
    h2d = sum(a.nbytes for a in info.weights + info.biases + info.arc_tm) + states.nbytes + points.nbytes
    clocks = sampler.stop() if rank == 0 else None
    export_time = e2e_parts["export_ply"]
    if rank == 0:
        ply_bytes = os.path.getsize(ply)
        os.remove(ply)

    # what was computed, as checksums every run prints (1 GPU vs N GPUs, this build vs the last): see cuam.digest
    digest = cuam.digest()
    digest.pop("raw")
    mesh_check = None
    if rank == 0:
        inc = cuam.edge_incidence()
        stc = cuam.stats()
        n_edges = inc["matched"] // 2 + inc["boundary"] + inc["neighbour_missing"] + inc["neighbour_without_edge"]
        mesh_check = {"edge_incidence": inc, "n_vertices": stc["n_vertices"], "n_faces": stc["n_faces"],
                      "n_edges": n_edges, "euler_characteristic": stc["n_vertices"] - n_edges + stc["n_faces"],
                      "n_stitch_miss": stc["n_stitch_miss"], "n_overflow": stc["n_overflow"],
                      "n_unbounded": stc["n_unbounded"], "n_inconsistent": stc["n_inconsistent"]}
    if world > 1:       # every rank holds the same mesh: compare the ordered digests
        mine = torch.tensor([int(digest["ordered"][:15], 16)], dtype=torch.int64, device=dev)
        lo, hi = mine.clone(), mine.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        digest["identical_on_all_ranks"] = bool(int(lo) == int(hi))

    if world > 1:
        t = torch.tensor([dt, e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, e_dt = float(t[0]), float(t[1])
        # the mesh is replicated: faces are counted once; the roofline is per GPU (mean over ranks)
        c = torch.tensor([launches, gemm_flops / max(gemm_ms, 1e-9), gemm_ms, gemm_launches], dtype=torch.float64,
                         device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        launches, gemm_ms, gemm_launches = int(c[0]), float(c[2]) / world, int(c[3]) // world
        gemm_flops = float(c[1]) / world * gemm_ms          # so that achieved = mean over ranks of flops/ms

    if rank == 0:
        achieved = gemm_flops / max(gemm_ms * 1e-3, 1e-12) / 1e12
        fpf = flops_per_face(info.nodes)
        out = {
            "metric": METRIC, "value": faces / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "faces_per_step_rank0": last["n_faces"],
                       "states_per_step_rank0": last["n_states"], "bfs_levels": last["n_levels"],
                       "l2_note": "every step streams >= 9 GB of keys and >100 GB of plane rows through a 126 MB L2; "
                                  "inputs are far larger than L2, no flush needed",
                       "parallelism": (f"1 process per GPU; compose+clip sharded by state owner over {world} GPUs, visited set "
                                       "sharded by key hash; per BFS level the polygons and winner masks are pushed into the "
                                       "peers' inboxes over NVLink by the engine's kernels (CUDA IPC peer memory, device-side "
                                       "flag barriers, no host sync, no NCCL); key arena + CSR replicated"
                                       if shard_mode != "nccl" else
                                       f"1 process per GPU; compose+clip sharded over {world} GPUs, ncclAllReduce of the level's "
                                       "polygons, frontier replicated (round-1 scheme)")
                       if world > 1 else "single GPU",
                       "mesh_time_s": {"init_point_time": init_point_time, "init_cuda_time": init_cuda_time,
                                       "am_time": dt / args.steps,
                                       "am_combine_export_host_buffers": e_dt / args.steps,
                                       "export_time": export_time, "ply_bytes": ply_bytes,
                                       "end_to_end": init_point_time + init_cuda_time + e_dt / args.steps},
                       "engine_stream_seconds_per_step": engine_s / args.steps,
                       "e2e_wall_seconds_per_step_rank0": e2e_parts,
                       "phase_seconds_last_step_rank0": phases,
                       "algorithmic_flops_per_face": fpf,
                       "arithmetic": (f"FP64 planes; contraction as {split_digits} x {split_digits} signed 8-bit digit planes on "
                                      "tcgen05 kind::i8 with exact int32 accumulation, FP64 recombination"
                                      if variant == 2 else "FP64 tensor-core DMMA")},
            "e2e": {"value": e_faces / e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "digest": digest,
            "mesh_check": mesh_check,
            "clocks": clocks,
            "roofline": roofline(variant, split_digits, achieved, peak, gemm_launches, gemm_ms, gemm_flops),
        }
        try:        # reported baseline: the reference's own CUDA build on a B200 (golden-vector run, same network)
            g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_mlp8x512s_cube.json")))
            out["reference_cuda_b200"] = {
                "value": g["timing"]["faces_per_sec"], "unit": UNIT, "faces": g["n_faces"],
                "am_time_s": min(g["timing"]["am_time"]),
                "note": "unmodified reference (baseline/build_ref.sh, sm_100) on one B200, same 8x512 network clipped "
                        "by a cube (80 493 faces; the reference's fixed 2^23 vertex arena cannot hold the full mesh), "
                        "measured by tests/golden/make_golden_ref.py"}
        except Exception:
            pass
        if world == 1 and not args.no_cpu_baseline:
            f, s, n = cpu_reference_sample(info, points, states, args.ref_states, threads=os.cpu_count() or 1)
            out["cpu_baseline"] = {"value": f / s, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"first {n} states of the same march (same network, same seeds), "
                                             f"{s:.1f} s of OpenMP CPU work"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
