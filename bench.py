"""bench.py -- faces/sec of one full Analytic Marching pass over the 8x512 SAL MLP (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference|reference-cuda] [--workload batch64]

One "step" = one complete march of the workload: seed states -> region BFS -> all polygons (the
reference's `am_time` phase, backend/main.py:455-465).  Prints ONE JSON line on rank 0.

  value      faces/s, whole job, inputs (weights, seed states/points) already resident in HBM
  e2e        same metric through the C ABI with HOST buffers: host->device copies of weights and
             seeds and the device->host read of the stitched mesh (am_combine) inside the timed region
  roofline   the dominant kernel -- split_gemm_kernel (tcgen05 kind::i8 split-integer composition GEMM) for layers
             >= 256 wide, compose_gemm_kernel (FP64 DMMA) for narrower networks: its algorithmic flops
             (2*M*K*4*S per launch) / its CUDA-event time measured live on the engine's stream, against
             the int8 tensor peak derived from MEASURED_PEAKS.json divided by the number of digit products (and, beside
             it, the instruction's own ceiling measured with tools/int8_peak.cu), or the FP64 DFMA/DMMA peak measured
             in the same process (am_fp64_peak_tflops; tools/fp64_peak.cu: 37.0 TFLOP/s)
  cpu_baseline  the CPU oracle (oracle/am_oracle.c, OpenMP) on a bounded sample of the same workload
  digest / mesh_check  device-side checksums of everything the march produced (numbering, keys, edge loops, every
             coordinate bit) and edge incidence / Euler characteristic of the stitched mesh: the same values at every N

N > 1: ONE march of the same network is spread over the N GPUs (am_set_shard_p2p, csrc/xchg.cuh): every rank composes
and clips the states it owns, the visited set is sharded by key hash, the level's polygons and winner masks are pushed
into the peers' exchange blocks over NVLink by the engine's own kernels (CUDA-IPC peer memory, device-side barriers; no
NCCL and no host synchronisation in the exchange); value = faces of the mesh / max-over-ranks time ("strong").
AM_B200_SHARD=nccl selects the round-1 scheme (ncclAllReduce of the level's polygons) for comparison.
--impl reference: the reference algorithm's CPU restatement on all host cores (rank 0 only), bounded sample per step.
--impl reference-cuda: the reference's own CUDA build (baseline/_ref) and this engine on the same inputs, one process.
--workload batch64: BASELINE config 5 (64 latent-conditioned shapes, replicas, shape k on GPU k mod N).
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
    burst = None
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        bf16, src = float(mp["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained"
        burst = float(mp["bf16_tflops"])
    except Exception:
        pass
    peak = 2.0 * bf16 / products
    instr = None
    try:      # the instruction's own ceiling, measured with tools/int8_peak.cu on a B200 of this pool (register-free MMA loop)
        ip = json.load(open(os.path.join(ROOT, "profiles", "r02_int8_peak.json")))
        instr = {"int8_tops_instruction_peak": ip["split_mix_tops"],
                 "frac_of_instruction_peak": achieved * products / ip["split_mix_tops"],
                 "source": "profiles/r02_int8_peak.json: tcgen05.mma kind::i8 issued back to back on resident shared-memory "
                           "operands (the 10 instructions per K step of this kernel), short burst, no power cap"}
    except Exception:
        pass
    return {"bound": "tensor", "kernel": "split_gemm_kernel", "achieved": achieved, "peak": peak, "instruction_peak": instr,
            "unit": "TFLOP/s", "frac": achieved / peak, "traffic": ncu_traffic("split_gemm_kernel", flops / max(launches, 1)),
            "traffic_source": "scaled from the committed ncu --set full capture (profiles/ncu_traffic.json), not measured in this run",
            "launches": launches, "avg_launch_ms": avg_ms,
            "int8_tops_achieved": achieved * products, "int8_tops_peak": 2.0 * bf16, "digit_products": products,
            # the sustained bf16 figure is a power-capped cuBLAS run: a fraction near (or above) 1 of the peak derived from
            # it means "as fast as a power-capped library GEMM", not "at the silicon limit" -- see the two other fractions
            "frac_of_burst_derived_peak": (achieved / (2.0 * burst / products)) if burst else None,
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
        "config": base_config(args),
        "details": {"note": "CPU restatement of the reference algorithm (oracle/am_oracle.c, OpenMP); the reference itself "
                            "has no CPU path; every step is a bounded SAMPLE of the march"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_reference_cuda_arm(args):
    """The reference's OWN CUDA build (baseline/build_ref.sh: unmodified sources compiled for sm_100, staged under
    baseline/_ref) and this engine on the SAME inputs in the same process, one B200.  The stock reference cannot hold
    the full 17.7 M-face mesh (fixed 2^23 vertex arena and 2^25-slot fingerprint tables are plain #defines,
    reference backend/inc/macro.h:33,48,50), so the like-for-like workload is the 8x512 network clipped by a cube
    (tests/golden/cases.py `mlp8x512s_cube`, the case the golden vectors were generated from) and a wider cube."""
    import importlib.util
    so = os.path.join(ROOT, "baseline", "_ref", "AnalyticMesh", "backend", "build", "cuam.so")
    if not os.path.exists(so):
        print(json.dumps({"impl": "reference-cuda", "unavailable": "baseline/_ref/.../cuam.so not built (baseline/build_ref.sh)"}))
        return
    spec = importlib.util.spec_from_file_location("make_golden_ref", os.path.join(ROOT, "tests", "golden", "make_golden_ref.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.import_reference()
    import importlib
    ref = importlib.import_module('build.cuam')
    from analyticmesh_b200 import cuam
    from analyticmesh_b200.utils import get_boundary
    from tests.golden.cases import build_case, extra_constraints, _auto_cube
    torch.cuda.set_device(0)
    base = build_case("mlp8x512s_cube")
    rows = []
    for label, half in (("cube_0.08", None), ("cube_0.30", 0.15)):
        c = dict(base)
        if half is not None:      # a wider cube around the same surface point; seeds of the small cube lie inside it
            cube = _auto_cube(c["model"], half=half)
            c["w_extra"], c["b_extra"] = extra_constraints(cube)
        model, info = c["model"], c["info"]
        dt = torch.float64
        mi = model.get_info()
        W = [w.detach().to(dt).cuda().contiguous() for w in mi['weights']]
        B = [b.detach().to(dt).cuda().contiguous() for b in mi['biases']]
        TM = [t.detach().to(dt).cuda().contiguous() for t in mi['arc_tm']]
        arc = mi['arc_table'].to(torch.int32).cpu().contiguous()
        st = torch.from_numpy(c["states"]).to(torch.bool).cuda().contiguous()
        pt = torch.from_numpy(c["points"]).to(dt).cuda().contiguous()
        we = torch.from_numpy(np.ascontiguousarray(c["w_extra"])).to(dt).reshape(-1, 3).cuda().contiguous()
        be = torch.from_numpy(np.ascontiguousarray(c["b_extra"])).to(dt).reshape(-1).cuda().contiguous()
        kw = dict(weights=W, biases=B, states=st, points=pt, arc_tm=TM, w_extra_constraints=we, b_extra_constraints=be,
                  iso=0.0, flip_insideout=False)
        # ---- this engine ----
        cuam.Init(float_type="float64", nodesnum=list(model.nodes), arc_table=arc.numpy(), num_extra_constraints=int(be.shape[0]))
        ours = []
        for _ in range(args.warmup + args.steps):
            cuam.AnalyticMarching(**kw)
            ours.append(cuam.stats())
        ours = ours[args.warmup:]
        o_faces = ours[-1]["n_faces"]
        o_time = float(np.mean([o["seconds_march"] for o in ours]))
        cuam.Destroy()
        if o_faces > 7_000_000:
            rows.append({"workload": label, "faces": o_faces, "skipped": "would overflow the reference's 2^23 arena"})
            continue
        # ---- the reference ----
        ref.Init(float_type="float64", nodesnum=list(model.nodes), arc_table=arc, num_extra_constraints=int(be.shape[0]))
        r_times = []
        for _ in range(max(1, args.warmup - 1) + args.steps):
            torch.cuda.synchronize()
            t0 = time.time()
            ref.AnalyticMarching(**kw)
            torch.cuda.synchronize()
            r_times.append(time.time() - t0)
        r_times = r_times[-args.steps:]
        ref.CombineMesh(scale=1.0, center=[0.0, 0.0, 0.0])
        ply = f"/tmp/am_ref_cuda_{label}.ply"
        ref.ExportMesh(file_path=ply, is_polymesh=True, is_float32=True)
        ref.Destroy()
        head = open(ply, "rb").read(400).decode(errors="ignore")
        r_faces = int([ln for ln in head.splitlines() if ln.startswith("element face")][0].split()[2])
        os.remove(ply)
        r_time = float(np.mean(r_times))
        rows.append({"workload": label, "faces_reference": r_faces, "faces_engine": o_faces,
                     "reference_am_time_s": r_time, "engine_am_time_s": o_time,
                     "reference_faces_per_s": r_faces / r_time, "engine_faces_per_s": o_faces / o_time,
                     "speedup": (o_faces / o_time) / (r_faces / r_time)})
    best = [r for r in rows if "speedup" in r]
    v = best[-1]["reference_faces_per_s"] if best else None
    print(json.dumps({
        "impl": "reference-cuda", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "8x512 SAL MLP with input skip (BASELINE config 4 network, tests/golden/cases.py seeds) clipped "
                               "by a cube; the stock reference cannot hold the unclipped mesh",
                   "reference": "unmodified reference sources compiled for sm_100 (baseline/build_ref.sh), driven through its "
                                "own cuam.Init/AnalyticMarching, am_time bracketed by torch.cuda.synchronize()",
                   "gpu": torch.cuda.get_device_name(0)},
        "same_inputs_same_process": rows}))


def run_batch(args, n_shapes):
    """BASELINE config 5: `n_shapes` latent-conditioned decoders (config-4 network, per-shape first-layer bias
    b'_0 = W_c c_k + b_0, reference README.md:181) marched in ONE environment per process (the reference's environment
    cache, backend/main.py:437-449); shape k runs on GPU k mod N -- replicas, no data-path collective.
    value = total faces / max-over-ranks time of the marches; e2e adds the per-shape host buffers, stitching, read-back
    and PLY export.  Only biases[0] differs between shapes, so a march re-stages one tensor (weight cache)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from analyticmesh_b200 import cuam, zoo
    from analyticmesh_b200.netinfo import NetInfo
    from analyticmesh_b200.initializers import dichotomy, states_of
    model = zoo.by_name(args.net)
    biases0 = zoo.latent_shapes(model, n_shapes)
    jobs = []
    t0 = time.time()
    for k in zoo.shapes_of_rank(n_shapes, rank, world):
        with torch.no_grad():
            model.linears[0].bias.copy_(biases0[k])
        pts = dichotomy(model, 0.0, args.seeds, generator=torch.Generator().manual_seed(k), rng=random.Random(k))
        info = NetInfo.from_model(model)
        jobs.append((k, info, np.ascontiguousarray(pts.double().numpy()), np.ascontiguousarray(states_of(model, pts).numpy())))
    init_point_time = time.time() - t0
    info0 = jobs[0][1]
    cuam.Init(float_type="float64", nodesnum=info0.nodes, arc_table=info0.arc_table, num_extra_constraints=0)
    E0w, E0b = np.zeros((0, 3)), np.zeros(0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def to_dev(info, pts, st):
        return dict(weights=[torch.from_numpy(w).to(dev) for w in info.weights],
                    biases=[torch.from_numpy(b).to(dev) for b in info.biases], states=torch.from_numpy(st).to(dev),
                    points=torch.from_numpy(pts).to(dev), arc_tm=[torch.from_numpy(t).to(dev) for t in info.arc_tm],
                    w_extra_constraints=torch.zeros((0, 3), dtype=torch.float64, device=dev),
                    b_extra_constraints=torch.zeros((0,), dtype=torch.float64, device=dev))

    k0, i0, p0, s0 = jobs[0]
    for _ in range(max(1, args.warmup)):            # cold start (arena growth) outside the timed region
        cuam.AnalyticMarching(iso=0.0, flip_insideout=False, **to_dev(i0, p0, s0))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    devs = [to_dev(i, p, s) for _, i, p, s in jobs]
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    faces, launches, per_shape, reloaded = 0, 0, [], 0
    for (k, info, pts, st), d in zip(jobs, devs):
        cuam.AnalyticMarching(iso=0.0, flip_insideout=False, **d)
        s = cuam.stats()
        dg = cuam.digest()
        faces += s["n_faces"]
        launches += s["n_launches"]
        reloaded += s["n_tensors_reloaded"]
        per_shape.append({"shape": k, "faces": s["n_faces"], "seconds": round(s["seconds_march"], 4),
                          "region_set": dg["region_set"][:16]})
    ev1.record()
    barrier()
    dt = ev0.elapsed_time(ev1) * 1e-3
    # ---- end to end: host buffers in, stitched mesh read back and written as PLY, per shape ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e_faces, h2d, d2h = 0, 0, 0
    ply = f"/tmp/am_b200_batch_rank{rank}.ply"
    for k, info, pts, st in jobs:
        cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=st, points=pts, arc_tm=info.arc_tm,
                              w_extra_constraints=E0w, b_extra_constraints=E0b, iso=0.0, flip_insideout=False)
        cuam.CombineMesh(scale=1.0, center=[0.0, 0.0, 0.0])
        cuam.ExportMesh(file_path=ply, is_polymesh=True, is_float32=True)
        s = cuam.stats()
        e_faces += s["n_faces"]
        h2d += sum(a.nbytes for a in info.weights + info.biases + info.arc_tm) + st.nbytes + pts.nbytes
        d2h += s["n_vertices"] * 24 + s["n_corners"] * 4 + (s["n_states"] + 1) * 8
    e1.record()
    barrier()
    e_dt = e0.elapsed_time(e1) * 1e-3
    if os.path.exists(ply):
        os.remove(ply)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([dt, e_dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, e_dt = float(t[0]), float(t[1])
        c = torch.tensor([faces, e_faces, launches, h2d, d2h, reloaded], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        faces, e_faces, launches, h2d, d2h, reloaded = (int(v) for v in c.tolist())
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": faces / dt, "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": max(1, args.warmup),
            "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"batch of {n_shapes} latent-conditioned shapes of {args.net} (BASELINE config 5: shared "
                                   f"weights, per-shape biases[0], latents N(0, 0.01^2 I_256), {args.seeds} dichotomy seeds "
                                   "each), one environment per process, shape k on GPU k mod N (replicas, no collective)",
                       "total_faces": faces, "init_point_time_rank0": init_point_time,
                       "weight_tensors_restaged_per_march": reloaded / max(1, n_shapes),
                       "rank0_shapes": per_shape,
                       "l2_note": "every march streams far more than the 126 MB L2"},
            "e2e": {"value": e_faces / e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def base_config(args):
    """The part of `config` that names the workload: identical in this engine's line and in the reference arm's."""
    return {"workload": workload_name(args),
            "l2_note": "every step streams >= 9 GB of keys and >100 GB of plane rows through a 126 MB L2; inputs are far "
                       "larger than L2, no flush needed"}


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
    ap.add_argument("--net", default="mlp8x512s", help="network of --workload batch<N>")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.impl == "reference-cuda":
        run_reference_cuda_arm(args)
        return
    if args.workload.startswith("batch"):
        run_batch(args, int(args.workload[5:] or 64))
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
    # the surface-point initialiser on the device (csrc/seeds.cuh), timed for the mesh-time breakdown; the marches below
    # keep the fixed torch-generated seeds so that digests stay comparable between builds
    native_init = None
    try:
        cuam.seed_dichotomy(devi["weights"], devi["biases"], devi["arc_tm"], devi["w_extra_constraints"],
                            devi["b_extra_constraints"], 0.0, init_num=args.seeds, seed=0)
        native_init = cuam.seed_dichotomy(devi["weights"], devi["biases"], devi["arc_tm"], devi["w_extra_constraints"],
                                          devi["b_extra_constraints"], 0.0, init_num=args.seeds, seed=0)
    except Exception as e:      # noqa: BLE001 - reported in the line, the bench itself does not depend on it
        native_init = {"error": repr(e)}
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
            "config": base_config(args),
            "details": {"faces_per_step_rank0": last["n_faces"],
                       "states_per_step_rank0": last["n_states"], "bfs_levels": last["n_levels"],
                       "parallelism": (f"1 process per GPU; compose+clip sharded by state owner over {world} GPUs, visited set "
                                       "sharded by key hash; per BFS level the polygons and winner masks are pushed into the "
                                       "peers' inboxes over NVLink by the engine's kernels (CUDA IPC peer memory, device-side "
                                       "flag barriers, no host sync, no NCCL); key arena + CSR replicated"
                                       if shard_mode != "nccl" else
                                       f"1 process per GPU; compose+clip sharded over {world} GPUs, ncclAllReduce of the level's "
                                       "polygons, frontier replicated (round-1 scheme)")
                       if world > 1 else "single GPU",
                       "mesh_time_s": {"init_point_time": init_point_time, "init_point_native": native_init,
                                       "init_cuda_time": init_cuda_time,
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
            # ~10 s of CPU work on a 16-core box (4 x the per-step sample of the reference arm)
            f, s, n = cpu_reference_sample(info, points, states, 4 * args.ref_states, threads=os.cpu_count() or 1)
            out["cpu_baseline"] = {"value": f / s, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"first {n} states of the same march (same network, same seeds), "
                                             f"{s:.1f} s of OpenMP CPU work"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
