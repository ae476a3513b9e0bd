"""Parity of the BENCHMARKED workloads (BASELINE.json configs 4 and 5) and of the sharded march.

The full 8x512 march (17.7 M states, 9 GB of keys) cannot be run by the CPU oracle as a whole, so it is pinned
three ways:
  * a random sample of its visited states is recomputed state by state with the oracle (same key, same point
    on the shared edge, same start edge -- exactly what the reference pops from its stack,
    backend/src/cuam_kernel.cu:53-98) and must give the same edge loop and the same vertices;
  * size-independent properties on the whole mesh: one face per visited state, every neuron edge shared by
    exactly two faces (exact visited-set lookups on the device), every vertex of valence 4, Euler
    characteristic 2, no stitching miss;
  * the device-side digests (cuam.digest()) that bench.py prints at every GPU count are checked against the
    oracle on the cases the oracle can run whole.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.golden.cases import build_case
from tests import parity

pytestmark = pytest.mark.gpu
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


@pytest.mark.parametrize("name", ["polytope", "chair_cube", "skipnet", "mlp3x256s_cube", "chair"])
def test_device_digest_equals_oracle_topology_sum(oracle_lib, name):
    """The order-independent checksum of (key, edge loop) computed on the device equals the one computed from
    the oracle's march; the edge-incidence counters agree with the host-side incidence histogram."""
    from analyticmesh_b200 import cuam
    case = build_case(name)
    eng = parity.run_engine(case)
    dg = cuam.digest()
    inc = cuam.edge_incidence()
    L = case["info"].state_len
    orc = oracle_lib.march(case["info"], case["states"], case["points"], case["w_extra"], case["b_extra"])
    assert dg["n_states"] == orc["n_states"]
    assert dg["topology_sum"] == oracle_lib.topology_sum(oracle_lib.canonical_faces(orc), L), name
    assert dg["topology_sum"] == oracle_lib.topology_sum(parity.engine_faces(eng, L), L)
    hist = parity.closed_manifold_report(eng, L)
    assert inc["matched"] == 2 * hist.get(2, 0), (inc, hist)
    assert inc["neighbour_missing"] + inc["neighbour_without_edge"] == hist.get(1, 0), (inc, hist)
    assert inc["boundary"] == int((eng["edges"] >= L).sum())
    # a second march of the same network re-stages nothing (weight cache) and reproduces every bit
    eng2 = parity.run_engine(case)
    assert cuam.digest()["raw"] == dg["raw"]


def _bench_workload(name, seeds):
    sys.path.insert(0, ROOT)
    import bench
    return bench.build_workload(name, 0, seeds)


def test_full_config4_sampled_against_oracle(oracle_lib):
    """BASELINE config 4 exactly as bench.py runs it (8x512 + skip, 1024 dichotomy seeds, float64)."""
    from analyticmesh_b200 import cuam
    info, points, states, _ = _bench_workload("mlp8x512s", 1024)
    L = info.state_len
    cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
    cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=states, points=points, arc_tm=info.arc_tm,
                          w_extra_constraints=np.zeros((0, 3)), b_extra_constraints=np.zeros(0), iso=0.0,
                          flip_insideout=False)
    st = cuam.stats()
    assert st["n_states"] > 10_000_000 and st["n_faces"] == st["n_states"], st
    assert st["n_overflow"] == st["n_unbounded"] == st["n_inconsistent"] == 0, st

    # ---- sample of the visited states, recomputed by the oracle ----
    rs = np.random.RandomState(0)
    ids = np.unique(np.concatenate([rs.randint(0, st["n_states"], 4096), np.arange(16),
                                    np.arange(st["n_states"] - 16, st["n_states"])]))
    g = cuam.gather_states(ids)
    orc = oracle_lib.process_states(info, g["keys"], g["seedpt"][:, :3], np.where(g["parent"] < 0, -1, g["via"]))
    of = oracle_lib.canonical_faces(orc)
    kb = parity.key_bytes(g["keys"], L)
    bad, max_err = [], 0.0
    for i, k in enumerate(kb):
        n = int(g["counts"][i])
        want = of[k]
        if want is None or n < 3:
            bad.append((int(ids[i]), "no polygon", n, want is None))
            continue
        if tuple(int(e) for e in g["edges"][i, :n]) != want[0]:
            bad.append((int(ids[i]), "loop", tuple(g["edges"][i, :n]), want[0]))
            continue
        max_err = max(max_err, float(np.abs(g["xyz"][i, :n] - want[1]).max()) / max(1.0, float(np.abs(want[1]).max())))
    assert not bad, bad[:5]
    assert max_err < 1e-9, max_err                    # north_star tolerance: 1e-5 relative

    # ---- whole-mesh properties ----
    inc = cuam.edge_incidence()
    assert inc["boundary"] == inc["neighbour_missing"] == inc["neighbour_without_edge"] == 0, inc
    assert inc["matched"] == st["n_corners"]
    cuam.CombineMesh(scale=1.0, center=[0.0, 0.0, 0.0])
    st = cuam.stats()
    assert st["n_stitch_miss"] == 0
    n_edges = inc["matched"] // 2
    assert st["n_vertices"] - n_edges + st["n_faces"] == 2, (st["n_vertices"], n_edges, st["n_faces"])
    v, fs, fi = cuam.mesh()
    val = np.bincount(fi, minlength=len(v))
    assert val.min() == 4 and val.max() == 4
    rsel = rs.randint(0, len(v), 200_000)
    f_abs = np.abs(info.forward(v[rsel])[0])
    assert f_abs.max() < 1e-9, f_abs.max()
    d1 = cuam.digest()
    # same march again: bit-identical numbering, topology and coordinates
    cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=states, points=points, arc_tm=info.arc_tm,
                          w_extra_constraints=np.zeros((0, 3)), b_extra_constraints=np.zeros(0), iso=0.0,
                          flip_insideout=False)
    assert cuam.stats()["n_tensors_reloaded"] == 0
    assert cuam.digest()["raw"] == d1["raw"]
    cuam.Destroy()


def test_config5_latent_shapes_sampled_against_oracle(oracle_lib):
    """BASELINE config 5 at full size: two of the latent-conditioned 8x512 shapes, marched back to back in ONE
    environment (only biases[0] changes: exactly one tensor is re-staged), each checked like config 4 -- a sample of
    its visited states recomputed by the oracle on that shape's network, edge incidence, Euler characteristic."""
    import random
    from analyticmesh_b200 import cuam, zoo
    from analyticmesh_b200.netinfo import NetInfo
    from analyticmesh_b200.initializers import dichotomy, states_of
    model = zoo.by_name("mlp8x512s")
    biases0 = zoo.latent_shapes(model, 64)
    rs = np.random.RandomState(1)
    first = True
    seen = set()
    for k in (0, 37):
        with torch.no_grad():
            model.linears[0].bias.copy_(biases0[k])
        pts = dichotomy(model, 0.0, 1024, generator=torch.Generator().manual_seed(k), rng=random.Random(k))
        info = NetInfo.from_model(model)
        st = np.ascontiguousarray(states_of(model, pts).numpy())
        if first:
            cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
        cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=st, points=pts.double().numpy(),
                              arc_tm=info.arc_tm, w_extra_constraints=np.zeros((0, 3)), b_extra_constraints=np.zeros(0),
                              iso=0.0, flip_insideout=False)
        s = cuam.stats()
        # unlike the bias-free config-4 network these surfaces have regions whose polygon degenerates (fewer than 3
        # vertices after the feasibility tolerance): visited, but without a face -- the reference drops them as well
        assert s["n_states"] > 10_000_000 and 0 <= s["n_states"] - s["n_faces"] < 1e-4 * s["n_states"], s
        assert s["n_overflow"] == s["n_unbounded"] == 0, s
        if not first:
            assert s["n_tensors_reloaded"] == 1, s                   # the layer-1 table only
        first = False
        ids = np.unique(rs.randint(0, s["n_states"], 768))
        g = cuam.gather_states(ids)
        orc = oracle_lib.process_states(info, g["keys"], g["seedpt"][:, :3], np.where(g["parent"] < 0, -1, g["via"]))
        of = oracle_lib.canonical_faces(orc)
        L = info.state_len
        n_checked = 0
        for i, kb in enumerate(parity.key_bytes(g["keys"], L)):
            n = int(g["counts"][i])
            want = of[kb]
            if n < 3 or want is None:                                # degenerate on both sides, or on neither
                assert n < 3 and want is None, (k, int(ids[i]), n, want is None)
                continue
            assert tuple(int(e) for e in g["edges"][i, :n]) == want[0], (k, int(ids[i]))
            assert np.abs(g["xyz"][i, :n] - want[1]).max() < 1e-9
            n_checked += 1
        assert n_checked > 700
        inc = cuam.edge_incidence()
        assert inc["boundary"] == inc["neighbour_missing"] == 0, inc              # the search is closed under adjacency
        assert inc["neighbour_without_edge"] < 1e-4 * inc["matched"], inc        # edges into the degenerate regions
        cuam.CombineMesh(scale=1.0, center=[0.0, 0.0, 0.0])
        seen.add(cuam.digest()["region_set"])
    assert len(seen) == 2                                            # the two shapes are different surfaces
    cuam.Destroy()


def _torchrun(n, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_sharded_march_two_ranks_sharing_one_gpu():
    """The multi-rank code path on a single-GPU box: two processes time-slice cuda:0, map each other's exchange
    block with CUDA IPC and run the same device-side barriers / peer pushes as over NVLink (csrc/xchg.cuh).
    Bit-identical to the unsharded march on a DMMA-path and a tcgen05-path network."""
    env = dict(os.environ, AM_SHARD_SAME_GPU="1", AM_B200_XCHG_TIMEOUT_MS="15000", AM_B200_XCHG_MIB="16",
               AM_B200_RESIDENT_GIB="8")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(ROOT, "tools/shard_check.py"), "chair_cube", "mlp3x256s_cube", "skipnet"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if "identical=" in ln]
    assert len(lines) == 3 and all("identical=True" in ln for ln in lines), lines


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_march_is_bit_identical_to_single_gpu():
    """One march spread over all visible GPUs == the single-GPU march: keys, numbering, polygons, stitched mesh
    (sha256 over every array, tools/shard_check.py), on a DMMA-path and a tcgen05-path network."""
    n = min(torch.cuda.device_count(), 8)
    r = _torchrun(n, "tools/shard_check.py", "chair_cube", "skipnet", "mlp3x256s_cube", "mlp4x128s")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if "identical=" in ln]
    assert len(lines) == 4 and all("identical=True" in ln for ln in lines), lines


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_bench_digest_is_the_same_at_every_gpu_count():
    """bench.py's own line: the digest printed by the sharded run equals the single-GPU run's."""
    import json
    outs = []
    for n in (1, min(torch.cuda.device_count(), 8)):
        args = ["--gpus", str(n), "--steps", "1", "--warmup", "1", "--workload", "mlp4x256s", "--seeds", "256",
                "--no-cpu-baseline"]
        if n == 1:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                               timeout=900, cwd=ROOT)
        else:
            r = _torchrun(n, "bench.py", *args)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        outs.append(json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]))
    assert outs[0]["digest"]["ordered"] == outs[1]["digest"]["ordered"], (outs[0]["digest"], outs[1]["digest"])
    assert outs[0]["details"]["faces_per_step_rank0"] == outs[1]["details"]["faces_per_step_rank0"]


def test_clip_bulk_copy_variant_is_bit_identical():
    """The clip kernel's TMA 1-D streaming variant (cp.async.bulk + mbarrier ring, bulk shared->global store of the
    inherited rows; AM_B200_CLIP_MINB=4) gives the same mesh bit for bit as the cp.async ring, on a DMMA-path network,
    a tcgen05-path network with read-through, and a case with extra constraints.  Subprocesses: a trapping kernel must
    not take the test session down."""
    import json
    cases = ["chair", "mlp3x256s_cube", "chair_cube", "mlp8x512s_cube"]
    digests = []
    for variant in ("2", "4"):
        env = dict(os.environ, AM_B200_CLIP_MINB=variant)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "variant_check.py"), *cases], capture_output=True,
                           text=True, timeout=600, cwd=ROOT, env=env)
        assert r.returncode == 0, (variant, r.stdout[-2000:], r.stderr[-3000:])
        digests.append(json.loads(r.stdout.strip().splitlines()[-1]))
    for c in cases:
        assert digests[0][c]["raw"] == digests[1][c]["raw"] and digests[0][c]["faces"] > 0, (c, digests)
