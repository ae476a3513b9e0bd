"""Variants of split_gemm_kernel (csrc/split.cuh): same bits, how fast?
    1 / 2 / 4   epilogue warps per TMEM lane quarter (AM_B200_SPLIT_EPI)
    bk64        64-byte K slabs per pipeline stage, SWIZZLE_64B boxes (AM_B200_SPLIT_BK=64)

    [EPI_VARIANTS=bk64,1] [EPI_PLANES=5] python tools/epi_check.py [out.json]

(1) the full 8x512 march (bench workload) once per variant: device time of the GEMM launches (per-kernel CUDA events),
    march / compose seconds and the ordered digest of everything the march produced -- the digests must be equal;
(2) plane rows of small cases (digit counts 6 / 7 / 8, padded M and K, skip connections) per variant, compared bit for bit
    with the rows of variant 1, whose kernel is the one tests/test_split_gpu.py pinned to the integer restatement.
Every result is printed as it is produced (one JSON object per line)."""
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

VARIANTS = os.environ.get("EPI_VARIANTS", "2,4,1").split(",")
N_PLANE_CASES = int(os.environ.get("EPI_PLANES", "7"))
LATE = [v for v in os.environ.get("EPI_LATE", "").split(",") if v]      # marched after the plane checks (when time is short)


def select(v):
    """environment of variant `v`"""
    os.environ.pop("AM_B200_SPLIT_BK", None)
    os.environ.pop("AM_B200_SPLIT_EPI", None)
    if v == "bk64":
        os.environ["AM_B200_SPLIT_BK"] = "64"
    else:
        os.environ["AM_B200_SPLIT_EPI"] = str(int(v))


out = {"march": {}, "planes": {}}


def say(**kw):
    print(json.dumps(kw), flush=True)


def march_once(info, st, pts, epi):
    select(epi)
    cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
    t0 = time.time()
    cuam.AnalyticMarching(weights=info.weights, biases=info.biases, states=st, points=pts, arc_tm=info.arc_tm,
                          w_extra_constraints=np.zeros((0, 3)), b_extra_constraints=np.zeros(0), iso=0.0,
                          flip_insideout=False)
    wall = time.time() - t0
    s = cuam.stats()
    g = cuam.kernel_profile(4)
    d = cuam.digest()
    d.pop("raw", None)
    variant, sd = cuam.gemm_variant()
    rec = dict(epi=epi, gemm_variant=variant, split_digits=sd, wall_march=wall, seconds_march=s["seconds_march"],
               seconds_compose=s["seconds_compose"], seconds_clip=s["seconds_clip"], n_faces=s["n_faces"],
               gemm_ms=g["ms_total"], gemm_launches=g["launches"],
               gemm_tflops_fp64_equiv=g["flops"] / max(g["ms_total"], 1e-9) / 1e9,
               digest_ordered=d["ordered"], region_set=d["region_set"])
    cuam.Destroy()
    return rec


def planes_of(case, n, sd, epi):
    select(epi)
    os.environ["AM_B200_GEMM_VARIANT"] = "2"
    os.environ["AM_B200_SPLIT_DIGITS"] = str(sd)
    info = case["info"]
    cuam.Init(float_type="float64", nodesnum=info.nodes, arc_table=info.arc_table, num_extra_constraints=0)
    cuam.load_weights(info.weights, info.biases, info.arc_tm)
    planes, equ = cuam.debug_planes(case["states"][:n], iso=0.125)
    cuam.Destroy()
    return planes, equ


def main():
    assert torch.cuda.is_available()
    name = os.environ.get("EPI_NET", "mlp8x512s")
    m = zoo.by_name(name)
    pts = dichotomy(m, 0.0, 1024, generator=torch.Generator().manual_seed(0), rng=random.Random(0))
    st = states_of(m, pts).numpy()
    pts = np.ascontiguousarray(pts.double().numpy())
    info = NetInfo.from_model(m)
    def marches(which):
        for epi in which:
            try:
                rec = march_once(info, st, pts, epi)
            except Exception as e:      # noqa: BLE001 - keep going: the other variants still tell something
                rec = dict(epi=epi, error=repr(e))
                try:
                    cuam.Destroy()
                except Exception:       # noqa: BLE001
                    pass
            out["march"][str(epi)] = rec
            say(stage="march", **rec)
        digs = {r.get("digest_ordered") for r in out["march"].values()}
        out["march_digests_identical"] = len(digs) == 1 and None not in digs
        say(stage="march_digests_identical", value=out["march_digests_identical"], n=len(out["march"]))

    marches(VARIANTS)

    for k in ("AM_B200_GEMM_VARIANT", "AM_B200_SPLIT_DIGITS"):
        os.environ.pop(k, None)
    from tests.golden.cases import build_case
    for cname, sd in (("mlp4x128s", 6), ("mlp4x128s", 8), ("skipnet", 7), ("sphere", 7), ("mlp3x256s_cube", 7),
                      ("chair_cube", 7), ("mlp8x512s_cube", 7))[:N_PLANE_CASES]:
        try:
            case = build_case(cname)
            n = 37 if case["info"].state_len < 2000 else 19
            ref_p, ref_e = planes_of(case, n, sd, "1")
            rec = dict(case=cname, sd=sd, n_states=int(n), finite=bool(np.isfinite(ref_p).all()),
                       absmax=float(np.abs(ref_p).max()))
            for epi in VARIANTS:
                if epi == "1":
                    continue
                p, e = planes_of(case, n, sd, epi)
                rec[f"epi{epi}_rows_bit_identical"] = bool(np.array_equal(p.view(np.uint64), ref_p.view(np.uint64)))
                rec[f"epi{epi}_level_plane_bit_identical"] = bool(np.array_equal(e.view(np.uint64), ref_e.view(np.uint64)))
        except Exception as e:      # noqa: BLE001
            rec = dict(case=cname, sd=sd, error=repr(e))
            try:
                cuam.Destroy()
            except Exception:       # noqa: BLE001
                pass
        out["planes"][f"{cname}/sd{sd}"] = rec
        say(stage="planes", **rec)
    for k in ("AM_B200_GEMM_VARIANT", "AM_B200_SPLIT_DIGITS"):
        os.environ.pop(k, None)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
    marches(LATE)
    for k in ("AM_B200_SPLIT_EPI", "AM_B200_SPLIT_BK"):
        os.environ.pop(k, None)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
