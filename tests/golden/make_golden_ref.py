"""Generates tests/golden/ref_<case>.json by running the UNMODIFIED reference CUDA build on a B200.

Run on the GPU box (the reference needs a GPU; this container has none):
    gpurun -- python tests/golden/make_golden_ref.py [case ...]
It imports the reference package staged under baseline/_ref/AnalyticMesh (a verbatim copy of
/root/reference plus its compiled `cuam.so`, git-ignored, built with the recipe in SURVEY App. C),
feeds it the fixed inputs of tests/golden/cases.py through the reference's own extension API
(`cuam.Init / AnalyticMarching / CombineMesh / ExportMesh`, reference backend/main.py:441-471),
and reduces the PLY it writes to a canonical, engine-independent form:

  * one activation-pattern key per polygon = pattern of the float64 network at the face centroid;
  * per face: vertex loop as written by the reference (float64);
  * digests (sha256 over the sorted keys) for the big cases, full listings for the small ones.

Outputs land in gpurun_out/golden/ and are copied into tests/golden/ by hand.
"""
import hashlib
import json
import os
import subprocess
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.path.join(REPO, "baseline", "_ref")
OUT = os.path.join(REPO, "gpurun_out", "golden")
sys.path.insert(0, REPO)

FULL_LISTING_MAX = 2000      # faces
SAMPLE_FACES = 512


def import_reference():
    onnx = types.ModuleType("onnx")
    nh = types.ModuleType("onnx.numpy_helper")
    nh.to_array = None
    hp = types.ModuleType("onnx.helper")
    onnx.numpy_helper, onnx.helper = nh, hp
    sys.modules.update({"onnx": onnx, "onnx.numpy_helper": nh, "onnx.helper": hp})
    sys.path.insert(0, os.path.join(REF, "AnalyticMesh", "backend"))
    sys.path.insert(0, REF)
    import AnalyticMesh as AM  # noqa
    return AM


def read_poly_ply(path):
    with open(path, "rb") as f:
        data = f.read()
    hdr_end = data.index(b"end_header\n") + len(b"end_header\n")
    head = data[:hdr_end].decode()
    nv = int([l for l in head.splitlines() if l.startswith("element vertex")][0].split()[2])
    nf = int([l for l in head.splitlines() if l.startswith("element face")][0].split()[2])
    dbl = "property double x" in head
    V = np.frombuffer(data, dtype="<f8" if dbl else "<f4", count=3 * nv, offset=hdr_end).reshape(nv, 3).astype(np.float64)
    off = hdr_end + 3 * nv * (8 if dbl else 4)
    body = np.frombuffer(data, dtype=np.uint8, offset=off)
    counts = np.zeros(nf, dtype=np.int64)
    starts = np.zeros(nf, dtype=np.int64)
    pos = 0
    for i in range(nf):
        k = int(body[pos])
        counts[i] = k
        starts[i] = pos + 1
        pos += 1 + 4 * k
    idx = np.concatenate([np.frombuffer(body[s:s + 4 * k].tobytes(), dtype="<i4") for s, k in zip(starts, counts)]) \
        if nf else np.zeros(0, dtype=np.int32)
    return V, counts, idx


def canonicalise(info, V, counts, idx):
    offs = np.concatenate([[0], np.cumsum(counts)])
    P = V[idx]
    cent = np.add.reduceat(P, offs[:-1], axis=0) / counts[:, None] if len(counts) else np.zeros((0, 3))
    _, bits = info.forward(cent)
    keys = np.packbits(bits, axis=1, bitorder="little")
    return offs, P, keys


def worker(name):
    import torch
    from tests.golden.cases import build_case
    AM = import_reference()
    import importlib
    cuam = importlib.import_module('build.cuam')
    c = build_case(name)
    model, info = c["model"], c["info"]
    dt = torch.float64
    mi = model.get_info()
    weights = [w.detach().to(dt).cuda().contiguous() for w in mi['weights']]
    biases = [b.detach().to(dt).cuda().contiguous() for b in mi['biases']]
    arc_tm = [t.detach().to(dt).cuda().contiguous() for t in mi['arc_tm']]
    arc_table = mi['arc_table'].to(torch.int32).cpu().contiguous()
    states = torch.from_numpy(c["states"]).to(torch.bool).cuda().contiguous()
    points = torch.from_numpy(c["points"]).to(dt).cuda().contiguous()
    w_extra = torch.from_numpy(np.ascontiguousarray(c["w_extra"])).to(dt).reshape(-1, 3).cuda().contiguous()
    b_extra = torch.from_numpy(np.ascontiguousarray(c["b_extra"])).to(dt).reshape(-1).cuda().contiguous()
    wsha = hashlib.sha256(b"".join(np.ascontiguousarray(w).tobytes() for w in info.weights)).hexdigest()

    t0 = time.time()
    cuam.Init(float_type="float64", nodesnum=list(model.nodes), arc_table=arc_table,
              num_extra_constraints=int(b_extra.shape[0]))
    torch.cuda.synchronize()
    init_time = time.time() - t0
    am_times = []
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.time()
        cuam.AnalyticMarching(weights=weights, biases=biases, states=states, points=points, arc_tm=arc_tm,
                              w_extra_constraints=w_extra, b_extra_constraints=b_extra, iso=0.0, flip_insideout=False)
        torch.cuda.synchronize()
        am_times.append(time.time() - t0)
    t0 = time.time()
    cuam.CombineMesh(scale=1.0, center=[0.0, 0.0, 0.0])
    ply = os.path.join(OUT, f"ref_{name}.ply")
    cuam.ExportMesh(file_path=ply, is_polymesh=True, is_float32=False)
    export_time = time.time() - t0
    cuam.Destroy()

    V, counts, idx = read_poly_ply(ply)
    offs, P, keys = canonicalise(info, V, counts, idx)
    nbytes = keys.shape[1] if len(keys) else 0
    key_list = [k.tobytes() for k in keys]
    order = sorted(range(len(key_list)), key=lambda i: key_list[i])
    digest = hashlib.sha256(b"".join(key_list[i] for i in order)).hexdigest()
    hist = {int(k): int(v) for k, v in zip(*np.unique(counts, return_counts=True))}
    f_abs = np.abs(info.forward(V)[0]) if len(V) else np.zeros(0)
    gold = dict(case=name, engine="reference CUDA build (sm_100) on " + torch.cuda.get_device_name(0),
                nodes=list(model.nodes), state_len=info.state_len, n_seeds=int(states.shape[0]),
                weights_sha256=wsha, n_faces=int(len(counts)), n_vertices=int(len(V)),
                n_unique_keys=len(set(key_list)), key_bytes=nbytes, keys_sha256=digest, poly_size_hist=hist,
                max_abs_f=float(f_abs.max()) if len(f_abs) else 0.0,
                timing=dict(init_cuda_time=init_time, am_time=am_times, export_time=export_time,
                            faces_per_sec=len(counts) / min(am_times)))
    if len(counts) <= FULL_LISTING_MAX:
        pick = order
    else:
        rng = np.random.RandomState(0)
        pick = sorted(rng.choice(len(counts), SAMPLE_FACES, replace=False).tolist(), key=lambda i: key_list[i])
    gold["faces"] = [dict(key=key_list[i].hex(), verts=P[offs[i]:offs[i + 1]].tolist()) for i in pick]
    gold["faces_are_complete"] = len(counts) <= FULL_LISTING_MAX
    with open(os.path.join(OUT, f"ref_{name}.json"), "w") as f:
        json.dump(gold, f)
    os.remove(ply)
    print("GOLD", json.dumps({k: v for k, v in gold.items() if k != "faces"}), flush=True)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        worker(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or ["polytope", "chair_cube", "skipnet", "chair", "mlp4x128s", "sphere", "mlp8x512s_cube"]
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], timeout=420,
                               capture_output=True, text=True)
            tail = p.stdout[-3000:] + "\n--stderr--\n" + p.stderr[-2000:]
            rc = p.returncode
        except subprocess.TimeoutExpired as e:
            tail, rc = "TIMEOUT " + str(e.stdout)[-2000:], "timeout"
        with open(os.path.join(OUT, f"ref_{name}.log"), "w") as f:
            f.write(tail)
        gl = [l for l in tail.splitlines() if l.startswith("GOLD")]
        print("==", name, rc, round(time.time() - t0, 1), gl[-1][:1500] if gl else tail[-1500:], flush=True)
