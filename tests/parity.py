"""Shared helpers of the parity tests: run the engine through the C ABI, canonical forms, comparisons."""
import numpy as np

from analyticmesh_b200 import cuam


def key_bytes(keys_u32, state_len):
    nbytes = (state_len + 7) // 8
    raw = np.ascontiguousarray(keys_u32).view(np.uint8).reshape(keys_u32.shape[0], -1)[:, :nbytes]
    return [r.tobytes() for r in raw]


def run_engine(case, flip=False, combine=True, iso=0.0, float_type="float64"):
    """March `case` (tests/golden/cases.py) with host numpy buffers through libam_b200.so."""
    info = case["info"]
    dt = np.float64 if float_type == "float64" else np.float32
    cuam.Init(float_type=float_type, nodesnum=info.nodes, arc_table=info.arc_table,
              num_extra_constraints=len(case["b_extra"]))
    W = [np.ascontiguousarray(w, dtype=dt) for w in info.weights]
    B = [np.ascontiguousarray(b, dtype=dt) for b in info.biases]
    TM = [np.ascontiguousarray(t, dtype=dt).reshape(t.shape[0] if t.size else 0, t.shape[1] if t.size else 0)
          for t in info.arc_tm]
    cuam.AnalyticMarching(weights=W, biases=B, states=np.ascontiguousarray(case["states"], dtype=bool),
                          points=np.ascontiguousarray(case["points"], dtype=dt), arc_tm=TM,
                          w_extra_constraints=np.ascontiguousarray(case["w_extra"], dtype=dt).reshape(-1, 3),
                          b_extra_constraints=np.ascontiguousarray(case["b_extra"], dtype=dt).reshape(-1),
                          iso=iso, flip_insideout=flip)
    keys, face_off, parent, via = cuam.states()
    edges, xyz = cuam.faces()
    out = dict(keys=keys, face_off=face_off, parent=parent, via=via, edges=edges, xyz=xyz, stats=cuam.stats(),
               profile=cuam.compose_profile())
    if combine:
        cuam.CombineMesh(scale=1.0, center=[0.0, 0.0, 0.0])
        out["mesh"] = cuam.mesh()
        out["stats"] = cuam.stats()
    return out


def engine_faces(eng, state_len):
    """{key bytes: (edge tuple, verts (k,3)) or None} in the shared canonical convention."""
    kb = key_bytes(eng["keys"], state_len)
    fo = eng["face_off"]
    out = {}
    for i, k in enumerate(kb):
        a, b = int(fo[i]), int(fo[i + 1])
        out[k] = None if b - a < 3 else (tuple(int(e) for e in eng["edges"][a:b]), eng["xyz"][a:b])
    return out


def compare_with_oracle(eng, orc_res, state_len):
    from oracle import am_oracle
    ef = engine_faces(eng, state_len)
    of = am_oracle.canonical_faces(orc_res)
    e_faces = {k for k, v in ef.items() if v is not None}
    o_faces = {k for k, v in of.items() if v is not None}
    rep = dict(n_states=len(ef), n_faces=len(e_faces), oracle_states=len(of), oracle_faces=len(o_faces),
               keys_equal=(set(ef) == set(of)), face_keys_equal=(e_faces == o_faces),
               only_engine=len(set(ef) - set(of)), only_oracle=len(set(of) - set(ef)))
    loops_equal, max_err, bad = True, 0.0, 0
    for k in e_faces & o_faces:
        ge, ve = ef[k]
        go, vo = of[k]
        if ge != go:
            loops_equal = False
            bad += 1
            continue
        scale = max(1.0, float(np.abs(vo).max()))
        max_err = max(max_err, float(np.abs(ve - vo).max()) / scale)
    rep.update(loops_equal=loops_equal and rep["face_keys_equal"], loops_different=bad, max_vertex_err=max_err)
    return rep


def closed_manifold_report(eng, state_len):
    """Topology self-check on the engine output: every neuron edge (key with that bit cleared, edge id)
    must be shared by exactly two faces on a closed surface."""
    ef = engine_faces(eng, state_len)
    inc = {}
    for k, v in ef.items():
        if v is None:
            continue
        for e in v[0]:
            if e >= state_len:
                continue
            kb = bytearray(k)
            kb[e >> 3] &= ~(1 << (e & 7)) & 0xFF
            inc[(bytes(kb), e)] = inc.get((bytes(kb), e), 0) + 1
    hist = {}
    for c in inc.values():
        hist[c] = hist.get(c, 0) + 1
    return hist
