"""ctypes wrapper around oracle/am_oracle.c.  TEST INFRASTRUCTURE ONLY (see the C file's header):
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libam_oracle.so")
VERT_MAX = 20
TAB_LEN = VERT_MAX + 3
VERT_LEN = 1 + 3 * VERT_MAX
_lib = None


def build(force=False):
    src = os.path.join(HERE, "am_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        for p in ("amo64", "amo32"):
            getattr(_lib, p + "_march").restype = ctypes.c_void_p
            getattr(_lib, p + "_process_states").restype = ctypes.c_void_p
            getattr(_lib, p + "_n_states").restype = ctypes.c_int64
            getattr(_lib, p + "_n_faces").restype = ctypes.c_int64
            getattr(_lib, p + "_seconds").restype = ctypes.c_double
            for f in ("_n_states", "_n_faces", "_seconds", "_key_words", "_free_result"):
                getattr(_lib, p + f).argtypes = [ctypes.c_void_p]
    return _lib


def _ptr_array(arrs):
    ptrs = (ctypes.c_void_p * max(len(arrs), 1))()
    for i, a in enumerate(arrs):
        ptrs[i] = a.ctypes.data if a.size else None
    return ptrs


def _net_args(info):
    dt = info.dtype
    W = [np.ascontiguousarray(w, dtype=dt) for w in info.weights]
    B = [np.ascontiguousarray(b, dtype=dt) for b in info.biases]
    TM = [np.ascontiguousarray(t, dtype=dt) for t in info.arc_tm]
    nodes = np.asarray(info.nodes, dtype=np.int32)
    arc = np.ascontiguousarray(info.arc_table, dtype=np.int32)
    tm_shape = np.asarray([[t.shape[0], t.shape[1]] if t.size else [0, 0] for t in TM], dtype=np.int32).reshape(-1)
    keep = (W, B, TM, nodes, arc, tm_shape)
    args = [nodes.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(nodes)), arc.ctypes.data_as(ctypes.c_void_p),
            ctypes.c_int(arc.shape[1]), _ptr_array(W), _ptr_array(B), _ptr_array(TM),
            tm_shape.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(TM))]
    return args, keep


def compose(info, state_bool, iso=0.0):
    """(planes [L,4] unsigned rows, equ [4]) of one activation pattern."""
    L = info.state_len
    pfx = "amo64" if info.dtype == np.float64 else "amo32"
    args, keep = _net_args(info)
    st = np.ascontiguousarray(state_bool, dtype=np.uint8).reshape(L)
    planes = np.zeros((L, 4), dtype=info.dtype)
    equ = np.zeros(4, dtype=info.dtype)
    getattr(lib(), pfx + "_compose")(*args, ctypes.c_double(iso), st.ctypes.data_as(ctypes.c_void_p),
                                      planes.ctypes.data_as(ctypes.c_void_p), equ.ctypes.data_as(ctypes.c_void_p))
    return planes, equ


def march(info, states_bool, points, w_extra=None, b_extra=None, iso=0.0, flip=False,
          quirk_drop_output_skip=False, max_states=0, threads=0):
    """Run the restated reference algorithm.  Returns a dict of numpy arrays in processing order."""
    dt = info.dtype
    pfx = "amo64" if dt == np.float64 else "amo32"
    L = info.state_len
    st = np.ascontiguousarray(states_bool, dtype=np.uint8).reshape(-1, L)
    pts = np.ascontiguousarray(points, dtype=dt).reshape(-1, 3)
    assert st.shape[0] == pts.shape[0]
    we = np.zeros((0, 3), dtype=dt) if w_extra is None else np.ascontiguousarray(w_extra, dtype=dt).reshape(-1, 3)
    be = np.zeros((0,), dtype=dt) if b_extra is None else np.ascontiguousarray(b_extra, dtype=dt).reshape(-1)
    args, keep = _net_args(info)
    L_ = lib()
    res = getattr(L_, pfx + "_march")(
        *args, st.ctypes.data_as(ctypes.c_void_p), pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(st.shape[0]),
        we.ctypes.data_as(ctypes.c_void_p), be.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(we.shape[0]),
        ctypes.c_double(iso), ctypes.c_int(int(flip)), ctypes.c_int(int(quirk_drop_output_skip)),
        ctypes.c_int64(max_states), ctypes.c_int(threads))
    res = ctypes.c_void_p(res)
    n = getattr(L_, pfx + "_n_states")(res)
    kw = getattr(L_, pfx + "_key_words")(res)
    out = dict(n_states=n, n_faces=getattr(L_, pfx + "_n_faces")(res), seconds=getattr(L_, pfx + "_seconds")(res),
               keys=np.zeros((n, kw), dtype=np.uint32), tab=np.zeros((n, TAB_LEN), dtype=np.int32),
               verts=np.zeros((n, VERT_LEN), dtype=dt), start=np.zeros(n, dtype=np.int32),
               equ=np.zeros((n, 4), dtype=dt), state_len=L)
    getattr(L_, pfx + "_copy_out")(res, *[out[k].ctypes.data_as(ctypes.c_void_p)
                                          for k in ("keys", "tab", "verts", "start", "equ")])
    getattr(L_, pfx + "_free_result")(res)
    return out


def _collect(L_, pfx, res, dt, L):
    res = ctypes.c_void_p(res)
    n = getattr(L_, pfx + "_n_states")(res)
    kw = getattr(L_, pfx + "_key_words")(res)
    out = dict(n_states=n, n_faces=getattr(L_, pfx + "_n_faces")(res), seconds=getattr(L_, pfx + "_seconds")(res),
               keys=np.zeros((n, kw), dtype=np.uint32), tab=np.zeros((n, TAB_LEN), dtype=np.int32),
               verts=np.zeros((n, VERT_LEN), dtype=dt), start=np.zeros(n, dtype=np.int32),
               equ=np.zeros((n, 4), dtype=dt), state_len=L)
    getattr(L_, pfx + "_copy_out")(res, *[out[k].ctypes.data_as(ctypes.c_void_p)
                                          for k in ("keys", "tab", "verts", "start", "equ")])
    getattr(L_, pfx + "_free_result")(res)
    return out


def process_states(info, keys_u32, midpoints, starts, w_extra=None, b_extra=None, iso=0.0, flip=False, threads=0):
    """The per-state stages of the reference for an explicit list of states: keys (n, >= ceil(L/32)) uint32,
    midpoints (n, 3) the point on the shared edge, starts (n,) start edge or -1.  Same result dict as march()."""
    dt = info.dtype
    pfx = "amo64" if dt == np.float64 else "amo32"
    L = info.state_len
    kw = (L + 31) // 32
    keys = np.ascontiguousarray(np.asarray(keys_u32, dtype=np.uint32)[:, :kw])
    mid = np.ascontiguousarray(midpoints, dtype=dt).reshape(-1, 3)
    st = np.ascontiguousarray(starts, dtype=np.int32).reshape(-1)
    assert keys.shape[0] == mid.shape[0] == st.shape[0]
    we = np.zeros((0, 3), dtype=dt) if w_extra is None else np.ascontiguousarray(w_extra, dtype=dt).reshape(-1, 3)
    be = np.zeros((0,), dtype=dt) if b_extra is None else np.ascontiguousarray(b_extra, dtype=dt).reshape(-1)
    args, keep = _net_args(info)
    L_ = lib()
    res = getattr(L_, pfx + "_process_states")(
        *args, keys.ctypes.data_as(ctypes.c_void_p), mid.ctypes.data_as(ctypes.c_void_p),
        st.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(keys.shape[0]), we.ctypes.data_as(ctypes.c_void_p),
        be.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(we.shape[0]), ctypes.c_double(iso), ctypes.c_int(int(flip)),
        ctypes.c_int(threads))
    return _collect(L_, pfx, res, dt, L)


_M64 = (1 << 64) - 1


def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def topology_sum(faces, state_len):
    """Order-independent checksum of {key: edge loop} -- the value the engine computes on the device
    (analyticmesh_b200/csrc/mesh.cuh digest_states_kernel, acc[0]); `faces` as returned by canonical_faces()
    or tests.parity.engine_faces(): {key bytes: (edge ids, vertices) or None}."""
    n_words = (state_len + 31) // 32
    total = 0
    for key, val in faces.items():
        words = np.frombuffer(key + b"\0" * (4 * n_words - len(key)), dtype="<u4")
        h = 0
        for w in range(n_words):
            h = (h + _splitmix64(((w + 1) << 32) | int(words[w]))) & _M64
        edges = () if val is None else val[0]
        for j, e in enumerate(edges):
            h = (h + _splitmix64(((int(e) & 0xFFFFFFFF) + _splitmix64(0xE0000000 + j)) & _M64)) & _M64
        h = (h + _splitmix64(len(edges))) & _M64
        total = (total + _splitmix64(h)) & _M64
    return total


def canonical_faces(res):
    """{key bytes: (edge ids tuple, vertices (k,3))} for every state with a closed polygon.

    Convention shared with the engine: vertices v_0..v_{k-1} in output orientation, edge id g_i
    is the constraint carrying the segment v_i -> v_{i+1}; the cycle is rotated so that g_0 is
    the smallest id.  States without a polygon map to None."""
    out = {}
    L = res["state_len"]
    nbytes = (L + 7) // 8
    for o in range(res["n_states"]):
        key = res["keys"][o].tobytes()[:nbytes]
        tab = res["tab"][o]
        n = int(tab[0])
        nv = int(res["verts"][o, 0])
        ent = [int(t) for t in tab[1:1 + n] if t >= 0]
        if nv < 3 or len(ent) != nv + 1 or ent[0] != ent[-1]:
            out[key] = None
            continue
        e = ent[:-1]
        k = nv
        v = res["verts"][o, 1:1 + 3 * k].reshape(k, 3)
        if tab[TAB_LEN - 1] == 1:   # vertex loop was reversed: v'_i = v_{k-1-i}, stored already reversed
            g = [e[(k - 1 - i) % k] for i in range(k)]
        else:
            g = [e[(i + 1) % k] for i in range(k)]
        r = int(np.argmin(g))
        out[key] = (tuple(g[r:] + g[:r]), np.concatenate([v[r:], v[:r]], axis=0))
    return out
