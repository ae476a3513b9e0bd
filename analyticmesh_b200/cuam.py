"""Drop-in for the reference's native extension module ``cuam`` on top of the C ABI (libam_b200.so).

Same five functions, same keyword names and the same process-global "environment" semantics as
reference backend/src/cuam.cpp:186-217 (Init / AnalyticMarching / CombineMesh / ExportMesh /
Destroy), bound with ctypes to include/am_b200.h.  Differences, all deliberate:
  * tensors may live on the GPU *or* on the host (torch tensors or numpy arrays);
  * violations raise ``RuntimeError`` (the reference's TORCH_CHECK does the same) and CUDA
    failures raise instead of exit()-ing the interpreter (reference inc/utilities.h:73-96);
  * the number of extra constraints is checked against Init;
  * extra accessors for parity tests: ``stats()``, ``states()``, ``faces()``, ``mesh()``,
    ``debug_planes()``.
There is no CPU fallback: if libam_b200.so is missing or CUDA is unavailable the call fails loudly.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libam_b200.so")
_lib = None
_handle = None          # process-global environment, like the reference's var_ptr<T>
_float_type = ""
_nodesnum = None
_arc_table = None
_num_extra = 0
_env_serial = 0         # bumped by every Init: lets a caching caller notice that the environment was replaced


class AmStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int64) for n in (
        "n_seeds", "n_unique_seeds", "n_states", "n_faces", "n_corners", "n_levels", "n_candidates", "n_unbounded",
        "n_overflow", "n_over_vertmax", "n_inconsistent", "n_vertices", "n_stitch_miss", "max_level_states",
        "n_launches")] + \
        [(n, ctypes.c_double) for n in ("seconds_march", "seconds_compose", "seconds_clip", "seconds_frontier",
                                        "compose_flops")] + [("n_tensors_reloaded", ctypes.c_int64),
                                                          ("seconds_host_wait", ctypes.c_double),
                                                          ("seconds_host_total", ctypes.c_double)]


EXPORTS = ("am_create", "am_march", "am_combine", "am_export", "am_destroy", "am_get_stats", "am_last_error",
           "am_key_words", "am_state_len", "am_copy_states", "am_copy_faces", "am_copy_mesh", "am_load_weights",
           "am_debug_planes", "am_compose_profile", "am_kernel_profile", "am_gemm_variant", "am_fp64_peak_tflops", "am_set_shard", "am_nccl_unique_id",
           "am_set_shard_nccl", "am_set_shard_p2p", "am_seed_dichotomy", "am_copy_seeds", "am_num_seeds", "am_gather_states", "am_digest", "am_edge_incidence", "am_ply_parse_faces",
           "am_ply_pack_faces")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                               "g.build()'` (there is no CPU fallback)")
        L = ctypes.CDLL(_LIB_PATH)
        L.am_last_error.restype = ctypes.c_char_p
        L.am_last_error.argtypes = [ctypes.c_void_p]
        L.am_destroy.restype = None
        L.am_destroy.argtypes = [ctypes.c_void_p]
        for f in ("am_key_words", "am_state_len"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _err(rc, what):
    if rc != 0:
        msg = lib().am_last_error(_handle)
        raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


# ---- duck-typed tensor access (torch.Tensor or numpy.ndarray) ----------------------------------

def _is_torch(t):
    return hasattr(t, "data_ptr")


def _shape(t):
    return tuple(int(s) for s in t.shape)


def _dtype_name(t):
    return str(t.dtype).replace("torch.", "")


def _ptr(t):
    if _is_torch(t):
        _check(t.is_contiguous(), "tensor must be contiguous")
        return ctypes.c_void_p(t.data_ptr() if t.numel() else None)
    _check(t.flags["C_CONTIGUOUS"], "array must be contiguous")
    return ctypes.c_void_p(t.ctypes.data if t.size else None)


def _ptr_array(ts):
    arr = (ctypes.c_void_p * max(len(ts), 1))()
    for i, t in enumerate(ts):
        arr[i] = _ptr(t)
    return arr


def Init(float_type, nodesnum, arc_table, num_extra_constraints):
    """Initialize environment (CUDA).  arc_table: CPU int32 (rows = #hidden layers, odd #cols)."""
    global _handle, _float_type, _nodesnum, _arc_table, _num_extra
    if float_type not in ("float32", "float64"):
        print("Error: `float_type` is either `float32` or `float64`!")
        return
    nodesnum = [int(v) for v in nodesnum]
    _check(nodesnum[0] == 3, "nodesnum.front() == 3")
    _check(nodesnum[-1] == 1, "nodesnum.back() == 1")
    _check(len(nodesnum) >= 3, "nodesnum.size() >= 3")
    if _is_torch(arc_table):
        _check(not arc_table.is_cuda, "arc_table must be a CPU tensor")
        _check(_dtype_name(arc_table) == "int32", "arc_table.dtype() == torch::kInt32")
        at = arc_table.contiguous().numpy()
    else:
        at = np.ascontiguousarray(arc_table)
        _check(at.dtype == np.int32, "arc_table.dtype() == int32")
    _check(at.ndim == 2, "arc_table.dim() == 2")
    _check(at.shape[0] == len(nodesnum) - 2, "arc_table.size(0) == nodesnum.size() - 2")
    _check(at.shape[1] >= 1 and (at.shape[1] + 1) % 2 == 0, "arc_table needs an odd number of columns")
    for r in range(at.shape[0]):
        _check(1 + 2 * int(at[r, 0]) <= at.shape[1], "arc_table row too short")
    _check(num_extra_constraints >= 0, "num_extra_constraints >= 0")
    if _handle is not None:
        Destroy()
    h = ctypes.c_void_p()
    nodes_c = (ctypes.c_int * len(nodesnum))(*nodesnum)
    rc = lib().am_create(ctypes.byref(h), 1 if float_type == "float64" else 0, nodes_c, len(nodesnum),
                         at.ctypes.data_as(ctypes.c_void_p), at.shape[0], at.shape[1], int(num_extra_constraints))
    if rc != 0:
        raise RuntimeError(f"Init failed ({rc}): {lib().am_last_error(None).decode()}")
    global _env_serial
    _env_serial += 1
    _handle, _float_type, _nodesnum, _arc_table, _num_extra = h, float_type, nodesnum, at.copy(), int(num_extra_constraints)


def environment_id():
    """0 when no environment is alive, else a number that changes with every Init (analyticmesh_b200.main uses it to
    validate its environment cache against direct cuam.Init / cuam.Destroy calls)."""
    return _env_serial if _handle is not None else 0


def AnalyticMarching(weights, biases, states, points, arc_tm, w_extra_constraints, b_extra_constraints, iso,
                     flip_insideout, stream=None):
    """AnalyticMarching (CUDA).  Argument contract of reference backend/src/cuam.cpp:97-184."""
    if _float_type == "":
        print("Environment must be initialized first!")
        return
    want = _float_type
    _check(len(weights) == len(biases), "weights.size() == biases.size()")
    n_fc = len(weights)
    _check(n_fc == len(_nodesnum) - 1, "fc_layers_num == nodesnum_.size() - 1")
    L = 0
    for i in range(n_fc):
        w, b = weights[i], biases[i]
        _check(len(_shape(w)) == 2 and len(_shape(b)) == 1, "weights[i].dim() == 2 and biases[i].dim() == 1")
        _check(_shape(w)[0] == _shape(b)[0], "weights[i].size(0) == biases[i].size(0)")
        if i:
            _check(_shape(w)[1] == _shape(weights[i - 1])[0], "weights[i].size(1) == weights[i-1].size(0)")
        L += _shape(w)[0]
        _check(_dtype_name(w) == want and _dtype_name(b) == want, f"weights/biases must be {want}")
        _check(_nodesnum[i + 1] == _shape(w)[0], "nodesnum_[i + 1] == weights[i].size(0)")
    _check(_shape(weights[0])[1] == 3, "weights[0].size(1) == 3")
    _check(_shape(weights[-1])[0] == 1, "weights.back().size(0) == 1")
    L -= 1
    stored = states is None and points is None          # march from the seeds of the last seed_dichotomy()
    if not stored:
        _check(len(_shape(states)) == 2 and _shape(states)[0] >= 1 and _shape(states)[1] == L,
               "states must have shape (N >= 1, hidden_states_vector_len)")
        _check(_dtype_name(states) == "bool", "states.dtype() == torch::kBool")
        _check(_shape(points) == (_shape(states)[0], 3), "points must have shape (N, 3)")
        _check(_dtype_name(points) == want, f"points must be {want}")
    tm_shapes = []
    for tm in arc_tm:
        _check(len(_shape(tm)) == 2 and _dtype_name(tm) == want, "arc_tm entries must be 2-D of the Init dtype")
        tm_shapes += list(_shape(tm))
    for i in range(_arc_table.shape[0]):
        for j in range(int(_arc_table[i, 0])):
            src, idx = int(_arc_table[i, 2 * j + 1]), int(_arc_table[i, 2 * j + 2])
            _check(idx < len(arc_tm), "arc_table references a transform that was not given")
            th, tw = _shape(arc_tm[idx])
            if th or tw:
                _check(th == _shape(weights[i + 1])[0], "arc_tm[idx].size(0) == weights[i + 1].size(0)")
                _check(tw == _shape(weights[src])[1], "arc_tm[idx].size(1) == weights[from].size(1)")
    _check(len(_shape(w_extra_constraints)) == 2 and _shape(w_extra_constraints)[1] == 3, "w_extra_constraints: (E, 3)")
    _check(len(_shape(b_extra_constraints)) == 1 and _shape(b_extra_constraints)[0] == _shape(w_extra_constraints)[0],
           "b_extra_constraints: (E,)")
    _check(_dtype_name(w_extra_constraints) == want and _dtype_name(b_extra_constraints) == want,
           f"extra constraints must be {want}")
    tm_c = (ctypes.c_int * max(len(tm_shapes), 1))(*tm_shapes)
    rc = lib().am_march(_handle, _ptr_array(weights), _ptr_array(biases), _ptr_array(arc_tm), tm_c, len(arc_tm),
                        None if stored else _ptr(states), None if stored else _ptr(points),
                        ctypes.c_int64(0 if stored else _shape(states)[0]), _ptr(w_extra_constraints),
                        _ptr(b_extra_constraints), _shape(w_extra_constraints)[0], ctypes.c_double(float(iso)),
                        int(bool(flip_insideout)), ctypes.c_void_p(stream))
    _err(rc, "AnalyticMarching")


def CombineMesh(scale, center):
    if _handle is None:
        print("AnalyticMarching must be done first!")
        return
    c = (ctypes.c_double * 3)(*[float(v) for v in center])
    rc = lib().am_combine(_handle, ctypes.c_double(float(scale)), c)
    if rc == -2:
        print("AnalyticMarching must be done first!")
        return
    _err(rc, "CombineMesh")


def ExportMesh(file_path, is_polymesh, is_float32):
    if _handle is None:
        print("CombineMesh must be done first!")
        return
    rc = lib().am_export(_handle, str(file_path).encode(), int(bool(is_polymesh)), int(bool(is_float32)))
    if rc == -2:
        print("CombineMesh must be done first!")
        return
    _err(rc, "ExportMesh")


def Destroy():
    global _handle, _float_type
    if _handle is None:
        print("Environment must be initialized first!")
        return
    lib().am_destroy(_handle)
    _handle = None
    _float_type = ""


# ---- accessors beyond the reference's five functions -------------------------------------------

def stats():
    s = AmStats()
    _err(lib().am_get_stats(_handle, ctypes.byref(s)), "stats")
    return {n: getattr(s, n) for n, _ in AmStats._fields_}


def compose_profile():
    ms, n, fl = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
    _err(lib().am_compose_profile(_handle, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl)), "compose_profile")
    return dict(ms_total=ms.value, launches=n.value, flops=fl.value)


def kernel_profile(kind):
    """Device ms / launches / flops of one timed span kind of the last march (include/am_b200.h: am_kernel_profile)."""
    ms, n, fl = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
    _err(lib().am_kernel_profile(_handle, int(kind), ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl)),
         "kernel_profile")
    return dict(ms_total=ms.value, launches=n.value, flops=fl.value)


def gemm_variant():
    """(variant, split digits): 0/1 FP64 DMMA tiles, 2 tcgen05 int8 split-integer path."""
    d = ctypes.c_int()
    lib().am_gemm_variant.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    return int(lib().am_gemm_variant(_handle, ctypes.byref(d))), d.value


def states():
    """keys (n, kw) uint32, face_off (n+1,) int64, parent (n,), via_edge (n,) of every visited state."""
    st = stats()
    n, kw = st["n_states"], lib().am_key_words(_handle)
    keys = np.zeros((n, kw), dtype=np.uint32)
    face_off = np.zeros(n + 1, dtype=np.int64)
    parent = np.zeros(n, dtype=np.int32)
    via = np.zeros(n, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    _err(lib().am_copy_states(_handle, p(keys), p(face_off), p(parent), p(via)), "states")
    return keys, face_off, parent, via


def faces():
    """edge ids (corners,) int32 and vertices (corners, 3) float64 of all polygons, CSR by states()[1]."""
    n = stats()["n_corners"]
    e = np.zeros(n, dtype=np.int32)
    v = np.zeros((n, 3), dtype=np.float64)
    _err(lib().am_copy_faces(_handle, e.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p)), "faces")
    return e, v


def mesh():
    """(vertices (V,3) float64, face sizes (F,), flat face indices) after CombineMesh."""
    st = stats()
    v = np.zeros((st["n_vertices"], 3), dtype=np.float64)
    fs = np.zeros(st["n_faces"], dtype=np.int32)
    fi = np.zeros(st["n_corners"], dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    _err(lib().am_copy_mesh(_handle, p(v), p(fs), p(fi)), "mesh")
    return v, fs, fi


class AmSeedReport(ctypes.Structure):
    _fields_ = [("n_points", ctypes.c_int64), ("rounds", ctypes.c_int64), ("iterations", ctypes.c_int64),
                ("avg_abs_error", ctypes.c_double), ("seconds", ctypes.c_double)]


def seed_dichotomy(weights, biases, arc_tm, w_extra_constraints, b_extra_constraints, iso, init_num=1024,
                   try_pts_num=4096, init_ball_radius=1.0, iter_max=100, avg_eps=1e-3, seed=0):
    """The surface-point initialiser on the device (include/am_b200.h: am_seed_dichotomy; reference
    backend/main.py:252-326).  The seeds stay on the device: AnalyticMarching(states=None, points=None, ...) marches
    from them; seeds() copies them out.  Returns the report dict."""
    tm_shapes = []
    for tm in arc_tm:
        tm_shapes += list(_shape(tm))
    tm_c = (ctypes.c_int * max(len(tm_shapes), 1))(*tm_shapes)
    rep = AmSeedReport()
    fn = lib().am_seed_dichotomy
    fn.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_uint64, ctypes.c_void_p]
    _err(fn(_handle, _ptr_array(weights), _ptr_array(biases), _ptr_array(arc_tm), tm_c, len(arc_tm),
            _ptr(w_extra_constraints), _ptr(b_extra_constraints), _shape(w_extra_constraints)[0], float(iso), int(init_num),
            int(try_pts_num), float(init_ball_radius), int(iter_max), float(avg_eps), int(seed) & (2**64 - 1),
            ctypes.byref(rep)), "seed_dichotomy")
    return {n: getattr(rep, n) for n, _ in AmSeedReport._fields_}


def seeds():
    """(points (N, 3) float64, states (N, L) bool) found by the last seed_dichotomy()."""
    L = lib().am_state_len(_handle)
    fn = lib().am_copy_seeds
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    nfn = lib().am_num_seeds
    nfn.argtypes = [ctypes.c_void_p]
    nfn.restype = ctypes.c_int64
    n = int(nfn(_handle))
    pts = np.zeros((n, 3), dtype=np.float64)
    st = np.zeros((n, L), dtype=np.uint8)
    _err(fn(_handle, pts.ctypes.data_as(ctypes.c_void_p), st.ctypes.data_as(ctypes.c_void_p)), "seeds")
    return pts, st.astype(bool)


def gather_states(ids):
    """The listed states only: dict(keys (n, kw), counts (n,), edges (n, 32), xyz (n, 32, 3), parent, via, seedpt (n, 4))."""
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    n, kw = len(ids), lib().am_key_words(_handle)
    out = dict(keys=np.zeros((n, kw), np.uint32), counts=np.zeros(n, np.int32), edges=np.zeros((n, 32), np.int32),
               xyz=np.zeros((n, 32, 3), np.float64), parent=np.zeros(n, np.int32), via=np.zeros(n, np.int32),
               seedpt=np.zeros((n, 4), np.float64))
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    fn = lib().am_gather_states
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 7
    _err(fn(_handle, p(ids), n, p(out["keys"]), p(out["counts"]), p(out["edges"]), p(out["xyz"]), p(out["parent"]),
            p(out["via"]), p(out["seedpt"])), "gather_states")
    return out


def digest():
    """Device-side checksums of the last march (include/am_b200.h: am_digest) as a dict; `ordered` / `region_set`
    are hex strings that two runs can compare (1 GPU vs N GPUs, two builds)."""
    import hashlib
    buf = (ctypes.c_uint64 * 8)()
    fn = lib().am_digest
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    _err(fn(_handle, buf), "digest")
    v = [int(x) for x in buf]
    hx = lambda xs: hashlib.sha256(b"".join(int(x).to_bytes(8, "little") for x in xs)).hexdigest()[:32]  # noqa: E731
    return dict(ordered=hx(v[0:4] + v[6:8]), region_set=hx([v[4], v[6], v[7]]), region_set_with_vertices=hx([v[5], v[6], v[7]]),
                topology_sum=v[4], n_states=v[6], n_corners=v[7], raw=v)


def edge_incidence():
    """dict(boundary, matched, neighbour_missing, neighbour_without_edge) -- include/am_b200.h: am_edge_incidence."""
    buf = (ctypes.c_int64 * 4)()
    fn = lib().am_edge_incidence
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    _err(fn(_handle, buf), "edge_incidence")
    return dict(boundary=int(buf[0]), matched=int(buf[1]), neighbour_missing=int(buf[2]), neighbour_without_edge=int(buf[3]))


ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)
_shard_cb = None   # keeps the ctypes callback alive


def set_shard(rank, world, allreduce):
    """Spread ONE march over `world` processes (one GPU each).  `allreduce(device_ptr, n_int32, stream)` must
    sum the int32 buffer over all ranks in place, ordered on the given CUDA stream; see
    analyticmesh_b200.parallel.make_allreduce for the torch.distributed implementation."""
    global _shard_cb

    def _cb(_user, ptr, n, stream):
        try:
            allreduce(ptr, n, stream)
            return 0
        except Exception as e:  # noqa: BLE001 - reported through the C ABI as a failed callback
            print(f"(cuam) all-reduce callback failed: {e!r}")
            return 1

    _shard_cb = ALLREDUCE_FN(_cb) if world > 1 else None
    fn = lib().am_set_shard
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    _err(fn(_handle, int(rank), int(world), ctypes.cast(_shard_cb, ctypes.c_void_p) if _shard_cb else None, None),
         "set_shard")


def set_shard_nccl(rank, world, broadcast_bytes):
    """Spread ONE march over `world` processes with the per-level all-reduce issued by the library itself
    (ncclAllReduce on the engine's stream).  `broadcast_bytes(b: bytes | None) -> bytes` must return rank 0's
    128-byte NCCL unique id on every rank (analyticmesh_b200.parallel.broadcast_bytes)."""
    buf = (ctypes.c_char * 128)()
    if rank == 0:
        _err(lib().am_nccl_unique_id(buf), "nccl_unique_id")
    uid = broadcast_bytes(bytes(buf) if rank == 0 else None)
    fn = lib().am_set_shard_nccl
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
    _err(fn(_handle, int(rank), int(world), uid), "set_shard_nccl")


ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64)
_gather_cb = None


def set_shard_p2p(rank, world, allgather):
    """Spread ONE march over `world` processes (one GPU each) with the engine's own exchange over NVLink peer
    memory (include/am_b200.h: am_set_shard_p2p).  `allgather(payload: bytes) -> list[bytes]` (rank order) is
    used during set-up only, to swap the CUDA IPC handles (analyticmesh_b200.parallel.make_allgather)."""
    global _gather_cb

    def _cb(_user, send, recv, nbytes):
        try:
            parts = allgather(ctypes.string_at(send, nbytes))
            assert len(parts) == world and all(len(b) == nbytes for b in parts)
            ctypes.memmove(recv, b"".join(parts), nbytes * world)
            return 0
        except Exception as e:  # noqa: BLE001 - reported through the C ABI as a failed callback
            print(f"(cuam) all-gather callback failed: {e!r}")
            return 1

    _gather_cb = ALLGATHER_FN(_cb)
    fn = lib().am_set_shard_p2p
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    _err(fn(_handle, int(rank), int(world), ctypes.cast(_gather_cb, ctypes.c_void_p), None), "set_shard_p2p")


def fp64_peak_tflops():
    f = lib().am_fp64_peak_tflops
    f.restype = ctypes.c_double
    return float(f())


def load_weights(weights, biases, arc_tm):
    tm_shapes = []
    for tm in arc_tm:
        tm_shapes += list(_shape(tm))
    tm_c = (ctypes.c_int * max(len(tm_shapes), 1))(*tm_shapes)
    _err(lib().am_load_weights(_handle, _ptr_array(weights), _ptr_array(biases), _ptr_array(arc_tm), tm_c, len(arc_tm)),
         "load_weights")


def debug_planes(states_bool, iso=0.0):
    """Unsigned plane rows (n, L, 4) and level planes (n, 4) from the composition kernels only."""
    st = np.ascontiguousarray(states_bool, dtype=np.uint8)
    n, L = st.shape
    dt = np.float64 if _float_type == "float64" else np.float32
    planes = np.zeros((n, L, 4), dtype=dt)
    equ = np.zeros((n, 4), dtype=dt)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    _err(lib().am_debug_planes(_handle, p(st), ctypes.c_int64(n), ctypes.c_double(iso), p(planes), p(equ)),
         "debug_planes")
    return planes, equ
