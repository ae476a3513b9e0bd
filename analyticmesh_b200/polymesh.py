"""``PolyMesh`` -- polygon-mesh container with the API of the reference's ``polylib.PolyMesh``
(reference backend/libpolytools/src/polylib.cpp:575-596): load/save binary-little-endian PLY and
ascii OFF, ``poly2tri`` (fan triangulation, index-degenerate triangles dropped), face counts.
Arrays inside (numpy), lists at the API like the pybind11 original; the two variable-length record
loops are native (csrc/polyio.cpp, exported by libam_b200.so)."""
import ctypes

import numpy as np

from . import cuam as _cuam


def _lib():
    L = _cuam.lib()
    L.am_ply_parse_faces.restype = ctypes.c_int64
    L.am_ply_pack_faces.restype = ctypes.c_int64
    return L


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class PolyMesh:
    def __init__(self, load_path=None, vertices=None, faces=None, colors=None):
        self.clear()
        if load_path is not None:
            self.load(load_path)
        elif vertices is not None:
            self.set_vertices(vertices)
            self.set_faces(faces or [])
            self.set_colors(colors or [])

    @classmethod
    def from_arrays(cls, vertices, face_sizes, face_index, colors=None):
        """Build from flat arrays (no python lists): vertices (V, 3), face_sizes (F,), face_index (sum of sizes,) --
        the layout cuam.mesh() returns, used by the voxel mode to merge the per-voxel meshes in memory."""
        m = cls()
        m._v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 3)
        m._cnt = np.ascontiguousarray(face_sizes, dtype=np.int32).reshape(-1)
        m._idx = np.ascontiguousarray(face_index, dtype=np.int32).reshape(-1)
        assert int(m._cnt.sum()) == len(m._idx)
        m._col = np.zeros((0, 3), dtype=np.uint8) if colors is None else np.ascontiguousarray(colors, dtype=np.uint8)
        return m

    # ------------------------------------------------------------------ storage
    def clear(self):
        self._v = np.zeros((0, 3), dtype=np.float32)
        self._cnt = np.zeros(0, dtype=np.int32)
        self._idx = np.zeros(0, dtype=np.int32)
        self._col = np.zeros((0, 3), dtype=np.uint8)

    def vertices(self):
        return self._v.tolist()

    def faces(self):
        off = np.concatenate([[0], np.cumsum(self._cnt)])
        idx = self._idx.tolist()
        return [idx[off[i]:off[i + 1]] for i in range(len(self._cnt))]

    def colors(self):
        return self._col.tolist()

    def set_vertices(self, vertices):
        self._v = np.asarray(vertices, dtype=np.float32).reshape(-1, 3)

    def set_faces(self, faces):
        self._cnt = np.asarray([len(f) for f in faces], dtype=np.int32)
        self._idx = np.asarray([i for f in faces for i in f], dtype=np.int32)

    def set_colors(self, colors):
        self._col = np.asarray(colors, dtype=np.uint8).reshape(-1, 3)

    # ------------------------------------------------------------------ queries
    def is_polymesh(self):
        return len(self._cnt) == 0 or bool((self._cnt > 3).any())

    def num_polyfaces(self):
        return int(len(self._cnt))

    def num_trifaces(self):
        # the reference does not drop degenerate triangles when counting (polylib.cpp:553-569)
        return int((self._cnt - 2).sum()) if self.is_polymesh() else int(len(self._cnt))

    def poly2tri(self):
        if len(self._cnt) == 0:
            return
        off = np.concatenate([[0], np.cumsum(self._cnt)])[:-1]
        ntri = np.maximum(self._cnt - 2, 0)
        face_of = np.repeat(np.arange(len(self._cnt)), ntri)
        j = np.arange(int(ntri.sum())) - np.repeat(np.concatenate([[0], np.cumsum(ntri)])[:-1], ntri)
        a = self._idx[off[face_of]]
        b = self._idx[off[face_of] + j + 1]
        c = self._idx[off[face_of] + j + 2]
        keep = (a != b) & (b != c) & (a != c)
        self._idx = np.stack([a[keep], b[keep], c[keep]], axis=1).reshape(-1).astype(np.int32)
        self._cnt = np.full(int(keep.sum()), 3, dtype=np.int32)
        if len(self._col):
            self._col = self._col[face_of[keep]]

    # ------------------------------------------------------------------ IO
    def load(self, load_path):
        ext = load_path.rsplit(".", 1)[-1].lower()
        if ext == "ply":
            self._load_ply(load_path)
        elif ext == "off":
            self._load_off(load_path)
        else:
            raise RuntimeError("Unsupported file format.")

    def save(self, save_path):
        ext = save_path.rsplit(".", 1)[-1].lower()
        if ext == "ply":
            self._save_ply(save_path)
        elif ext == "off":
            self._save_off(save_path)
        else:
            raise RuntimeError("Unsupported file format.")

    def _load_ply(self, path):
        self.clear()
        with open(path, "rb") as f:
            data = f.read()
        end = data.find(b"end_header")
        if not data[:3].lower() == b"ply" or end < 0:
            raise RuntimeError("It should be a ply file!")
        end = data.index(b"\n", end) + 1
        head = data[:end].decode("ascii", "replace").lower().splitlines()
        if not any("binary_little_endian" in l for l in head):
            raise RuntimeError("We only support binary_little_endian ply file!")
        nv = nf = 0
        vtype, colours, in_face = "float", 0, False
        for i, l in enumerate(head):
            t = l.split()
            if len(t) == 3 and t[0] == "element" and t[1] == "vertex":
                nv = int(t[2])
                vtype = "double" if "double" in head[i + 1] else "float"
            elif len(t) == 3 and t[0] == "element" and t[1] == "face":
                nf, in_face = int(t[2]), True
            elif in_face and t[:1] == ["property"] and t[-1] in ("red", "green", "blue"):
                colours += 1
        if colours not in (0, 3):
            raise RuntimeError("Incomplete colors.")
        vs = 8 if vtype == "double" else 4
        self._v = np.frombuffer(data, dtype="<f8" if vs == 8 else "<f4", count=3 * nv, offset=end) \
            .reshape(nv, 3).astype(np.float32)
        body = np.frombuffer(data, dtype=np.uint8, offset=end + 3 * nv * vs)
        L = _lib()
        cnt = np.zeros(nf, dtype=np.int32)
        n_idx = ctypes.c_int64()
        if L.am_ply_parse_faces(_p(body), ctypes.c_int64(len(body)), ctypes.c_int64(nf), int(colours == 3), _p(cnt),
                                None, ctypes.c_int64(0), None, ctypes.byref(n_idx)) < 0:
            raise RuntimeError("truncated ply file")
        idx = np.zeros(n_idx.value, dtype=np.int32)
        col = np.zeros((nf if colours else 0, 3), dtype=np.uint8)
        L.am_ply_parse_faces(_p(body), ctypes.c_int64(len(body)), ctypes.c_int64(nf), int(colours == 3), _p(cnt),
                             _p(idx), ctypes.c_int64(len(idx)), _p(col) if colours else None, ctypes.byref(n_idx))
        self._cnt, self._idx, self._col = cnt, idx, col

    def _save_ply(self, path):
        has_col = len(self._col) > 0
        head = ("ply\nformat binary_little_endian 1.0\n"
                f"element vertex {len(self._v)}\nproperty float x\nproperty float y\nproperty float z\n"
                f"element face {len(self._cnt)}\nproperty list uchar int vertex_index\n")
        if has_col:
            head += "property uchar red\nproperty uchar green\nproperty uchar blue\n"
        head += "end_header\n"
        L = _lib()
        col = np.ascontiguousarray(self._col) if has_col else None
        n = L.am_ply_pack_faces(_p(self._cnt), _p(self._idx), ctypes.c_int64(len(self._cnt)), _p(col), None)
        out = np.zeros(n, dtype=np.uint8)
        L.am_ply_pack_faces(_p(self._cnt), _p(self._idx), ctypes.c_int64(len(self._cnt)), _p(col), _p(out))
        with open(path, "wb") as f:
            f.write(head.encode())
            f.write(np.ascontiguousarray(self._v, dtype="<f4").tobytes())
            f.write(out.tobytes())

    def _load_off(self, path):
        self.clear()
        with open(path) as f:
            lines = [l.strip() for l in f.read().splitlines()]
        if lines[0].lower() != "off":
            raise RuntimeError("It should be an ascii off file!")
        nv, nf = (int(x) for x in lines[1].split()[:2])
        self._v = np.asarray([[float(x) for x in l.split()[:3]] for l in lines[2:2 + nv]], dtype=np.float32).reshape(-1, 3)
        faces, cols = [], []
        for l in lines[2 + nv:2 + nv + nf]:
            t = l.split()
            k = int(t[0])
            faces.append([int(x) for x in t[1:1 + k]])
            if len(t) > k + 1:
                cols.append([int(round(float(x) * 255)) for x in t[k + 1:k + 4]])
        self.set_faces(faces)
        self.set_colors(cols)

    def _save_off(self, path):
        off = np.concatenate([[0], np.cumsum(self._cnt)])
        with open(path, "w") as f:
            f.write(f"OFF\n{len(self._v)} {len(self._cnt)} 0\n")
            for p in self._v:
                f.write(f"{p[0]:g} {p[1]:g} {p[2]:g}\n")
            for i in range(len(self._cnt)):
                row = [str(self._cnt[i])] + [str(v) for v in self._idx[off[i]:off[i + 1]]]
                if len(self._col):
                    row += [f"{int(c) / 255:g}" for c in self._col[i]]   # OFF colours are floats in [0, 1]
                f.write(" ".join(row) + "\n")


def poly2tri(src_file, dst_file):
    """convert the polygonal mesh to a triangular one (reference libpolytools/poly2tri.py)"""
    mesh = PolyMesh(src_file)
    mesh.poly2tri()
    mesh.save(dst_file)


def get_faces_num(mesh_file):
    """{'poly': #polygon faces, 'tri': #fan triangles} (reference libpolytools/cntfaces.py)"""
    mesh = PolyMesh(mesh_file)
    return {'poly': mesh.num_polyfaces(), 'tri': mesh.num_trifaces()}


def load_ply_header(ply_path):
    """parsed PLY header: storing_type, storing_version, vertex_num, face_num (reference libpolytools/header.py)"""
    content, size = b"", 1000
    while b"end_header" not in content.lower():
        with open(ply_path, "rb") as f:
            content = f.read(size)
        if len(content) < size:
            break
        size *= 2
    info = {}
    for raw in content.split(b"\n"):
        try:
            s = raw.decode("utf-8").lower().strip()
        except UnicodeDecodeError:
            continue
        if s.startswith("format"):
            info["storing_type"] = s.split(" ")[1]
            info["storing_version"] = float(s.split(" ")[2])
        elif s.startswith("element"):
            info[f"{s.split(' ')[1]}_num"] = int(s.split(" ")[2])
        elif s.startswith("end_header"):
            break
    return info
