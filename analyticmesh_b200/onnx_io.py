"""ONNX <-> MLP codec with no dependency on the ``onnx`` package.

Replaces reference backend/onnx_io.py:11-285 (``save_model`` / ``load_model``).  The reference
walks the text produced by ``onnx.helper.printable_graph`` with regular expressions; this module
reads the protobuf wire format directly (ModelProto.graph -> nodes, initializers, inputs,
outputs) and applies the same acceptance rules and the same graph -> (nodes, arc_table, arc_tm)
mapping:

  * ops allowed: Gemm (alpha = beta = 1, transB = 1), Relu, MatMul, Add, Transpose(perm=[1,0])
    feeding a MatMul                                   (reference onnx_io.py:85-125)
  * widths come from the Gemm weight shapes            (reference onnx_io.py:132-160)
  * every Add is a skip connection: its destination is the next Relu (or the graph output)
    downstream, its source is the Relu / graph input that feeds it directly (identity) or
    through a MatMul (linear transform)                (reference onnx_io.py:163-219)
  * a MatMul initializer is stored (in, out) unless it is routed through a Transpose
                                                       (reference onnx_io.py:222-233)

``save_model`` writes the five-op subset itself for ``MLP`` instances and falls back to
``torch.onnx.export`` (TorchScript exporter) for arbitrary ``nn.Module``s, like the reference.
"""
import os
import struct

import numpy as np
import torch

from .model import MLP

# ----------------------------------------------------------------------------------------------
# protobuf wire format (just enough for ONNX ModelProto / GraphProto / NodeProto / TensorProto)
# ----------------------------------------------------------------------------------------------


def _read_varint(buf, pos):
    result = 0
    shift = 0
    while True:
        byte = buf[pos]
        pos += 1
        result |= (byte & 0x7F) << shift
        if not byte & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) for one message."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _read_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            value, pos = _read_varint(buf, pos)
        elif wire == 1:
            value = buf[pos:pos + 8]
            pos += 8
        elif wire == 2:
            size, pos = _read_varint(buf, pos)
            value = buf[pos:pos + size]
            pos += size
        elif wire == 5:
            value = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wire}")
        yield field, wire, value


def _varint(n):
    if n < 0:
        n += 1 << 64
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field, wire):
    return _varint((field << 3) | wire)


def _ld(field, payload):  # length-delimited
    if isinstance(payload, str):
        payload = payload.encode()
    return _key(field, 2) + _varint(len(payload)) + payload


def _vi(field, n):
    return _key(field, 0) + _varint(n)


def _packed_varints(buf):
    out, pos = [], 0
    while pos < len(buf):
        v, pos = _read_varint(buf, pos)
        out.append(v)
    return out


_ONNX_FLOAT, _ONNX_DOUBLE, _ONNX_INT64 = 1, 11, 7


def _parse_tensor(buf):
    dims, dtype, name, raw, floats, doubles = [], _ONNX_FLOAT, "", None, [], []
    for f, w, v in _fields(buf):
        if f == 1:
            dims += _packed_varints(v) if w == 2 else [v]
        elif f == 2:
            dtype = v
        elif f == 4:  # float_data
            floats += list(struct.unpack(f"<{len(v) // 4}f", v)) if w == 2 else [struct.unpack("<f", v)[0]]
        elif f == 10:  # double_data
            doubles += list(struct.unpack(f"<{len(v) // 8}d", v)) if w == 2 else [struct.unpack("<d", v)[0]]
        elif f == 8:
            name = bytes(v).decode()
        elif f == 9:
            raw = bytes(v)
    if dtype == _ONNX_FLOAT:
        arr = np.frombuffer(raw, dtype="<f4") if raw is not None else np.asarray(floats, dtype=np.float32)
    elif dtype == _ONNX_DOUBLE:
        arr = np.frombuffer(raw, dtype="<f8") if raw is not None else np.asarray(doubles, dtype=np.float64)
    else:
        return name, dtype, None
    return name, dtype, np.array(arr).reshape(dims)


def _parse_attr(buf):
    name, val = "", None
    ints = []
    for f, w, v in _fields(buf):
        if f == 1:
            name = bytes(v).decode()
        elif f == 2:
            val = struct.unpack("<f", v)[0]
        elif f == 3:
            val = v if v < (1 << 63) else v - (1 << 64)
        elif f == 8:
            ints += _packed_varints(v) if w == 2 else [v]
    if ints:
        val = ints
    return name, val


def _parse_node(buf):
    node = dict(inputs=[], outputs=[], op="", attrs={})
    for f, _, v in _fields(buf):
        if f == 1:
            node["inputs"].append(bytes(v).decode())
        elif f == 2:
            node["outputs"].append(bytes(v).decode())
        elif f == 4:
            node["op"] = bytes(v).decode()
        elif f == 5:
            k, a = _parse_attr(v)
            node["attrs"][k] = a
    return node


def _parse_value_info(buf):
    name, elem, shape = "", None, []
    for f, _, v in _fields(buf):
        if f == 1:
            name = bytes(v).decode()
        elif f == 2:  # TypeProto
            for f2, _, v2 in _fields(v):
                if f2 == 1:  # tensor_type
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            elem = v3
                        elif f3 == 2:  # TensorShapeProto
                            for f4, _, v4 in _fields(v3):
                                if f4 == 1:
                                    dim = None
                                    for f5, _, v5 in _fields(v4):
                                        if f5 == 1:
                                            dim = v5
                                    shape.append(dim)
    return dict(name=name, elem=elem, shape=shape)


def _parse_model(data):
    graph_buf = None
    for f, _, v in _fields(memoryview(data)):
        if f == 7:
            graph_buf = v
    if graph_buf is None:
        raise Exception("Error: not an ONNX ModelProto (no graph)")
    g = dict(nodes=[], init={}, init_dtype={}, inputs=[], outputs=[])
    for f, _, v in _fields(graph_buf):
        if f == 1:
            g["nodes"].append(_parse_node(v))
        elif f == 5:
            name, dtype, arr = _parse_tensor(v)
            g["init"][name] = arr
            g["init_dtype"][name] = dtype
        elif f == 11:
            g["inputs"].append(_parse_value_info(v))
        elif f == 12:
            g["outputs"].append(_parse_value_info(v))
    return g


# ----------------------------------------------------------------------------------------------
# graph -> MLP
# ----------------------------------------------------------------------------------------------

_ALLOWED = ("Gemm", "Relu", "MatMul", "Add", "Transpose")


def _graph_to_mlp(g):
    real_inputs = [vi for vi in g["inputs"] if vi["name"] not in g["init"]]
    assert len(real_inputs) == 1, "the network must have exactly one input"
    assert len(g["outputs"]) == 1, "the network must have exactly one output"
    inp, ret = real_inputs[0], g["outputs"][0]["name"]
    assert len(inp["shape"]) == 2 and inp["shape"][1] == 3, "input must have shape [N, 3]"
    assert inp["elem"] == _ONNX_FLOAT, "input must be FLOAT"
    for name, dt in g["init_dtype"].items():
        assert dt == _ONNX_FLOAT, f"initializer {name} must be FLOAT"

    nodes = []
    for n in g["nodes"]:
        assert n["op"] in _ALLOWED, f"Found unrecognized function: {n['op']}"
        nodes.append(n)
    # Transpose(perm=[1,0]) must feed a MatMul; fold it away and remember which MatMuls it fed
    transposed = {}
    for n in nodes:
        if n["op"] == "Transpose":
            assert n["attrs"].get("perm") == [1, 0], "Transpose must be perm=[1,0]"
            transposed[n["outputs"][0]] = n["inputs"][0]
    folded = []
    for i, n in enumerate(nodes):
        if n["op"] == "Transpose":
            nxt = nodes[i + 1] if i + 1 < len(nodes) else None
            assert nxt is not None and nxt["op"] == "MatMul", "Transpose must follow by MatMul"
            continue
        if n["op"] == "MatMul" and n["inputs"][1] in transposed:
            n = dict(n, inputs=[n["inputs"][0], transposed[n["inputs"][1]]], tm_is_out_in=True)
        elif n["op"] == "MatMul":
            n = dict(n, tm_is_out_in=False)
        if n["op"] == "Gemm":
            a = n["attrs"]
            assert a.get("alpha", 1.0) == 1 and a.get("beta", 1.0) == 1 and a.get("transB", 0) == 1, \
                "Gemm must have alpha=1, beta=1, transB=1"
        folded.append(n)
    nodes = folded

    gemms = [n for n in nodes if n["op"] == "Gemm"]
    relus = [n for n in nodes if n["op"] == "Relu"]
    assert len(relus) == len(gemms) - 1, "expected one Relu after every Gemm but the last"
    widths = [3] + [int(g["init"][gm["inputs"][1]].shape[0]) for gm in gemms[:-1]] + [1]
    assert g["init"][gemms[-1]["inputs"][1]].shape[0] == 1, "the last Gemm must have one output"

    # aux index: 0 = graph input, i = output of Relu i, last = graph output
    aux = [inp["name"]] + [r["outputs"][0] for r in relus] + [ret]
    producer = {n["outputs"][0]: n for n in nodes}

    def consumer_of(value):
        for n in nodes:
            if value in n["inputs"]:
                return n
        return None

    arc_table = [[0] for _ in range(len(widths) - 2)]
    tm_names, tm_shapes = [], []
    for n in nodes:
        if n["op"] != "Add":
            continue
        dst, hop = n["outputs"][0], None
        while dst != ret and (hop is None or hop["op"] != "Relu"):
            hop = consumer_of(dst)
            assert hop is not None, "dangling Add output"
            dst = hop["outputs"][0]
        dst_index = aux.index(dst)
        src_index = tm_name = tm_out_in = None
        for arg in n["inputs"]:
            p = producer.get(arg)
            if p is None:  # graph input used directly: identity skip from the raw input
                if arg == inp["name"]:
                    src_index, tm_name = 0, "identity_matrix"
                    break
                continue
            if p["op"] in ("Gemm", "Add"):
                continue
            if p["op"] == "Relu":
                src_index, tm_name = aux.index(arg), "identity_matrix"
                break
            if p["op"] == "MatMul":
                src_index, tm_name, tm_out_in = aux.index(p["inputs"][0]), p["inputs"][1], p["tm_is_out_in"]
                break
        assert src_index is not None, "could not resolve the source of an Add"
        if tm_name not in tm_names:
            tm_names.append(tm_name)
            if tm_name == "identity_matrix":
                tm_shapes.append(([0, 0], None))
            else:
                w = g["init"][tm_name]
                w = w if tm_out_in else w.T
                tm_shapes.append(([int(w.shape[0]), int(w.shape[1])], np.ascontiguousarray(w)))
        row = arc_table[dst_index - 2]
        row[0] += 1
        row.extend([src_index, tm_names.index(tm_name)])

    model = MLP(nodes=widths, arc_table=arc_table, arc_tm_shape=[s for s, _ in tm_shapes],
                initialization=None, enable_print=False)
    with torch.no_grad():
        for i, gm in enumerate(gemms):
            model.linears[i].weight.copy_(torch.from_numpy(np.array(g["init"][gm["inputs"][1]])))
            model.linears[i].bias.copy_(torch.from_numpy(np.array(g["init"][gm["inputs"][2]])))
        for i, (_, w) in enumerate(tm_shapes):
            if w is not None:
                model.tms[i].weight.copy_(torch.from_numpy(np.array(w)))
    return model


def load_model(model_file_str):
    """Load an ONNX ReLU-MLP as an ``MLP``; accepts a path (str) or the file content (bytes)."""
    if isinstance(model_file_str, str) and os.path.exists(model_file_str):
        with open(model_file_str, "rb") as f:
            data = f.read()
    elif isinstance(model_file_str, (bytes, bytearray)):
        data = bytes(model_file_str)
    else:
        raise Exception(f"Error: unknown model_file_str = {model_file_str}")
    return _graph_to_mlp(_parse_model(data))


# ----------------------------------------------------------------------------------------------
# MLP -> ONNX
# ----------------------------------------------------------------------------------------------


def _tensor_proto(name, arr):
    arr = np.ascontiguousarray(arr, dtype="<f4")
    out = b"".join(_vi(1, int(d)) for d in arr.shape)
    out += _vi(2, _ONNX_FLOAT) + _ld(8, name) + _ld(9, arr.tobytes())
    return out


def _attr_float(name, v):
    return _ld(1, name) + _key(2, 5) + struct.pack("<f", v) + _vi(20, 1)


def _attr_int(name, v):
    return _ld(1, name) + _vi(3, v) + _vi(20, 2)


def _node_proto(op, inputs, outputs, attrs=()):
    out = b"".join(_ld(1, i) for i in inputs) + b"".join(_ld(2, o) for o in outputs)
    out += _ld(3, outputs[0] + "_node") + _ld(4, op) + b"".join(_ld(5, a) for a in attrs)
    return out


def _value_info(name, shape):
    dims = b"".join(_ld(1, _vi(1, d)) for d in shape)
    tensor_type = _vi(1, _ONNX_FLOAT) + _ld(2, dims)
    return _ld(1, name) + _ld(2, _ld(1, tensor_type))


def _mlp_to_onnx_bytes(model):
    info = model.get_info()
    nodes_pb, inits = [], []
    last = model.num_of_linears - 1
    taps = {}
    x = "input"
    for i in range(model.num_of_linears):
        taps[i] = x
        wname, bname = f"linears.{i}.weight", f"linears.{i}.bias"
        inits.append(_tensor_proto(wname, info["weights"][i].detach().cpu().float().numpy()))
        inits.append(_tensor_proto(bname, info["biases"][i].detach().cpu().float().numpy()))
        y = f"fc{i}"
        nodes_pb.append(_node_proto("Gemm", [x, wname, bname], [y],
                                    [_attr_float("alpha", 1.0), _attr_float("beta", 1.0), _attr_int("transB", 1)]))
        if i >= 1:
            row = model.arc_table[i - 1]
            for j in range(row[0]):
                src, tm = row[1 + 2 * j], row[2 + 2 * j]
                s = taps[src]
                if model.arc_tm_shape[tm][0] != 0 or model.arc_tm_shape[tm][1] != 0:
                    tname = f"tms.{tm}.weight_t"
                    if not any(tname.encode() in b for b in inits):
                        inits.append(_tensor_proto(tname, info["arc_tm"][tm].detach().cpu().float().numpy().T))
                    s2 = f"skip{i}_{j}"
                    nodes_pb.append(_node_proto("MatMul", [s, tname], [s2]))
                    s = s2
                y2 = f"add{i}_{j}"
                nodes_pb.append(_node_proto("Add", [y, s], [y2]))
                y = y2
        if i != last:
            r = f"relu{i}"
            nodes_pb.append(_node_proto("Relu", [y], [r]))
            x = r
        else:
            x = y
    graph = b"".join(_ld(1, n) for n in nodes_pb) + _ld(2, "analyticmesh_b200_mlp")
    graph += b"".join(_ld(5, t) for t in inits)
    graph += _ld(11, _value_info("input", [1, 3])) + _ld(12, _value_info(x, [1, 1]))
    opset = _ld(1, "") + _vi(2, 9)
    return _vi(1, 4) + _ld(2, "analyticmesh_b200") + _ld(3, "1") + _ld(7, graph) + _ld(8, opset)


def save_model(model, model_path):
    """Export to ONNX.  ``MLP`` instances are written directly; any other ``nn.Module`` goes through
    ``torch.onnx.export`` exactly like reference backend/onnx_io.py:262-265."""
    model.cpu().float()
    if isinstance(model, MLP):
        with open(model_path, "wb") as f:
            f.write(_mlp_to_onnx_bytes(model))
        return
    dummy_input = torch.randn([1, 3])
    try:
        import onnx  # noqa: F401
    except ImportError:
        # The TorchScript exporter's C++ serializer has produced the bytes before its last
        # post-pass imports `onnx`; without that package the post-pass is an identity here.
        from torch.onnx._internal.torchscript_exporter import onnx_proto_utils
        onnx_proto_utils._add_onnxscript_fn = lambda proto, *a, **k: proto
    torch.onnx.export(model, dummy_input, model_path, dynamo=False)
