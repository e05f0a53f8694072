"""Test infrastructure: a minimal ONNX *writer* (protobuf wire format by hand; the `onnx` package is not in this image).
Used to build the model files the reader / executor tests load: the reference's MNIST test model re-encoded from
tests/golden/mnist.npz, and small transformer / quantised graphs with seeded weights."""
import struct

import numpy as np

FLOAT, UINT8, INT8, INT32, INT64 = 1, 2, 3, 6, 7
_NP2ONNX = {np.dtype(np.float32): FLOAT, np.dtype(np.uint8): UINT8, np.dtype(np.int8): INT8, np.dtype(np.int32): INT32,
            np.dtype(np.int64): INT64}


def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wt: int) -> bytes:
    return _varint((field << 3) | wt)


def _ld(field: int, payload: bytes) -> bytes:
    return _key(field, 2) + _varint(len(payload)) + payload


def _vi(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(v)


def tensor(name: str, arr: np.ndarray, raw: bool = True) -> bytes:
    arr = np.asarray(arr)
    dt = _NP2ONNX[arr.dtype]
    out = b"".join(_vi(1, int(d)) for d in arr.shape) + _vi(2, dt)
    if raw:
        out += _ld(9, np.ascontiguousarray(arr).tobytes())
    elif dt == FLOAT:
        out += _ld(4, np.ascontiguousarray(arr, "<f4").tobytes())      # packed float_data
    elif dt == INT64:
        out += _ld(7, b"".join(_varint(int(x)) for x in arr.reshape(-1)))  # packed int64_data
    else:
        out += _ld(5, b"".join(_varint(int(x)) for x in arr.reshape(-1)))  # packed int32_data
    return out + _ld(8, name.encode())


def attribute(name: str, value) -> bytes:
    out = _ld(1, name.encode())
    if isinstance(value, bool):
        value = int(value)
    if isinstance(value, int):
        return out + _vi(3, value) + _vi(20, 2)
    if isinstance(value, float):
        return out + _key(2, 5) + struct.pack("<f", value) + _vi(20, 1)
    if isinstance(value, str):
        return out + _ld(4, value.encode()) + _vi(20, 3)
    if isinstance(value, np.ndarray):
        return out + _ld(5, tensor("", value)) + _vi(20, 4)
    if isinstance(value, (list, tuple)) and all(isinstance(x, int) for x in value):
        return out + b"".join(_vi(8, x) for x in value) + _vi(20, 7)
    if isinstance(value, (list, tuple)):
        return out + b"".join(_key(7, 5) + struct.pack("<f", float(x)) for x in value) + _vi(20, 6)
    raise TypeError(type(value))


def node(op: str, inputs, outputs, name: str = "", domain: str = "", **attrs) -> bytes:
    out = b"".join(_ld(1, i.encode()) for i in inputs) + b"".join(_ld(2, o.encode()) for o in outputs)
    if name:
        out += _ld(3, name.encode())
    out += _ld(4, op.encode())
    out += b"".join(_ld(5, attribute(k, v)) for k, v in attrs.items())
    if domain:
        out += _ld(7, domain.encode())
    return out


def value_info(name: str, elem_type: int, shape) -> bytes:
    dims = b"".join(_ld(1, _vi(1, int(d)) if isinstance(d, int) else _ld(2, str(d).encode())) for d in shape)
    ttype = _vi(1, elem_type) + _ld(2, dims)
    return _ld(1, name.encode()) + _ld(2, _ld(1, ttype))


def model(nodes, initializers, inputs, outputs, opset: int = 18, ir_version: int = 8, graph_name: str = "g", extra_opsets=()) -> bytes:
    g = b"".join(_ld(1, n) for n in nodes) + _ld(2, graph_name.encode()) + b"".join(_ld(5, t) for t in initializers)
    g += b"".join(_ld(11, v) for v in inputs) + b"".join(_ld(12, v) for v in outputs)
    out = _vi(1, ir_version) + _ld(7, g) + _ld(8, _ld(1, b"") + _vi(2, opset))
    for dom, ver in extra_opsets:
        out += _ld(8, _ld(1, dom.encode()) + _vi(2, ver))
    return out


# ------------------------------------------------------------------------------------------
def mnist_from_fixture(npz_path: str) -> bytes:
    """The reference's MNIST test model (rten-onnx/test-data/mnist.onnx) re-encoded from the fixture's node list and
    weights: same 13 nodes (incl. the two Constant nodes), same 8 initialisers, opset 18, input `input`, output `logits`."""
    import json
    z = np.load(npz_path)
    nodes_j = json.loads(str(z["nodes"]))
    weights = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    nodes = []
    for n in nodes_j:
        attrs = {}
        for k, v in n["attrs"].items():
            if n["op"] == "Constant":
                attrs[k] = np.asarray(v, np.int64)
            elif isinstance(v, list):
                attrs[k] = [int(x) for x in v]
            elif v is None:
                attrs[k] = "NOTSET"  # (auto_pad: the fixture's walker did not decode string attributes)
            else:
                attrs[k] = v
        nodes.append(node(n["op"], n["inputs"], n["outputs"], **attrs))
    inits = [tensor(k, v) for k, v in weights.items()]
    return model(nodes, inits, [value_info("input", FLOAT, ["batch", 1, 28, 28])], [value_info("logits", FLOAT, ["batch", 10])])
