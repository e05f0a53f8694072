"""Host-side mirror of `rten::Model` (src/model.rs) over rten_b200_model_*: load an ONNX file, run it by input / output
names.  The graph executor itself is native (csrc/model.cu); this module only marshals descriptors."""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import _lib
from .ops import Context, DeviceTensor, OpError, _Args, _RT2NP
from ._lib import RtenTensor


def onnx_summary(data: bytes) -> dict:
    """The reader alone (no GPU): decoded structure of an ONNX file as a dict."""
    lib = _lib.load()
    need = C.c_size_t(0)
    st = lib.rten_b200_onnx_summary(data, len(data), None, 0, C.byref(need))
    if st != 0:
        raise OpError(st, "ONNX decode failed")
    buf = C.create_string_buffer(need.value)
    st = lib.rten_b200_onnx_summary(data, len(data), buf, need.value, None)
    if st != 0:
        raise OpError(st, "ONNX decode failed")
    return json.loads(buf.value.decode())


class Model:
    """`Model::load` / `Model::run` (src/model.rs:300-760): inputs and outputs are addressed by name."""

    def __init__(self, ctx: Context, data: Union[bytes, str]):
        if isinstance(data, str):
            data = open(data, "rb").read()
        self.ctx = ctx
        self._bytes = data
        h = C.c_void_p()
        ctx.check(ctx.lib.rten_b200_model_load(ctx.handle, data, len(data), C.byref(h)))
        self.handle = h
        lib = ctx.lib
        self.input_names = [lib.rten_b200_model_input_name(h, i).decode() for i in range(lib.rten_b200_model_num_inputs(h))]
        self.output_names = [lib.rten_b200_model_output_name(h, i).decode() for i in range(lib.rten_b200_model_num_outputs(h))]
        self.node_ops = [lib.rten_b200_model_node_op(h, i).decode() for i in range(lib.rten_b200_model_num_nodes(h))]

    @property
    def summary(self) -> dict:
        return json.loads(self.ctx.lib.rten_b200_model_summary(self.handle).decode())

    def run(self, inputs: Dict[str, Union[np.ndarray, DeviceTensor]], outputs: Optional[Sequence[str]] = None) -> List[DeviceTensor]:
        outputs = list(outputs) if outputs is not None else self.output_names
        A = _Args(self.ctx)
        names = list(inputs)
        in_names = (C.c_char_p * len(names))(*[n.encode() for n in names])
        in_t = (RtenTensor * max(len(names), 1))()
        for i, n in enumerate(names):
            ref = A.t(inputs[n])
            C.memmove(C.byref(in_t, i * C.sizeof(RtenTensor)), ref, C.sizeof(RtenTensor))
        out_names = (C.c_char_p * len(outputs))(*[n.encode() for n in outputs])
        out_t = (RtenTensor * len(outputs))()
        self.ctx.check(self.ctx.lib.rten_b200_model_run(self.handle, len(names), in_names, in_t, len(outputs), out_names, out_t))
        res = []
        for d in out_t:
            shape = tuple(d.shape[i] for i in range(d.ndim))
            strides = tuple(d.strides[i] for i in range(d.ndim))
            res.append(DeviceTensor(self.ctx, d.data, shape, strides, _RT2NP[d.dtype], owner=True))
        return res

    def close(self):
        if getattr(self, "handle", None) and self.ctx.handle:
            self.ctx.lib.rten_b200_model_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
