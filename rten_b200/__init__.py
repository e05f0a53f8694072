"""rten_b200 -- B200 (sm_100a) operator execution backend for the RTen hot path.

The product is `librten_b200.so` (hand-written CUDA behind the C ABI in include/rten_b200.h);
`rten_b200.ops` is the host-side mirror of RTen's operator interface used by the tests, the model
runners and bench.py.  There is no CPU implementation in this package."""
from . import _lib  # noqa: F401
from .ops import (  # noqa: F401
    ACT_GELU, ACT_GELU_TANH, ACT_NONE, ACT_RELU, Add, AddSoftmax, Attention, Comm, Context, Conv, ConvInteger, ConvIntegerToFloat,
    DeviceTensor, DynamicQuantizeLinear, Erf, FusedMatMul, GatherRows, Gelu, Gemm, GlobalAveragePool,
    LayerNormalization, MatMul, MatMulInteger, MatMulIntegerToFloat, MaxPool, Mul, OpError, Packed, QuantizedLinear, Relu, ScatterRows, Softmax, from_torch,
)
