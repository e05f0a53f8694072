/*
 * rten_b200.h -- C ABI of librten_b200.so: the B200 (sm_100a) operator execution backend for RTen.
 *
 * This is the drop-in boundary for ONE path of robertknight/rten: `Operator::run` /
 * `run_in_place` / `prepack` (src/operator.rs:486-613) of the dense operators of ResNet-50 /
 * BERT-base / GPT-2, replacing rten-gemm's packed GEMM (rten-gemm/src/lib.rs:199-391) and the
 * rten-vecmath row kernels.  A Rust `impl Operator` shim fills `rten_tensor` descriptors from
 * `ValueView`s (src/value.rs:299) and calls one function below per operator.
 *
 * Conventions
 *  - Every entry point takes a context (one per host thread / stream: rten's ops are `Send + Sync`
 *    and `Model::run` may be re-entered concurrently, src/operator.rs:622).
 *  - `rten_tensor.strides` are ELEMENT strides like rten-tensor layouts; arbitrary (non-negative)
 *    strides are accepted, including the permuted views `TransformInputs` hands to MatMul
 *    (src/ops/transform_inputs.rs:23-34).
 *  - `device >= 0`: `data` is a device pointer on that CUDA ordinal (buffers resident in HBM).
 *    `device == RTEN_DEVICE_HOST`: `data` is host memory; the library stages it through HBM on the
 *    context's stream (host<->device copies are part of the call) -- this is how an unmodified
 *    rten `Vec<T>`-backed tensor crosses the boundary.
 *  - Output tensors: `out->data == NULL` => the library allocates from the context pool (plays
 *    `ctx.pool()`, src/operator.rs:360) on device and fills shape/strides/device; the caller later
 *    returns it with rten_b200_free().  Otherwise `out` must already have the result shape.
 *  - Calls enqueue work on the context stream and return without synchronising unless a host
 *    tensor is involved; asynchronous CUDA errors surface on the next call or rten_b200_sync().
 *  - Errors: status codes mirror `OpError` (src/operator.rs:116-144); rten_b200_last_error()
 *    returns the reference's static message string for that error.
 *  - No CPU fallback exists: without a B200 + driver every op fails with RTEN_ERR_CUDA.
 */
#ifndef RTEN_B200_H
#define RTEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTEN_MAX_DIMS 8
#define RTEN_DEVICE_HOST (-1)

/* = `DataType` (src/value.rs:20-40) */
typedef enum { RTEN_F32 = 0, RTEN_I32 = 1, RTEN_I8 = 2, RTEN_U8 = 3 } rten_dtype;

typedef struct {
    void* data;
    int32_t dtype; /* rten_dtype */
    int32_t ndim;
    int64_t shape[RTEN_MAX_DIMS];
    int64_t strides[RTEN_MAX_DIMS]; /* elements */
    int32_t device;                 /* RTEN_DEVICE_HOST or CUDA ordinal */
    int32_t _reserved;
} rten_tensor;

/* = `OpError` variants (src/operator.rs:116-144) + device errors */
typedef enum {
    RTEN_OK = 0,
    RTEN_ERR_CAST_FAILED = 1,
    RTEN_ERR_UNSUPPORTED_TYPE = 2,
    RTEN_ERR_INCOMPATIBLE_SHAPES = 3, /* IncompatibleInputShapes(&'static str) */
    RTEN_ERR_MISSING_INPUTS = 4,
    RTEN_ERR_INVALID_VALUE = 5,      /* InvalidValue(&'static str) */
    RTEN_ERR_UNSUPPORTED_VALUE = 6,  /* UnsupportedValue(&'static str) */
    RTEN_ERR_UNSUPPORTED_OUTPUT = 7, /* UnsupportedOutput(&'static str) */
    RTEN_ERR_CUDA = 100,
    RTEN_ERR_NCCL = 101
} rten_status;

typedef struct rten_ctx rten_ctx;
/* = `PrepackedInput` (src/operator.rs:25-31): a weight re-laid out once for the tensor cores. */
typedef struct rten_packed rten_packed;

/* fp32 GEMM/Conv arithmetic mode (SURVEY.md hard part A). */
typedef enum {
    RTEN_F32_TF32 = 0,  /* single tcgen05 kind::tf32 pass: operands rounded to 10 mantissa bits.  EXPLICIT OPT-IN
                           (rten_b200_set_f32_mode or env RTEN_B200_F32_MODE=tf32); tolerance in DESIGN.md */
    RTEN_F32_TF32X3 = 1 /* DEFAULT: 3-pass error-compensated split (hi*hi + hi*lo + lo*hi, f32 accumulation): meets the
                           reference's own f32 tolerance (rten-tensor/src/test_util.rs:47-92) at 1/3 of the tensor rate */
} rten_f32_mode;

/* ---- context, memory, diagnostics ----------------------------------------------------------- */
rten_status rten_b200_ctx_create(int device, void* cuda_stream_or_null, size_t workspace_bytes, rten_ctx** out);
void rten_b200_ctx_destroy(rten_ctx* ctx);
const char* rten_b200_last_error(rten_ctx* ctx);
rten_status rten_b200_sync(rten_ctx* ctx);
rten_status rten_b200_set_f32_mode(rten_ctx* ctx, int mode /* rten_f32_mode */);
/* Plan autotuning (off by default; env RTEN_B200_AUTOTUNE=1 turns it on at context creation).  When on, the FIRST
 * MatMul / Conv launch of every distinct problem (shape, layout, epilogue) outside graph capture times the cost
 * model's best launch plans on the device and caches the winner in the context -- the role rten-gemm's per-arch
 * kernel selection and blocking heuristics play (rten-gemm/src/lib.rs:199-391), decided by measurement.  Integer
 * results do not depend on the plan; f32 results stay within the TF32 tolerance but may differ in the last bits
 * between plans (split-K changes the summation order).  Measured plans are used for the rest of the context's life
 * even after autotuning is switched off again, and env RTEN_B200_TUNE_FILE=<path> keeps them across processes
 * (read at context creation, rewritten at destruction when new problems were measured). */
rten_status rten_b200_set_autotune(rten_ctx* ctx, int enable);
rten_status rten_b200_save_plans(rten_ctx* ctx, const char* path);
rten_status rten_b200_load_plans(rten_ctx* ctx, const char* path);
/* Caching, stream-ordered device allocator = `BufferPool` (src/buffer_pool.rs:1-140). */
rten_status rten_b200_alloc(rten_ctx* ctx, size_t bytes, void** dev_ptr);
rten_status rten_b200_free(rten_ctx* ctx, void* dev_ptr);
/* Pinned host buffers for callers that want asynchronous staging of host tensors. */
rten_status rten_b200_host_alloc(rten_ctx* ctx, size_t bytes, void** host_ptr);
rten_status rten_b200_host_free(rten_ctx* ctx, void* host_ptr);
/* Copies between host and device tensors of identical shape (strided on both sides). */
rten_status rten_b200_copy(rten_ctx* ctx, const rten_tensor* src, rten_tensor* dst);
/* Number of CUDA kernels this context has launched so far (bench.py `gpu_launches`). */
uint64_t rten_b200_launch_count(rten_ctx* ctx);
const char* rten_b200_version(void);
/* Debug aid for plan sweeps (env RTEN_B200_FORCE_BN / _PAIR / _KATOMS / _SPLITK / _CTA2): how many tensor-core launches
 * found a valid plan matching every forced field, and how many fell back to the cost model's choice because none did.
 * With env RTEN_B200_FORCE_STRICT=1 such a launch fails with RTEN_ERR_INVALID_VALUE instead. */
rten_status rten_b200_debug_forced_plans(rten_ctx* ctx, uint64_t* matched, uint64_t* unmatched);
/* Debug aid: when enabled, CTA 0 of every GEMM/conv launch records clock64() at its pipeline hand-offs into a
 * 4 x 2048 int64 buffer (rows: producer slot acquired, MMA operands landed, epilogue start, epilogue end).
 * `host_out_8192_or_null` receives the current contents before the state change. */
rten_status rten_b200_debug_trace(rten_ctx* ctx, int enable, int64_t* host_out_8192_or_null);
/* Capture everything enqueued between begin/end into a CUDA graph; replay with graph_launch.
 * (launch-bound op lists: the `Graph::run_plan` loop, src/graph.rs:880-1286, as one graph) */
typedef struct rten_graph rten_graph;
rten_status rten_b200_graph_begin(rten_ctx* ctx);
rten_status rten_b200_graph_end(rten_ctx* ctx, rten_graph** out);
rten_status rten_b200_graph_launch(rten_ctx* ctx, rten_graph* g);
void rten_b200_graph_destroy(rten_graph* g);

/* ---- prepack == `Operator::prepack` (src/operator.rs:587-601) ------------------------------- */
/* MatMul/FusedMatMul/MatMulInteger input 1 (src/ops/matmul.rs:410-420,687-697): B [K,N] (f32, i8 or
 * u8) -> K-major [N,K] tensor-core layout (+ per-column sums for the int8 zero-point epilogue). */
rten_status rten_b200_prepack_b(rten_ctx* ctx, const rten_tensor* b, rten_packed** out);
/* Conv/ConvInteger kernel OIHW (the reference prepacks it per call, src/ops/conv.rs:302-315):
 * -> [O, kh, kw, C/g] K-major (+ per-output-channel sums for ConvInteger). */
rten_status rten_b200_prepack_conv_weight(rten_ctx* ctx, const rten_tensor* w, int groups, rten_packed** out);
void rten_b200_packed_free(rten_ctx* ctx, rten_packed* p);

/* ---- operators == `Operator::run` ------------------------------------------------------------ */
/* Gemm (src/ops/matmul.rs:32-167): out = alpha * op(a) @ op(b) + beta * broadcast(c). */
rten_status rten_b200_gemm(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, const rten_tensor* c_or_null,
                           float alpha, float beta, int trans_a, int trans_b, rten_tensor* out);

/* MatMul (src/ops/matmul.rs:390-434) / FusedMatMul (:462-507): numpy-matmul broadcasting; optional row
 * bias over N; alpha scales the product before the bias is added.  `packed_b_or_null` = the
 * PrepackedInput for input 1 (then `b` is only consulted for its shape). */
rten_status rten_b200_matmul(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b,
                             const rten_packed* packed_b_or_null, const rten_tensor* row_bias_or_null, float alpha,
                             rten_tensor* out);
/* Extension used by whole-model runners: fused activation after bias (0 none, 1 relu, 2 gelu(erf),
 * 3 gelu(tanh)) and optional residual add (same shape as out) before the activation. */
rten_status rten_b200_matmul_ex(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b,
                                const rten_packed* packed_b_or_null, const rten_tensor* row_bias_or_null, float alpha,
                                const rten_tensor* residual_or_null, int activation, rten_tensor* out);

/* MatMulInteger (src/ops/matmul.rs:582-697): a u8|i8, b u8|i8, zero points scalar or vector, out i32
 * exact.  `scale_or_null` != NULL => MatMulIntegerToFloat (:776-811): out f32 = f32(acc) * scale
 * (scalar or per column). */
rten_status rten_b200_matmul_integer(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b,
                                     const rten_packed* packed_b_or_null, const rten_tensor* a_zero_point_or_null,
                                     const rten_tensor* b_zero_point_or_null, const rten_tensor* scale_or_null,
                                     rten_tensor* out);
/* MatMulIntegerToFloat followed by the graph's Add(bias [N]), Add(residual, same shape as the output) and activation
 * (rten_activation) in the epilogue, as separate exactly rounded f32 operations in that order -- bit-identical to the
 * unfused operators.  Requires `scale`. */
rten_status rten_b200_matmul_integer_ex(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, const rten_packed* packed_b,
                                        const rten_tensor* a_zero_point, const rten_tensor* b_zero_point,
                                        const rten_tensor* scale, const rten_tensor* scale_b_or_null /* scalar: effective scale = scale_b * scale */,
                                        const rten_tensor* bias, const rten_tensor* residual, int activation,
                                        rten_tensor* out_range_or_null, rten_tensor* out);

/* Conv (src/ops/conv.rs:124-419).  x NCHW (or NCW), w OIHW, bias [O].  pads = {top,left,bottom,right};
 * auto_pad_same != 0 => `Padding::Same` (pads ignored).  n_spatial = 1 or 2 gives the expected
 * number of stride/dilation values (error strings as the reference). */
typedef struct {
    int32_t pads[4];
    int32_t auto_pad_same;
    int32_t groups;
    int32_t strides[2];
    int32_t dilations[2];
    int32_t n_strides;   /* number of valid entries in strides (reference validates == spatial dims) */
    int32_t n_dilations; /* idem */
} rten_conv_params;
rten_status rten_b200_conv2d(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* w,
                             const rten_packed* packed_w_or_null, const rten_tensor* bias_or_null,
                             const rten_conv_params* p, rten_tensor* out);
/* Extension: fused residual add (same shape as out) + activation (see matmul_ex). */
rten_status rten_b200_conv2d_ex(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* w,
                                const rten_packed* packed_w_or_null, const rten_tensor* bias_or_null,
                                const rten_conv_params* p, const rten_tensor* residual_or_null, int activation,
                                rten_tensor* out);
/* ConvInteger (src/ops/conv.rs:421-533); scale_or_null != NULL => ConvIntegerToFloat (:535-587). */
rten_status rten_b200_conv_integer(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* w,
                                   const rten_packed* packed_w_or_null, const rten_tensor* x_zero_point_or_null,
                                   const rten_tensor* w_zero_point_or_null, const rten_tensor* scale_or_null,
                                   const rten_conv_params* p, rten_tensor* out);
/* ConvIntegerToFloat with the graph nodes around it folded into the kernel epilogue, each as the same exactly rounded
 * f32 operation the separate operator would perform (bit-identical results): the Mul that forms the scale
 * (`scale_b_or_null`: scalar, effective scale = scale_b * scale), then Add(bias [O]), Add(residual, same shape as the
 * output) and Relu (activation 0 / 1).  Requires `scale`. */
rten_status rten_b200_conv_integer_ex(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* w, const rten_packed* packed_w,
                                      const rten_tensor* x_zero_point, const rten_tensor* w_zero_point,
                                      const rten_tensor* scale, const rten_tensor* scale_b_or_null,
                                      const rten_conv_params* params, const rten_tensor* bias, const rten_tensor* residual,
                                      int activation, rten_tensor* out_range_or_null, rten_tensor* out);

/* ---- autoregressive decode path (rten-generate's loop: one token per sequence and step) -------------------------- */
/* The per-token linear layer of a dynamically quantised transformer as ONE call:
 *   [LayerNormalization(x, ln_scale, ln_bias, axis -1)] -> DynamicQuantizeLinear -> Mul(x_scale, w_scale) ->
 *   MatMulIntegerToFloat(x_q, w, x_zp, w_zero_point, scale) -> Add(bias) -> Add(residual) -> activation
 * -- the node chain `tools/ort-quantize.py` + RTen's fusions (src/optimize/fusions.rs:966-1058) leave around every
 * MatMul of GPT-2.  x [.., K] f32, w [K, N] i8 | u8 (packed_w = its rten_b200_prepack_b handle), w_scale scalar or [N].
 * For M = prod(leading dims) <= 16 this is the skinny-M kernel that plays rten-gemm's gemv path
 * (rten-gemm/src/lib.rs:668-747, kernels simd_generic.rs:795-1129): the weights stream from HBM exactly once, the
 * quantised activations live in shared memory, nothing else is launched.  Larger M runs the separate operators.  Either
 * way every stage performs the operators' exactly rounded arithmetic: results are bit-identical to the unfused graph. */
rten_status rten_b200_quantized_linear(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* ln_scale_or_null,
                                       const rten_tensor* ln_bias_or_null, float ln_epsilon, const rten_tensor* w,
                                       const rten_packed* packed_w_or_null, const rten_tensor* w_zero_point_or_null,
                                       const rten_tensor* w_scale, const rten_tensor* bias_or_null,
                                       const rten_tensor* residual_or_null, int activation, rten_tensor* out);
/* Attention (src/ops/attention.rs:645-905, the ONNX `Attention` operator) on 4-D inputs: query [batch, q_heads, q_seq,
 * head], key / value [batch, kv_heads, total_seq, head] with any strides (a transposed value cache is just a view),
 * attn_mask float broadcastable to [batch, q_heads, q_seq, total_seq], nonpad_kv_seqlen i32 [batch] = number of valid
 * key / value positions when the caller manages a right-padded KV cache (`:817-832`; read on the device, so a decode
 * step stays a fixed launch list).  out [batch, q_heads, q_seq, head]; fully masked rows give zeros (sdpa_head :548-552).
 * q_seq = 1 (decode) with head size 64 / 128 is ONE kernel: scores, softmax and the value product stream the cache once,
 * split over the sequence to fill the SMs.  `new_key` / `new_value` [batch, kv_heads, 1, head] (optional) are written
 * into the caches at position nonpad_kv_seqlen[b] - 1 by the same kernel first (the cache append of rten-generate,
 * rten-generate/src/generator.rs:858-886, without a separate launch).  Other shapes compose MatMul / Softmax / MatMul. */
typedef struct {
    int32_t is_causal;
    int32_t q_num_heads;  /* informative for 4-D inputs */
    int32_t kv_num_heads;
    float scale;          /* <= 0: 1 / sqrt(head size) */
    float softcap;        /* > 0 unsupported */
} rten_attention_params;
rten_status rten_b200_attention(rten_ctx* ctx, const rten_tensor* query, const rten_tensor* key, const rten_tensor* value,
                                const rten_tensor* attn_mask_or_null, const rten_tensor* nonpad_kv_seqlen_or_null,
                                const rten_attention_params* params, const rten_tensor* new_key_or_null,
                                const rten_tensor* new_value_or_null, rten_tensor* out);

/* Softmax (src/ops/norm.rs:825-899) and AddSoftmax (src/ops/attention.rs:30-165) when mask != NULL
 * (mask broadcast to x, added lane-wise before the softmax over `axis`; AddSoftmax uses axis -1).
 * `out` may alias `x` (= run_in_place). */
rten_status rten_b200_softmax(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* mask_or_null, int axis,
                              int flush_nans_to_zero, rten_tensor* out);
/* LayerNormalization (src/ops/norm.rs:437-569); epsilon < 0 => default 1e-5. */
rten_status rten_b200_layer_norm(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* scale,
                                 const rten_tensor* bias_or_null, int axis, float epsilon, rten_tensor* out);
/* Erf / Gelu (src/ops/unary_elementwise.rs:384-435); approximate != 0 => tanh form. */
rten_status rten_b200_erf(rten_ctx* ctx, const rten_tensor* x, rten_tensor* out);
rten_status rten_b200_gelu(rten_ctx* ctx, const rten_tensor* x, int approximate, rten_tensor* out);
/* Communicator of a batch-sharded run (one process per GPU).  The reference has no counterpart (single process, rayon
 * threads); it exists so that DynamicQuantizeLinear can use the range of the WHOLE tensor when the batch is split over
 * ranks.  NCCL (libnccl.so.2) is resolved with dlopen at the first call; rank 0 creates the 128-byte id, the host
 * distributes it (any side channel), every rank calls comm_create with the same id. */
typedef struct rten_comm rten_comm;
rten_status rten_b200_comm_unique_id(void* id_out_128_bytes);
rten_status rten_b200_comm_create(rten_ctx* ctx, const void* id_128_bytes, int rank, int world_size, rten_comm** out);
void rten_b200_comm_destroy(rten_comm* comm);
/* 1: the range exchange runs as one kernel over NVLink peer memory (mailboxes opened through CUDA IPC at comm_create);
 * 0: the peers' memory could not be opened (or RTEN_B200_NCCL_RANGES=1) and two ncclAllReduce calls are used.  Both are exact. */
int rten_b200_comm_uses_peer_memory(const rten_comm* comm);
/* Exchanges that gave up waiting for a peer (~30 s) since comm_create: 0 in a healthy run, -1 if the device cannot be read. */
int rten_b200_comm_timeouts(const rten_comm* comm);

/* DynamicQuantizeLinear (src/ops/quantize.rs:352-468): y u8, scale f32 scalar, zero_point u8 scalar.
 * comm_or_null (a rten_comm*): when the batch is sharded over ranks, the local (min, max) is all-reduced over the ranks
 * first -- integer min / max on an order-preserving encoding, exact -- so every rank picks the unsharded tensor's scale
 * and zero point and the sharded outputs stay bit-identical to the unsharded reference's. */
rten_status rten_b200_dynamic_quantize_linear(rten_ctx* ctx, const rten_tensor* x, rten_tensor* y,
                                              rten_tensor* scale, rten_tensor* zero_point, void* comm_or_null);
/* Producer-computed ranges.  `out_range_or_null` of the *_integer_ex functions is a device i32[2] in which the kernel
 * epilogue accumulates (min, max) of the f32 output it writes (order-preserving integer encoding, atomicMin / atomicMax:
 * exact, order independent); `rten_b200_range_reset` re-arms any number of such pairs in one launch; the ranged
 * DynamicQuantizeLinear then skips its own pass over x.  Same bits as the plain operator (NaN inputs excepted). */
rten_status rten_b200_range_reset(rten_ctx* ctx, rten_tensor* ranges_i32_n_by_2);
rten_status rten_b200_dynamic_quantize_linear_ranged(rten_ctx* ctx, const rten_tensor* x, const rten_tensor* range_or_null,
                                                     rten_tensor* y, rten_tensor* scale, rten_tensor* zero_point,
                                                     void* comm_or_null);

/* ---- residency glue so whole models stay in HBM ------------------------------ */
rten_status rten_b200_relu(rten_ctx* ctx, const rten_tensor* x, rten_tensor* out);
/* Add with numpy broadcasting (src/ops/binary_elementwise.rs). */
rten_status rten_b200_add(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, rten_tensor* out);
/* Mul (src/ops/binary_elementwise.rs), f32, numpy broadcasting. */
rten_status rten_b200_mul(rten_ctx* ctx, const rten_tensor* a, const rten_tensor* b, rten_tensor* out);
/* MaxPool 2-D (src/ops/pooling.rs): kernel {kh,kw}; pads/strides as conv; padding never wins. */
rten_status rten_b200_max_pool(rten_ctx* ctx, const rten_tensor* x, const int32_t kernel[2], const int32_t pads[4],
                               const int32_t strides[2], rten_tensor* out);
rten_status rten_b200_global_average_pool(rten_ctx* ctx, const rten_tensor* x, rten_tensor* out);
/* Gather along axis 0 of a 2-D table with i32 indices (embedding lookups, src/ops/gather.rs). */
rten_status rten_b200_gather_rows(rten_ctx* ctx, const rten_tensor* table, const rten_tensor* indices_i32,
                                  rten_tensor* out);
/* table[indices[r], :] = updates[r, :] in place (ScatterElements / ScatterND restricted to whole rows of a 2-D f32 table,
 * distinct indices): the KV-cache append when the write position is a device-resident value, so that a decode step is
 * a fixed list of launches and can be replayed as a CUDA graph. */
rten_status rten_b200_scatter_rows(rten_ctx* ctx, rten_tensor* table, const rten_tensor* indices, const rten_tensor* updates);

/* ---- model loading and graph execution ------------------------------------------------- */
/* `Model::load` + `Graph::run_plan` (src/model.rs, src/graph.rs:880-1286) for the hot-path operator set: the ONNX file is
 * decoded by a hand-written wire-format reader (rten-onnx/src/onnx.rs), int64 tensors become i32 as in rten's loader,
 * constants are uploaded to HBM once, Conv + Relu and MatMul + Add(bias) are fused at load (the subset of
 * src/optimize.rs these models need), constant weights are prepacked once (`Operator::prepack`, src/graph.rs:488-565).
 * A run executes the nodes in topological order, one operator call of this library each; temporaries are reference
 * counted and return to the context pool after their last consumer; Relu / Gelu / Erf / Softmax run in place when the
 * executor holds the last reference to their input (src/graph.rs:973-1049); Reshape / Flatten / Squeeze / Unsqueeze /
 * Transpose / Identity are views.  Operators: Conv, ConvInteger, Relu, MaxPool, GlobalAveragePool, ReduceMean (spatial
 * axes), Gemm, MatMul, MatMulInteger, Add, Mul, Softmax, LayerNormalization, Gelu, Erf, Gather (rows), Cast (i32 -> f32),
 * DynamicQuantizeLinear, Attention (4-D), Constant and the view operators; anything else fails the LOAD with
 * RTEN_ERR_UNSUPPORTED_VALUE ("unsupported operator <name>"). */
typedef struct rten_model rten_model;
rten_status rten_b200_model_load(rten_ctx* ctx, const void* onnx_bytes, size_t len, rten_model** out);
void rten_b200_model_free(rten_model* model);
int32_t rten_b200_model_num_inputs(const rten_model* model);
int32_t rten_b200_model_num_outputs(const rten_model* model);
const char* rten_b200_model_input_name(const rten_model* model, int32_t index);
const char* rten_b200_model_output_name(const rten_model* model, int32_t index);
int32_t rten_b200_model_num_nodes(const rten_model* model); /* after the load-time fusions */
const char* rten_b200_model_node_op(const rten_model* model, int32_t index);
const char* rten_b200_model_summary(const rten_model* model); /* JSON: the decoded file (before fusion) */
/* Inputs by name (device tensors, or host tensors staged for the run; integer inputs are i32).  Outputs by name: any
 * value of the graph may be requested (`Model::run` with arbitrary output nodes); each comes back as a contiguous device
 * tensor the caller owns (rten_b200_free). */
rten_status rten_b200_model_run(rten_model* model, int32_t n_inputs, const char* const* input_names, const rten_tensor* inputs,
                                int32_t n_outputs, const char* const* output_names, rten_tensor* outputs);
/* The reader alone -- no context, no GPU: JSON description (opset, nodes with operator / inputs / outputs / attribute
 * names, initialisers with type and shape, graph inputs / outputs) of an ONNX file.  `needed` receives the size of the
 * full text incl. the terminator. */
rten_status rten_b200_onnx_summary(const void* onnx_bytes, size_t len, char* json_out, size_t capacity, size_t* needed);

#ifdef __cplusplus
}
#endif
#endif /* RTEN_B200_H */
