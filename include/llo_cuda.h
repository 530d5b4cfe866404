/*
 * llo_cuda.h — C-ABI of the B200 (sm_100a) LLO operator library `libllo_cuda.so`.
 *
 * This is the drop-in boundary for the hot path of mingkaic/cortenn's LLO
 * evaluator: every entry point replaces one CPU loop (or one generated
 * dispatch) of the reference and is what a reference-side FFI for the path
 * would bind (INTEGRATION.md shows the C++ shim).  Citations are relative to
 * the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary
 *   - every `void* out` / `data` pointer is a DEVICE pointer unless the name
 *     says host; shapes are rank-8, dimension 0 fastest (ade::Shape)
 *   - `coorder` is the 9x9 row-major matrix of the argument's ade::CoordMap:
 *     out[j] = sum_{i<8} in[i]*M[i][j] + M[8][j]
 *   - dtype / opcode integers are age::_GENERATED_DTYPE / _GENERATED_OPCODE
 *     (LLO_DT_*, LLO_OP_* below; same order as llo/cfg/llo.json:13-124)
 *   - return value 0 = ok, non-zero = error; the message is available from
 *     llo_cuda_last_error() (thread-local).  LLO_ERR_UNSUPPORTED_TYPED_OP
 *     mirrors the reference's std::bad_function_call
 *     (llo/src/operator.cpp:39-60, llo/operator.hpp:352-355); every other
 *     failure mirrors logs::fatal (std::runtime_error).
 *   - one calling thread per context; work is enqueued on the context's
 *     stream and is asynchronous until llo_cuda_sync / llo_cuda_d2h.
 */
#ifndef LLO_CUDA_H
#define LLO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLO_CUDA_ABI_VERSION 1
#define LLO_RANK 8
#define LLO_MAT 81

/* age::_GENERATED_DTYPE (llo/cfg/llo.json:13-24) */
enum {
	LLO_DT_BAD = 0, LLO_DT_DOUBLE, LLO_DT_FLOAT, LLO_DT_INT8, LLO_DT_UINT8,
	LLO_DT_INT16, LLO_DT_UINT16, LLO_DT_INT32, LLO_DT_UINT32, LLO_DT_INT64,
	LLO_DT_UINT64, LLO_DT_COUNT
};

/* age::_GENERATED_OPCODE (llo/cfg/llo.json:31-124) */
enum {
	LLO_OP_BAD = 0, LLO_OP_ABS, LLO_OP_NEG, LLO_OP_SIN, LLO_OP_COS, LLO_OP_TAN,
	LLO_OP_EXP, LLO_OP_LOG, LLO_OP_SQRT, LLO_OP_ROUND, LLO_OP_POW, LLO_OP_SUM,
	LLO_OP_SUB, LLO_OP_PROD, LLO_OP_DIV, LLO_OP_MIN, LLO_OP_MAX, LLO_OP_EQ,
	LLO_OP_NEQ, LLO_OP_LT, LLO_OP_GT, LLO_OP_RAND_BINO, LLO_OP_RAND_UNIF,
	LLO_OP_RAND_NORM, LLO_OP_COUNT
};

enum {
	LLO_OK = 0,
	LLO_ERR_FATAL = 1,               /* logs::fatal equivalent */
	LLO_ERR_UNSUPPORTED_TYPED_OP = 2, /* std::bad_function_call equivalent */
	LLO_ERR_CUDA = 3                 /* CUDA runtime failure */
};

typedef struct llo_cuda_ctx llo_cuda_ctx;

/* One evaluated argument: the information in llo::DataArg
 * (llo/data.hpp:224-238). */
typedef struct llo_cuda_arg
{
	const void* data;         /* device pointer to the argument's elements */
	uint64_t shape[LLO_RANK]; /* argument shape */
	double coorder[LLO_MAT];  /* DataArg::mapper_ */
	int32_t push;             /* DataArg::fwd_: 1 = maps arg coords to out coords */
	int32_t reserved;
} llo_cuda_arg;

/* ---- context, memory, transfer (device-resident llo data store;
 *      replaces malloc/free of llo/src/data.cpp:8-18) ---- */
int llo_cuda_init (int device, llo_cuda_ctx** ctx);
int llo_cuda_destroy (llo_cuda_ctx* ctx);
const char* llo_cuda_last_error (void);
int llo_cuda_abi_version (void);
/* adopt an external cudaStream_t (e.g. torch's current stream); NULL = own */
int llo_cuda_set_stream (llo_cuda_ctx* ctx, void* cuda_stream);
void* llo_cuda_get_stream (llo_cuda_ctx* ctx);
int llo_cuda_device (llo_cuda_ctx* ctx);
int llo_cuda_sm_count (llo_cuda_ctx* ctx);
int llo_cuda_alloc (llo_cuda_ctx* ctx, size_t bytes, void** dptr);
int llo_cuda_free (llo_cuda_ctx* ctx, void* dptr);
int llo_cuda_host_alloc (llo_cuda_ctx* ctx, size_t bytes, void** hptr); /* pinned */
int llo_cuda_host_free (llo_cuda_ctx* ctx, void* hptr);
int llo_cuda_h2d (llo_cuda_ctx* ctx, void* dst, const void* src_host, size_t bytes);
int llo_cuda_d2h (llo_cuda_ctx* ctx, void* dst_host, const void* src, size_t bytes);
int llo_cuda_d2d (llo_cuda_ctx* ctx, void* dst, const void* src, size_t bytes);
/* Copy-engine lanes: transfers that overlap each other (PCIe is full duplex)
 * and the kernels already queued on the context's stream.  The reference
 * copies synchronously (llo/python/llo.cpp:101-147); these serve a pipelined
 * caller that uploads the next input while the last result drains.
 *   h2d_async: `dst` must be a fresh allocation (nothing queued reads or
 *     writes it).  The copy starts once the work queued on the stream so far
 *     has finished; everything queued afterwards waits for it.  `src_host`
 *     (pinned) must stay untouched until the next llo_cuda_sync /
 *     llo_cuda_event_sync of a later download.
 *   d2h_async: the copy starts once the work queued on the stream so far has
 *     finished and does not hold up later work; `*done` receives an event to
 *     pass to llo_cuda_event_sync (then llo_cuda_event_destroy).  `src` must
 *     stay allocated until then.  A device-side validation error raised by
 *     the work before the copy is reported by that llo_cuda_event_sync (the
 *     flag is snapshotted on the download lane right behind the copy). */
int llo_cuda_h2d_async (llo_cuda_ctx* ctx, void* dst, const void* src_host, size_t bytes);
int llo_cuda_d2h_async (llo_cuda_ctx* ctx, void* dst_host, const void* src, size_t bytes, void** done);
int llo_cuda_event_sync (llo_cuda_ctx* ctx, void* event);
int llo_cuda_memset (llo_cuda_ctx* ctx, void* dst, int byte, size_t bytes);
/* order the context's stream behind another cudaStream_t (0 = the legacy
 * default stream): device buffers produced by another library may then be
 * read by calls made afterwards (zero-copy interop, llo/python/llo.cpp:36-99
 * only knows host arrays) */
int llo_cuda_wait_stream (llo_cuda_ctx* ctx, void* cuda_stream);
int llo_cuda_sync (llo_cuda_ctx* ctx);
/* CUDA-graph replay of a fixed sequence of calls on this context (the same
 * plan over the same buffers, e.g. one gradient evaluation per training step):
 * llo_cuda_graph_begin starts capturing everything queued on the context's
 * stream (kernels, stream-ordered allocations and frees, copies, collectives
 * on the second lane), llo_cuda_graph_end instantiates what was captured,
 * llo_cuda_graph_launch replays it (`launches` = kernels inside, for
 * llo_cuda_launch_count).  Nothing runs while capturing.  Calls that
 * synchronise (h2d / d2h / sync) must not be made between begin and end;
 * llo_cuda_graph_abort ends a capture that went wrong. */
int llo_cuda_graph_begin (llo_cuda_ctx* ctx);
int llo_cuda_graph_end (llo_cuda_ctx* ctx, void** graph_exec);
int llo_cuda_graph_abort (llo_cuda_ctx* ctx);
int llo_cuda_graph_launch (llo_cuda_ctx* ctx, void* graph_exec, uint64_t launches);
int llo_cuda_graph_destroy (llo_cuda_ctx* ctx, void* graph_exec);
/* number of kernels this context has launched so far (bench `gpu_launches`) */
uint64_t llo_cuda_launch_count (llo_cuda_ctx* ctx);
/* CUDA events on the context's stream, for device-side timing */
int llo_cuda_event_create (llo_cuda_ctx* ctx, void** event);
int llo_cuda_event_record (llo_cuda_ctx* ctx, void* event);
int llo_cuda_event_elapsed_ms (llo_cuda_ctx* ctx, void* start, void* stop, float* ms);
int llo_cuda_event_destroy (llo_cuda_ctx* ctx, void* event);

/* Run-time specialisation of the fused elementwise engine: launches of at
 * least `elements` outputs (default 2^20; 0 = every launch, UINT64_MAX = none)
 * run a build of the kernel compiled for their program with NVRTC (sm_100a,
 * cached in memory and under LLO_KERNEL_CACHE / beside the library); smaller
 * ones, and any launch whose build fails, run the interpreted kernel.  Both
 * apply the same operator code in the same order: results are bit-identical.
 * stats = {specialised launches, NVRTC compilations, disk-cache hits, failed builds} */
int llo_cuda_set_specialize_min (llo_cuda_ctx* ctx, uint64_t elements);
int llo_cuda_specialize_stats (llo_cuda_ctx* ctx, uint64_t stats[4]);

/* ---- K0: leaf dtype conversion, GenericData::copyover
 *      (llo/src/data.cpp:20-72) ---- */
int llo_cuda_convert (llo_cuda_ctx* ctx, void* out, int out_dtype,
	const void* in, int in_dtype, uint64_t n);

/* ---- the three generic loops of llo/operator.hpp ---- */
/* llo::unary (operator.hpp:31-61) with the typed lambdas of :66-165.
 * As in the reference wrappers the output shape is the argument's shape
 * for identity maps; `outshape` is used for mapped arguments. */
int llo_cuda_unary (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK], const llo_cuda_arg* in);
/* llo::binary (operator.hpp:169-231) with pow sub div eq neq lt gt (:237-301).
 * RAND_* opcodes route to llo_cuda_rand. */
int llo_cuda_binary (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const llo_cuda_arg* a, const llo_cuda_arg* b);
/* llo::nnary (operator.hpp:367-414) with add mul min max (:419-451):
 * first touch assigns, later touches accumulate, argument order then
 * element order.  Outputs no argument touches are zero (reference:
 * uninitialised). */
int llo_cuda_nnary (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const llo_cuda_arg* args, int nargs);
/* age::op_exec (call site llo/eval.hpp:90): opcode switch over the above */
int llo_cuda_op_exec (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const llo_cuda_arg* args, int nargs);

/* ---- K9: rand_binom / rand_uniform / rand_normal
 *      (operator.hpp:307-363, operator.cpp:63-133).  Counter-based Philox
 *      stream; statistical parity only (test_api.cpp:1016-1159).
 *      b is DOUBLE for RAND_BINO (eval.hpp:50-74), `dtype` otherwise. ---- */
int llo_cuda_seed (llo_cuda_ctx* ctx, uint64_t seed);
int llo_cuda_rand (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const llo_cuda_arg* a, const llo_cuda_arg* b);

/* ---- K4: reductions = nnary with one push argument whose coorder is a
 *      reduce map (llo/src/helper.cpp:57-65).  Reduces dims >= first_dim of
 *      `inshape`.  mode 0 = fixed-shape tree (deterministic), 1 = reference
 *      sequential order (bit-parity for floating sums). ---- */
enum { LLO_REDUCE_TREE = 0, LLO_REDUCE_SEQUENTIAL = 1 };
int llo_cuda_reduce (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const void* in, const uint64_t inshape[LLO_RANK], int first_dim, int mode);
int llo_cuda_set_reduce_mode (llo_cuda_ctx* ctx, int mode);

/* ---- strided views + fused elementwise programs (planner back end) ---- */
/* A view reads element  data[offset + sum_i coord_i * stride[i]]  for output
 * coordinate coord (strides in elements, may be 0 or negative). */
typedef struct llo_cuda_view
{
	const void* data;
	int64_t stride[LLO_RANK];
	int64_t offset;
} llo_cuda_view;

/* Stack-machine instruction: LLO_VM_LOAD pushes input `arg`; LLO_VM_CONST
 * pushes consts[arg]; an LLO_OP_* unary opcode replaces the top; a binary
 * opcode (POW SUB DIV EQ NEQ LT GT, and SUM PROD MIN MAX as 2-ary) pops two
 * and pushes one.  The value left on the stack is stored to `out`. */
enum { LLO_VM_LOAD = 64, LLO_VM_CONST = 65 };
typedef struct llo_cuda_insn { uint8_t op; uint8_t arg; } llo_cuda_insn;
#define LLO_VM_MAX_INPUTS 8
#define LLO_VM_MAX_INSNS 48
#define LLO_VM_MAX_CONSTS 8
#define LLO_VM_MAX_DEPTH 4
int llo_cuda_elementwise (llo_cuda_ctx* ctx, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const llo_cuda_view* inputs, int ninputs,
	const double* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns);
/* Reduction straight from a view: out[o] = fold over r of
 * src.data[src.offset + sum_d o_d*src.stride[d] + sum_k r_k*fold_stride[k]],
 * o over `outshape`, r over `fold_size` with fold dim 0 fastest (the
 * reference's accumulation order).  This is nnary's push branch
 * (operator.hpp:376-393) for a reduce map whose argument is itself a
 * permuted / strided view (e.g. reduce_sum(permute(x,{1,0}),1)), without
 * materialising the view.  accumulate: 1 = fold into the values in `out`. */
int llo_cuda_reduce_view (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK], const llo_cuda_view* src,
	int nfold, const uint64_t* fold_size, const int64_t* fold_stride,
	int accumulate, int mode);
/* ---- K6/K7: GEMM the composed matmul sub-graph lowers to
 *      (llo/src/helper.cpp:67-97).  C[m,n] = sum_k A[m,k]*B[k,n], k ascending
 *      per output in FFMA/DFMA order within a k-tile.  Element (m,k) of A is
 *      A[m*a_rs + k*a_cs], likewise B (k,n) and C (m,n); one of each stride
 *      pair must be 1.  precision: 0 = exact FMA pipe (fp32/fp64/int),
 *      1 = opt-in 3xTF32 tensor-core path (fp32 only). ---- */
enum { LLO_GEMM_EXACT = 0, LLO_GEMM_3XTF32 = 1 };
int llo_cuda_gemm (llo_cuda_ctx* ctx, int dtype,
	uint64_t M, uint64_t N, uint64_t K,
	const void* A, int64_t a_rs, int64_t a_cs,
	const void* B, int64_t b_rs, int64_t b_cs,
	void* C, int64_t c_rs, int64_t c_cs,
	int precision);

/* ---- K8: the exchange step of batch-sharded evaluation (no counterpart in
 *      the reference, which has no communication layer: SURVEY.md 2.2).  One
 *      context per GPU / process; rank 0 creates a 128-byte NCCL unique id,
 *      the host passes it to every rank (any out-of-band channel), every rank
 *      calls llo_cuda_comm_init.  llo_cuda_allreduce folds `n` elements of
 *      `buf` in place over all ranks with SUM / PROD / MIN / MAX on the
 *      context's stream.  mode LLO_ALLREDUCE_FAST = ncclAllReduce (ring / tree
 *      / NVLS as NCCL chooses; run-to-run deterministic for a fixed topology),
 *      LLO_ALLREDUCE_ORDERED = all-gather + fold in rank order (bit-identical
 *      on every rank and independent of NCCL's algorithm).  nranks == 1 is a
 *      no-op and needs no NCCL. ---- */
enum { LLO_ALLREDUCE_FAST = 0, LLO_ALLREDUCE_ORDERED = 1 };
int llo_cuda_comm_unique_id (void* id128);
int llo_cuda_comm_init (llo_cuda_ctx* ctx, int rank, int nranks, const void* id128);
int llo_cuda_comm_destroy (llo_cuda_ctx* ctx);
int llo_cuda_comm_rank (llo_cuda_ctx* ctx);
int llo_cuda_comm_size (llo_cuda_ctx* ctx);
int llo_cuda_allreduce (llo_cuda_ctx* ctx, void* buf, uint64_t n, int dtype,
	int opcode, int mode);
/* The same fold (LLO_ALLREDUCE_FAST) on a second lane: it starts once the work
 * queued on the context's stream so far has finished and runs beside the work
 * queued afterwards, so the gradients of the last layers travel over NVLink
 * while the earlier layers' backward GEMMs still run.  llo_cuda_comm_join makes
 * the stream's later work (and llo_cuda_sync / d2h) wait for every collective
 * started so far.  Calls must be made in the same order on every rank. */
int llo_cuda_allreduce_async (llo_cuda_ctx* ctx, void* buf, uint64_t n,
	int dtype, int opcode);
int llo_cuda_comm_join (llo_cuda_ctx* ctx);

#ifdef __cplusplus
}
#endif

#endif /* LLO_CUDA_H */
