///
/// plan.hpp
/// llo
///
/// Purpose:
/// Lower an ade graph to device work.  The reference interprets a graph node
/// by node and re-evaluates shared sub-graphs (llo/eval.hpp:41-91); every
/// data movement (extend / permute / flip / reduce) is its own n-ary functor
/// over a full-size tensor (llo/cfg/llo.json:199-208, 393-460), and matmul is
/// five [M,N,C] tensors (llo/src/helper.cpp:92-96).  The planner evaluates
/// the same graph with the same per-element semantics, but
///   * every distinct sub-graph once (structural hashing, so the copies the
///     derivative rules make of forward nodes are shared),
///   * single-argument SUM / PROD / MIN / MAX functors over an integer
///     coordinate map as VIEWS: strides folded into the consumer's loads,
///   * constant leaves (llo::get_scalar) as immediates,
///   * chains of elementwise functors as ONE fused program per output
///     (llo_cuda_elementwise),
///   * reductions straight from a view of their source (llo_cuda reduce),
///   * reduce_sum(mul(view(A), view(B))) contractions — matmul and its
///     gradients — as GEMM calls (llo_cuda_gemm), never materialising the
///     [M,N,C] product,
///   * anything else through the generic operator call (age::op_exec).
/// Optional algebra on derived graphs ("simplify" option): x*1 -> x (exact)
/// and (x*y)/x -> y (<= 1 ulp, defined at x == 0), which the reference's
/// product rule introduces (llo/src/helper.cpp:18-33).
///

#include <functional>

#include "llo/data.hpp"

#ifndef LLO_PLAN_HPP
#define LLO_PLAN_HPP

namespace llo
{

/// Counters of the last planned evaluation (exposed for tests / docs)
struct PlanStats
{
	size_t nodes_ = 0;        // distinct ade nodes reached
	size_t views_ = 0;        // functors folded into views
	size_t consts_ = 0;       // constant leaves folded into immediates
	size_t fused_ = 0;        // elementwise functors inlined into a consumer
	size_t programs_ = 0;     // elementwise program launches
	size_t reduces_ = 0;      // reduction launches
	size_t gemms_ = 0;        // contractions lowered to GEMM
	size_t gathers_ = 0;      // GEMM operands gathered into a dense matrix first (im2col)
	size_t generic_ = 0;      // functors run through age::op_exec
	size_t simplified_ = 0;   // algebraic rewrites applied
	size_t shared_ = 0;       // plan nodes merged with a structurally equal one
};

/// Evaluate `root` as `dtype` on the device; the result stays in HBM
GenericData planned_eval (ade::TensptrT root, age::_GENERATED_DTYPE dtype);

/// Evaluate several roots in one plan (shared sub-graphs computed once)
/// `dests` (optional, one per root, null entries allowed): device blocks the
/// results are to end up in -- a root that computes writes there directly (no
/// copy; the packed gradient buffer of the sharded evaluator).  `on_done(i)`
/// (optional) is called when everything root i needs has been enqueued, in
/// root order: the hook that starts root i's collective while the later
/// roots are still being computed.
std::vector<GenericData> planned_eval_many (const ade::TensT& roots,
	age::_GENERATED_DTYPE dtype,
	const std::vector<std::shared_ptr<char>>* dests = nullptr,
	const std::function<void(size_t)>& on_done = nullptr);

/// Statistics of the most recent planned_eval on this thread
const PlanStats& last_plan_stats (void);

/// What the planner makes of the graphs WITHOUT evaluating them (no device
/// needed): nodes, views, constants, sharing and how many functors fall back
/// to the generic per-functor path.  Launch counts (programs / reduces /
/// gemms) are only known after an evaluation.
PlanStats plan_preview (const ade::TensT& roots, age::_GENERATED_DTYPE dtype);

}

#endif // LLO_PLAN_HPP
