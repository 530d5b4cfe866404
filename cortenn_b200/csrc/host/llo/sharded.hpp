///
/// sharded.hpp
/// llo
///
/// Purpose:
/// Batch-sharded gradient evaluation (north star item 5): the
/// one multi-GPU executor of the evaluator, in C++ like the rest of the host
/// layer.  The reference evaluates one graph in one thread on one host and has
/// no communication layer.
///
/// One process drives one GPU.  Every rank builds the same graph over its
/// contiguous slice of the batched leaves (in ade order the batch dimension is
/// the slowest one: llo/python/llo.cpp:18-22), scaled by the GLOBAL batch, and
/// hands it to a ShardedEvaluator, which
///   1. derives the gradients once (llo::derive, llo/src/zprune.cpp:109-116)
///      and rewrites them with the named llo::optimize passes,
///   2. evaluates all of them in ONE plan, every gradient written straight
///      into its slot of one packed device buffer (no pack copies),
///   3. folds the buffer over the ranks -- bucket by bucket on a second lane
///      (llo_cuda_allreduce_async), so a layer's gradients travel over NVLink
///      while the other layers' weight-gradient GEMMs still run (see
///      ShardedOptions::deep_first_ for the order); or with one ordered fold
///      (bit-identical on every rank),
///   4. from the second call on replays the whole step from one CUDA graph.
///

#ifndef LLO_SHARDED_HPP
#define LLO_SHARDED_HPP

#include "llo/data.hpp"
#include "llo/optimize.hpp"

namespace llo
{

/// Contiguous split of `batch` rows over `world` ranks; the remainder goes to
/// the first ranks one row each
struct ShardPlan final
{
	ShardPlan (size_t batch, size_t world);

	/// [first, last) rows of `rank`
	std::pair<size_t,size_t> bounds (size_t rank) const;

	size_t rows (size_t rank) const
	{
		auto b = bounds(rank);
		return b.second - b.first;
	}

	size_t batch_;

	size_t world_;
};

/// Layout of several tensors in one flat buffer: offsets in elements, every
/// slot rounded up to 16 bytes so each tensor stays vector-aligned
struct GradientPack final
{
	GradientPack (void) = default;

	GradientPack (const std::vector<ade::Shape>& shapes,
		age::_GENERATED_DTYPE dtype);

	size_t nbytes (void) const
	{
		return count_ * age::type_size(dtype_);
	}

	age::_GENERATED_DTYPE dtype_ = age::BAD_TYPE;

	std::vector<ade::Shape> shapes_;

	std::vector<size_t> sizes_;    // elements of each tensor

	std::vector<size_t> offsets_;  // first element of each slot

	size_t count_ = 0;             // elements of the whole buffer
};

struct ShardedOptions
{
	/// llo::optimize passes applied to [loss] + gradients (0 = none).
	/// OPT_DIV_CANCEL is what lets the matmul gradients lower to GEMMs.
	unsigned passes_ = OPT_ALL;

	/// one all-gather + rank-ordered fold instead of ncclAllReduce:
	/// bit-identical sums on every rank, not overlapped
	bool ordered_ = false;

	/// start each gradient's fold as soon as it is computed
	bool overlap_ = true;

	/// Evaluation order of the gradients when the folds overlap.
	/// true : the FIRST parameters' gradients first.  Their derivative chain
	///        runs through every layer, so the forward pass and the whole
	///        backward chain -- the big persistent GEMMs, one CTA per SM with a
	///        fixed unit schedule, which a collective's CTAs landing on some SMs
	///        would delay as a whole -- finish before the first fold starts;
	///        the folds then run beside the per-layer weight-gradient GEMMs,
	///        whose CTAs the hardware hands out as slots free up.
	/// false: backward order (last layer's gradients first, each fold starts
	///        as early as possible, beside the rest of the chain).
	/// Measured at 8 GPUs, MLP 784-1024-1024-10 / batch 65536: backward order
	/// 1.933 ms -- slower than no overlap at all, 1.887 ms (tools/_ab/n8_prio.sh).
	bool deep_first_ = true;

	/// evaluate on this rank alone whatever the communicator says (the
	/// single-GPU result a multi-rank run checks itself against)
	bool solo_ = false;

	/// From the second evaluation on, replay the whole step -- every kernel,
	/// the stream-ordered allocations of the intermediates, the collectives
	/// on the second lane -- from ONE CUDA graph (llo_cuda_graph_*), for as
	/// long as no leaf has been assigned to: the planner and ~40 launches
	/// per step leave the host's critical path.  Graphs with random draws
	/// are never captured (a replay would repeat the draw).
	bool graph_ = true;
};

/// Gradient evaluation of one loss w.r.t. a set of parameters, summed over
/// the ranks of the context's communicator (llo_cuda_comm_init)
struct ShardedEvaluator final
{
	ShardedEvaluator (ade::TensptrT loss, ade::TensT params,
		ShardedOptions options = ShardedOptions());

	/// One gradient evaluation.  Returns the packed device buffer: the
	/// summed gradients in parameter order (preceded by the summed loss when
	/// with_loss), laid out as pack(dtype, with_loss) says.  The buffer is
	/// reused by the next call.
	std::shared_ptr<char> grad_eval_device (age::_GENERATED_DTYPE dtype,
		bool with_loss = false);

	/// The same, fetched: one GenericData (host mirror filled) per tensor
	std::vector<GenericData> grad_eval (age::_GENERATED_DTYPE dtype,
		bool with_loss = false);

	const GradientPack& pack (age::_GENERATED_DTYPE dtype, bool with_loss);

	/// Report of the optimize passes applied at construction
	OptReport report_;

	/// The (rewritten) roots: loss first, then one gradient per parameter
	ade::TensT roots_;

private:
	/// what a captured step read: a leaf, its version, its device mirror
	struct LeafState
	{
		Variable* leaf_;
		uint64_t version_;
		const char* mirror_;

		bool operator == (const LeafState& o) const
		{
			return leaf_ == o.leaf_ && version_ == o.version_ && mirror_ == o.mirror_;
		}
	};

	struct Layout
	{
		GradientPack pack_;
		std::shared_ptr<char> buffer_;
		std::vector<std::shared_ptr<char>> slots_;  // aliases into buffer_
		size_t evals_ = 0;                          // eager evaluations with warm_options_
		long warm_options_ = -1;
		bool no_graph_ = false;                     // a capture failed: stay eager
		std::shared_ptr<void> graph_;               // captured step (cudaGraphExec_t)
		uint64_t graph_launches_ = 0;
		std::vector<LeafState> graph_leaves_;
		long graph_options_ = 0;
	};

	Layout& layout (age::_GENERATED_DTYPE dtype, bool with_loss);

	/// enqueue one step (plan, kernels, collectives) on the context's stream
	void enqueue (Layout& lay, age::_GENERATED_DTYPE dtype, bool with_loss);

	std::vector<LeafState> leaf_states (age::_GENERATED_DTYPE dtype);

	std::vector<Variable*> leaves_;

	bool capturable_ = true;    // only Variable leaves, no random draws

public:
	/// steps replayed from a captured graph so far (tests, bench)
	size_t replays_ = 0;

private:

	ade::TensT params_;

	ShardedOptions options_;

	bool independent_ = true;   // no root is an input of another root

	std::vector<Layout> layouts_;

	std::vector<std::pair<int,bool>> layout_keys_;
};

}

#endif // LLO_SHARDED_HPP
