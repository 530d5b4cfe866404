#include <algorithm>
#include <unordered_set>

#include "llo/plan.hpp"
#include "llo/session.hpp"
#include "llo/sharded.hpp"
#include "llo/zprune.hpp"

#ifdef LLO_SHARDED_HPP

namespace llo
{

ShardPlan::ShardPlan (size_t batch, size_t world) : batch_(batch), world_(world)
{
	if (world < 1 || batch < world)
	{
		logs::fatalf("cannot shard %d rows over %d ranks", (int) batch, (int) world);
	}
}

std::pair<size_t,size_t> ShardPlan::bounds (size_t rank) const
{
	if (rank >= world_)
	{
		logs::fatalf("rank %d outside [0, %d)", (int) rank, (int) world_);
	}
	size_t base = batch_ / world_;
	size_t extra = batch_ % world_;
	size_t lo = rank * base + std::min(rank, extra);
	return {lo, lo + base + (rank < extra ? 1 : 0)};
}

GradientPack::GradientPack (const std::vector<ade::Shape>& shapes,
	age::_GENERATED_DTYPE dtype) : dtype_(dtype), shapes_(shapes)
{
	size_t align = std::max<size_t>(1, 16 / age::type_size(dtype));
	for (const ade::Shape& shape : shapes)
	{
		size_t n = shape.n_elems();
		offsets_.push_back(count_);
		sizes_.push_back(n);
		count_ += (n + align - 1) / align * align;
	}
}

/// true when `needle` is reachable from `root` (other than root itself)
static bool reaches (ade::iTensor* root, const std::unordered_set<ade::iTensor*>& needles)
{
	std::unordered_set<ade::iTensor*> seen;
	std::vector<ade::iTensor*> todo;
	auto push_children = [&](ade::iTensor* t)
	{
		if (auto f = dynamic_cast<ade::iFunctor*>(t))
		{
			for (const ade::MappedTensor& child : f->get_children())
			{
				todo.push_back(child.get_tensor().get());
			}
		}
	};
	push_children(root);
	while (false == todo.empty())
	{
		ade::iTensor* t = todo.back();
		todo.pop_back();
		if (false == seen.emplace(t).second)
		{
			continue;
		}
		if (needles.end() != needles.find(t))
		{
			return true;
		}
		push_children(t);
	}
	return false;
}

ShardedEvaluator::ShardedEvaluator (ade::TensptrT loss, ade::TensT params,
	ShardedOptions options) : params_(params), options_(options)
{
	ade::TensT roots = {loss};
	for (const ade::TensptrT& param : params)
	{
		roots.push_back(derive(loss, param.get()));
	}
	if (0 != options.passes_)
	{
		// the loss is rewritten together with the gradients so that they keep
		// sharing the forward pass
		roots = optimize(roots, options.passes_, &report_);
	}
	roots_ = roots;
	// folding a gradient in place while later roots are still computed is only
	// safe if no later root reads it
	std::unordered_set<ade::iTensor*> grads;
	for (size_t i = 1; i < roots_.size(); ++i)
	{
		if (nullptr != dynamic_cast<ade::iFunctor*>(roots_[i].get()))
		{
			grads.emplace(roots_[i].get());
		}
	}
	for (size_t i = 0; i < roots_.size() && independent_; ++i)
	{
		independent_ = false == reaches(roots_[i].get(), grads);
	}
	// what a captured step depends on: every leaf, and whether it may be captured
	std::unordered_set<ade::iTensor*> seen;
	std::vector<ade::iTensor*> todo;
	for (const ade::TensptrT& root : roots_)
	{
		todo.push_back(root.get());
	}
	while (false == todo.empty())
	{
		ade::iTensor* t = todo.back();
		todo.pop_back();
		if (false == seen.emplace(t).second)
		{
			continue;
		}
		if (auto f = dynamic_cast<ade::iFunctor*>(t))
		{
			size_t code = f->get_opcode().code_;
			if (age::RAND_BINO == code || age::RAND_UNIF == code || age::RAND_NORM == code)
			{
				capturable_ = false;
			}
			for (const ade::MappedTensor& child : f->get_children())
			{
				todo.push_back(child.get_tensor().get());
			}
		}
		else if (auto var = dynamic_cast<Variable*>(t))
		{
			leaves_.push_back(var);
		}
		else
		{
			capturable_ = false; // a foreign leaf is copied over on every visit
		}
	}
}

std::vector<ShardedEvaluator::LeafState> ShardedEvaluator::leaf_states (
	age::_GENERATED_DTYPE dtype)
{
	std::vector<LeafState> out;
	for (Variable* leaf : leaves_)
	{
		// (makes the mirror current: uploads and conversions happen here, not
		// inside a capture; constant leaves are immediates and have no mirror)
		const char* mirror = leaf->is_filled() ? nullptr : leaf->device_data(dtype).get();
		out.push_back(LeafState{leaf, leaf->version(), mirror});
	}
	return out;
}

ShardedEvaluator::Layout& ShardedEvaluator::layout (
	age::_GENERATED_DTYPE dtype, bool with_loss)
{
	std::pair<int,bool> key((int) dtype, with_loss);
	for (size_t i = 0; i < layout_keys_.size(); ++i)
	{
		if (layout_keys_[i] == key)
		{
			return layouts_[i];
		}
	}
	std::vector<ade::Shape> shapes;
	if (with_loss)
	{
		shapes.push_back(roots_[0]->shape());
	}
	for (const ade::TensptrT& param : params_)
	{
		shapes.push_back(param->shape());
	}
	Layout lay;
	lay.pack_ = GradientPack(shapes, dtype);
	lay.buffer_ = device::alloc(lay.pack_.nbytes());
	// padding between slots takes part in the fold: keep it defined
	device::check(llo_cuda_memset(device::ctx(), lay.buffer_.get(), 0,
		lay.pack_.nbytes()));
	size_t esize = age::type_size(dtype);
	for (size_t offset : lay.pack_.offsets_)
	{
		// aliasing pointers: the slot keeps the buffer alive
		lay.slots_.push_back(std::shared_ptr<char>(lay.buffer_,
			lay.buffer_.get() + offset * esize));
	}
	layout_keys_.push_back(key);
	layouts_.push_back(lay);
	return layouts_.back();
}

const GradientPack& ShardedEvaluator::pack (age::_GENERATED_DTYPE dtype,
	bool with_loss)
{
	return layout(dtype, with_loss).pack_;
}

std::shared_ptr<char> ShardedEvaluator::grad_eval_device (
	age::_GENERATED_DTYPE dtype, bool with_loss)
{
	Layout& lay = layout(dtype, with_loss);
	llo_cuda_ctx* ctx = device::ctx();
	bool use_graph = options_.graph_ && capturable_ && false == lay.no_graph_ &&
		0 != get_option("plan");
	if (false == use_graph)
	{
		enqueue(lay, dtype, with_loss);
		return lay.buffer_;
	}
	std::vector<LeafState> now = leaf_states(dtype);
	long opts = get_option("tf32x3") * 2 + get_option("sequential_reduce");
	if (nullptr != lay.graph_ && now == lay.graph_leaves_ && opts == lay.graph_options_)
	{
		device::check(llo_cuda_graph_launch(ctx, lay.graph_.get(), lay.graph_launches_));
		++replays_;
		return lay.buffer_;
	}
	lay.graph_ = nullptr;
	if (opts != lay.warm_options_)
	{
		lay.warm_options_ = opts;
		lay.evals_ = 0;
	}
	if (lay.evals_ < 1)
	{
		// the first step under a set of options runs eagerly: kernels are
		// specialised and loaded, the scratch arena and the memory pool reach
		// their working size (none of which may happen inside a capture)
		enqueue(lay, dtype, with_loss);
		++lay.evals_;
		return lay.buffer_;
	}
	uint64_t before = llo_cuda_launch_count(ctx);
	device::check(llo_cuda_graph_begin(ctx));
	void* exec = nullptr;
	bool captured = true;
	try
	{
		enqueue(lay, dtype, with_loss);
		device::check(llo_cuda_graph_end(ctx, &exec));
	}
	catch (const std::exception&)
	{
		// something in the step cannot be captured: stay eager from here on
		llo_cuda_graph_abort(ctx);
		captured = false;
	}
	if (false == captured || nullptr == exec)
	{
		lay.no_graph_ = true;
		enqueue(lay, dtype, with_loss);
		return lay.buffer_;
	}
	lay.graph_ = std::shared_ptr<void>(exec,
		[ctx](void* e) { llo_cuda_graph_destroy(ctx, e); });
	lay.graph_launches_ = llo_cuda_launch_count(ctx) - before;
	lay.graph_leaves_ = now;
	lay.graph_options_ = opts;
	// nothing ran while capturing: this step is the graph's first launch
	device::check(llo_cuda_graph_launch(ctx, exec, 0));
	return lay.buffer_;
}

void ShardedEvaluator::enqueue (Layout& lay, age::_GENERATED_DTYPE dtype,
	bool with_loss)
{
	llo_cuda_ctx* ctx = device::ctx();
	size_t esize = age::type_size(dtype);
	size_t nparams = params_.size();
	bool many = false == options_.solo_ && llo_cuda_comm_size(ctx) > 1;
	bool overlap = many && options_.overlap_ && false == options_.ordered_ &&
		independent_ && nparams > 0;

	// evaluation order: the loss, then the gradients -- of the first
	// parameters first (deep_first_, see ShardedOptions) or of the last
	// parameters first (backward order: they are complete first)
	ade::TensT roots;
	std::vector<std::shared_ptr<char>> dests;
	std::vector<size_t> slot_of;
	if (with_loss)
	{
		roots.push_back(roots_[0]);
		dests.push_back(lay.slots_[0]);
		slot_of.push_back(0);
	}
	bool deep = overlap && options_.deep_first_;
	for (size_t j = 0; j < nparams; ++j)
	{
		size_t k = deep ? j : nparams - 1 - j;
		size_t slot = k + (with_loss ? 1 : 0);
		roots.push_back(roots_[1 + k]);
		dests.push_back(lay.slots_[slot]);
		slot_of.push_back(slot);
	}
	// Buckets: slots are contiguous in the buffer, so a run of roots that is
	// complete folds with one call.  Two gradients (weights + bias of a
	// layer) per bucket keeps the calls few; the loss (slot 0, evaluated
	// first) joins the bucket that holds slot 1.
	size_t first_grad = with_loss ? 1 : 0;
	std::function<void(size_t)> on_done = nullptr;
	if (overlap)
	{
		on_done = [&](size_t i)
		{
			if (i < first_grad)
			{
				return; // the loss travels with a gradient bucket
			}
			size_t g = i - first_grad;       // g-th gradient in evaluation order
			bool last = g + 1 == nparams;
			if (0 == g % 2 && false == last)
			{
				return; // first of a pair: wait for its partner
			}
			size_t a_slot = slot_of[i];
			size_t b_slot = slot_of[i - g % 2];
			size_t lo_slot = std::min(a_slot, b_slot);
			size_t hi_slot = std::max(a_slot, b_slot);
			if (with_loss && 1 == lo_slot)
			{
				lo_slot = 0;
			}
			size_t lo = lay.pack_.offsets_[lo_slot];
			size_t hi = hi_slot + 1 < lay.pack_.offsets_.size() ?
				lay.pack_.offsets_[hi_slot + 1] : lay.pack_.count_;
			device::check(llo_cuda_allreduce_async(ctx,
				lay.buffer_.get() + lo * esize, hi - lo, (int) dtype, age::SUM));
		};
	}
	planned_eval_many(roots, dtype, &dests, on_done);
	if (overlap)
	{
		device::check(llo_cuda_comm_join(ctx));
	}
	else if (many)
	{
		device::check(llo_cuda_allreduce(ctx, lay.buffer_.get(),
			lay.pack_.count_, (int) dtype, age::SUM,
			options_.ordered_ ? LLO_ALLREDUCE_ORDERED : LLO_ALLREDUCE_FAST));
	}
}

std::vector<GenericData> ShardedEvaluator::grad_eval (
	age::_GENERATED_DTYPE dtype, bool with_loss)
{
	std::shared_ptr<char> buffer = grad_eval_device(dtype, with_loss);
	const GradientPack& pk = pack(dtype, with_loss);
	size_t esize = age::type_size(dtype);
	std::vector<GenericData> outs;
	for (size_t i = 0; i < pk.shapes_.size(); ++i)
	{
		GenericData out;
		out.shape_ = pk.shapes_[i];
		out.dtype_ = dtype;
		out.device_ = std::shared_ptr<char>(buffer,
			buffer.get() + pk.offsets_[i] * esize);
		out.fetch();
		outs.push_back(out);
	}
	return outs;
}

}

#endif
