///
/// grader.hpp
/// bwd
///
/// Reverse-mode derivative construction over ade graphs.  The reference uses
/// Tenncor's `bwd` module (not vendored); this restates the behaviour its
/// call sites and gradient golden vectors pin:
///  * the rule for argument i of a functor receives
///      fwd  = the functor,
///      bwd  = MappedTensor(upstream gradient, child.shaper.reverse(),
///                          !child.map_io, child.coorder)
///      args = the functor's arguments expressed in argument i's space
///    (llo/cfg/llo.json:32-123 "derivative" strings use exactly these names)
///  * leaf derivatives are rules->data(1|0, target shape), whose labels
///    "1"/"0" drive llo::zero_prune (llo/src/zprune.cpp:96-105)
///  * golden vectors: llo/test/test_api.cpp:153-269 (elementwise),
///    :655-736 (reductions), :786-826 (extend), :829-916 (matmul)
///

#ifndef BWD_GRADER_HPP
#define BWD_GRADER_HPP

#include <list>

#include "ade/ade.hpp"

namespace bwd
{

/// Per-library derivative rules
struct iRuleSet
{
	virtual ~iRuleSet (void) = default;

	/// Leaf of given shape filled with scalar
	virtual ade::TensptrT data (double scalar, ade::Shape shape) = 0;

	/// Opcode of n-ary summation
	virtual ade::Opcode sum_opcode (void) = 0;

	/// Opcode of n-ary product
	virtual ade::Opcode prod_opcode (void) = 0;

	/// Chain rule of fwd with respect to its argument idx
	virtual ade::TensptrT grad_rule (ade::iFunctor* fwd,
		ade::MappedTensor bwd, ade::TensT args, size_t idx) = 0;
};

/// Traveler building d(visited)/d(target)
struct Grader final : public ade::iTraveler
{
	Grader (const ade::iTensor* target, std::shared_ptr<iRuleSet> rules) :
		target_(target), rules_(rules)
	{
		if (nullptr == target_)
		{
			logs::fatal("cannot derive with respect to null");
		}
		if (nullptr == rules_)
		{
			logs::fatal("cannot derive without ruleset");
		}
	}

	void visit (ade::iLeaf* leaf) override
	{
		derivatives_.emplace(leaf, rules_->data(
			leaf == target_ ? 1 : 0, target_->shape()));
	}

	void visit (ade::iFunctor* func) override
	{
		if (func == target_)
		{
			derivatives_.emplace(func, rules_->data(1, target_->shape()));
			return;
		}

		ade::PathFinder finder(target_);
		func->accept(finder);
		auto& pathmap = finder.parents_;
		if (pathmap.empty())
		{
			// target does not feed func
			derivatives_.emplace(func, rules_->data(0, target_->shape()));
			return;
		}

		// walk path functors from the root down: taller sub-graphs first,
		// later-discovered first among equals (deterministic)
		ade::GraphStat stat;
		func->accept(stat);
		std::vector<ade::iFunctor*> parents(
			finder.order_.rbegin(), finder.order_.rend());
		std::stable_sort(parents.begin(), parents.end(),
			[&](ade::iFunctor* a, ade::iFunctor* b)
			{
				return stat.graphsize_[a] > stat.graphsize_[b];
			});

		const ade::Opcode sum = rules_->sum_opcode();
		std::unordered_map<const ade::iTensor*,ade::TensT> grads = {
			{func, {rules_->data(1, func->shape())}},
		};
		for (ade::iFunctor* parent : parents)
		{
			ade::TensT& incoming = grads[parent];
			ade::TensptrT upstream = incoming.size() > 1 ?
				ade::TensptrT(ade::Functor::get(sum,
					ade::to_args(incoming))) : incoming[0];

			const ade::ArgsT& children = parent->get_children();
			std::vector<size_t> indices(
				pathmap[parent].begin(), pathmap[parent].end());
			std::sort(indices.begin(), indices.end());
			for (size_t i : indices)
			{
				const ade::MappedTensor& child = children[i];
				ade::CoordptrT revshaper(child.get_shaper()->reverse());
				bool plain = ade::is_identity(child.get_shaper().get()) &&
					ade::is_identity(child.get_coorder().get());

				ade::TensT args;
				for (size_t j = 0, n = children.size(); j < n; ++j)
				{
					const ade::MappedTensor& kid = children[j];
					if (j == i)
					{
						args.push_back(kid.get_tensor());
					}
					else if (plain &&
						ade::is_identity(kid.get_shaper().get()) &&
						ade::is_identity(kid.get_coorder().get()))
					{
						args.push_back(kid.get_tensor());
					}
					else
					{
						// sibling seen from the functor, then brought into
						// child i's coordinate space
						ade::TensptrT seen(ade::Functor::get(sum, {kid}));
						args.push_back(ade::TensptrT(ade::Functor::get(sum, {
							ade::MappedTensor(seen, revshaper,
								!child.map_io(), child.get_coorder())
						})));
					}
				}
				ade::TensptrT grad = rules_->grad_rule(parent,
					ade::MappedTensor(upstream, revshaper,
						!child.map_io(), child.get_coorder()), args, i);
				grads[child.get_tensor().get()].push_back(grad);
			}
		}

		ade::TensT& finals = grads[target_];
		derivatives_.emplace(func, finals.size() > 1 ?
			ade::TensptrT(ade::Functor::get(sum, ade::to_args(finals))) :
			finals[0]);
	}

	/// Target of tensor all visited nodes are derived with respect to
	const ade::iTensor* target_;

	/// Map visited nodes to their derivatives
	std::unordered_map<const ade::iTensor*,ade::TensptrT> derivatives_;

private:
	std::shared_ptr<iRuleSet> rules_;
};

}

#endif // BWD_GRADER_HPP
