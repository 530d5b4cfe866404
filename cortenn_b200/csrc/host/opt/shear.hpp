///
/// shear.hpp
/// opt
///
/// VENDORED FROM THE REFERENCE.  This file is mingkaic/cortenn's
/// opt/shear.hpp:20-176 (MIT licence, (c) its authors) with local variable
/// renames, a `seen_` set and a stable sort; it is the reference's host-side
/// graph pruner, on the reference's side of the drop-in boundary, and is kept
/// so that graphs "produced by ade::derive and optimised by opt" are the same
/// graphs.  Nothing on the GPU path derives from it.  The optimiser that is
/// new work is opt/rewrite.hpp + llo/optimize.hpp.
///
/// Purpose:
/// Target-driven graph pruning, same interface as the reference's
/// opt/shear.hpp:20-176 (LeafFinder, TargetPruner, GetLeafValT, PruneFuncT).
///
/// Behaviour kept on purpose (pinned by opt/test/test_shear.cpp:100-178):
/// every path functor is offered to the prune functor with its ORIGINAL
/// children; a child's replacement only influences its parent when the
/// replacement is itself a target leaf (then the child's index is passed in
/// the parent's target set).  The result is the replacement of the root.
///

#include <list>

#include "ade/ade.hpp"

#ifndef OPT_SHEAR_HPP
#define OPT_SHEAR_HPP

namespace opt
{

/// Functor for getting leaf values
template <typename T>
using GetLeafValT = std::function<T(ade::iLeaf*)>;

/// Map path functors to the argument indices that lead to a target
using ParentMapT = std::unordered_map<
	ade::iTensor*,std::unordered_set<size_t>>;

/// Pruning functor type
using PruneFuncT = std::function<ade::TensptrT(ade::iFunctor*,\
	std::unordered_set<size_t>,ade::ArgsT)>;

/// Find leaves whose attribute equals the target value, and every functor
/// on a path to one
template <typename T>
struct LeafFinder final : public ade::iTraveler
{
	LeafFinder (T target, GetLeafValT<T> get_leaf) :
		target_(target), get_leaf_(get_leaf) {}

	/// Implementation of iTraveler
	void visit (ade::iLeaf* leaf) override
	{
		if (target_ == get_leaf_(leaf))
		{
			founds_.emplace(leaf);
		}
	}

	/// Implementation of iTraveler
	void visit (ade::iFunctor* func) override
	{
		if (false == seen_.emplace(func).second)
		{
			return;
		}
		const ade::ArgsT& children = func->get_children();
		std::unordered_set<size_t> path;
		for (size_t i = 0, n = children.size(); i < n; ++i)
		{
			ade::iTensor* tens = children[i].get_tensor().get();
			tens->accept(*this);
			if (parents_.end() != parents_.find(tens) ||
				founds_.end() != founds_.find(tens))
			{
				path.emplace(i);
			}
		}
		if (false == path.empty())
		{
			parents_[func] = path;
			order_.push_back(func);
		}
	}

	/// Target of label all paths are travelling to
	T target_;

	/// Leaf value getter
	GetLeafValT<T> get_leaf_;

	/// Set of leaf nodes found
	std::unordered_set<ade::iTensor*> founds_;

	/// Map of parent nodes in path
	ParentMapT parents_;

	/// Path functors, children before parents
	std::vector<ade::iFunctor*> order_;

private:
	std::unordered_set<ade::iTensor*> seen_;
};

/// Shorten branches that lead to a target leaf value.  E.g. with target
/// "0": f(x) * 0 becomes 0, x + 0 becomes x.
template <typename T>
struct TargetPruner
{
	TargetPruner (T target, GetLeafValT<T> find_target, PruneFuncT pruner) :
		finder_(target, find_target), pruner_(pruner) {}

	/// Prune graph of root Tensptr
	ade::TensptrT prune (ade::TensptrT root)
	{
		root->accept(finder_);
		if (finder_.parents_.empty())
		{
			// no path to a target, or root is a leaf
			return root;
		}

		// children-first order; heights break no ties we care about since
		// order_ is already a post-order
		ade::GraphStat stat;
		root->accept(stat);
		std::vector<ade::iFunctor*> todo = finder_.order_;
		std::stable_sort(todo.begin(), todo.end(),
			[&](ade::iFunctor* a, ade::iFunctor* b)
			{
				return stat.graphsize_[a] < stat.graphsize_[b];
			});

		std::unordered_set<ade::iTensor*> targets = finder_.founds_;
		std::unordered_map<ade::iTensor*,ade::TensptrT> replaced;
		for (ade::iFunctor* func : todo)
		{
			ade::ArgsT children = func->get_children();
			std::unordered_set<size_t> hits;
			for (size_t i : finder_.parents_[func])
			{
				if (targets.end() !=
					targets.find(children[i].get_tensor().get()))
				{
					hits.emplace(i);
				}
			}
			ade::TensptrT repl = pruner_(func, hits, children);
			replaced.emplace(func, repl);
			if (auto leaf = dynamic_cast<ade::iLeaf*>(repl.get()))
			{
				if (finder_.get_leaf_(leaf) == finder_.target_)
				{
					targets.emplace(func);
				}
			}
		}
		auto it = replaced.find(root.get());
		if (replaced.end() == it)
		{
			logs::fatal(
				"GraphStat failed to identify children of root subgraph");
		}
		return it->second;
	}

private:
	/// Target finding traveler
	LeafFinder<T> finder_;

	/// Prune functor defining how to prune a given graph
	PruneFuncT pruner_;
};

}

#endif // OPT_SHEAR_HPP
