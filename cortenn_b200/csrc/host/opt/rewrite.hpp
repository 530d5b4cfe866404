///
/// rewrite.hpp
/// opt
///
/// Purpose:
/// Graph -> graph rewriting of ade graphs: the optimiser the reference lists
/// as an open item ("implement optimization options to decrease equation
/// graph size in opt", /root/reference/todo:12) next to its pruner
/// (opt/shear.hpp:91-176).  Like opt::TargetPruner the rewriter is opcode
/// agnostic: rules are callables in the shape of opt::PruneFuncT and decide
/// what a functor becomes once its children have been rewritten.
///
/// The source graph is never touched; rewritten graphs share every node that
/// did not change.
///

#include <cstring>
#include <functional>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "ade/ade.hpp"

#ifndef OPT_REWRITE_HPP
#define OPT_REWRITE_HPP

namespace opt
{

/// What `func` becomes given its already rewritten children, or null to leave
/// the decision to the next rule (no rule: the functor is rebuilt over the new
/// children, or kept when none of them changed)
using RewriteRuleT = std::function<ade::TensptrT(ade::iFunctor*,const ade::ArgsT&)>;

/// True when two coordinate maps hold the same matrix
inline bool same_map (const ade::CoordptrT& a, const ade::CoordptrT& b)
{
	if (a.get() == b.get())
	{
		return true;
	}
	bool same = true;
	a->access([&](const ade::MatrixT& ma)
	{
		b->access([&](const ade::MatrixT& mb)
		{
			same = 0 == std::memcmp(ma, mb, sizeof(ade::MatrixT));
		});
	});
	return same;
}

/// Number of distinct functors reachable from the roots
inline size_t count_functors (const ade::TensT& roots)
{
	std::unordered_set<ade::iTensor*> seen;
	std::vector<ade::iTensor*> todo;
	for (const ade::TensptrT& root : roots)
	{
		todo.push_back(root.get());
	}
	size_t n = 0;
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
			++n;
			for (const ade::MappedTensor& child : f->get_children())
			{
				todo.push_back(child.get_tensor().get());
			}
		}
	}
	return n;
}

/// Bottom-up rewriter with optional common sub-expression sharing
struct Rewriter final
{
	/// `volatile_ops`: opcodes whose every evaluation is a fresh value (random
	/// draws); such functors and everything above them are never shared
	Rewriter (std::vector<RewriteRuleT> rules, bool share,
		std::unordered_set<size_t> volatile_ops = {}) :
		rules_(rules), share_(share), volatile_ops_(volatile_ops) {}

	ade::TensptrT rewrite (ade::TensptrT root)
	{
		auto it = done_.find(root.get());
		if (done_.end() != it)
		{
			return it->second;
		}
		ade::TensptrT out = root;
		if (auto func = dynamic_cast<ade::iFunctor*>(root.get()))
		{
			const ade::ArgsT& children = func->get_children();
			ade::ArgsT args;
			bool changed = false;
			for (const ade::MappedTensor& child : children)
			{
				ade::TensptrT tens = rewrite(child.get_tensor());
				changed = changed || tens.get() != child.get_tensor().get();
				args.push_back(ade::MappedTensor(tens, child.get_shaper(),
					child.map_io(), child.get_coorder()));
			}
			out = apply(root, func, args, changed);
		}
		done_.emplace(root.get(), out);
		keep_.push_back(root); // keys stay alive, so addresses stay unique
		return out;
	}

	ade::TensT rewrite (const ade::TensT& roots)
	{
		ade::TensT out;
		for (const ade::TensptrT& root : roots)
		{
			out.push_back(rewrite(root));
		}
		return out;
	}

	/// Number of functors merged into an earlier, structurally equal one
	size_t shared_ = 0;

private:
	ade::TensptrT apply (ade::TensptrT original, ade::iFunctor* func,
		const ade::ArgsT& args, bool changed)
	{
		for (const RewriteRuleT& rule : rules_)
		{
			ade::TensptrT repl = rule(func, args);
			if (nullptr != repl)
			{
				if (repl.get() == original.get())
				{
					break;
				}
				// the replacement is built from rewritten parts; give the
				// rules one more look at it unless it is one of those parts
				for (const ade::MappedTensor& arg : args)
				{
					if (arg.get_tensor().get() == repl.get())
					{
						return repl;
					}
				}
				if (auto rf = dynamic_cast<ade::iFunctor*>(repl.get()))
				{
					if (++depth_ < 16)
					{
						ade::TensptrT again = apply(repl, rf, rf->get_children(), false);
						--depth_;
						return again;
					}
					--depth_;
				}
				return repl;
			}
		}
		ade::TensptrT out = changed ?
			ade::TensptrT(ade::Functor::get(func->get_opcode(), args)) : original;
		return share(out);
	}

	/// Structural key of a functor over canonical children
	ade::TensptrT share (ade::TensptrT tens)
	{
		auto func = dynamic_cast<ade::iFunctor*>(tens.get());
		if (false == share_ || nullptr == func)
		{
			return tens;
		}
		size_t code = func->get_opcode().code_;
		bool fresh = volatile_ops_.end() != volatile_ops_.find(code);
		std::string key;
		key.append((const char*) &code, sizeof(code));
		for (const ade::MappedTensor& child : func->get_children())
		{
			ade::iTensor* ptr = child.get_tensor().get();
			fresh = fresh || volatile_.end() != volatile_.find(ptr);
			bool push = child.map_io();
			key.append((const char*) &ptr, sizeof(ptr));
			key.append((const char*) &push, sizeof(push));
			for (const ade::CoordptrT& map : {child.get_shaper(), child.get_coorder()})
			{
				map->access([&](const ade::MatrixT& m)
				{
					key.append((const char*) m, sizeof(ade::MatrixT));
				});
			}
		}
		if (fresh)
		{
			volatile_.emplace(tens.get());
			keep_.push_back(tens);
			return tens;
		}
		auto it = canon_.find(key);
		if (canon_.end() != it)
		{
			if (it->second.get() != tens.get())
			{
				++shared_;
			}
			return it->second;
		}
		canon_.emplace(std::move(key), tens);
		return tens;
	}

	std::vector<RewriteRuleT> rules_;

	bool share_;

	std::unordered_set<size_t> volatile_ops_;

	std::unordered_map<ade::iTensor*,ade::TensptrT> done_;

	std::unordered_map<std::string,ade::TensptrT> canon_;

	std::unordered_set<ade::iTensor*> volatile_;

	std::vector<ade::TensptrT> keep_;

	int depth_ = 0;
};

}

#endif // OPT_REWRITE_HPP
