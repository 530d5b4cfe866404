///
/// ade.hpp
/// dbg
///
/// PrettyEquation: ascii tree rendering of an ade graph, the format the
/// reference's golden trees use (pbm/data/graph.txt, opt/test/test_shear.cpp:160-176;
/// Tenncor `dbg` is not vendored).
///

#ifndef DBG_ADE_HPP
#define DBG_ADE_HPP

#include <iostream>

#include "ade/ade.hpp"

struct PrettyEquation final
{
	void print (std::ostream& out, const ade::TensptrT& root)
	{
		render(out, root.get(), "", true);
	}

	void print (std::ostream& out, ade::iTensor* root)
	{
		render(out, root, "", true);
	}

	/// Print shapes next to node names
	bool showshape_ = false;

	/// Display names for specific nodes
	std::unordered_map<ade::iTensor*,std::string> labels_;

private:
	void render (std::ostream& out, ade::iTensor* node,
		std::string prefix, bool is_root)
	{
		out << "(";
		auto it = labels_.find(node);
		auto func = dynamic_cast<ade::iFunctor*>(node);
		if (labels_.end() != it)
		{
			out << it->second << "=";
		}
		if (nullptr != func)
		{
			out << func->to_string();
			if (showshape_)
			{
				out << func->shape().to_string();
			}
		}
		else
		{
			// leaves print as their shape (pbm/data/graph.txt:2)
			out << node->shape().to_string();
		}
		out << ")\n";
		if (nullptr == func)
		{
			return;
		}
		const ade::ArgsT& children = func->get_children();
		for (size_t i = 0, n = children.size(); i < n; ++i)
		{
			out << prefix << " `--";
			std::string next = prefix +
				(i + 1 < n ? " |  " : "    ");
			render(out, children[i].get_tensor().get(), next, false);
		}
	}
};

#endif // DBG_ADE_HPP
