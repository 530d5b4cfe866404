///
/// pbm.hpp
/// pbm
///
/// Purpose:
/// Marshal and unmarshal ade graphs, the reference's pbm module
/// (pbm/save.hpp:24-93, pbm/src/save.cpp:13-90, pbm/load.hpp, pbm/src/load.cpp:33-108)
/// over the hand-written codec of graph.hpp: same interface names
/// (GraphSaver, load_graph, GraphInfo, PathedTens, DataSaverT / DataLoaderT),
/// same node order on the wire (leaves in visiting order, then functors by
/// ascending sub-graph height), same labels.
///

#include <functional>
#include <list>
#include <map>
#include <unordered_map>
#include <unordered_set>

#include "ade/ade.hpp"

#include "pbm/graph.hpp"

#ifndef PBM_PBM_HPP
#define PBM_PBM_HPP

namespace pbm
{

using TensT = std::vector<ade::TensptrT>;

/// bytes of a leaf -> Source.data
using DataSaverT = std::function<std::string(const char*,size_t,size_t)>;

/// Source.data, shape, type code, label -> leaf
using DataLoaderT = std::function<ade::TensptrT(const char*,ade::Shape,
	size_t,std::string)>;

using StringsT = std::list<std::string>;

using PathedMapT = std::unordered_map<ade::TensptrT,StringsT>;

/// Collects the nodes under the visited roots and writes them as a Graph
struct GraphSaver final : public ade::iTraveler
{
	GraphSaver (DataSaverT saver) : saver_(saver) {}

	void visit (ade::iLeaf* leaf) override
	{
		if (visited_.emplace(leaf).second)
		{
			leaf->accept(stat_);
			leaves_.push_back(leaf);
		}
	}

	void visit (ade::iFunctor* func) override
	{
		if (visited_.emplace(func).second)
		{
			func->accept(stat_);
			funcs_.push_back(func);
			for (const ade::MappedTensor& child : func->get_children())
			{
				child.get_tensor()->accept(*this);
			}
		}
	}

	void save (Graph& out, PathedMapT labels = PathedMapT())
	{
		std::unordered_map<ade::iTensor*,StringsT> by_node;
		for (auto& entry : labels)
		{
			by_node[entry.first.get()] = entry.second;
		}
		// children before parents, like the order the nodes were created in
		std::vector<ade::iFunctor*> funcs(funcs_.begin(), funcs_.end());
		std::stable_sort(funcs.begin(), funcs.end(),
			[&](ade::iFunctor* a, ade::iFunctor* b)
			{
				return stat_.graphsize_[a] < stat_.graphsize_[b];
			});
		std::unordered_map<ade::iTensor*,size_t> position;
		auto add_node = [&](ade::iTensor* tens) -> Node&
		{
			position[tens] = out.nodes_.size();
			out.nodes_.emplace_back();
			Node& node = out.nodes_.back();
			auto it = by_node.find(tens);
			if (by_node.end() != it)
			{
				node.labels_.assign(it->second.begin(), it->second.end());
			}
			return node;
		};
		for (ade::iLeaf* leaf : leaves_)
		{
			Node& node = add_node(leaf);
			node.has_source_ = true;
			save_data(node.source_, leaf);
		}
		for (ade::iFunctor* func : funcs)
		{
			Node& node = add_node(func);
			node.has_functor_ = true;
			ade::Opcode opcode = func->get_opcode();
			node.functor_.opname_ = opcode.name_;
			node.functor_.opcode_ = (uint32_t) opcode.code_;
			for (const ade::MappedTensor& child : func->get_children())
			{
				NodeArg arg;
				arg.idx_ = (uint32_t) position[child.get_tensor().get()];
				save_coord(arg.coord_, child.get_coorder());
				if (child.get_shaper() != child.get_coorder())
				{
					save_coord(arg.shaper_, child.get_shaper());
				}
				arg.fwd_ = child.map_io();
				node.functor_.args_.push_back(arg);
			}
		}
	}

	std::list<ade::iLeaf*> leaves_;

	std::list<ade::iFunctor*> funcs_;

	std::unordered_set<ade::iTensor*> visited_;

	ade::GraphStat stat_;

private:
	static void save_coord (std::vector<double>& out, const ade::CoordptrT& mapper)
	{
		mapper->access([&](const ade::MatrixT& mat)
		{
			for (ade::RankT i = 0; i < ade::mat_dim; ++i)
			{
				out.insert(out.end(), mat[i], mat[i] + ade::mat_dim);
			}
		});
	}

	void save_data (Source& out, ade::iLeaf* leaf)
	{
		const ade::Shape& shape = leaf->shape();
		bool narrow = std::all_of(shape.begin(), shape.end(),
			[](ade::DimT d) { return d <= 255; });
		if (narrow)
		{
			// one byte per dimension, as the reference writes it
			out.shape_.assign(shape.begin(), shape.end());
		}
		else
		{
			out.wide_shape_.assign(shape.begin(), shape.end());
		}
		size_t tcode = leaf->type_code();
		out.data_ = saver_((const char*) ((const ade::iLeaf*) leaf)->data(),
			shape.n_elems(), tcode);
		out.typecode_ = (uint32_t) tcode;
	}

	DataSaverT saver_;
};

/// Tree of labels -> tensors (pbm/load.hpp:19-160)
struct PathedTens final
{
	~PathedTens (void)
	{
		for (auto& child : children_)
		{
			delete child.second;
		}
	}

	PathedTens (void) = default;

	PathedTens (const PathedTens&) = delete;

	PathedTens& operator = (const PathedTens&) = delete;

	ade::TensptrT get_labelled (StringsT path) const
	{
		return get_labelled(path.begin(), path.end());
	}

	void set_labelled (StringsT path, ade::TensptrT tens)
	{
		set_labelled(path.begin(), path.end(), tens);
	}

	template <typename IT>
	ade::TensptrT get_labelled (IT begin, IT end) const
	{
		if (begin == end)
		{
			return nullptr;
		}
		IT next = begin;
		++next;
		if (next == end)
		{
			auto it = tens_.find(*begin);
			return tens_.end() == it ? nullptr : it->second;
		}
		auto it = children_.find(*begin);
		return children_.end() == it ? nullptr : it->second->get_labelled(next, end);
	}

	template <typename IT>
	void set_labelled (IT begin, IT end, ade::TensptrT tens)
	{
		if (begin == end)
		{
			return;
		}
		IT next = begin;
		++next;
		if (next == end)
		{
			tens_[*begin] = tens;
			return;
		}
		PathedTens*& child = children_[*begin];
		if (nullptr == child)
		{
			child = new PathedTens();
		}
		child->set_labelled(next, end, tens);
	}

	std::map<std::string,PathedTens*> children_;

	std::map<std::string,ade::TensptrT> tens_;
};

struct GraphInfo final
{
	/// nodes no other node of the graph uses
	std::unordered_set<ade::TensptrT> roots_;

	PathedTens tens_;

	/// every node in file order (what NodeArg.idx refers to)
	TensT nodes_;
};

inline ade::CoordptrT load_coord (const std::vector<double>& coord)
{
	if ((size_t) ade::mat_dim * ade::mat_dim != coord.size())
	{
		logs::fatal("cannot deserialize non-matrix coordinate map");
	}
	return std::make_shared<ade::CoordMap>([&](ade::MatrixT fwd)
	{
		for (ade::RankT i = 0; i < ade::mat_dim; ++i)
		{
			for (ade::RankT j = 0; j < ade::mat_dim; ++j)
			{
				fwd[i][j] = coord[i * ade::mat_dim + j];
			}
		}
	});
}

inline ade::Shape load_shape (const Source& source)
{
	std::vector<ade::DimT> dims;
	if (false == source.wide_shape_.empty())
	{
		dims.assign(source.wide_shape_.begin(), source.wide_shape_.end());
	}
	else
	{
		for (char c : source.shape_)
		{
			dims.push_back((ade::DimT) (uint8_t) c);
		}
	}
	return ade::Shape(dims);
}

inline void load_graph (GraphInfo& out, const Graph& in, DataLoaderT dataloader)
{
	for (const Node& node : in.nodes_)
	{
		ade::TensptrT tens;
		if (node.has_source_)
		{
			std::string label = node.labels_.empty() ? "" : node.labels_.back();
			tens = dataloader(node.source_.data_.c_str(), load_shape(node.source_),
				node.source_.typecode_, label);
		}
		else
		{
			ade::ArgsT args;
			for (const NodeArg& arg : node.functor_.args_)
			{
				if (arg.idx_ >= out.nodes_.size())
				{
					logs::fatalf("cannot deserialize argument %d before its node",
						(int) arg.idx_);
				}
				ade::CoordptrT coord = load_coord(arg.coord_);
				ade::CoordptrT shaper = arg.shaper_.empty() ? coord :
					load_coord(arg.shaper_);
				args.push_back(ade::MappedTensor(out.nodes_[arg.idx_], shaper,
					arg.fwd_, coord));
				out.roots_.erase(out.nodes_[arg.idx_]);
			}
			tens = ade::TensptrT(ade::Functor::get(ade::Opcode{
				node.functor_.opname_, node.functor_.opcode_}, args));
		}
		out.nodes_.push_back(tens);
		if (false == node.labels_.empty())
		{
			out.tens_.set_labelled(node.labels_.begin(), node.labels_.end(), tens);
		}
		out.roots_.emplace(tens);
	}
}

}

#endif // PBM_PBM_HPP
