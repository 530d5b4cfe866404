///
/// graph.hpp
/// pbm
///
/// Purpose:
/// The messages of the reference's pbm/graph.proto:7-44 as plain structs with
/// a hand-written codec (wire.hpp).  Field numbers, types and defaults are
/// the reference's, so its files load here and files written here for graphs
/// it can represent load there.
///
/// One extension: ade::DimT is wider than the
/// reference's uint8_t here, and Source.shape holds ONE BYTE per dimension
/// (pbm/save.hpp:82, pbm/src/load.cpp:50).  A shape with a dimension above
/// 255 is written to the new field `wide_shape` (= 4, packed uint64) instead
/// and `shape` is left empty; shapes that fit are written exactly as before.
/// A reference reader skips the unknown field (proto3) and sees an empty
/// shape for tensors it could not have represented anyway.
///

#include "pbm/wire.hpp"

#ifndef PBM_GRAPH_HPP
#define PBM_GRAPH_HPP

namespace pbm
{

/// message Source { bytes shape = 1; bytes data = 2; uint32 typecode = 3;
///                  repeated uint64 wide_shape = 4; }
struct Source final
{
	std::string shape_;                // one byte per dimension
	std::string data_;
	uint32_t typecode_ = 0;
	std::vector<uint64_t> wide_shape_; // extension, see above
};

/// message NodeArg { uint32 idx = 1; repeated double coord = 2;
///                   repeated double shaper = 3; bool fwd = 4; }
struct NodeArg final
{
	uint32_t idx_ = 0;
	std::vector<double> coord_;
	std::vector<double> shaper_;
	bool fwd_ = false;
};

/// message Functor { string opname = 1; uint32 opcode = 2; repeated NodeArg args = 3; }
struct FunctorMsg final
{
	std::string opname_;
	uint32_t opcode_ = 0;
	std::vector<NodeArg> args_;
};

/// message Node { repeated string labels = 1; oneof detail { Source source = 2;
///                Functor functor = 3; } }
struct Node final
{
	std::vector<std::string> labels_;
	bool has_source_ = false;
	bool has_functor_ = false;
	Source source_;
	FunctorMsg functor_;
};

/// message Graph { string label = 1; repeated Node nodes = 2; }
struct Graph final
{
	std::string label_;
	std::vector<Node> nodes_;
};

inline std::string encode (const Source& msg)
{
	wire::Writer w;
	w.bytes_field(1, msg.shape_);
	w.bytes_field(2, msg.data_);
	w.uint_field(3, msg.typecode_);
	w.packed_varints(4, msg.wide_shape_);
	return w.out_;
}

inline std::string encode (const NodeArg& msg)
{
	wire::Writer w;
	w.uint_field(1, msg.idx_);
	w.packed_doubles(2, msg.coord_);
	w.packed_doubles(3, msg.shaper_);
	w.uint_field(4, msg.fwd_ ? 1 : 0);
	return w.out_;
}

inline std::string encode (const FunctorMsg& msg)
{
	wire::Writer w;
	w.bytes_field(1, msg.opname_);
	w.uint_field(2, msg.opcode_);
	for (const NodeArg& arg : msg.args_)
	{
		w.bytes_field(3, encode(arg), true);
	}
	return w.out_;
}

inline std::string encode (const Node& msg)
{
	wire::Writer w;
	for (const std::string& label : msg.labels_)
	{
		w.bytes_field(1, label, true);
	}
	if (msg.has_source_)
	{
		w.bytes_field(2, encode(msg.source_), true);
	}
	if (msg.has_functor_)
	{
		w.bytes_field(3, encode(msg.functor_), true);
	}
	return w.out_;
}

inline std::string encode (const Graph& msg)
{
	wire::Writer w;
	w.bytes_field(1, msg.label_);
	for (const Node& node : msg.nodes_)
	{
		w.bytes_field(2, encode(node), true);
	}
	return w.out_;
}

inline void decode (Source& out, const std::string& bytes)
{
	wire::Reader r(bytes);
	while (false == r.done())
	{
		uint32_t field;
		wire::Type type;
		r.tag(field, type);
		switch (field)
		{
			case 1: out.shape_ = r.bytes(); break;
			case 2: out.data_ = r.bytes(); break;
			case 3: out.typecode_ = (uint32_t) r.varint(); break;
			case 4: r.varints(out.wide_shape_, type); break;
			default: r.skip(type);
		}
	}
}

inline void decode (NodeArg& out, const std::string& bytes)
{
	wire::Reader r(bytes);
	while (false == r.done())
	{
		uint32_t field;
		wire::Type type;
		r.tag(field, type);
		switch (field)
		{
			case 1: out.idx_ = (uint32_t) r.varint(); break;
			case 2: r.doubles(out.coord_, type); break;
			case 3: r.doubles(out.shaper_, type); break;
			case 4: out.fwd_ = 0 != r.varint(); break;
			default: r.skip(type);
		}
	}
}

inline void decode (FunctorMsg& out, const std::string& bytes)
{
	wire::Reader r(bytes);
	while (false == r.done())
	{
		uint32_t field;
		wire::Type type;
		r.tag(field, type);
		switch (field)
		{
			case 1: out.opname_ = r.bytes(); break;
			case 2: out.opcode_ = (uint32_t) r.varint(); break;
			case 3:
				out.args_.emplace_back();
				decode(out.args_.back(), r.bytes());
				break;
			default: r.skip(type);
		}
	}
}

inline void decode (Node& out, const std::string& bytes)
{
	wire::Reader r(bytes);
	while (false == r.done())
	{
		uint32_t field;
		wire::Type type;
		r.tag(field, type);
		switch (field)
		{
			case 1: out.labels_.push_back(r.bytes()); break;
			case 2:
				out.has_source_ = true;
				out.has_functor_ = false;
				decode(out.source_, r.bytes());
				break;
			case 3:
				out.has_functor_ = true;
				out.has_source_ = false;
				decode(out.functor_, r.bytes());
				break;
			default: r.skip(type);
		}
	}
}

inline void decode (Graph& out, const std::string& bytes)
{
	wire::Reader r(bytes);
	while (false == r.done())
	{
		uint32_t field;
		wire::Type type;
		r.tag(field, type);
		switch (field)
		{
			case 1: out.label_ = r.bytes(); break;
			case 2:
				out.nodes_.emplace_back();
				decode(out.nodes_.back(), r.bytes());
				break;
			default: r.skip(type);
		}
	}
}

}

#endif // PBM_GRAPH_HPP
