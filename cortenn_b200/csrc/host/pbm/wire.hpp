///
/// wire.hpp
/// pbm
///
/// Purpose:
/// The part of the protocol-buffers wire format pbm/graph.proto needs, written
/// by hand: there is no protoc in this build and the five
/// messages only use varints, length-delimited fields and packed doubles.
/// Encoding follows what the C++ protobuf runtime emits for proto3 (fields in
/// number order, defaults omitted, packed repeated scalars), so files written
/// here are byte-identical to the reference's for the same graph.
///

#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "logs/logs.hpp"

#ifndef PBM_WIRE_HPP
#define PBM_WIRE_HPP

namespace pbm
{

namespace wire
{

enum Type { VARINT = 0, FIXED64 = 1, LEN = 2, FIXED32 = 5 };

struct Writer final
{
	void varint (uint64_t v)
	{
		while (v >= 0x80)
		{
			out_.push_back((char) (v | 0x80));
			v >>= 7;
		}
		out_.push_back((char) v);
	}

	void tag (uint32_t field, Type type)
	{
		varint(((uint64_t) field << 3) | (uint64_t) type);
	}

	/// proto3 scalar: omitted when zero
	void uint_field (uint32_t field, uint64_t v)
	{
		if (0 != v)
		{
			tag(field, VARINT);
			varint(v);
		}
	}

	/// string / bytes / embedded message; `always`: emit even when empty
	/// (oneof members and repeated elements are present by being there)
	void bytes_field (uint32_t field, const std::string& v, bool always = false)
	{
		if (always || false == v.empty())
		{
			tag(field, LEN);
			varint(v.size());
			out_ += v;
		}
	}

	void packed_doubles (uint32_t field, const std::vector<double>& v)
	{
		if (v.empty())
		{
			return;
		}
		tag(field, LEN);
		varint(v.size() * sizeof(double));
		size_t at = out_.size();
		out_.resize(at + v.size() * sizeof(double));
		std::memcpy(&out_[at], v.data(), v.size() * sizeof(double)); // little endian hosts
	}

	void packed_varints (uint32_t field, const std::vector<uint64_t>& v)
	{
		if (v.empty())
		{
			return;
		}
		Writer body;
		for (uint64_t e : v)
		{
			body.varint(e);
		}
		bytes_field(field, body.out_, true);
	}

	std::string out_;
};

struct Reader final
{
	Reader (const char* data, size_t size) : at_(data), end_(data + size) {}

	Reader (const std::string& s) : Reader(s.data(), s.size()) {}

	bool done (void) const { return at_ >= end_; }

	uint64_t varint (void)
	{
		uint64_t v = 0;
		for (int shift = 0; shift < 64; shift += 7)
		{
			if (at_ >= end_)
			{
				logs::fatal("cannot parse truncated varint");
			}
			uint8_t b = (uint8_t) *at_++;
			v |= (uint64_t) (b & 0x7f) << shift;
			if (0 == (b & 0x80))
			{
				return v;
			}
		}
		logs::fatal("cannot parse overlong varint");
		return 0;
	}

	/// next field: number and wire type
	void tag (uint32_t& field, Type& type)
	{
		uint64_t t = varint();
		field = (uint32_t) (t >> 3);
		type = (Type) (t & 7);
	}

	std::string bytes (void)
	{
		uint64_t n = varint();
		if ((uint64_t) (end_ - at_) < n)
		{
			logs::fatal("cannot parse truncated length-delimited field");
		}
		std::string out(at_, at_ + n);
		at_ += n;
		return out;
	}

	double fixed64 (void)
	{
		if (end_ - at_ < 8)
		{
			logs::fatal("cannot parse truncated fixed64");
		}
		double v;
		std::memcpy(&v, at_, 8);
		at_ += 8;
		return v;
	}

	/// skip a field this reader has no use for (forward compatibility)
	void skip (Type type)
	{
		switch (type)
		{
			case VARINT: varint(); break;
			case FIXED64: fixed64(); break;
			case LEN: bytes(); break;
			case FIXED32:
				if (end_ - at_ < 4)
				{
					logs::fatal("cannot parse truncated fixed32");
				}
				at_ += 4;
				break;
			default: logs::fatalf("cannot parse wire type %d", (int) type);
		}
	}

	/// repeated double, packed or not
	void doubles (std::vector<double>& out, Type type)
	{
		if (FIXED64 == type)
		{
			out.push_back(fixed64());
			return;
		}
		std::string body = bytes();
		if (0 != body.size() % 8)
		{
			logs::fatal("cannot parse packed doubles of odd size");
		}
		size_t n = body.size() / 8;
		size_t at = out.size();
		out.resize(at + n);
		std::memcpy(out.data() + at, body.data(), body.size());
	}

	void varints (std::vector<uint64_t>& out, Type type)
	{
		if (VARINT == type)
		{
			out.push_back(varint());
			return;
		}
		std::string packed = bytes();
		Reader body(packed);
		while (false == body.done())
		{
			out.push_back(body.varint());
		}
	}

	const char* at_;

	const char* end_;
};

}

}

#endif // PBM_WIRE_HPP
