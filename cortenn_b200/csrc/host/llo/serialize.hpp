///
/// serialize.hpp
/// llo
///
/// Purpose:
/// Marshal and unmarshal the data of llo leaves for pbm (reference:
/// llo/serialize.hpp:15-25, llo/src/serialize.cpp:22-63): element bytes in
/// little-endian order.  Variable::data() refreshes the host bytes from the
/// device mirror when the newest values only live in HBM.
///

#include "llo/data.hpp"

#ifndef LLO_SERIALIZE_HPP
#define LLO_SERIALIZE_HPP

namespace llo
{

/// Marshal `nelems` elements of type `typecode` to Source.data
std::string serialize (const char* in, size_t nelems, size_t typecode);

/// Unmarshal Source.data as a Variable
ade::TensptrT deserialize (const char* pb, ade::Shape shape,
	size_t typecode, std::string label);

}

#endif // LLO_SERIALIZE_HPP
