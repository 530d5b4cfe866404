///
/// fmts.hpp
///
/// Minimal stand-in for Tenncor's `fmts` module: stream-based to_string
/// (llo/data.hpp:214 uses it to label scalars: 0 -> "0", 1 -> "1") and
/// range printing used by test helpers (llo/test/common.hpp:3-7).
///

#ifndef CORTENN_FMTS_HPP
#define CORTENN_FMTS_HPP

#include <sstream>
#include <string>
#include <type_traits>

namespace fmts
{

template <typename T>
inline void put (std::ostream& os, const T& v)
{
	// 8-bit integers must print as numbers, not characters
	if constexpr (std::is_same<T,int8_t>::value ||
		std::is_same<T,uint8_t>::value)
	{
		os << (int) v;
	}
	else
	{
		os << v;
	}
}

template <typename T>
inline std::string to_string (const T& v)
{
	std::stringstream ss;
	put(ss, v);
	return ss.str();
}

template <typename IT>
inline void to_stream (std::ostream& os, IT begin, IT end,
	const char* delim = "\\")
{
	os << "[";
	bool first = true;
	for (IT it = begin; it != end; ++it)
	{
		if (!first)
		{
			os << delim;
		}
		put(os, *it);
		first = false;
	}
	os << "]";
}

template <typename IT>
inline std::string to_string (IT begin, IT end)
{
	std::stringstream ss;
	to_stream(ss, begin, end);
	return ss.str();
}

}

#endif // CORTENN_FMTS_HPP
