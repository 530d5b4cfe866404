#include "llo/serialize.hpp"

#ifdef LLO_SERIALIZE_HPP

namespace llo
{

static bool big_endian (void)
{
	const uint16_t probe = 1;
	return 0 == *(const char*) &probe;
}

/// element bytes reversed in place (the wire is little endian)
static std::string swapped (const char* in, size_t nelems, size_t esize)
{
	std::string out(nelems * esize, '\0');
	for (size_t e = 0; e < nelems; ++e)
	{
		for (size_t b = 0; b < esize; ++b)
		{
			out[e * esize + b] = in[e * esize + esize - 1 - b];
		}
	}
	return out;
}

std::string serialize (const char* in, size_t nelems, size_t typecode)
{
	size_t esize = age::type_size((age::_GENERATED_DTYPE) typecode);
	if (big_endian() && esize > 1)
	{
		return swapped(in, nelems, esize);
	}
	return std::string(in, nelems * esize);
}

ade::TensptrT deserialize (const char* pb, ade::Shape shape,
	size_t typecode, std::string label)
{
	age::_GENERATED_DTYPE dtype = (age::_GENERATED_DTYPE) typecode;
	size_t esize = age::type_size(dtype);
	if (big_endian() && esize > 1)
	{
		std::string host = swapped(pb, shape.n_elems(), esize);
		return ade::TensptrT(new Variable(host.c_str(), dtype, shape, label));
	}
	return ade::TensptrT(new Variable(pb, dtype, shape, label));
}

}

#endif
