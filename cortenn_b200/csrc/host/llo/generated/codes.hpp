///
/// codes.hpp
///
/// Opcode / dtype enumerations of the `age` glue layer.  In the reference
/// these are emitted by Tenncor's age generator from llo/cfg/llo.json:13-124
/// (dtypes, opcodes); the generator is not vendored, so the table is written
/// out statically here in the same order.  Numeric values are not pinned by
/// any reference fixture (SURVEY.md 8c "Unpinned").
///
/// The numeric values are ALSO the ones the C-ABI uses (include/llo_cuda.h).
///

#ifndef CORTENN_AGE_CODES_HPP
#define CORTENN_AGE_CODES_HPP

#include <cstdint>
#include <string>

#include "logs/logs.hpp"

// X(NAME, ctype)
#define AGE_DTYPES(X)\
	X(DOUBLE, double)\
	X(FLOAT, float)\
	X(INT8, int8_t)\
	X(UINT8, uint8_t)\
	X(INT16, int16_t)\
	X(UINT16, uint16_t)\
	X(INT32, int32_t)\
	X(UINT32, uint32_t)\
	X(INT64, int64_t)\
	X(UINT64, uint64_t)

#define AGE_OPCODES(X)\
	X(ABS) X(NEG) X(SIN) X(COS) X(TAN) X(EXP) X(LOG) X(SQRT) X(ROUND)\
	X(POW) X(SUM) X(SUB) X(PROD) X(DIV) X(MIN) X(MAX)\
	X(EQ) X(NEQ) X(LT) X(GT)\
	X(RAND_BINO) X(RAND_UNIF) X(RAND_NORM)

namespace age
{

enum _GENERATED_OPCODE
{
	BAD_OP = 0,
#define AGE_X(NAME) NAME,
	AGE_OPCODES(AGE_X)
#undef AGE_X
	_N_GENERATED_OPCODES,
};

enum _GENERATED_DTYPE
{
	BAD_TYPE = 0,
#define AGE_X(NAME, CTYPE) NAME,
	AGE_DTYPES(AGE_X)
#undef AGE_X
	_N_GENERATED_DTYPES,
};

inline std::string name_op (_GENERATED_OPCODE code)
{
	switch (code)
	{
#define AGE_X(NAME) case NAME: return #NAME;
		AGE_OPCODES(AGE_X)
#undef AGE_X
		default: break;
	}
	return "BAD_OP";
}

inline _GENERATED_OPCODE get_op (const std::string& name)
{
#define AGE_X(NAME) if (name == #NAME) return NAME;
	AGE_OPCODES(AGE_X)
#undef AGE_X
	return BAD_OP;
}

inline std::string name_type (_GENERATED_DTYPE type)
{
	switch (type)
	{
#define AGE_X(NAME, CTYPE) case NAME: return #NAME;
		AGE_DTYPES(AGE_X)
#undef AGE_X
		default: break;
	}
	return "BAD_TYPE";
}

inline _GENERATED_DTYPE get_type (const std::string& name)
{
#define AGE_X(NAME, CTYPE) if (name == #NAME) return NAME;
	AGE_DTYPES(AGE_X)
#undef AGE_X
	return BAD_TYPE;
}

inline uint8_t type_size (_GENERATED_DTYPE type)
{
	switch (type)
	{
#define AGE_X(NAME, CTYPE) case NAME: return sizeof(CTYPE);
		AGE_DTYPES(AGE_X)
#undef AGE_X
		default: logs::fatal("cannot get size of bad type");
	}
	return 0;
}

template <typename T>
_GENERATED_DTYPE get_type (void)
{
	return BAD_TYPE;
}

#define AGE_X(NAME, CTYPE)\
template <> inline _GENERATED_DTYPE get_type<CTYPE> (void) { return NAME; }
AGE_DTYPES(AGE_X)
#undef AGE_X

// `long long` / `unsigned long long` are distinct from int64_t / uint64_t on
// LP64 but have the same representation (age::n_elems passes NElemT)
template <> inline _GENERATED_DTYPE get_type<long long> (void) { return INT64; }
template <> inline _GENERATED_DTYPE get_type<unsigned long long> (void) { return UINT64; }

}

#endif // CORTENN_AGE_CODES_HPP
