///
/// logs.hpp
///
/// Minimal stand-in for Tenncor's `logs` module (not vendored by the reference:
/// third_party/repos/tenncor.bzl:3-8). Contract relied on by cortenn:
/// fatal/fatalf throw std::runtime_error (llo/test/common.hpp:15-16 catches
/// exactly that), warnf only reports.
///

#ifndef CORTENN_LOGS_HPP
#define CORTENN_LOGS_HPP

#include <cstdarg>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace logs
{

inline std::string vformat (const char* fmt, va_list args)
{
	va_list probe;
	va_copy(probe, args);
	int need = std::vsnprintf(nullptr, 0, fmt, probe);
	va_end(probe);
	if (need < 0)
	{
		return fmt;
	}
	std::string out((size_t) need + 1, '\0');
	std::vsnprintf(&out[0], out.size(), fmt, args);
	out.resize((size_t) need);
	return out;
}

[[noreturn]] inline void fatal (const std::string& msg)
{
	throw std::runtime_error(msg);
}

[[noreturn]] inline void fatalf (const char* fmt, ...)
{
	va_list args;
	va_start(args, fmt);
	std::string msg = vformat(fmt, args);
	va_end(args);
	throw std::runtime_error(msg);
}

inline void warn (const std::string& msg)
{
	std::fprintf(stderr, "[cortenn warn] %s\n", msg.c_str());
}

inline void warnf (const char* fmt, ...)
{
	va_list args;
	va_start(args, fmt);
	std::string msg = vformat(fmt, args);
	va_end(args);
	warn(msg);
}

}

#endif // CORTENN_LOGS_HPP
