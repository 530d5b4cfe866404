#include <cstdlib>
#include <map>

#include "logs/logs.hpp"

#include "llo/session.hpp"

#ifdef LLO_SESSION_HPP

namespace llo
{

static long env_default (const char* name, long fallback)
{
	if (const char* env = std::getenv(name))
	{
		return std::atol(env);
	}
	return fallback;
}

static std::map<std::string,long>& options (void)
{
	static std::map<std::string,long> opts = {
		{"plan", env_default("LLO_PLAN", 1)},
		{"simplify", env_default("LLO_SIMPLIFY", 0)},
		{"recompute", env_default("LLO_RECOMPUTE", 0)},
		{"tf32x3", env_default("LLO_TF32X3", 0)},
		{"sequential_reduce", env_default("LLO_SEQUENTIAL_REDUCE", 0)},
	};
	return opts;
}

void set_option (std::string name, long value)
{
	auto& opts = options();
	auto it = opts.find(name);
	if (opts.end() == it)
	{
		logs::fatalf("unknown evaluator option %s", name.c_str());
	}
	it->second = value;
}

long get_option (std::string name)
{
	auto& opts = options();
	auto it = opts.find(name);
	if (opts.end() == it)
	{
		logs::fatalf("unknown evaluator option %s", name.c_str());
	}
	return it->second;
}

}

#endif
