///
/// session.hpp
/// llo
///
/// Purpose:
/// Process-wide evaluator options.  The reference has no runtime flags
/// (SURVEY.md 5); these only select between mechanisms that give the same
/// results within the stated tolerances.
///
///   "plan"        1 = lower graphs through the planner (views, constant
///                 folding, fused elementwise programs, GEMM lowering);
///                 0 = one C-ABI operator call per functor
///   "simplify"    1 = allow x*1 -> x and (x*y)/x -> y in derived graphs
///                 (the second is inexact by <= 1 ulp and defined at x == 0,
///                 where the reference yields NaN)
///   "recompute"   n > 0 = an elementwise producer with several consumers is
///                 recomputed inside each of them when its fully inlined
///                 program has at most n instructions (and 2 inputs) instead
///                 of being stored and read back; 0 = every shared producer
///                 is materialised once
///   "tf32x3"      1 = opt-in 3xTF32 tensor-core GEMM for fp32
///   "sequential_reduce" 1 = reference summation order for reductions
///

#ifndef LLO_SESSION_HPP
#define LLO_SESSION_HPP

#include <string>

namespace llo
{

void set_option (std::string name, long value);

long get_option (std::string name);

}

#endif // LLO_SESSION_HPP
