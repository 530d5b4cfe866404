///
/// api.hpp
///
/// Graph builders of the `age` namespace: one function per entry of the
/// reference's op table (llo/cfg/llo.json:125-531, 42 entries, 38 distinct
/// names).  Written out statically; the reference emits them with Tenncor's
/// generator at build time (llo/BUILD.bazel:31-50).
///

#ifndef CORTENN_AGE_API_HPP
#define CORTENN_AGE_API_HPP

#include "ade/functor.hpp"

#include "llo/generated/codes.hpp"

namespace age
{

ade::TensptrT abs (ade::TensptrT arg1);
ade::TensptrT neg (ade::TensptrT arg1);
ade::TensptrT sin (ade::TensptrT arg1);
ade::TensptrT cos (ade::TensptrT arg1);
ade::TensptrT tan (ade::TensptrT arg1);
ade::TensptrT exp (ade::TensptrT arg1);
ade::TensptrT log (ade::TensptrT arg1);
ade::TensptrT sqrt (ade::TensptrT arg1);
ade::TensptrT round (ade::TensptrT arg1);
ade::TensptrT flip (ade::TensptrT arg1, uint8_t arg2);
ade::TensptrT pow (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT add (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT sub (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT mul (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT div (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT eq (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT neq (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT lt (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT gt (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT rand_bino (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT rand_unif (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT rand_norm (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT n_elems (ade::TensptrT arg);
ade::TensptrT n_dims (ade::TensptrT arg, uint8_t rank);
ade::TensptrT sum (ade::TensT arg1);
ade::TensptrT prod (ade::TensT arg1);
ade::TensptrT min (ade::TensT arg1);
ade::TensptrT max (ade::TensT arg1);
ade::TensptrT reduce_sum (ade::TensptrT arg1, uint8_t arg2);
ade::TensptrT reduce_prod (ade::TensptrT arg1, uint8_t arg2);
ade::TensptrT reduce_min (ade::TensptrT arg1, uint8_t arg2);
ade::TensptrT reduce_max (ade::TensptrT arg1, uint8_t arg2);
ade::TensptrT permute (ade::TensptrT arg1, std::vector<uint8_t> arg2);
ade::TensptrT extend (ade::TensptrT arg1, uint8_t arg2,
	std::vector<ade::DimT> arg3);
ade::TensptrT reduce_sum (ade::TensptrT arg1);
ade::TensptrT reduce_prod (ade::TensptrT arg1);
ade::TensptrT reduce_min (ade::TensptrT arg1);
ade::TensptrT reduce_max (ade::TensptrT arg1);
ade::TensptrT transpose (ade::TensptrT arg1);
ade::TensptrT reduce_mean (ade::TensptrT arg1);
ade::TensptrT matmul (ade::TensptrT arg1, ade::TensptrT arg2);
ade::TensptrT convolution (ade::TensptrT arg1, ade::TensptrT arg2);

}

#endif // CORTENN_AGE_API_HPP
