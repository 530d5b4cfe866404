///
/// mapview.hpp — host-side classification of ade::CoordMap matrices.
///
/// The reference walks every element through a virtual 9x9 double mat-vec
/// (`coordinate -> mapper->forward -> index`, llo/operator.hpp:46-58,
/// 181-227, 380-400).  Here each argument's matrix is analysed ONCE per
/// operator call and folded into integer strides that the kernels add to the
/// load address:
///   pull (out coords -> arg coords): arg index = offset + sum_i o_i*stride_i
///   push (arg coords -> out coords): out index = offset + sum_i a_i*stride_i
/// Classification is exact: it only succeeds when, for every coordinate in
/// the iteration box, the reference's double arithmetic + truncation gives
/// that same integer (integer matrix columns, or size-1 target dimensions
/// whose coordinate truncates/wraps to 0).  Anything else (fractional
/// scales onto real dimensions, mixed-sign wraps) takes the general path,
/// which evaluates the matrix in double on the device.
///

#ifndef LLO_CUDA_MAPVIEW_HPP
#define LLO_CUDA_MAPVIEW_HPP

#include <cmath>
#include <cstdint>

#include "llo_cuda.h"

namespace llo_cuda
{

enum MapKind
{
	MAP_LINEAR = 0, // integer strides + offset reproduce the map exactly
	MAP_GENERAL,    // needs the per-element double evaluation
	MAP_INVALID,    // some coordinate provably leaves the target shape
};

struct LinearMap
{
	/// stride (in target elements) per source dimension, offset in elements
	int64_t stride[LLO_RANK];
	int64_t offset;
};

inline bool is_identity_matrix (const double m[LLO_MAT])
{
	for (int i = 0; i < 9; ++i)
	{
		for (int j = 0; j < 9; ++j)
		{
			if (m[i * 9 + j] != (i == j ? 1.0 : 0.0))
			{
				return false;
			}
		}
	}
	return true;
}

/// Classify the map M taking coordinates of a box `from` (iteration space)
/// to coordinates of a tensor of shape `to`; on MAP_LINEAR fill `out` so
/// that index(to, M(coordinate(from, i))) == offset + sum_k c_k * stride_k.
inline MapKind classify_map (const double m[LLO_MAT],
	const uint64_t from[LLO_RANK], const uint64_t to[LLO_RANK],
	LinearMap& out)
{
	int64_t tostride[LLO_RANK];
	int64_t acc = 1;
	for (int j = 0; j < LLO_RANK; ++j)
	{
		tostride[j] = acc;
		acc *= (int64_t) to[j];
	}
	for (int i = 0; i < LLO_RANK; ++i)
	{
		out.stride[i] = 0;
	}
	out.offset = 0;
	for (int j = 0; j < LLO_RANK; ++j)
	{
		// range of target coordinate j over the box
		double lo = m[8 * 9 + j];
		double hi = lo;
		bool integral = (std::floor(lo) == lo);
		for (int i = 0; i < LLO_RANK; ++i)
		{
			if (from[i] <= 1)
			{
				continue; // coordinate i is always 0
			}
			double w = m[i * 9 + j];
			if (0 == w)
			{
				continue;
			}
			integral = integral && (std::floor(w) == w);
			double span = w * (double) (from[i] - 1);
			if (span < 0)
			{
				lo += span;
			}
			else
			{
				hi += span;
			}
		}
		double limit = (double) to[j];
		if (hi >= limit)
		{
			return MAP_INVALID;
		}
		if (1 == to[j])
		{
			// every value below 1 wraps / truncates to coordinate 0
			continue;
		}
		if (false == integral)
		{
			return MAP_GENERAL;
		}
		int64_t shift = 0;
		if (lo < 0)
		{
			if (hi >= 0 || lo < -limit)
			{
				return MAP_GENERAL; // the range wraps unevenly
			}
			shift = (int64_t) to[j]; // whole range wraps once from the end
		}
		out.offset += ((int64_t) m[8 * 9 + j] + shift) * tostride[j];
		for (int i = 0; i < LLO_RANK; ++i)
		{
			if (from[i] > 1)
			{
				out.stride[i] += (int64_t) m[i * 9 + j] * tostride[j];
			}
		}
	}
	return MAP_LINEAR;
}

/// Dense strides of a shape
inline void dense_strides (const uint64_t shape[LLO_RANK],
	int64_t stride[LLO_RANK])
{
	int64_t acc = 1;
	for (int i = 0; i < LLO_RANK; ++i)
	{
		stride[i] = acc;
		acc *= (int64_t) shape[i];
	}
}

/// Structure of a push map whose index function is linear: which source
/// dims are folded away (stride 0) and whether the others hit every output
/// exactly once
struct PushShape
{
	bool clean;             // kept dims biject onto the output elements
	bool bijective;         // clean and nothing is folded
	int nfold;
	uint64_t fold_size[LLO_RANK];   // folded source dims, fastest first
	int64_t fold_stride[LLO_RANK];  // their strides in the SOURCE tensor
	/// pull view of the kept part: source index of output coordinate o
	/// (with every folded coordinate 0) = offset + sum_d o_d * stride[d]
	LinearMap inverse;
};

inline PushShape analyse_push (const LinearMap& push,
	const uint64_t src[LLO_RANK], const uint64_t dst[LLO_RANK])
{
	PushShape ps;
	ps.clean = false;
	ps.bijective = false;
	ps.nfold = 0;
	int64_t srcstride[LLO_RANK];
	int64_t dststride[LLO_RANK];
	dense_strides(src, srcstride);
	dense_strides(dst, dststride);
	for (int d = 0; d < LLO_RANK; ++d)
	{
		ps.inverse.stride[d] = 0;
	}
	ps.inverse.offset = 0;
	bool used[LLO_RANK] = {false};
	int64_t remaining = push.offset;
	for (int i = 0; i < LLO_RANK; ++i)
	{
		if (src[i] <= 1)
		{
			continue;
		}
		int64_t s = push.stride[i];
		if (0 == s)
		{
			ps.fold_size[ps.nfold] = src[i];
			ps.fold_stride[ps.nfold] = srcstride[i];
			++ps.nfold;
			continue;
		}
		// a kept source dim must own exactly one output dim
		int match = -1;
		for (int d = 0; d < LLO_RANK; ++d)
		{
			if (dst[d] > 1 && (s == dststride[d] || s == -dststride[d]))
			{
				match = d;
				break;
			}
		}
		if (match < 0 || used[match] || dst[match] != src[i])
		{
			return ps;
		}
		used[match] = true;
		if (s > 0)
		{
			ps.inverse.stride[match] = srcstride[i];
		}
		else
		{
			// mirrored: out coord o = (n-1) - a  ->  a = (n-1) - o
			ps.inverse.stride[match] = -srcstride[i];
			ps.inverse.offset += (int64_t) (src[i] - 1) * srcstride[i];
			remaining -= (int64_t) (dst[match] - 1) * dststride[match];
		}
	}
	for (int d = 0; d < LLO_RANK; ++d)
	{
		if (dst[d] > 1 && false == used[d])
		{
			return ps; // some output dim is never reached
		}
	}
	if (0 != remaining)
	{
		return ps; // shifted: not every output is covered
	}
	ps.clean = true;
	ps.bijective = (0 == ps.nfold);
	return ps;
}

}

#endif // LLO_CUDA_MAPVIEW_HPP
