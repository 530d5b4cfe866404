#include "llo/generated/api.hpp"
#include "llo/generated/codes.hpp"
#include "llo/data.hpp"
#include "llo/helper.hpp"

#ifdef LLO_HELPER_HPP

namespace llo
{

static inline ade::Opcode opcode_of (age::_GENERATED_OPCODE code)
{
	return ade::Opcode{age::name_op(code), (size_t) code};
}

// reference: llo/src/helper.cpp:11-16
ade::TensptrT mtens_mul (ade::TensptrT lhs, ade::MappedTensor rhs)
{
	return ade::TensptrT(ade::Functor::get(opcode_of(age::PROD),
		{ade::identity_map(lhs), rhs}));
}

/// Bring the forward value of `fwd` back into the coordinate space of its
/// argument `gradidx`: the argument's shaper reversed, direction flipped
static ade::TensptrT fwd_in_arg_space (ade::iFunctor* fwd, size_t gradidx)
{
	const ade::ArgsT& children = fwd->get_children();
	ade::TensptrT again(ade::Functor::get(fwd->get_opcode(), children));
	const ade::MappedTensor& child = children[gradidx];
	ade::MappedTensor back(again,
		ade::CoordptrT(child.get_shaper()->reverse()),
		!child.map_io(), child.get_coorder());
	return ade::TensptrT(ade::Functor::get(opcode_of(age::SUM), {back}));
}

// reference: llo/src/helper.cpp:18-33 -- d(prod)/d(arg) = prod / arg
ade::TensptrT grad_prod (ade::iFunctor* fwd, size_t gradidx, ade::TensT tens)
{
	return age::div(fwd_in_arg_space(fwd, gradidx), tens[gradidx]);
}

/// reference: llo/src/helper.cpp:35-55 -- 1 where the argument attains the
/// extremum.  The reference maps fwd back with the 2-argument MappedTensor
/// (direction inferred from element counts).
static ade::TensptrT grad_extremum (ade::iFunctor* fwd, size_t gradidx,
	ade::TensT& tens)
{
	const ade::ArgsT& children = fwd->get_children();
	ade::TensptrT again(ade::Functor::get(fwd->get_opcode(), children));
	ade::CoordptrT shaper(children[gradidx].get_shaper()->reverse());
	ade::TensptrT back(ade::Functor::get(opcode_of(age::SUM),
		{ade::MappedTensor(again, shaper)}));
	return age::eq(back, tens[gradidx]);
}

ade::TensptrT grad_min (ade::iFunctor* fwd, size_t gradidx, ade::TensT tens)
{
	return grad_extremum(fwd, gradidx, tens);
}

ade::TensptrT grad_max (ade::iFunctor* fwd, size_t gradidx, ade::TensT tens)
{
	return grad_extremum(fwd, gradidx, tens);
}

// reference: llo/src/helper.cpp:57-65 -- folds ALL dimensions >= dim
ade::TensptrT reduce (ade::Opcode opcode, ade::TensptrT tens, uint8_t dim)
{
	const ade::Shape& shape = tens->shape();
	if (dim >= ade::rank_cap)
	{
		logs::fatalf("cannot reduce dimensions beyond rank_cap %d",
			(int) ade::rank_cap);
	}
	std::vector<ade::DimT> folded(shape.begin() + dim, shape.end());
	return ade::TensptrT(ade::Functor::get(opcode,
		{ade::reduce_map(tens, dim, folded)}));
}

// reference: llo/src/helper.cpp:67-97
// a: [C,N], b: [M,C] (dim 0 fastest) -> [M,N], out[m,n] = sum_c a[c,n]*b[m,c]
ade::TensptrT matmul (ade::TensptrT a, ade::TensptrT b)
{
	const ade::Shape& ashape = a->shape();
	const ade::Shape& bshape = b->shape();

	ade::DimT ncommon = ashape.at(0);
	ade::DimT nrow = ashape.at(1);
	ade::DimT ncol = bshape.at(0);
	if (ncommon != bshape.at(1))
	{
		logs::fatalf("cannot matmul shapes of incompatible common dimension "
			"%s and %s", ashape.to_string().c_str(),
			bshape.to_string().c_str());
	}
	if ((ade::NElemT) ncommon * nrow != ashape.n_elems())
	{
		logs::fatalf("cannot matmul ashape %s of dimension "
			"higher than 2-D", ashape.to_string().c_str());
	}
	if ((ade::NElemT) ncommon * ncol != bshape.n_elems())
	{
		logs::fatalf("cannot matmul bshape %s of dimension "
			"higher than 2-D", bshape.to_string().c_str());
	}

	// both operands are broadcast to [M,N,C], multiplied, and folded over C
	ade::TensptrT lhs = age::permute(age::extend(a, 2, {ncol}), {2, 1, 0});
	ade::TensptrT rhs = age::permute(age::extend(b, 2, {nrow}), {0, 2, 1});
	return age::reduce_sum(age::mul(lhs, rhs), 2);
}

// reference: llo/src/helper.cpp:102-202 (tf.nn.conv2d "VALID" layout)
// img: [in_channels, in_height, in_width, nbatch]
// kernel: [out_channels, in_channels, kernel_height, kernel_width]
ade::TensptrT convolution (ade::TensptrT img, ade::TensptrT kernel)
{
	const ade::Shape& ishape = img->shape();
	const ade::Shape& kshape = kernel->shape();

	ade::DimT nbatch = ishape.at(3);
	ade::DimT iw = ishape.at(2);
	ade::DimT ih = ishape.at(1);
	ade::DimT ic = ishape.at(0);
	ade::DimT kw = kshape.at(3);
	ade::DimT kh = kshape.at(2);
	ade::DimT oc = kshape.at(0);
	if (ic != kshape.at(1))
	{
		logs::fatalf("cannot convolution with mismatch img %s and kernel %s "
			"in_channel (dim=0 for img, dim=1 for kernel)",
			ishape.to_string().c_str(), kshape.to_string().c_str());
	}
	if (iw < kw || ih < kh)
	{
		logs::fatalf("cannot convolution kernel %s against a smaller image %s",
			kshape.to_string().c_str(), ishape.to_string().c_str());
	}

	// rows / columns lost to "VALID" padding
	double lost_h = 2 * (kh / 2);
	double lost_w = 2 * (kw / 2);
	double oh = ih - lost_h;
	double ow = iw - lost_w;
	const ade::RankT last = ade::mat_dim - 1;

	// both operands map to
	// [oc, oh, ow, nbatch, ic, kh, kw]; the product is folded over dims >= 4
	ade::CoordptrT img_shaper(new ade::CoordMap(
		[&](ade::MatrixT m)
		{
			m[0][4] = m[1][1] = m[2][2] = m[3][3] = 1;
			m[4][0] = oc;
			m[5][5] = kh;
			m[6][6] = kw;
			m[last][1] = -lost_h;
			m[last][2] = -lost_w;
			m[7][7] = m[last][last] = 1;
		}));
	// pull: image pixel (h + kh, w + kw) of channel ic for output (h, w)
	ade::CoordptrT img_coorder(new ade::CoordMap(
		[&](ade::MatrixT m)
		{
			m[1][1] = m[2][2] = m[3][3] = m[4][0] = m[5][1] = m[6][2] = 1;
			m[0][4] = 1.0 / oc;
			m[1][5] = 1.0 / oh;
			m[2][6] = 1.0 / ow;
			m[7][7] = m[last][last] = 1;
		}));
	ade::CoordptrT kernel_shaper(new ade::CoordMap(
		[&](ade::MatrixT m)
		{
			m[0][0] = m[1][4] = m[2][5] = m[3][6] = 1;
			m[4][1] = m[5][2] = m[6][3] = 1;
			m[last][1] = oh - 1;
			m[last][2] = ow - 1;
			m[last][3] = (double) nbatch - 1;
			m[7][7] = m[last][last] = 1;
		}));

	ade::TensptrT prod(ade::Functor::get(opcode_of(age::PROD), {
		ade::MappedTensor(img, img_shaper, false, img_coorder),
		ade::MappedTensor(kernel, kernel_shaper),
	}));
	return age::reduce_sum(prod, 4);
}

}

#endif
