///
/// ade.hpp
///
/// Host graph substrate: an `ade`-compatible expression-graph layer written
/// for this repo.  The reference builds on Tenncor's `ade` module, pinned at
/// commit 959d51e (third_party/repos/tenncor.bzl:3-8) and NOT vendored, so the
/// contract below is the one cortenn's own call sites and golden vectors pin
/// (SURVEY.md Appendix A):
///   * Shape: rank 8, dim 0 fastest            (llo/test/test_api.cpp:38-51)
///   * CoordMap: 9x9 double matrix, row-vector * matrix, row 8 = translation
///                                              (llo/test/test_operator.cpp:29-38)
///   * MappedTensor(tensor, shaper [, map_io, coorder])
///                                              (llo/src/helper.cpp:18-55)
///   * Functor::get(opcode, args), iTraveler double dispatch
///                                              (llo/eval.hpp:26-91)
/// The one deliberate difference: DimT is 32-bit ("wide mode") instead of the
/// reference's uint8_t, so [16384,16384] and 8192^3 shapes are expressible.
///

#ifndef CORTENN_ADE_HPP
#define CORTENN_ADE_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <numeric>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "logs/logs.hpp"
#include "fmts/fmts.hpp"

namespace ade
{

/// Dimension value type (reference: uint8_t; widened here)
using DimT = uint32_t;

/// Coordinate value type: coordinates pass through matrices, so are real
using CDimT = double;

/// Rank index type
using RankT = uint8_t;

/// Element count / linear index type
using NElemT = uint64_t;

/// Number of dimensions every shape carries
const RankT rank_cap = 8;

/// Homogeneous matrix dimension
const RankT mat_dim = rank_cap + 1;

using CoordT = std::array<CDimT,rank_cap>;

using MatrixT = double[mat_dim][mat_dim];

/// Fixed-rank shape, unspecified trailing dimensions are 1
struct Shape final
{
	using iterator = std::array<DimT,rank_cap>::iterator;
	using const_iterator = std::array<DimT,rank_cap>::const_iterator;

	Shape (void)
	{
		dims_.fill(1);
	}

	Shape (const std::vector<DimT>& dims)
	{
		if (dims.size() > rank_cap)
		{
			logs::fatalf("cannot create shape of rank %d beyond rank cap %d",
				(int) dims.size(), (int) rank_cap);
		}
		dims_.fill(1);
		for (size_t i = 0; i < dims.size(); ++i)
		{
			if (0 == dims[i])
			{
				logs::fatalf("cannot create shape with zero dimension %d",
					(int) i);
			}
			dims_[i] = dims[i];
		}
	}

	DimT at (RankT idx) const
	{
		if (idx >= rank_cap)
		{
			logs::fatalf("cannot access out of bounds index %d", (int) idx);
		}
		return dims_[idx];
	}

	NElemT n_elems (void) const
	{
		NElemT n = 1;
		for (DimT d : dims_)
		{
			n *= d;
		}
		return n;
	}

	/// Number of leading dimensions up to the last non-1 dimension
	RankT n_rank (void) const
	{
		RankT r = rank_cap;
		while (r > 0 && dims_[r - 1] == 1)
		{
			--r;
		}
		return r;
	}

	/// True if dims [0, idx) match
	bool compatible_before (const Shape& other, RankT idx) const
	{
		idx = std::min(idx, rank_cap);
		return std::equal(dims_.begin(), dims_.begin() + idx,
			other.dims_.begin());
	}

	/// True if dims [idx, rank_cap) match
	bool compatible_after (const Shape& other, RankT idx) const
	{
		if (idx >= rank_cap)
		{
			return true;
		}
		return std::equal(dims_.begin() + idx, dims_.end(),
			other.dims_.begin() + idx);
	}

	/// "[d0\d1\...\d7]" (pbm/data/graph.txt:2, opt/test/test_shear.cpp:165)
	std::string to_string (void) const
	{
		return fmts::to_string(dims_.begin(), dims_.end());
	}

	iterator begin (void) { return dims_.begin(); }
	iterator end (void) { return dims_.end(); }
	const_iterator begin (void) const { return dims_.begin(); }
	const_iterator end (void) const { return dims_.end(); }

	bool operator == (const Shape& other) const { return dims_ == other.dims_; }
	bool operator != (const Shape& other) const { return dims_ != other.dims_; }

private:
	std::array<DimT,rank_cap> dims_;
};

/// Linear index -> coordinate, dimension 0 varies fastest
inline CoordT coordinate (const Shape& shape, NElemT idx)
{
	CoordT coord;
	NElemT rem = idx;
	for (RankT i = 0; i < rank_cap; ++i)
	{
		DimT d = shape.at(i);
		coord[i] = (CDimT) (rem % d);
		rem /= d;
	}
	if (rem > 0)
	{
		logs::fatalf("cannot get coordinate of index %llu beyond shape %s",
			(unsigned long long) idx, shape.to_string().c_str());
	}
	return coord;
}

/// Coordinate -> linear index.  Real coordinates are accepted: a negative
/// coordinate wraps from the end of its dimension (floored modulo), then
/// the value is truncated toward zero (0.5 -> 0, 1.5 -> 1: llo/test/test_operator.cpp:42-48)
inline NElemT index (const Shape& shape, const CoordT& coord)
{
	NElemT idx = 0;
	for (RankT k = rank_cap; k-- > 0;)
	{
		CDimT limit = (CDimT) shape.at(k);
		CDimT c = coord[k];
		if (c >= limit)
		{
			logs::fatalf("cannot get index of bad coordinate %s for shape %s",
				fmts::to_string(coord.begin(), coord.end()).c_str(),
				shape.to_string().c_str());
		}
		if (c < 0)
		{
			// python-style wrap from the end of the dimension, any number
			// of periods (flip maps use -c - 1; the convolution's kernel
			// map reaches below -limit on size-1 dims)
			c = c - limit * std::floor(c / limit);
			if (c >= limit)
			{
				c = limit - 1;
			}
		}
		idx = idx * shape.at(k) + (NElemT) c;
	}
	return idx;
}

/// Affine coordinate transformation interface
struct iCoordMap
{
	virtual ~iCoordMap (void) = default;

	/// Return this composed with rhs: forward = rhs.forward(this.forward(x))
	virtual iCoordMap* connect (const iCoordMap& rhs) const = 0;

	/// out[j] = sum_{i<8} in[i] * M[i][j] + M[8][j]
	virtual void forward (CoordT::iterator out,
		CoordT::const_iterator in) const = 0;

	/// Return newly allocated inverse map
	virtual iCoordMap* reverse (void) const = 0;

	virtual std::string to_string (void) const = 0;

	/// Read-only visit of the underlying matrix
	virtual void access (
		std::function<void(const MatrixT&)> cb) const = 0;
};

using CoordptrT = std::shared_ptr<iCoordMap>;

/// Matrix-backed coordinate map
struct CoordMap final : public iCoordMap
{
	/// Matrix starts from zeros and is filled by init
	CoordMap (std::function<void(MatrixT)> init)
	{
		for (RankT i = 0; i < mat_dim; ++i)
		{
			for (RankT j = 0; j < mat_dim; ++j)
			{
				fwd_[i][j] = 0;
			}
		}
		init(fwd_);
	}

	iCoordMap* connect (const iCoordMap& rhs) const override
	{
		return new CoordMap([&](MatrixT out)
		{
			rhs.access([&](const MatrixT& r)
			{
				for (RankT i = 0; i < mat_dim; ++i)
				{
					for (RankT j = 0; j < mat_dim; ++j)
					{
						double acc = 0;
						for (RankT k = 0; k < mat_dim; ++k)
						{
							acc += fwd_[i][k] * r[k][j];
						}
						out[i][j] = acc;
					}
				}
			});
		});
	}

	void forward (CoordT::iterator out,
		CoordT::const_iterator in) const override
	{
		// `in` and `out` may alias, so stage the result
		CDimT tmp[rank_cap];
		for (RankT j = 0; j < rank_cap; ++j)
		{
			CDimT acc = 0;
			for (RankT i = 0; i < rank_cap; ++i)
			{
				acc += in[i] * fwd_[i][j];
			}
			tmp[j] = acc + fwd_[rank_cap][j];
		}
		std::copy(tmp, tmp + rank_cap, out);
	}

	iCoordMap* reverse (void) const override;

	std::string to_string (void) const override
	{
		std::stringstream ss;
		ss << "[";
		for (RankT i = 0; i < mat_dim; ++i)
		{
			if (i > 0)
			{
				ss << ",\n";
			}
			fmts::to_stream(ss, fwd_[i], fwd_[i] + mat_dim);
		}
		ss << "]";
		return ss.str();
	}

	void access (std::function<void(const MatrixT&)> cb) const override
	{
		cb(fwd_);
	}

private:
	MatrixT fwd_;
};

/// Gauss-Jordan inverse with partial pivoting.  Exact for the permutation /
/// diagonal / flip matrices the map builders produce.
inline iCoordMap* CoordMap::reverse (void) const
{
	double aug[mat_dim][2 * mat_dim];
	for (RankT i = 0; i < mat_dim; ++i)
	{
		for (RankT j = 0; j < mat_dim; ++j)
		{
			aug[i][j] = fwd_[i][j];
			aug[i][mat_dim + j] = (i == j) ? 1 : 0;
		}
	}
	for (RankT col = 0; col < mat_dim; ++col)
	{
		RankT pivot = col;
		for (RankT r = col + 1; r < mat_dim; ++r)
		{
			if (std::fabs(aug[r][col]) > std::fabs(aug[pivot][col]))
			{
				pivot = r;
			}
		}
		if (aug[pivot][col] == 0)
		{
			logs::fatalf("cannot invert singular coordinate map %s",
				to_string().c_str());
		}
		if (pivot != col)
		{
			std::swap_ranges(aug[pivot], aug[pivot] + 2 * mat_dim, aug[col]);
		}
		double p = aug[col][col];
		for (RankT j = 0; j < 2 * mat_dim; ++j)
		{
			aug[col][j] /= p;
		}
		for (RankT r = 0; r < mat_dim; ++r)
		{
			double f = aug[r][col];
			if (r == col || f == 0)
			{
				continue;
			}
			for (RankT j = 0; j < 2 * mat_dim; ++j)
			{
				aug[r][j] -= f * aug[col][j];
			}
		}
	}
	return new CoordMap([&](MatrixT out)
	{
		for (RankT i = 0; i < mat_dim; ++i)
		{
			for (RankT j = 0; j < mat_dim; ++j)
			{
				double v = aug[i][mat_dim + j];
				out[i][j] = (v == 0) ? 0 : v; // no negative zeros
			}
		}
	});
}

/// The identity map.  Pointer identity is meaningful: the reference's unary
/// fast path tests `mapper == ade::identity` (llo/operator.hpp:34)
inline CoordptrT make_identity_ (void)
{
	return std::make_shared<CoordMap>([](MatrixT m)
	{
		for (RankT i = 0; i < mat_dim; ++i)
		{
			m[i][i] = 1;
		}
	});
}

inline const CoordptrT identity = make_identity_();

/// Value test for the identity map: pointer identity does not survive
/// separately loaded shared objects (each gets its own `identity`)
inline bool is_identity (const iCoordMap* map)
{
	if (map == identity.get())
	{
		return true;
	}
	bool same = true;
	map->access([&](const MatrixT& m)
	{
		for (RankT i = 0; i < mat_dim && same; ++i)
		{
			for (RankT j = 0; j < mat_dim; ++j)
			{
				if (m[i][j] != ((i == j) ? 1.0 : 0.0))
				{
					same = false;
					break;
				}
			}
		}
	});
	return same;
}

/// Dimensions [rank, rank+|red|) shrink by red[i]  (M[d][d] = 1/red[i])
inline CoordptrT reduce (RankT rank, std::vector<DimT> red)
{
	size_t n = red.size();
	if (rank + n > rank_cap)
	{
		logs::fatalf("cannot reduce shape rank %d beyond rank_cap with n_red %d",
			(int) rank, (int) n);
	}
	if (std::any_of(red.begin(), red.end(), [](DimT d) { return 0 == d; }))
	{
		logs::fatalf("cannot reduce using zero dimensions %s",
			fmts::to_string(red.begin(), red.end()).c_str());
	}
	return std::make_shared<CoordMap>([&](MatrixT m)
	{
		for (RankT i = 0; i < mat_dim; ++i)
		{
			m[i][i] = 1;
		}
		for (size_t i = 0; i < n; ++i)
		{
			m[rank + i][rank + i] = 1.0 / red[i];
		}
	});
}

/// Dimensions [rank, rank+|ext|) grow by ext[i]  (M[d][d] = ext[i])
inline CoordptrT extend (RankT rank, std::vector<DimT> ext)
{
	size_t n = ext.size();
	if (rank + n > rank_cap)
	{
		logs::fatalf("cannot extend shape rank %d beyond rank_cap with n_ext %d",
			(int) rank, (int) n);
	}
	if (std::any_of(ext.begin(), ext.end(), [](DimT d) { return 0 == d; }))
	{
		logs::fatalf("cannot extend using zero dimensions %s",
			fmts::to_string(ext.begin(), ext.end()).c_str());
	}
	return std::make_shared<CoordMap>([&](MatrixT m)
	{
		for (RankT i = 0; i < mat_dim; ++i)
		{
			m[i][i] = 1;
		}
		for (size_t i = 0; i < n; ++i)
		{
			m[rank + i][rank + i] = ext[i];
		}
	});
}

/// out_shape[j] = in_shape[order[j]]; unlisted dims follow in ascending order
/// (llo/test/test_api.cpp:739-783; matrices decoded from pbm/data/graph.pb)
inline CoordptrT permute (std::vector<RankT> order)
{
	bool seen[rank_cap];
	std::fill(seen, seen + rank_cap, false);
	for (RankT d : order)
	{
		if (d >= rank_cap)
		{
			logs::fatalf("cannot permute dimension %d beyond rank_cap %d",
				(int) d, (int) rank_cap);
		}
		if (seen[d])
		{
			logs::fatalf("cannot permute with repeated dimension %d", (int) d);
		}
		seen[d] = true;
	}
	for (RankT d = 0; d < rank_cap; ++d)
	{
		if (!seen[d])
		{
			order.push_back(d);
		}
	}
	return std::make_shared<CoordMap>([&](MatrixT m)
	{
		for (RankT i = 0; i < rank_cap; ++i)
		{
			m[order[i]][i] = 1;
		}
		m[rank_cap][rank_cap] = 1;
	});
}

/// Mirror coordinates along dim: c -> -c - 1, which `index` wraps to
/// dim_size - 1 - c  (llo/test/test_api.cpp:448-456)
inline CoordptrT flip (RankT dim)
{
	if (dim >= rank_cap)
	{
		logs::warnf("flipping dimension %d beyond rank_cap %d is a no-op",
			(int) dim, (int) rank_cap);
		return identity;
	}
	return std::make_shared<CoordMap>([&](MatrixT m)
	{
		for (RankT i = 0; i < mat_dim; ++i)
		{
			m[i][i] = 1;
		}
		m[dim][dim] = -1;
		m[rank_cap][dim] = -1;
	});
}

struct iLeaf;
struct iFunctor;

/// Graph visitor interface
struct iTraveler
{
	virtual ~iTraveler (void) = default;

	virtual void visit (iLeaf* leaf) = 0;

	virtual void visit (iFunctor* func) = 0;
};

/// Graph node interface
struct iTensor
{
	virtual ~iTensor (void) = default;

	virtual void accept (iTraveler& visiter) = 0;

	virtual const Shape& shape (void) const = 0;

	virtual std::string to_string (void) const = 0;
};

using TensptrT = std::shared_ptr<iTensor>;

using TensT = std::vector<TensptrT>;

/// Data-holding node interface
struct iLeaf : public iTensor
{
	virtual ~iLeaf (void) = default;

	void accept (iTraveler& visiter) override
	{
		visiter.visit(this);
	}

	virtual void* data (void) = 0;

	virtual const void* data (void) const = 0;

	virtual size_t type_code (void) const = 0;
};

/// Tensor argument with its shape transform (shaper) and its coordinate
/// transform (coorder).  map_io true: coorder maps argument coordinates to
/// functor coordinates ("push"); false: functor coordinates to argument
/// coordinates ("pull").
struct MappedTensor final
{
	MappedTensor (TensptrT tensor, CoordptrT shaper) :
		tensor_(tensor), shaper_(shaper)
	{
		check_();
		map_io_ = tensor_->shape().n_elems() > shape().n_elems();
		if (map_io_ || is_identity(shaper_.get()))
		{
			coorder_ = shaper_;
		}
		else
		{
			coorder_ = CoordptrT(shaper_->reverse());
		}
	}

	MappedTensor (TensptrT tensor, CoordptrT shaper,
		bool map_io, CoordptrT coorder) :
		tensor_(tensor), shaper_(shaper),
		map_io_(map_io), coorder_(coorder)
	{
		check_();
		if (nullptr == coorder_)
		{
			logs::fatal("cannot map a tensor with a null coorder");
		}
	}

	/// Shape of the argument as seen by the functor: shaper applied to the
	/// argument's shape.  A negative result (flip) encodes -(d) - 1.
	Shape shape (void) const
	{
		const Shape& in = tensor_->shape();
		CoordT cin;
		CoordT cout;
		std::copy(in.begin(), in.end(), cin.begin());
		shaper_->forward(cout.begin(), cin.begin());
		std::vector<DimT> slist(rank_cap);
		for (RankT i = 0; i < rank_cap; ++i)
		{
			CDimT cd = cout[i];
			if (cd < 0)
			{
				cd = -cd - 1;
			}
			slist[i] = (DimT) std::llround(cd);
		}
		return Shape(slist);
	}

	TensptrT get_tensor (void) const { return tensor_; }

	CoordptrT get_shaper (void) const { return shaper_; }

	bool map_io (void) const { return map_io_; }

	CoordptrT get_coorder (void) const { return coorder_; }

private:
	void check_ (void) const
	{
		if (nullptr == tensor_)
		{
			logs::fatal("cannot map a null tensor");
		}
		if (nullptr == shaper_)
		{
			logs::fatal("cannot map a tensor with a null shaper");
		}
	}

	TensptrT tensor_;

	CoordptrT shaper_;

	bool map_io_;

	CoordptrT coorder_;
};

using ArgsT = std::vector<MappedTensor>;

/// Operation identifier
struct Opcode final
{
	std::string name_;

	size_t code_;
};

/// Operation node interface
struct iFunctor : public iTensor
{
	virtual ~iFunctor (void) = default;

	void accept (iTraveler& visiter) override
	{
		visiter.visit(this);
	}

	virtual Opcode get_opcode (void) const = 0;

	virtual const ArgsT& get_children (void) const = 0;
};

/// Operation node: opcode over mapped arguments; shape inferred from the
/// first argument, every other argument must map to the same shape
struct Functor final : public iFunctor
{
	static Functor* get (Opcode opcode, ArgsT args)
	{
		if (args.empty())
		{
			logs::fatalf("cannot perform %s with no arguments",
				opcode.name_.c_str());
		}
		Shape shape = args[0].shape();
		for (size_t i = 1, n = args.size(); i < n; ++i)
		{
			Shape ishape = args[i].shape();
			if (false == ishape.compatible_after(shape, 0))
			{
				logs::fatalf("cannot perform %s with incompatible shapes %s "
					"and %s", opcode.name_.c_str(), shape.to_string().c_str(),
					ishape.to_string().c_str());
			}
		}
		return new Functor(opcode, shape, args);
	}

	const Shape& shape (void) const override { return shape_; }

	std::string to_string (void) const override { return opcode_.name_; }

	Opcode get_opcode (void) const override { return opcode_; }

	const ArgsT& get_children (void) const override { return args_; }

private:
	Functor (Opcode opcode, Shape shape, ArgsT args) :
		opcode_(opcode), shape_(shape), args_(args) {}

	Opcode opcode_;

	Shape shape_;

	ArgsT args_;
};

inline MappedTensor identity_map (TensptrT tensor)
{
	return MappedTensor(tensor, identity);
}

inline MappedTensor reduce_map (TensptrT tensor, RankT rank,
	std::vector<DimT> red)
{
	return MappedTensor(tensor, reduce(rank, red));
}

inline MappedTensor extend_map (TensptrT tensor, RankT rank,
	std::vector<DimT> ext)
{
	return MappedTensor(tensor, extend(rank, ext));
}

inline MappedTensor permute_map (TensptrT tensor, std::vector<RankT> order)
{
	return MappedTensor(tensor, permute(order));
}

inline MappedTensor flip_map (TensptrT tensor, RankT dim)
{
	return MappedTensor(tensor, flip(dim));
}

inline ArgsT to_args (TensT tens)
{
	ArgsT args;
	for (TensptrT& t : tens)
	{
		args.push_back(identity_map(t));
	}
	return args;
}

/// Height of every sub-graph: leaf 0, functor 1 + max(children).  Parents
/// are strictly taller than their children, so sorting by it is topological
/// (opt/shear.hpp:107-121, pbm/src/save.cpp:22-26)
struct GraphStat final : public iTraveler
{
	void visit (iLeaf* leaf) override
	{
		graphsize_.emplace(leaf, 0);
	}

	void visit (iFunctor* func) override
	{
		if (graphsize_.end() != graphsize_.find(func))
		{
			return;
		}
		size_t height = 0;
		for (const MappedTensor& child : func->get_children())
		{
			iTensor* tens = child.get_tensor().get();
			tens->accept(*this);
			height = std::max(height, graphsize_[tens]);
		}
		graphsize_[func] = height + 1;
	}

	std::unordered_map<iTensor*,size_t> graphsize_;
};

/// Records, for every functor with a path to target, which argument
/// indices lead there
struct PathFinder final : public iTraveler
{
	using ParentMapT = std::unordered_map<iTensor*,std::unordered_set<size_t>>;

	PathFinder (const iTensor* target) : target_(target) {}

	void visit (iLeaf*) override {}

	void visit (iFunctor* func) override
	{
		if (visited_.end() != visited_.find(func))
		{
			return;
		}
		visited_.emplace(func);
		const ArgsT& children = func->get_children();
		std::unordered_set<size_t> path;
		for (size_t i = 0, n = children.size(); i < n; ++i)
		{
			iTensor* tens = children[i].get_tensor().get();
			if (tens == target_)
			{
				path.emplace(i);
			}
			else
			{
				tens->accept(*this);
				if (parents_.end() != parents_.find(tens))
				{
					path.emplace(i);
				}
			}
		}
		if (false == path.empty())
		{
			parents_[func] = path;
			order_.push_back(func);
		}
	}

	const iTensor* target_;

	ParentMapT parents_;

	/// Path functors in post-order (children before parents): a
	/// deterministic tie-break for equal-height parents
	std::vector<iFunctor*> order_;

private:
	std::unordered_set<iTensor*> visited_;
};

}

#endif // CORTENN_ADE_HPP
