#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <unordered_map>

#include "llo/plan.hpp"
#include "llo/session.hpp"

#include "llo/generated/opmap.hpp"

#ifdef LLO_PLAN_HPP

namespace llo
{

namespace
{

using DT = age::_GENERATED_DTYPE;

constexpr int R = ade::rank_cap;

// ---- coordinate views -------------------------------------------------------

/// target coordinate j = sum_i coef[i] * c[i] + offset: one column of an
/// integer affine coordinate map.  Most columns have a single source (extend,
/// permute, flip); a convolution's image argument has two (h = oh + kh,
/// llo/src/helper.cpp:149-165).
struct DimMap
{
	int64_t coef_[R] = {0, 0, 0, 0, 0, 0, 0, 0};
	int64_t offset_ = 0;

	bool operator == (const DimMap& o) const
	{
		return offset_ == o.offset_ &&
			0 == std::memcmp(coef_, o.coef_, sizeof(coef_));
	}

	/// number of source coordinates the column depends on
	int nsrc (void) const
	{
		int n = 0;
		for (int i = 0; i < R; ++i)
		{
			n += 0 != coef_[i] ? 1 : 0;
		}
		return n;
	}

	/// the only source of a single-source column (-1: none or several)
	int src (void) const
	{
		int found = -1;
		for (int i = 0; i < R; ++i)
		{
			if (0 != coef_[i])
			{
				if (found >= 0)
				{
					return -1;
				}
				found = i;
			}
		}
		return found;
	}
};

/// Integer coordinate map from a consumer's coordinates to a node's
/// coordinates: the exact integer form of an ade::CoordMap
struct CoordView
{
	DimMap t_[R];

	static CoordView identity (void)
	{
		CoordView v;
		for (int j = 0; j < R; ++j)
		{
			v.t_[j].coef_[j] = 1;
		}
		return v;
	}

	bool operator == (const CoordView& o) const
	{
		for (int j = 0; j < R; ++j)
		{
			if (false == (t_[j] == o.t_[j]))
			{
				return false;
			}
		}
		return true;
	}
};

/// outer: consumer -> mid coordinates; inner: mid -> node coordinates
static CoordView compose (const CoordView& outer, const CoordView& inner)
{
	CoordView out;
	for (int k = 0; k < R; ++k)
	{
		const DimMap& in = inner.t_[k];
		DimMap& o = out.t_[k];
		o.offset_ = in.offset_;
		for (int m = 0; m < R; ++m)
		{
			if (0 == in.coef_[m])
			{
				continue;
			}
			const DimMap& mid = outer.t_[m];
			o.offset_ += in.coef_[m] * mid.offset_;
			for (int i = 0; i < R; ++i)
			{
				o.coef_[i] += in.coef_[m] * mid.coef_[i];
			}
		}
	}
	return out;
}

/// Normal form: coordinates of size-1 consumer dims are always 0, so drop
/// them; target dims of size 1 are always coordinate 0
static void normalise (CoordView& v, const ade::Shape& from, const ade::Shape& to)
{
	for (int j = 0; j < R; ++j)
	{
		DimMap& d = v.t_[j];
		if (to.at(j) <= 1)
		{
			d = DimMap();
			continue;
		}
		for (int i = 0; i < R; ++i)
		{
			if (from.at(i) <= 1)
			{
				d.coef_[i] = 0;
			}
		}
	}
}

enum class MapClass
{
	VIEW,     // pull (or bijective push) expressible as a CoordView
	REDUCE,   // push folding whole dims, the kept dims a bijection
	GENERAL,  // needs the per-element double evaluation
};

struct MapInfo
{
	MapClass cls_ = MapClass::GENERAL;
	CoordView view_;          // VIEW: consumer -> arg; REDUCE: out -> arg (folded coords 0)
	std::vector<int> fold_;   // REDUCE: folded arg dims, ascending
};

static bool is_integral (double v)
{
	return std::floor(v) == v;
}

/// Column-wise analysis of matrix m taking box `from` to tensor `to`.
/// Returns false when some target coordinate is not an integer affine
/// function of the source coordinates.  Throws when a coordinate provably
/// leaves `to`.
static bool affine_columns (DimMap (&cols)[R], const ade::CoordptrT& mapper,
	const ade::Shape& from, const ade::Shape& to)
{
	bool ok = true;
	mapper->access([&](const ade::MatrixT& m)
	{
		for (int j = 0; j < R; ++j)
		{
			double lo = m[R][j];
			double hi = lo;
			bool integral = is_integral(lo);
			for (int i = 0; i < R; ++i)
			{
				if (from.at(i) <= 1 || 0 == m[i][j])
				{
					continue;
				}
				integral = integral && is_integral(m[i][j]);
				double span = m[i][j] * (double) (from.at(i) - 1);
				(span < 0 ? lo : hi) += span;
			}
			double limit = (double) to.at(j);
			if (hi >= limit)
			{
				logs::fatalf("cannot get index of bad coordinate: a "
					"coordinate map takes shape %s outside shape %s",
					from.to_string().c_str(), to.to_string().c_str());
			}
			cols[j] = DimMap();
			if (to.at(j) <= 1)
			{
				continue; // always wraps / truncates to 0
			}
			if (false == integral)
			{
				ok = false;
				continue;
			}
			int64_t shift = 0;
			if (lo < 0)
			{
				if (hi >= 0 || lo < -limit)
				{
					ok = false; // the range wraps unevenly
					continue;
				}
				shift = (int64_t) to.at(j);
			}
			cols[j].offset_ = (int64_t) m[R][j] + shift;
			for (int i = 0; i < R; ++i)
			{
				if (from.at(i) > 1 && 0 != m[i][j])
				{
					cols[j].coef_[i] = (int64_t) m[i][j];
				}
			}
		}
	});
	return ok;
}

/// Classify the mapping of one functor argument
static MapInfo classify (const ade::MappedTensor& arg, const ade::Shape& outshape)
{
	MapInfo info;
	const ade::Shape& argshape = arg.get_tensor()->shape();
	ade::CoordptrT coorder = arg.get_coorder();
	if (ade::is_identity(coorder.get()) && argshape == outshape)
	{
		info.cls_ = MapClass::VIEW;
		info.view_ = CoordView::identity();
		normalise(info.view_, outshape, argshape);
		return info;
	}
	DimMap cols[R];
	if (false == arg.map_io())
	{
		// pull: out coordinates -> arg coordinates
		if (affine_columns(cols, coorder, outshape, argshape))
		{
			info.cls_ = MapClass::VIEW;
			for (int j = 0; j < R; ++j)
			{
				info.view_.t_[j] = cols[j];
			}
			normalise(info.view_, outshape, argshape);
		}
		return info;
	}
	// push: arg coordinates -> out coordinates
	if (false == affine_columns(cols, coorder, argshape, outshape))
	{
		return info;
	}
	// invert: every out dim (size > 1) must be fed by its own arg dim with
	// scale +-1 and cover it exactly; arg dims feeding nothing are folded
	bool fed[R];
	std::fill(fed, fed + R, false);
	CoordView inv;
	for (int j = 0; j < R; ++j)
	{
		if (outshape.at(j) <= 1)
		{
			continue;
		}
		const DimMap& c = cols[j];
		int csrc = c.src();
		int64_t cscale = csrc >= 0 ? c.coef_[csrc] : 0;
		if (csrc < 0 || (1 != cscale && -1 != cscale) ||
			fed[csrc] || argshape.at(csrc) != outshape.at(j))
		{
			return info;
		}
		// o = s * a + b  ->  a = s * o - s * b; must cover [0, n)
		int64_t n = (int64_t) outshape.at(j);
		if ((1 == cscale && 0 != c.offset_) ||
			(-1 == cscale && n - 1 != c.offset_))
		{
			return info;
		}
		fed[csrc] = true;
		inv.t_[csrc].coef_[j] = cscale;
		inv.t_[csrc].offset_ = -cscale * c.offset_;
	}
	for (int i = 0; i < R; ++i)
	{
		if (argshape.at(i) > 1 && false == fed[i])
		{
			info.fold_.push_back(i);
		}
	}
	info.view_ = inv;
	normalise(info.view_, outshape, argshape);
	info.cls_ = info.fold_.empty() ? MapClass::VIEW : MapClass::REDUCE;
	return info;
}

// ---- plan nodes ---------------------------------------------------------------

enum class Kind
{
	LEAF,     // device mirror of a leaf
	CONST,    // every element equals value_
	VIEW,     // args_[0] seen through its view
	ELEM,     // elementwise opcode over view arguments
	REDUCE,   // fold of args_[0] over fold_ dims
	GENERIC,  // age::op_exec over materialised arguments
};

struct PNode;

struct PArg
{
	PNode* node_ = nullptr;
	CoordView view_;               // consumer (or REDUCE output) -> node coords
	// GENERIC consumers keep the original mapping
	ade::CoordptrT coorder_;
	bool push_ = false;
};

struct PNode
{
	Kind kind_ = Kind::GENERIC;
	int opcode_ = 0;
	ade::Shape shape_;
	DT dtype_ = age::BAD_TYPE;
	std::vector<PArg> args_;
	ade::iLeaf* leaf_ = nullptr;
	double value_ = 0;
	// REDUCE: args_[0] = the tensor actually read, its view_ maps the
	// coordinates of the functor's argument (shape argshape_) onto it;
	// kept_ maps output coordinates to argument coordinates (folded ones 0)
	std::vector<int> fold_;
	ade::Shape argshape_;
	CoordView kept_;
	bool has_rand_ = false;
	int uses_ = 0;
	bool counted_ = false;
	std::shared_ptr<char> data_;   // device result once computed
	std::shared_ptr<char> dest_;   // where the caller wants the result written, if anywhere
	// cost of the fully inlined elementwise form
	int need_insns_ = 0;
	int need_inputs_ = 0;
	int need_consts_ = 0;
	int need_depth_ = 0;
};

static bool is_unary (int op) { return op >= age::ABS && op <= age::ROUND; }

static bool is_binary (int op)
{
	return age::POW == op || age::SUB == op || age::DIV == op ||
		(op >= age::EQ && op <= age::GT);
}

static bool is_nnary (int op)
{
	return age::SUM == op || age::PROD == op || age::MIN == op || age::MAX == op;
}

static bool is_rand (int op) { return op >= age::RAND_BINO && op <= age::RAND_NORM; }

static bool is_unsigned (DT t)
{
	return age::UINT8 == t || age::UINT16 == t || age::UINT32 == t ||
		age::UINT64 == t;
}

struct Plan
{
	Plan (DT dtype) : dtype_(dtype), simplify_(0 != get_option("simplify")),
		recompute_((int) get_option("recompute")) {}

	PNode* make (Kind kind)
	{
		arena_.push_back(std::unique_ptr<PNode>(new PNode()));
		PNode* n = arena_.back().get();
		n->kind_ = kind;
		n->dtype_ = dtype_;
		return n;
	}

	PNode* make_const (double value, const ade::Shape& shape, DT dtype)
	{
		PNode* n = make(Kind::CONST);
		n->value_ = value;
		n->shape_ = shape;
		n->dtype_ = dtype;
		++stats_.consts_;
		return intern(n);
	}

	// ---- common sub-expression sharing ---------------------------------------
	//
	// llo::derive builds one gradient graph per target; the graphs of
	// different targets (and the paths inside one graph) repeat the same
	// backward chains as distinct ade nodes.  Plan nodes are therefore
	// hash-consed on their structure, so every distinct computation is
	// planned - and run - once per evaluation.

	template <typename V>
	static void put (std::string& key, const V& v)
	{
		key.append((const char*) &v, sizeof(V));
	}

	static void put_view (std::string& key, const CoordView& v)
	{
		for (int j = 0; j < R; ++j)
		{
			put(key, v.t_[j].coef_);
			put(key, v.t_[j].offset_);
		}
	}

	static void put_shape (std::string& key, const ade::Shape& shape)
	{
		for (int j = 0; j < R; ++j)
		{
			put(key, shape.at(j));
		}
	}

	PNode* intern (PNode* n)
	{
		if (n->has_rand_)
		{
			return n; // random draws are never shared
		}
		std::string key;
		put(key, n->kind_);
		put(key, n->opcode_);
		put(key, n->dtype_);
		put_shape(key, n->shape_);
		put(key, n->leaf_);
		if (Kind::CONST == n->kind_)
		{
			put(key, n->value_);
		}
		if (Kind::REDUCE == n->kind_)
		{
			put_shape(key, n->argshape_);
			put_view(key, n->kept_);
			for (int dim : n->fold_)
			{
				put(key, dim);
			}
		}
		for (const PArg& a : n->args_)
		{
			put(key, a.node_);
			put_view(key, a.view_);
			if (Kind::GENERIC == n->kind_)
			{
				put(key, a.push_);
				a.coorder_->access([&](const ade::MatrixT& m)
				{
					for (int i = 0; i <= R; ++i)
					{
						for (int j = 0; j <= R; ++j)
						{
							put(key, m[i][j]);
						}
					}
				});
			}
		}
		auto it = interned_.find(key);
		if (interned_.end() != it)
		{
			++stats_.shared_;
			return it->second;
		}
		interned_.emplace(std::move(key), n);
		return n;
	}

	// ---- building ------------------------------------------------------------

	PNode* build (ade::iTensor* tens, DT dtype)
	{
		auto key = std::make_pair(tens, (int) dtype);
		auto it = memo_.find(key);
		if (memo_.end() != it)
		{
			return it->second;
		}
		PNode* out = nullptr;
		if (auto leaf = dynamic_cast<ade::iLeaf*>(tens))
		{
			out = build_leaf(leaf, dtype);
		}
		else
		{
			out = build_functor(static_cast<ade::iFunctor*>(tens), dtype);
		}
		++stats_.nodes_;
		if (false == out->has_rand_)
		{
			// random draws are fresh per visit (llo/eval.hpp:77-87 re-runs
			// every child), everything else is computed once
			memo_[key] = out;
		}
		return out;
	}

	PNode* build_leaf (ade::iLeaf* leaf, DT dtype)
	{
		auto var = dynamic_cast<Variable*>(leaf);
		if (nullptr != var && var->is_filled())
		{
			double v = var->fill_value();
			// immediates travel as doubles: only fold what a double holds
			if (std::fabs(v) <= 9007199254740992.0)
			{
				return make_const(leaf_to_dtype(v, (DT) var->type_code(), dtype),
					leaf->shape(), dtype);
			}
		}
		PNode* n = make(Kind::LEAF);
		n->leaf_ = leaf;
		n->shape_ = leaf->shape();
		n->dtype_ = dtype;
		return intern(n);
	}

	/// value of a leaf constant after the reference's leaf conversion
	/// (leaf type -> evaluation type, C++ conversion semantics)
	static double leaf_to_dtype (double v, DT from, DT to)
	{
		(void) from;
		switch (to)
		{
			case age::DOUBLE: return v;
			case age::FLOAT: return (double) (float) v;
			case age::INT8: return (double) (int8_t) (int64_t) v;
			case age::UINT8: return (double) (uint8_t) (int64_t) v;
			case age::INT16: return (double) (int16_t) (int64_t) v;
			case age::UINT16: return (double) (uint16_t) (int64_t) v;
			case age::INT32: return (double) (int32_t) (int64_t) v;
			case age::UINT32: return (double) (uint32_t) (int64_t) v;
			case age::INT64: return (double) (int64_t) v;
			case age::UINT64: return v < 0 ? (double) (uint64_t) (int64_t) v :
				(double) (uint64_t) v;
			default: break;
		}
		return v;
	}

	/// Argument as seen through `view`, with views and constants folded
	PArg resolve (PNode* node, const CoordView& view, const ade::Shape& from)
	{
		PArg out;
		out.node_ = node;
		out.view_ = view;
		while (Kind::VIEW == out.node_->kind_)
		{
			const PArg& inner = out.node_->args_[0];
			out.view_ = compose(out.view_, inner.view_);
			out.node_ = inner.node_;
		}
		normalise(out.view_, from, out.node_->shape_);
		return out;
	}

	PNode* generic (ade::iFunctor* func, DT dtype)
	{
		PNode* n = make(Kind::GENERIC);
		n->opcode_ = (int) func->get_opcode().code_;
		n->shape_ = func->shape();
		n->dtype_ = dtype;
		const ade::ArgsT& children = func->get_children();
		for (size_t i = 0, nargs = children.size(); i < nargs; ++i)
		{
			DT childtype = (age::RAND_BINO == n->opcode_ && 1 == i) ?
				age::DOUBLE : dtype;
			PArg arg;
			arg.node_ = build(children[i].get_tensor().get(), childtype);
			arg.coorder_ = children[i].get_coorder();
			arg.push_ = children[i].map_io();
			n->has_rand_ = n->has_rand_ || arg.node_->has_rand_;
			n->args_.push_back(arg);
		}
		n->has_rand_ = n->has_rand_ || is_rand(n->opcode_);
		++stats_.generic_;
		return intern(n);
	}

	PNode* build_functor (ade::iFunctor* func, DT dtype)
	{
		int op = (int) func->get_opcode().code_;
		const ade::ArgsT& children = func->get_children();
		size_t nargs = children.size();
		const ade::Shape& shape = func->shape();
		if (is_rand(op) || 0 == nargs ||
			(age::NEG == op && is_unsigned(dtype)))
		{
			return generic(func, dtype);
		}
		if (age::RAND_BINO == op && 2 != nargs)
		{
			logs::fatalf("cannot RAND_BINO without exactly 2 arguments: "
				"using %d arguments", (int) nargs);
		}
		if (is_unary(op))
		{
			// the reference's unary wrappers only make sense for identity
			// maps (llo/operator.hpp:66-69); anything else keeps its quirks
			if (false == ade::is_identity(children[0].get_coorder().get()) ||
				false == (children[0].get_tensor()->shape() == shape))
			{
				return generic(func, dtype);
			}
		}
		else if (false == is_binary(op) && false == is_nnary(op))
		{
			return generic(func, dtype);
		}
		if ((is_binary(op) && 2 != nargs))
		{
			return generic(func, dtype);
		}

		std::vector<MapInfo> infos;
		for (size_t i = 0; i < nargs; ++i)
		{
			infos.push_back(classify(children[i], shape));
			if (MapClass::GENERAL == infos.back().cls_ ||
				(MapClass::REDUCE == infos.back().cls_ &&
				false == is_nnary(op)))
			{
				return generic(func, dtype);
			}
		}

		std::vector<PArg> args;
		bool has_rand = false;
		for (size_t i = 0; i < nargs; ++i)
		{
			PNode* child = build(children[i].get_tensor().get(), dtype);
			has_rand = has_rand || child->has_rand_;
			if (MapClass::REDUCE == infos[i].cls_)
			{
				PNode* red = make_reduce(op, shape, child, infos[i], dtype);
				args.push_back(resolve(red, CoordView::identity(), shape));
			}
			else
			{
				args.push_back(resolve(child, infos[i].view_, shape));
			}
		}
		PNode* out = make_elem(op, shape, args, dtype);
		out->has_rand_ = out->has_rand_ || has_rand;
		return out;
	}

	PNode* make_reduce (int op, const ade::Shape& outshape, PNode* child,
		const MapInfo& info, DT dtype)
	{
		PNode* n = make(Kind::REDUCE);
		n->opcode_ = op;
		n->shape_ = outshape;
		n->dtype_ = dtype;
		n->fold_ = info.fold_;
		n->argshape_ = child->shape_;
		n->kept_ = info.view_;
		n->args_.push_back(resolve(child, CoordView::identity(),
			child->shape_));
		n->has_rand_ = child->has_rand_;
		return intern(n);
	}

	PNode* make_view (const ade::Shape& shape, const PArg& arg, DT dtype)
	{
		if (Kind::CONST == arg.node_->kind_)
		{
			return make_const(arg.node_->value_, shape, dtype);
		}
		CoordView same = CoordView::identity();
		normalise(same, shape, shape);
		if (arg.node_->shape_ == shape && arg.view_ == same)
		{
			return arg.node_; // the identity view of a node is the node
		}
		PNode* n = make(Kind::VIEW);
		n->shape_ = shape;
		n->dtype_ = dtype;
		n->args_.push_back(arg);
		n->has_rand_ = arg.node_->has_rand_;
		++stats_.views_;
		return intern(n);
	}

	static bool is_const (const PArg& a, double v)
	{
		return Kind::CONST == a.node_->kind_ && v == a.node_->value_;
	}

	static bool same_arg (const PArg& a, const PArg& b)
	{
		return a.node_ == b.node_ && a.view_ == b.view_;
	}

	/// op over constant arguments in the arithmetic of T, argument order as
	/// the device applies it (only operations IEEE defines exactly)
	template <typename T>
	static bool fold_consts (double& out, int op, const std::vector<PArg>& args)
	{
		T acc = (T) args[0].node_->value_;
		if (age::NEG == op) { out = (double) (T) -acc; return true; }
		if (age::ABS == op) { out = (double) (T) std::fabs(acc); return true; }
		for (size_t i = 1; i < args.size(); ++i)
		{
			T x = (T) args[i].node_->value_;
			switch (op)
			{
				case age::SUM: acc = acc + x; break;
				case age::PROD: acc = acc * x; break;
				case age::SUB: acc = acc - x; break;
				case age::DIV: acc = acc / x; break;
				default: return false;
			}
		}
		if (args.size() < 2)
		{
			return false;
		}
		out = (double) acc;
		return true;
	}

	PNode* make_elem (int op, const ade::Shape& shape, std::vector<PArg> args,
		DT dtype)
	{
		// constant sub-expressions (the derivative rules build -1 as neg(1),
		// 2 - 1, 1 / batch, ...) never reach the device
		if ((age::FLOAT == dtype || age::DOUBLE == dtype) && false == args.empty())
		{
			bool all_const = true;
			for (const PArg& a : args)
			{
				all_const = all_const && Kind::CONST == a.node_->kind_;
			}
			double value = 0;
			if (all_const && (age::FLOAT == dtype ?
				fold_consts<float>(value, op, args) :
				fold_consts<double>(value, op, args)))
			{
				++stats_.simplified_;
				return make_const(value, shape, dtype);
			}
		}
		if (simplify_ && age::POW == op && 2 == args.size() && is_const(args[1], 1))
		{
			// pow(x, 1) -> x (exact)
			++stats_.simplified_;
			return make_view(shape, args[0], dtype);
		}
		if (simplify_)
		{
			if (age::PROD == op || age::SUM == op)
			{
				// x * 1 -> x, x + 0 -> x
				double neutral = age::PROD == op ? 1 : 0;
				std::vector<PArg> kept;
				for (PArg& a : args)
				{
					if (is_const(a, neutral))
					{
						++stats_.simplified_;
					}
					else
					{
						kept.push_back(a);
					}
				}
				if (kept.empty())
				{
					return make_const(neutral, shape, dtype);
				}
				args = kept;
			}
			else if (age::DIV == op && (age::FLOAT == dtype || age::DOUBLE == dtype) &&
				Kind::ELEM == args[0].node_->kind_ &&
				age::PROD == args[0].node_->opcode_)
			{
				// (x * y) / x -> y: the reference's product rule
				// (llo/src/helper.cpp:18-33).  Floating types only: integer
				// products wrap and integer division truncates, so the
				// quotient is not the other factor there.
				PNode* prod = args[0].node_;
				std::vector<PArg> factors;
				for (const PArg& f : prod->args_)
				{
					PArg seen = f;
					seen.view_ = compose(args[0].view_, f.view_);
					normalise(seen.view_, shape, f.node_->shape_);
					factors.push_back(seen);
				}
				for (size_t i = 0; i < factors.size(); ++i)
				{
					if (same_arg(factors[i], args[1]))
					{
						factors.erase(factors.begin() + i);
						++stats_.simplified_;
						if (factors.empty())
						{
							return make_const(1, shape, dtype);
						}
						return make_elem(age::PROD, shape, factors, dtype);
					}
				}
			}
		}
		if (is_nnary(op) && 1 == args.size())
		{
			// a single-argument accumulation just copies through its map
			return make_view(shape, args[0], dtype);
		}
		if (is_nnary(op) && args.size() > 6)
		{
			// keep every program within the engine's input limit: fold the
			// first arguments (same left-to-right order), then the rest
			std::vector<PArg> first(args.begin(), args.begin() + 6);
			PNode* partial = make_elem(op, shape, first, dtype);
			std::vector<PArg> rest;
			rest.push_back(resolve(partial, CoordView::identity(), shape));
			rest.insert(rest.end(), args.begin() + 6, args.end());
			return make_elem(op, shape, rest, dtype);
		}
		PNode* n = make(Kind::ELEM);
		n->opcode_ = op;
		n->shape_ = shape;
		n->dtype_ = dtype;
		n->args_ = args;
		for (const PArg& a : args)
		{
			n->has_rand_ = n->has_rand_ || a.node_->has_rand_;
		}
		return intern(n);
	}

	// ---- use counting and inlining costs --------------------------------------

	void count_uses (PNode* n)
	{
		++n->uses_;
		if (n->counted_)
		{
			return;
		}
		n->counted_ = true;
		for (PArg& a : n->args_)
		{
			count_uses(a.node_);
		}
	}

	/// A producer is emitted inside its consumer's program when nothing else
	/// needs its value -- or ("recompute" option) when its fully inlined form
	/// is so short that recomputing it in every consumer costs less than a
	/// pass over HBM to store it and one per consumer to read it back
	bool inlinable (PNode* n)
	{
		if (Kind::ELEM != n->kind_ || nullptr != n->data_ || n->has_rand_)
		{
			return false;
		}
		if (1 == n->uses_)
		{
			return true;
		}
		if (recompute_ <= 0)
		{
			return false;
		}
		measure(n);
		return n->need_insns_ <= recompute_ && n->need_inputs_ <= 2;
	}

	/// Size of node n's program when every inlinable producer is inlined
	void measure (PNode* n)
	{
		if (n->need_insns_ > 0)
		{
			return;
		}
		int insns = 0;
		int inputs = 0;
		int consts = 0;
		int depth = 0;
		for (size_t i = 0; i < n->args_.size(); ++i)
		{
			PNode* c = n->args_[i].node_;
			int ci = 1, cin = 0, cc = 0, cd = 1;
			if (Kind::CONST == c->kind_)
			{
				cc = 1;
			}
			else if (inlinable(c))
			{
				measure(c);
				if (fits(c->need_insns_, c->need_inputs_, c->need_consts_,
					c->need_depth_ + 1))
				{
					ci = c->need_insns_;
					cin = c->need_inputs_;
					cc = c->need_consts_;
					cd = c->need_depth_;
				}
				else
				{
					cin = 1;
				}
			}
			else
			{
				cin = 1;
			}
			insns += ci;
			inputs += cin;
			consts += cc;
			depth = std::max(depth, (i > 0 ? 1 : 0) + cd);
		}
		insns += is_nnary(n->opcode_) ? (int) n->args_.size() - 1 : 1;
		n->need_insns_ = insns;
		n->need_inputs_ = inputs;
		n->need_consts_ = consts;
		n->need_depth_ = depth;
	}

	static bool fits (int insns, int inputs, int consts, int depth)
	{
		return insns <= LLO_VM_MAX_INSNS - 4 && inputs <= LLO_VM_MAX_INPUTS - 1 &&
			consts <= LLO_VM_MAX_CONSTS - 1 && depth <= LLO_VM_MAX_DEPTH;
	}

	// ---- execution --------------------------------------------------------------

	struct Program
	{
		std::vector<llo_cuda_view> inputs_;
		std::vector<PNode*> input_nodes_;
		std::vector<double> consts_;
		std::vector<llo_cuda_insn> code_;
		int depth_ = 0;
		int peak_ = 0;
	};

	/// strides of `view` (consumer coords -> node coords) over node's memory
	static void to_strides (llo_cuda_view& out, const CoordView& view,
		const ade::Shape& nodeshape)
	{
		int64_t dense = 1;
		out.offset = 0;
		for (int d = 0; d < R; ++d)
		{
			out.stride[d] = 0;
		}
		for (int j = 0; j < R; ++j)
		{
			const DimMap& m = view.t_[j];
			out.offset += m.offset_ * dense;
			for (int i = 0; i < R; ++i)
			{
				out.stride[i] += m.coef_[i] * dense;
			}
			dense *= (int64_t) nodeshape.at(j);
		}
	}

	void push_insn (Program& prog, uint8_t op, uint8_t arg, int delta)
	{
		prog.code_.push_back(llo_cuda_insn{op, arg});
		prog.depth_ += delta;
		prog.peak_ = std::max(prog.peak_, prog.depth_);
	}

	/// Emit code leaving the value of `arg` (seen through `outer`) on the stack
	void emit_arg (Program& prog, const PArg& arg, const CoordView& outer,
		const ade::Shape& rootshape)
	{
		CoordView view = compose(outer, arg.view_);
		normalise(view, rootshape, arg.node_->shape_);
		PNode* node = arg.node_;
		if (Kind::CONST == node->kind_)
		{
			size_t k = 0;
			for (; k < prog.consts_.size(); ++k)
			{
				// bit patterns: -0.0 and 0.0 are different immediates
				if (0 == std::memcmp(&prog.consts_[k], &node->value_, sizeof(double)))
				{
					break;
				}
			}
			if (k == prog.consts_.size())
			{
				prog.consts_.push_back(node->value_);
			}
			push_insn(prog, LLO_VM_CONST, (uint8_t) k, 1);
			return;
		}
		if (inlinable(node))
		{
			measure(node);
			bool room = fits(
				(int) prog.code_.size() + node->need_insns_,
				(int) prog.inputs_.size() + node->need_inputs_,
				(int) prog.consts_.size() + node->need_consts_,
				prog.depth_ + node->need_depth_);
			if (room)
			{
				++stats_.fused_;
				emit_node(prog, node, view, rootshape);
				return;
			}
		}
		evaluate(node);
		llo_cuda_view in;
		in.data = node->data_.get();
		to_strides(in, view, node->shape_);
		size_t k = 0;
		for (; k < prog.inputs_.size(); ++k)
		{
			if (prog.inputs_[k].data == in.data &&
				prog.inputs_[k].offset == in.offset &&
				0 == std::memcmp(prog.inputs_[k].stride, in.stride,
					sizeof(in.stride)))
			{
				break;
			}
		}
		if (k == prog.inputs_.size())
		{
			prog.inputs_.push_back(in);
			prog.input_nodes_.push_back(node);
		}
		push_insn(prog, LLO_VM_LOAD, (uint8_t) k, 1);
	}

	void emit_node (Program& prog, PNode* n, const CoordView& outer,
		const ade::Shape& rootshape)
	{
		for (size_t i = 0; i < n->args_.size(); ++i)
		{
			emit_arg(prog, n->args_[i], outer, rootshape);
			if (i > 0 && is_nnary(n->opcode_))
			{
				push_insn(prog, (uint8_t) n->opcode_, 0, -1);
			}
		}
		if (is_unary(n->opcode_))
		{
			push_insn(prog, (uint8_t) n->opcode_, 0, 0);
		}
		else if (is_binary(n->opcode_))
		{
			push_insn(prog, (uint8_t) n->opcode_, 0, -1);
		}
	}

	/// LLO_PLAN_TRACE=1: one line per device step on stderr
	static bool tracing (void)
	{
		static const bool on = nullptr != std::getenv("LLO_PLAN_TRACE");
		return on;
	}

	void trace_program (const Program& prog, const PNode* n) const
	{
		std::string text = "program " + n->shape_.to_string() + " in=" +
			std::to_string(prog.inputs_.size()) + " :";
		for (const llo_cuda_insn& insn : prog.code_)
		{
			if (LLO_VM_LOAD == insn.op)
			{
				text += " L" + std::to_string((int) insn.arg);
			}
			else if (LLO_VM_CONST == insn.op)
			{
				text += " C(" + std::to_string(prog.consts_[insn.arg]) + ")";
			}
			else
			{
				text += " " + age::name_op((age::_GENERATED_OPCODE) insn.op);
			}
		}
		std::fprintf(stderr, "[plan] %s\n", text.c_str());
	}

	/// Storage for the result of n: the caller's destination when the node
	/// is a root that was given one, a fresh block otherwise
	static std::shared_ptr<char> out_buffer (PNode* n)
	{
		if (nullptr != n->dest_)
		{
			return n->dest_;
		}
		return device::alloc(n->shape_.n_elems() * age::type_size(n->dtype_));
	}

	void run_program (Program& prog, PNode* n)
	{
		if (tracing())
		{
			trace_program(prog, n);
		}
		n->data_ = out_buffer(n);
		uint64_t outshape[R];
		std::copy(n->shape_.begin(), n->shape_.end(), outshape);
		device::check(llo_cuda_elementwise(device::ctx(), (int) n->dtype_,
			n->data_.get(), outshape,
			prog.inputs_.data(), (int) prog.inputs_.size(),
			prog.consts_.data(), (int) prog.consts_.size(),
			prog.code_.data(), (int) prog.code_.size()));
		++stats_.programs_;
	}

	/// Materialise the value of a VIEW / CONST argument as a dense tensor
	std::shared_ptr<char> materialise (const PArg& arg, const ade::Shape& shape,
		DT dtype)
	{
		PNode holder;
		holder.kind_ = Kind::ELEM;
		holder.shape_ = shape;
		holder.dtype_ = dtype;
		Program prog;
		emit_arg(prog, arg, CoordView::identity(), shape);
		run_program(prog, &holder);
		return holder.data_;
	}

	void evaluate (PNode* n)
	{
		if (nullptr != n->data_)
		{
			return;
		}
		switch (n->kind_)
		{
			case Kind::LEAF:
				eval_leaf(n);
				break;
			case Kind::CONST:
			case Kind::VIEW:
			{
				PArg self;
				self.node_ = Kind::VIEW == n->kind_ ? n->args_[0].node_ : n;
				self.view_ = Kind::VIEW == n->kind_ ? n->args_[0].view_ :
					CoordView::identity();
				n->data_ = materialise(self, n->shape_, n->dtype_);
			}
				break;
			case Kind::ELEM:
			{
				Program prog;
				emit_node(prog, n, CoordView::identity(), n->shape_);
				run_program(prog, n);
			}
				break;
			case Kind::REDUCE:
				eval_reduce(n);
				break;
			case Kind::GENERIC:
				eval_generic(n);
				break;
		}
	}

	void eval_leaf (PNode* n)
	{
		if (auto var = dynamic_cast<Variable*>(n->leaf_))
		{
			n->data_ = var->device_data(n->dtype_);
			return;
		}
		GenericData staged(n->shape_, n->dtype_);
		staged.copyover((const char*) ((const ade::iLeaf*) n->leaf_)->data(),
			(DT) n->leaf_->type_code());
		n->data_ = staged.device_;
	}

	void eval_generic (PNode* n)
	{
		DataArgsT argdata;
		for (PArg& a : n->args_)
		{
			evaluate(a.node_);
			argdata.push_back(DataArg{a.node_->data_, a.node_->shape_,
				a.coorder_, a.push_});
		}
		n->data_ = out_buffer(n);
		age::op_exec((age::_GENERATED_OPCODE) n->opcode_, n->dtype_,
			n->data_.get(), n->shape_, argdata);
	}

	void eval_reduce (PNode* n)
	{
		if (age::SUM == n->opcode_ && try_gemm(n))
		{
			if (tracing())
			{
				std::fprintf(stderr, "[plan] gemm -> %s\n", n->shape_.to_string().c_str());
			}
			return;
		}
		if (tracing())
		{
			std::fprintf(stderr, "[plan] reduce %s %s -> %s\n",
				age::name_op((age::_GENERATED_OPCODE) n->opcode_).c_str(),
				n->argshape_.to_string().c_str(), n->shape_.to_string().c_str());
		}
		PArg& src = n->args_[0];
		std::shared_ptr<char> data;
		CoordView through = src.view_;
		if (Kind::CONST == src.node_->kind_ || inlinable(src.node_))
		{
			// the fold reads a dense copy of its argument
			data = materialise(src, n->argshape_, n->dtype_);
			through = CoordView::identity();
			normalise(through, n->argshape_, n->argshape_);
			reduce_from(n, data.get(), through, n->argshape_);
			return;
		}
		evaluate(src.node_);
		reduce_from(n, src.node_->data_.get(), through, src.node_->shape_);
	}

	/// Fold `data` (a tensor of shape srcshape reached from the argument's
	/// coordinates through `through`) into n
	void reduce_from (PNode* n, const char* data, const CoordView& through,
		const ade::Shape& srcshape)
	{
		llo_cuda_view whole;   // argument coords -> source memory
		to_strides(whole, through, srcshape);
		CoordView kept = compose(n->kept_, through);
		llo_cuda_view src;     // output coords -> source memory (folded coords 0)
		to_strides(src, kept, srcshape);
		src.data = data;
		uint64_t fold_size[R];
		int64_t fold_stride[R];
		int nfold = 0;
		for (int dim : n->fold_)
		{
			fold_size[nfold] = n->argshape_.at(dim);
			fold_stride[nfold] = whole.stride[dim];
			++nfold;
		}
		uint64_t outshape[R];
		std::copy(n->shape_.begin(), n->shape_.end(), outshape);
		n->data_ = out_buffer(n);
		int mode = 0 != get_option("sequential_reduce") ?
			LLO_REDUCE_SEQUENTIAL : LLO_REDUCE_TREE;
		device::check(llo_cuda_reduce_view(device::ctx(), n->opcode_,
			(int) n->dtype_, n->data_.get(), outshape, &src, nfold,
			fold_size, fold_stride, 0, mode));
		++stats_.reduces_;
	}

	/// reduce_sum(mul(view(A), view(B))) over whole dims is a contraction:
	/// lower it to one GEMM when its dims split cleanly into M (A only),
	/// N (B only) and K (folded, both) groups
	bool try_gemm (PNode* n)
	{
		PNode* prod = n->args_[0].node_;
		if (Kind::ELEM != prod->kind_ || age::PROD != prod->opcode_ ||
			2 != prod->args_.size() || false == inlinable(prod))
		{
			return false;
		}
		const CoordView& through = n->args_[0].view_;
		PArg fa = prod->args_[0];
		PArg fb = prod->args_[1];
		if (Kind::CONST == fa.node_->kind_ || Kind::CONST == fb.node_->kind_)
		{
			return false;
		}
		// argument coordinates -> operand memory
		auto operand_strides = [&](llo_cuda_view& out, const PArg& f)
		{
			CoordView full = compose(through, f.view_);
			normalise(full, n->argshape_, f.node_->shape_);
			to_strides(out, full, f.node_->shape_);
		};
		llo_cuda_view sa, sb;
		operand_strides(sa, fa);
		operand_strides(sb, fb);
		// argument dim -> output stride (kept dims only)
		int64_t sc[R];
		bool folded[R];
		std::fill(sc, sc + R, 0);
		std::fill(folded, folded + R, false);
		for (int dim : n->fold_)
		{
			folded[dim] = true;
		}
		int64_t dense = 1;
		for (int j = 0; j < R; ++j)
		{
			// kept_: arg dim i = scale * o_j + offset
			for (int i = 0; i < R; ++i)
			{
				const DimMap& m = n->kept_.t_[i];
				if (m.src() == j && n->shape_.at(j) > 1)
				{
					if (1 != m.coef_[j] || 0 != m.offset_)
					{
						return false;
					}
					sc[i] = dense;
				}
			}
			dense *= (int64_t) n->shape_.at(j);
		}
		// dims of the argument space by role: M (A only), N (B only), K (folded,
		// both); sizes and the strides each tensor has along them
		struct Role
		{
			std::vector<int> dims_;
			uint64_t size_ = 1;
		};
		Role rm, rn, rk;
		for (int i = 0; i < R; ++i)
		{
			uint64_t size = n->argshape_.at(i);
			if (size <= 1)
			{
				continue;
			}
			int64_t a = sa.stride[i];
			int64_t b = sb.stride[i];
			Role* role = nullptr;
			if (folded[i])
			{
				role = (0 != a && 0 != b) ? &rk : nullptr;
			}
			else if (0 != a && 0 == b)
			{
				role = &rm;
			}
			else if (0 == a && 0 != b)
			{
				role = &rn;
			}
			if (nullptr == role)
			{
				return false;
			}
			role->dims_.push_back(i);
			role->size_ *= size;
		}
		if (rk.dims_.empty())
		{
			return false;
		}
		// a role's dims merge into ONE stride of a tensor when each dim's
		// stride is the previous one's times its size
		auto merged = [&](const Role& role, const int64_t* stride, int64_t& out) -> bool
		{
			out = 0;
			uint64_t run = 1;
			for (size_t q = 0; q < role.dims_.size(); ++q)
			{
				int i = role.dims_[q];
				if (0 == q)
				{
					out = stride[i];
				}
				else if (stride[i] != out * (int64_t) run)
				{
					return false;
				}
				run *= n->argshape_.at(i);
			}
			return true;
		};
		int64_t c_m = 0, c_n = 0;
		if (false == merged(rm, sc, c_m) || false == merged(rn, sc, c_n))
		{
			return false; // the output itself is not a matrix of the two roles
		}
		evaluate(fa.node_);
		evaluate(fb.node_);
		size_t esize = age::type_size(n->dtype_);
		// An operand whose dims do not merge -- the image of a convolution: the
		// window offsets kh, kw move along the same axes as oh, ow
		// (llo/src/helper.cpp:149-165) -- is gathered once into a dense
		// [rows][k] matrix (im2col) by the elementwise engine; everything else
		// is multiplied where it lies.
		struct Operand
		{
			const char* base_ = nullptr;
			int64_t row_ = 0;   // stride along its M / N role
			int64_t k_ = 0;     // stride along K
			std::shared_ptr<char> dense_;
		};
		auto operand = [&](Operand& out, const PArg& f, const llo_cuda_view& strides,
			const Role& rows) -> bool
		{
			if (merged(rows, strides.stride, out.row_) && merged(rk, strides.stride, out.k_))
			{
				out.base_ = f.node_->data_.get() + strides.offset * (int64_t) esize;
				return true;
			}
			if (rk.dims_.size() + rows.dims_.size() > (size_t) R)
			{
				return false;
			}
			// dense copy with coordinates (k dims..., row dims...): k fastest
			std::vector<ade::DimT> dims;
			CoordView pick;   // copy coords -> argument-space coords
			int pos = 0;
			const Role* order[2] = {&rk, &rows};
			for (const Role* role : order)
			{
				for (int i : role->dims_)
				{
					dims.push_back(n->argshape_.at(i));
					pick.t_[i].coef_[pos] = 1;
					++pos;
				}
			}
			ade::Shape dense_shape(dims);
			PArg gathered = f;
			gathered.view_ = compose(pick, compose(through, f.view_));
			normalise(gathered.view_, dense_shape, f.node_->shape_);
			out.dense_ = materialise(gathered, dense_shape, n->dtype_);
			out.base_ = out.dense_.get();
			out.k_ = 1;
			out.row_ = (int64_t) rk.size_;
			++stats_.gathers_;
			return true;
		};
		Operand oa, ob;
		if (false == operand(oa, fa, sa, rm) || false == operand(ob, fb, sb, rn))
		{
			return false;
		}
		n->data_ = out_buffer(n);
		int precision = (age::FLOAT == n->dtype_ && 0 != get_option("tf32x3")) ?
			LLO_GEMM_3XTF32 : LLO_GEMM_EXACT;
		device::check(llo_cuda_gemm(device::ctx(), (int) n->dtype_,
			rm.size_, rn.size_, rk.size_,
			oa.base_, oa.row_, oa.k_, ob.base_, ob.k_, ob.row_,
			n->data_.get(), c_m, c_n, precision));
		++stats_.gemms_;
		++stats_.fused_;
		return true;
	}

	DT dtype_;
	bool simplify_;
	int recompute_;
	PlanStats stats_;
	std::vector<std::unique_ptr<PNode>> arena_;
	std::map<std::pair<ade::iTensor*,int>,PNode*> memo_;
	std::unordered_map<std::string,PNode*> interned_;
};

static thread_local PlanStats last_stats;

}

std::vector<GenericData> planned_eval_many (const ade::TensT& roots,
	age::_GENERATED_DTYPE dtype, const std::vector<std::shared_ptr<char>>* dests,
	const std::function<void(size_t)>& on_done)
{
	// one plan for every root: sub-graphs the roots share (the forward pass
	// under a set of gradients) are computed once
	Plan plan(dtype);
	std::vector<PNode*> tops;
	for (const ade::TensptrT& root : roots)
	{
		tops.push_back(plan.build(root.get(), dtype));
	}
	for (PNode* top : tops)
	{
		plan.count_uses(top);
	}
	if (nullptr != dests)
	{
		if (dests->size() != roots.size())
		{
			logs::fatalf("cannot evaluate %d roots into %d destinations",
				(int) roots.size(), (int) dests->size());
		}
		// a root that computes (elementwise program, reduction, GEMM, generic
		// functor) writes straight into its destination; leaves, views,
		// constants and roots that share a node are copied there afterwards
		for (size_t i = 0; i < tops.size(); ++i)
		{
			PNode* top = tops[i];
			bool computes = Kind::ELEM == top->kind_ || Kind::REDUCE == top->kind_ ||
				Kind::GENERIC == top->kind_;
			if (computes && nullptr == top->dest_ && nullptr == top->data_ &&
				nullptr != (*dests)[i])
			{
				top->dest_ = (*dests)[i];
			}
		}
	}
	std::vector<GenericData> outs;
	for (size_t i = 0; i < tops.size(); ++i)
	{
		PNode* top = tops[i];
		plan.evaluate(top);
		GenericData out;
		out.shape_ = top->shape_;
		out.dtype_ = dtype;
		out.device_ = top->data_;
		if (nullptr != dests && nullptr != (*dests)[i] &&
			top->data_.get() != (*dests)[i].get())
		{
			device::check(llo_cuda_d2d(device::ctx(), (*dests)[i].get(),
				top->data_.get(), out.nbytes()));
			out.device_ = (*dests)[i];
		}
		outs.push_back(out);
		if (on_done)
		{
			on_done(i);
		}
	}
	last_stats = plan.stats_;
	return outs;
}

PlanStats plan_preview (const ade::TensT& roots, age::_GENERATED_DTYPE dtype)
{
	Plan plan(dtype);
	for (const ade::TensptrT& root : roots)
	{
		plan.build(root.get(), dtype);
	}
	return plan.stats_;
}

GenericData planned_eval (ade::TensptrT root, age::_GENERATED_DTYPE dtype)
{
	return planned_eval_many(ade::TensT{root}, dtype, nullptr, nullptr)[0];
}

const PlanStats& last_plan_stats (void)
{
	return last_stats;
}

}

#endif
