///
/// module.cpp
///
/// Python surface of the evaluator, kept identical to the reference's two
/// pybind modules:
///   `age`  — one graph builder per op-table entry, the second overload of a
///            name gets the suffix 0 (reduce_sum / reduce_sum0 ...), class
///            Tensor                      (pybinder/pyapi_plugin.py:19-54)
///   `llo`  — variable, Variable.assign, evaluate, derive, seed, print_graph,
///            Tensor.__str__/shape()/children()   (llo/python/llo.cpp:156-215)
/// numpy shapes are the reverse of ade shapes (llo/python/llo.cpp:18-34).
///
/// Threading: like the reference's binding (which never releases the GIL,
/// llo/python/llo.cpp) every call here holds the GIL for its whole duration.
/// The evaluator context, the planner and Variable's device mirrors are
/// single-threaded by contract (SURVEY.md 8b "Threading"); only comm_init,
/// which blocks on the other ranks' processes, releases it.
///
/// Additions: evaluate() honours the exact numpy dtype
/// (the reference only distinguishes 'f' -> DOUBLE and 'i' -> INT64, which
/// the float64 / int64 defaults still select), unsigned dtypes are accepted,
/// and evaluate_device() leaves the result in HBM.
///

#include "pybind11/pybind11.h"
#include "pybind11/numpy.h"
#include "pybind11/stl.h"

#include "llo/llo.hpp"
#include "llo/session.hpp"
#include "llo/plan.hpp"

#include "dbg/ade.hpp"

#include "llo/serialize.hpp"
#include "pbm/pbm.hpp"

namespace py = pybind11;

namespace pyllo
{

static ade::Shape p2cshape (const std::vector<py::ssize_t>& pyshape)
{
	return ade::Shape(std::vector<ade::DimT>(
		pyshape.rbegin(), pyshape.rend()));
}

static std::vector<py::ssize_t> c2pshape (const ade::Shape& cshape)
{
	auto it = cshape.begin();
	auto et = cshape.end();
	while (it != et && *(et - 1) == 1)
	{
		--et;
	}
	std::vector<py::ssize_t> fwd(it, et);
	return std::vector<py::ssize_t>(fwd.rbegin(), fwd.rend());
}

static age::_GENERATED_DTYPE exact_dtype (py::dtype dtype, const char* what)
{
	char kind = dtype.kind();
	py::ssize_t tbytes = dtype.itemsize();
	switch (kind)
	{
		case 'f':
			switch (tbytes)
			{
				case 4: return age::FLOAT;
				case 8: return age::DOUBLE;
				default:
					logs::fatalf("unsupported float type with %d bytes",
						(int) tbytes);
			}
		case 'i':
			switch (tbytes)
			{
				case 1: return age::INT8;
				case 2: return age::INT16;
				case 4: return age::INT32;
				case 8: return age::INT64;
				default:
					logs::fatalf("unsupported integer type with %d bytes",
						(int) tbytes);
			}
		case 'u':
			switch (tbytes)
			{
				case 1: return age::UINT8;
				case 2: return age::UINT16;
				case 4: return age::UINT32;
				case 8: return age::UINT64;
				default:
					logs::fatalf("unsupported integer type with %d bytes",
						(int) tbytes);
			}
		default:
			logs::fatalf("unknown dtype %c", kind);
	}
	(void) what;
	return age::BAD_TYPE;
}


// ---- zero-copy device interop --------------------------------
//
// Results leave as __cuda_array_interface__ (version 3) and DLPack capsules;
// leaves accept any object that offers __cuda_array_interface__ (a torch / cupy
// CUDA tensor) and copy device-to-device.  The reference's binding only knows
// host numpy arrays (llo/python/llo.cpp:36-147).

static const char* typestr_of (age::_GENERATED_DTYPE dtype)
{
	switch (dtype)
	{
		case age::DOUBLE: return "<f8";
		case age::FLOAT: return "<f4";
		case age::INT8: return "|i1";
		case age::UINT8: return "|u1";
		case age::INT16: return "<i2";
		case age::UINT16: return "<u2";
		case age::INT32: return "<i4";
		case age::UINT32: return "<u4";
		case age::INT64: return "<i8";
		case age::UINT64: return "<u8";
		default: logs::fatal("cannot describe bad type");
	}
}

static age::_GENERATED_DTYPE dtype_of_typestr (const std::string& ts)
{
	if (ts.size() < 3 || '>' == ts[0])
	{
		logs::fatalf("unsupported device array type %s", ts.c_str());
	}
	std::string kind = ts.substr(1);
	if ("f8" == kind) return age::DOUBLE;
	if ("f4" == kind) return age::FLOAT;
	if ("i1" == kind) return age::INT8;
	if ("u1" == kind) return age::UINT8;
	if ("i2" == kind) return age::INT16;
	if ("u2" == kind) return age::UINT16;
	if ("i4" == kind) return age::INT32;
	if ("u4" == kind) return age::UINT32;
	if ("i8" == kind) return age::INT64;
	if ("u8" == kind) return age::UINT64;
	logs::fatalf("unsupported device array type %s", ts.c_str());
}

static py::dict cuda_array_interface (const std::shared_ptr<char>& data,
	const std::vector<py::ssize_t>& pshape, age::_GENERATED_DTYPE dtype)
{
	py::dict out;
	out["shape"] = py::tuple(py::cast(pshape));
	out["typestr"] = typestr_of(dtype);
	out["data"] = py::make_tuple((uintptr_t) data.get(), false);
	out["version"] = 3;
	out["strides"] = py::none();
	// consumers order themselves behind the evaluator's stream (1 would be
	// the legacy default stream; a real handle is passed as an integer)
	uintptr_t stream = (uintptr_t) llo_cuda_get_stream(llo::device::ctx());
	out["stream"] = 0 == stream ? py::object(py::int_(1)) : py::object(py::int_(stream));
	return out;
}

// the slice of dlpack.h (v0.8) this module needs
struct DLDevice { int32_t device_type; int32_t device_id; };
struct DLDataType { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DLTensor
{
	void* data;
	DLDevice device;
	int32_t ndim;
	DLDataType dtype;
	int64_t* shape;
	int64_t* strides;
	uint64_t byte_offset;
};
struct DLManagedTensor
{
	DLTensor dl_tensor;
	void* manager_ctx;
	void (*deleter) (DLManagedTensor*);
};

struct DLOwner
{
	std::shared_ptr<char> data_;
	std::vector<int64_t> shape_;
	DLManagedTensor managed_;
};

static py::capsule dlpack_of (const std::shared_ptr<char>& data,
	const std::vector<py::ssize_t>& pshape, age::_GENERATED_DTYPE dtype)
{
	// the consumer may use the memory on any stream: finish ours first
	llo::device::check(llo_cuda_sync(llo::device::ctx()));
	DLOwner* owner = new DLOwner();
	owner->data_ = data;
	owner->shape_.assign(pshape.begin(), pshape.end());
	DLTensor& t = owner->managed_.dl_tensor;
	t.data = data.get();
	t.device = DLDevice{2 /* kDLCUDA */, llo_cuda_device(llo::device::ctx())};
	t.ndim = (int32_t) owner->shape_.size();
	bool fp = age::DOUBLE == dtype || age::FLOAT == dtype;
	bool uns = age::UINT8 == dtype || age::UINT16 == dtype ||
		age::UINT32 == dtype || age::UINT64 == dtype;
	t.dtype = DLDataType{(uint8_t) (fp ? 2 : (uns ? 1 : 0)),
		(uint8_t) (8 * age::type_size(dtype)), 1};
	t.shape = owner->shape_.data();
	t.strides = nullptr;
	t.byte_offset = 0;
	owner->managed_.manager_ctx = owner;
	owner->managed_.deleter = [](DLManagedTensor* self)
	{
		delete static_cast<DLOwner*>(self->manager_ctx);
	};
	return py::capsule(&owner->managed_, "dltensor", [](PyObject* cap)
	{
		// still named "dltensor": nobody consumed it, so it is ours to free
		if (PyCapsule_IsValid(cap, "dltensor"))
		{
			auto managed = static_cast<DLManagedTensor*>(
				PyCapsule_GetPointer(cap, "dltensor"));
			managed->deleter(managed);
		}
	});
}

/// A device array offered through __cuda_array_interface__
struct DeviceArrayView
{
	const void* data_ = nullptr;
	ade::Shape shape_;
	age::_GENERATED_DTYPE dtype_ = age::BAD_TYPE;
};

static bool is_device_array (const py::object& obj)
{
	return py::hasattr(obj, "__cuda_array_interface__");
}

static DeviceArrayView device_array (const py::object& obj)
{
	py::dict cai = obj.attr("__cuda_array_interface__");
	DeviceArrayView out;
	std::vector<py::ssize_t> pshape = cai["shape"].cast<std::vector<py::ssize_t>>();
	out.dtype_ = dtype_of_typestr(cai["typestr"].cast<std::string>());
	if (cai.contains("strides") && false == cai["strides"].is_none())
	{
		// only dense C order is accepted (what .contiguous() gives)
		std::vector<py::ssize_t> strides = cai["strides"].cast<std::vector<py::ssize_t>>();
		py::ssize_t expect = (py::ssize_t) age::type_size(out.dtype_);
		for (size_t i = pshape.size(); i-- > 0;)
		{
			if (pshape[i] > 1 && strides[i] != expect)
			{
				logs::fatal("device arrays must be C-contiguous");
			}
			expect *= pshape[i];
		}
	}
	out.shape_ = p2cshape(pshape);
	py::tuple data = cai["data"];
	out.data_ = (const void*) data[0].cast<uintptr_t>();
	// order our stream behind the producer's
	if (cai.contains("stream") && false == cai["stream"].is_none())
	{
		uintptr_t stream = cai["stream"].cast<uintptr_t>();
		llo::device::check(llo_cuda_wait_stream(llo::device::ctx(),
			(void*) (stream <= 2 ? (uintptr_t) (2 == stream ? 2 : 0) : stream)));
	}
	return out;
}

static llo::VarptrT variable_from_device (py::object obj, std::string label)
{
	DeviceArrayView view = device_array(obj);
	llo::VarptrT out(new llo::Variable(nullptr, view.dtype_, view.shape_, label));
	out->assign_device(view.data_, view.shape_, view.dtype_);
	return out;
}

static llo::VarptrT variable (py::array data, std::string label)
{
	py::array dense = py::array::ensure(data, py::array::c_style);
	py::buffer_info info = dense.request();
	ade::Shape shape = p2cshape(info.shape);
	age::_GENERATED_DTYPE dtype = exact_dtype(dense.dtype(), "variable");
	if ((size_t) info.size != shape.n_elems())
	{
		logs::fatalf("cannot create variable with data size %d "
			"against shape %s", (int) info.size, shape.to_string().c_str());
	}
	return llo::VarptrT(new llo::Variable(
		(const char*) info.ptr, dtype, shape, label));
}

// reference: llo/python/llo.cpp:101-124 reads every 'f' array as double and
// every 'i' array as int64; here the array is cast to the variable's own
// type first, which is the same thing whenever the reference's read is valid
static void assign (llo::Variable* target, py::array data)
{
	age::_GENERATED_DTYPE want =
		(age::_GENERATED_DTYPE) target->type_code();
	age::_GENERATED_DTYPE have = exact_dtype(data.dtype(), "assign");
	py::array dense = py::array::ensure(data, py::array::c_style);
	py::buffer_info info = dense.request();
	ade::Shape shape = p2cshape(info.shape);
	if (have != want)
	{
		logs::fatalf("cannot assign data of incompatible types %s "
			"(external) and %s (internal)",
			age::name_type(have).c_str(), age::name_type(want).c_str());
	}
	*target = llo::GenericRef((char*) info.ptr, shape, want);
}

/// assign, on the upload lane (the array must be pinned: llo.pinned_empty)
static void assign_async (llo::Variable* target, py::array data)
{
	age::_GENERATED_DTYPE want =
		(age::_GENERATED_DTYPE) target->type_code();
	age::_GENERATED_DTYPE have = exact_dtype(data.dtype(), "assign");
	if (0 == (data.flags() & py::array::c_style))
	{
		logs::fatal("assign_async needs a C-contiguous (pinned) array");
	}
	py::buffer_info info = data.request();
	ade::Shape shape = p2cshape(info.shape);
	if (have != want)
	{
		logs::fatalf("cannot assign data of incompatible types %s "
			"(external) and %s (internal)",
			age::name_type(have).c_str(), age::name_type(want).c_str());
	}
	target->assign_async(llo::GenericRef((char*) info.ptr, shape, want));
}

static age::_GENERATED_DTYPE eval_dtype (py::dtype dtype)
{
	return exact_dtype(dtype, "evaluate");
}

/// numpy array over a host block the array keeps alive: the result's pinned
/// mirror is handed to numpy as is (the reference lets numpy copy the buffer,
/// llo/python/llo.cpp:143-146; the array is the block's only owner either way)
static py::array wrap_host (py::dtype dtype, std::vector<py::ssize_t> pshape,
	std::shared_ptr<char> block)
{
	auto holder = new std::shared_ptr<char>(block);
	py::capsule base(holder, [](void* p)
		{ delete static_cast<std::shared_ptr<char>*>(p); });
	return py::array(dtype, pshape, block.get(), base);
}

static py::array evaluate (ade::TensptrT tens, py::dtype dtype)
{
	age::_GENERATED_DTYPE ctype = eval_dtype(dtype);
	llo::GenericData gdata = llo::eval(tens, ctype);
	auto pshape = c2pshape(gdata.shape_);
	return wrap_host(dtype, pshape, gdata.data_);
}

/// Result whose device->host copy is in flight (evaluate_async)
struct PendingResult
{
	llo::GenericData data_;
	py::dtype dtype_;
};

/// Device-resident result handle
struct DeviceResult
{
	llo::GenericData data_;
};

/// Flat device buffer: the packed gradient vector of the batch-sharded
/// executor (one allreduce per gradient evaluation instead of one per tensor)
struct DeviceBuffer
{
	DeviceBuffer (size_t nbytes) :
		nbytes_(nbytes), data_(llo::device::alloc(nbytes)) {}

	DeviceBuffer (std::shared_ptr<char> data, size_t nbytes) :
		nbytes_(nbytes), data_(data) {}

	size_t nbytes_;

	std::shared_ptr<char> data_;
};

}

PYBIND11_MODULE(_C, m)
{
	m.doc() = "cortenn_b200: B200-native llo evaluator";

	py::module_ age_m = m.def_submodule("age", "pybind ade generator");
	py::module_ llo_m = m.def_submodule("llo", "llo variables");

	py::class_<ade::iTensor,ade::TensptrT> tensor(age_m, "Tensor");
	tensor
		.def("__str__", &ade::iTensor::to_string)
		.def("shape", [](ade::TensptrT self)
		{
			auto pshape = pyllo::c2pshape(self->shape());
			std::vector<int> ipshape(pshape.begin(), pshape.end());
			return py::array(ipshape.size(), ipshape.data());
		})
		.def("children", [](ade::TensptrT self)
		{
			std::vector<ade::TensptrT> tens;
			if (auto f = dynamic_cast<ade::iFunctor*>(self.get()))
			{
				for (const ade::MappedTensor& arg : f->get_children())
				{
					tens.push_back(arg.get_tensor());
				}
			}
			return tens;
		})
		.def("opname", [](ade::TensptrT self) -> std::string
		{
			if (auto f = dynamic_cast<ade::iFunctor*>(self.get()))
			{
				return f->get_opcode().name_;
			}
			return "";
		});

	py::class_<llo::Variable,ade::iTensor,llo::VarptrT> variable(
		llo_m, "Variable");
	variable
		.def("assign", [](llo::VarptrT self, py::object data)
		{
			if (pyllo::is_device_array(data))
			{
				pyllo::DeviceArrayView view = pyllo::device_array(data);
				self->assign_device(view.data_, view.shape_, view.dtype_);
				return;
			}
			pyllo::assign(self.get(), py::array::ensure(data));
		}, "assign to variable (numpy array, or a device array: copied on the GPU)")
		.def("assign_async", [](llo::VarptrT self, py::array data)
		{
			pyllo::assign_async(self.get(), data);
		}, "assign from a pinned array without waiting: the upload runs on "
			"the copy lane, evaluations queued afterwards see the new values")
		.def_readwrite("label", &llo::Variable::label_);

	// ---- age: graph builders (llo/cfg/llo.json:125-531) ----
	using T = ade::TensptrT;
#define AGE_UNARY(NAME) age_m.def(#NAME, [](T a) { return age::NAME(a); },\
	"ade::TensptrT " #NAME " (ade::TensptrT arg1)");
	AGE_UNARY(abs) AGE_UNARY(neg) AGE_UNARY(sin) AGE_UNARY(cos) AGE_UNARY(tan)
	AGE_UNARY(exp) AGE_UNARY(log) AGE_UNARY(sqrt) AGE_UNARY(round)
#undef AGE_UNARY
	age_m.def("flip", [](T a, uint8_t dim) { return age::flip(a, dim); });
#define AGE_BINARY(NAME) age_m.def(#NAME, [](T a, T b) { return age::NAME(a, b); },\
	"ade::TensptrT " #NAME " (ade::TensptrT arg1, ade::TensptrT arg2)");
	AGE_BINARY(pow) AGE_BINARY(add) AGE_BINARY(sub) AGE_BINARY(mul)
	AGE_BINARY(div) AGE_BINARY(eq) AGE_BINARY(neq) AGE_BINARY(lt) AGE_BINARY(gt)
	AGE_BINARY(rand_bino) AGE_BINARY(rand_unif) AGE_BINARY(rand_norm)
#undef AGE_BINARY
	age_m.def("n_elems", [](T a) { return age::n_elems(a); });
	age_m.def("n_dims", [](T a, uint8_t rank) { return age::n_dims(a, rank); });
#define AGE_NARY(NAME) age_m.def(#NAME, [](ade::TensT a) { return age::NAME(a); });
	AGE_NARY(sum) AGE_NARY(prod) AGE_NARY(min) AGE_NARY(max)
#undef AGE_NARY
#define AGE_REDUCE(NAME)\
	age_m.def(#NAME, [](T a, uint8_t dim) { return age::NAME(a, dim); });\
	age_m.def(#NAME "0", [](T a) { return age::NAME(a); });
	AGE_REDUCE(reduce_sum) AGE_REDUCE(reduce_prod)
	AGE_REDUCE(reduce_min) AGE_REDUCE(reduce_max)
#undef AGE_REDUCE
	age_m.def("permute", [](T a, std::vector<uint8_t> order)
		{ return age::permute(a, order); });
	age_m.def("extend", [](T a, uint8_t rank, std::vector<ade::DimT> ext)
		{ return age::extend(a, rank, ext); });
	age_m.def("transpose", [](T a) { return age::transpose(a); });
	age_m.def("reduce_mean", [](T a) { return age::reduce_mean(a); });
	age_m.def("matmul", [](T a, T b) { return age::matmul(a, b); });
	age_m.def("convolution", [](T a, T b) { return age::convolution(a, b); });

	// raw functor construction with explicit 9x9 maps, for tests that
	// exercise arbitrary CoordMaps (llo/test/test_operator.cpp:29-38)
	age_m.def("functor", [](std::string opname, std::vector<T> args,
		std::vector<py::array_t<double>> shapers,
		std::vector<py::array_t<double>> coorders,
		std::vector<bool> pushes)
	{
		age::_GENERATED_OPCODE code = age::get_op(opname);
		if (age::BAD_OP == code)
		{
			logs::fatalf("unknown opcode %s", opname.c_str());
		}
		auto to_map = [](py::array_t<double>& mat) -> ade::CoordptrT
		{
			auto r = mat.unchecked<2>();
			if (r.shape(0) != ade::mat_dim || r.shape(1) != ade::mat_dim)
			{
				logs::fatal("coordinate maps are 9x9");
			}
			return std::make_shared<ade::CoordMap>([&](ade::MatrixT m)
			{
				for (int i = 0; i < ade::mat_dim; ++i)
				{
					for (int j = 0; j < ade::mat_dim; ++j)
					{
						m[i][j] = r(i, j);
					}
				}
			});
		};
		ade::ArgsT margs;
		for (size_t i = 0; i < args.size(); ++i)
		{
			margs.push_back(ade::MappedTensor(args[i], to_map(shapers[i]),
				pushes[i], to_map(coorders[i])));
		}
		return T(ade::Functor::get(
			ade::Opcode{opname, (size_t) code}, margs));
	});
	// functor over extend-mapped arguments (llo/test/test_api.cpp:1024-1027)
	age_m.def("functor_extended", [](std::string opname, std::vector<T> args,
		uint8_t rank, std::vector<ade::DimT> ext)
	{
		age::_GENERATED_OPCODE code = age::get_op(opname);
		ade::ArgsT margs;
		for (T& a : args)
		{
			margs.push_back(ade::extend_map(a, rank, ext));
		}
		return T(ade::Functor::get(
			ade::Opcode{opname, (size_t) code}, margs));
	});

	// ---- llo ----
	llo_m.def("variable", [](py::object data, std::string label)
		{
			if (pyllo::is_device_array(data))
			{
				// a CUDA tensor of another library: device-to-device, no host bounce
				return pyllo::variable_from_device(data, label);
			}
			return pyllo::variable(py::array::ensure(data), label);
		}, "create tensor variable from a numpy array or from any object with "
		"__cuda_array_interface__ (torch / cupy CUDA tensors)",
		py::arg("data"), py::arg("label") = "");
	llo_m.def("scalar", [](double value, std::vector<py::ssize_t> shape,
		std::string label)
	{
		return llo::get_scalar<double>(value, pyllo::p2cshape(shape), label);
	}, py::arg("value"), py::arg("shape"), py::arg("label") = "");
	llo_m.def("evaluate", &pyllo::evaluate,
		"evaluate data of tens according to dtype",
		py::arg("tens"), py::arg("dtype") = py::dtype::of<double>());
	llo_m.def("derive", [](T root, T target)
		{ return llo::derive(root, target.get()); },
		"derive tensor with respect to some derive");
	// graph -> graph optimisation (llo/optimize.hpp); explicit, never implied
	llo_m.attr("OPT_ONE_MUL") = (unsigned) llo::OPT_ONE_MUL;
	llo_m.attr("OPT_ZERO_ADD") = (unsigned) llo::OPT_ZERO_ADD;
	llo_m.attr("OPT_POW_ONE") = (unsigned) llo::OPT_POW_ONE;
	llo_m.attr("OPT_SINGLE") = (unsigned) llo::OPT_SINGLE;
	llo_m.attr("OPT_CSE") = (unsigned) llo::OPT_CSE;
	llo_m.attr("OPT_DIV_CANCEL") = (unsigned) llo::OPT_DIV_CANCEL;
	llo_m.attr("OPT_EXACT") = (unsigned) llo::OPT_EXACT;
	llo_m.attr("OPT_ALL") = (unsigned) llo::OPT_ALL;
	llo_m.def("optimize", [](std::vector<T> roots, unsigned passes)
		{
			llo::OptReport report;
			ade::TensT out = llo::optimize(roots, passes, &report);
			py::dict rep;
			rep["functors_before"] = report.functors_before_;
			rep["functors_after"] = report.functors_after_;
			rep["one_mul"] = report.one_mul_;
			rep["zero_add"] = report.zero_add_;
			rep["pow_one"] = report.pow_one_;
			rep["single"] = report.single_;
			rep["shared"] = report.shared_;
			rep["div_cancel"] = report.div_cancel_;
			return py::make_tuple(out, rep);
		}, "rewrite the graphs under `roots` with the selected passes "
		"(OPT_EXACT: value-preserving; OPT_DIV_CANCEL: (x*y)/x -> y, inexact, "
		"floating point only); returns (new roots, report)",
		py::arg("roots"), py::arg("passes") = (unsigned) llo::OPT_ALL);
	llo_m.def("seed", [](size_t seed)
		{
			llo::device::check(llo_cuda_seed(llo::device::ctx(), seed));
		}, "seed internal rng");
	llo_m.def("print_graph", [](T root, bool showshape)
		{
			PrettyEquation peq;
			peq.showshape_ = showshape;
			std::stringstream ss;
			peq.print(ss, root);
			py::print(ss.str(), py::arg("end") = "");
		}, "print graph of root tensor",
		py::arg("root"), py::arg("showshape") = false);
	llo_m.def("graph_string", [](T root, bool showshape)
		{
			PrettyEquation peq;
			peq.showshape_ = showshape;
			std::stringstream ss;
			peq.print(ss, root);
			return ss.str();
		}, py::arg("root"), py::arg("showshape") = false);

	// ---- pbm: marshal / unmarshal graphs (pbm/save.hpp, pbm/load.hpp) ----
	llo_m.def("save_graph", [](std::vector<T> roots,
		std::vector<std::pair<T,std::vector<std::string>>> labels, std::string graph_label)
		{
			pbm::GraphSaver saver(llo::serialize);
			for (T& root : roots)
			{
				root->accept(saver);
			}
			pbm::PathedMapT paths;
			for (auto& entry : labels)
			{
				paths[entry.first] = pbm::StringsT(entry.second.begin(), entry.second.end());
			}
			pbm::Graph graph;
			graph.label_ = graph_label;
			saver.save(graph, paths);
			return py::bytes(pbm::encode(graph));
		}, "marshal the graphs under `roots` to the bytes of a cortenn.Graph message "
		"(pbm/graph.proto); `labels` = [(tensor, [path, ...]), ...]",
		py::arg("roots"), py::arg("labels") = std::vector<std::pair<T,std::vector<std::string>>>(),
		py::arg("label") = "");
	llo_m.def("load_graph", [](py::bytes data)
		{
			pbm::Graph graph;
			pbm::decode(graph, std::string(data));
			pbm::GraphInfo info;
			pbm::load_graph(info, graph, llo::deserialize);
			py::dict out;
			out["label"] = graph.label_;
			out["nodes"] = info.nodes_;
			std::vector<T> roots;
			for (const T& node : info.nodes_)
			{
				if (info.roots_.end() != info.roots_.find(node))
				{
					roots.push_back(node);
				}
			}
			out["roots"] = roots;
			py::list labelled;
			for (size_t i = 0; i < graph.nodes_.size(); ++i)
			{
				if (false == graph.nodes_[i].labels_.empty())
				{
					labelled.append(py::make_tuple(graph.nodes_[i].labels_, info.nodes_[i]));
				}
			}
			out["labelled"] = labelled;
			return out;
		}, "unmarshal a cortenn.Graph message: {label, nodes (file order), roots, "
		"labelled: [(path, tensor)]}; leaves become llo Variables");
	llo_m.def("pbm_decode", [](py::bytes data)
		{
			pbm::Graph graph;
			pbm::decode(graph, std::string(data));
			py::list nodes;
			for (const pbm::Node& node : graph.nodes_)
			{
				py::dict n;
				n["labels"] = node.labels_;
				if (node.has_source_)
				{
					py::dict src;
					src["shape"] = py::bytes(node.source_.shape_);
					src["wide_shape"] = node.source_.wide_shape_;
					src["data"] = py::bytes(node.source_.data_);
					src["typecode"] = node.source_.typecode_;
					n["source"] = src;
				}
				if (node.has_functor_)
				{
					py::dict fn;
					fn["opname"] = node.functor_.opname_;
					fn["opcode"] = node.functor_.opcode_;
					py::list args;
					for (const pbm::NodeArg& arg : node.functor_.args_)
					{
						py::dict a;
						a["idx"] = arg.idx_;
						a["coord"] = arg.coord_;
						a["shaper"] = arg.shaper_;
						a["fwd"] = arg.fwd_;
						args.append(a);
					}
					fn["args"] = args;
					n["functor"] = fn;
				}
				nodes.append(n);
			}
			py::dict out;
			out["label"] = graph.label_;
			out["nodes"] = nodes;
			return out;
		}, "the messages of a cortenn.Graph file as dictionaries (no tensors are built)");
	llo_m.def("pbm_reencode", [](py::bytes data)
		{
			pbm::Graph graph;
			pbm::decode(graph, std::string(data));
			return py::bytes(pbm::encode(graph));
		}, "decode and encode again (byte-identical for files of the reference's writer)");

	// ---- device-resident additions ----
	py::class_<pyllo::DeviceResult>(llo_m, "DeviceResult")
		.def("ptr", [](pyllo::DeviceResult& self)
			{ return (uintptr_t) self.data_.device_.get(); })
		.def("nbytes", [](pyllo::DeviceResult& self)
			{ return self.data_.nbytes(); })
		.def("shape", [](pyllo::DeviceResult& self)
			{ return pyllo::c2pshape(self.data_.shape_); })
		.def("numpy", [](pyllo::DeviceResult& self, py::dtype dtype)
			{
				auto pshape = pyllo::c2pshape(self.data_.shape_);
				return py::array(dtype, pshape, self.data_.fetch());
			})
		.def_property_readonly("__cuda_array_interface__", [](pyllo::DeviceResult& self)
			{
				return pyllo::cuda_array_interface(self.data_.device_,
					pyllo::c2pshape(self.data_.shape_), self.data_.dtype_);
			}, "the result where it lies in HBM (torch.as_tensor / cupy.asarray: no copy)")
		.def("__dlpack__", [](pyllo::DeviceResult& self, py::object stream)
			{
				(void) stream;
				return pyllo::dlpack_of(self.data_.device_,
					pyllo::c2pshape(self.data_.shape_), self.data_.dtype_);
			}, py::arg("stream") = py::none())
		.def("__dlpack_device__", [](pyllo::DeviceResult&)
			{
				return py::make_tuple(2, llo_cuda_device(llo::device::ctx()));
			});
	llo_m.def("pinned_empty", [](std::vector<py::ssize_t> shape, py::dtype dtype)
		{
			size_t n = (size_t) dtype.itemsize();
			for (py::ssize_t d : shape)
			{
				n *= (size_t) d;
			}
			return pyllo::wrap_host(dtype, shape, llo::device::host_alloc(n));
		}, "uninitialised numpy array in page-locked host memory: "
		"Variable.assign from it and evaluate() move data by DMA",
		py::arg("shape"), py::arg("dtype") = py::dtype::of<double>());
	py::class_<pyllo::PendingResult>(llo_m, "PendingResult")
		.def("result", [](pyllo::PendingResult& self)
		{
			self.data_.wait();
			return pyllo::wrap_host(self.dtype_, pyllo::c2pshape(self.data_.shape_),
				self.data_.data_);
		}, "wait for the device->host copy and return the array");
	llo_m.def("evaluate_async", [](T tens, py::dtype dtype)
		{
			age::_GENERATED_DTYPE ctype = pyllo::eval_dtype(dtype);
			pyllo::PendingResult out;
			out.dtype_ = dtype;
			out.data_ = llo::eval_device(tens, ctype);
			out.data_.fetch_async();
			return out;
		}, "queue the evaluation and the copy of its result to pinned host "
			"memory; .result() waits.  With Variable.assign_async the upload of "
			"the next input overlaps the download of the last result",
		py::arg("tens"), py::arg("dtype") = py::dtype::of<double>());
	llo_m.def("evaluate_device", [](T tens, py::dtype dtype)
		{
			age::_GENERATED_DTYPE ctype = pyllo::eval_dtype(dtype);
			pyllo::DeviceResult out;
			out.data_ = llo::eval_device(tens, ctype);
			return out;
		}, py::arg("tens"), py::arg("dtype") = py::dtype::of<double>());
	llo_m.def("evaluate_device_many", [](std::vector<T> roots, py::dtype dtype)
		{
			age::_GENERATED_DTYPE ctype = pyllo::eval_dtype(dtype);
			std::vector<llo::GenericData> datas;
			datas = llo::eval_device_many(roots, ctype);
			std::vector<pyllo::DeviceResult> outs;
			for (llo::GenericData& d : datas)
			{
				outs.push_back(pyllo::DeviceResult{d});
			}
			return outs;
		}, "evaluate several roots of one graph together; shared sub-graphs "
		"are computed once and the results stay on the device",
		py::arg("roots"), py::arg("dtype") = py::dtype::of<double>());

	py::class_<pyllo::DeviceBuffer>(llo_m, "DeviceBuffer")
		.def(py::init<size_t>(), py::arg("nbytes"))
		.def("ptr", [](pyllo::DeviceBuffer& self)
			{ return (uintptr_t) self.data_.get(); })
		.def("nbytes", [](pyllo::DeviceBuffer& self) { return self.nbytes_; })
		.def("store", [](pyllo::DeviceBuffer& self, size_t offset,
			pyllo::DeviceResult& src)
			{
				size_t n = src.data_.nbytes();
				if (offset + n > self.nbytes_)
				{
					logs::fatalf("cannot store %d bytes at offset %d of a "
						"%d-byte buffer", (int) n, (int) offset, (int) self.nbytes_);
				}
				llo::device::check(llo_cuda_d2d(llo::device::ctx(),
					self.data_.get() + offset, src.data_.device_.get(), n));
			}, "device-to-device copy of a result into the buffer",
			py::arg("offset"), py::arg("src"))
		.def("upload", [](pyllo::DeviceBuffer& self, size_t offset, py::array data)
			{
				py::array dense = py::array::ensure(data, py::array::c_style);
				size_t n = (size_t) dense.nbytes();
				if (offset + n > self.nbytes_)
				{
					logs::fatal("upload past the end of the buffer");
				}
				llo::device::check(llo_cuda_h2d(llo::device::ctx(),
					self.data_.get() + offset, dense.data(), n));
			}, py::arg("offset"), py::arg("data"))
		.def("numpy", [](pyllo::DeviceBuffer& self, py::dtype dtype,
			size_t offset, size_t count)
			{
				py::array out(dtype, std::vector<py::ssize_t>{(py::ssize_t) count});
				size_t n = count * (size_t) dtype.itemsize();
				if (offset + n > self.nbytes_)
				{
					logs::fatal("read past the end of the buffer");
				}
				llo::device::check(llo_cuda_d2h(llo::device::ctx(),
					out.mutable_data(), self.data_.get() + offset, n));
				return out;
			}, py::arg("dtype"), py::arg("offset"), py::arg("count"))
		.def("allreduce", [](pyllo::DeviceBuffer& self, py::dtype dtype,
			std::string opname, bool ordered)
			{
				age::_GENERATED_DTYPE ctype = pyllo::eval_dtype(dtype);
				age::_GENERATED_OPCODE code = age::get_op(opname);
				size_t count = self.nbytes_ / (size_t) dtype.itemsize();
				llo::device::check(llo_cuda_allreduce(llo::device::ctx(),
					self.data_.get(), count, (int) ctype, (int) code,
					ordered ? LLO_ALLREDUCE_ORDERED : LLO_ALLREDUCE_FAST));
			}, "fold the buffer in place over every rank of the context's "
			"communicator (K8)",
			py::arg("dtype"), py::arg("op") = "SUM", py::arg("ordered") = false);

	py::class_<llo::ShardedEvaluator>(llo_m, "ShardedEvaluator")
		.def(py::init([](T loss, std::vector<T> params, unsigned passes,
			bool ordered, bool overlap, bool solo, bool graph, bool deep_first)
			{
				llo::ShardedOptions opts;
				opts.passes_ = passes;
				opts.ordered_ = ordered;
				opts.overlap_ = overlap;
				opts.solo_ = solo;
				opts.graph_ = graph;
				opts.deep_first_ = deep_first;
				return new llo::ShardedEvaluator(loss, params, opts);
			}), "gradient evaluation of `loss` w.r.t. `params`, summed over the "
			"ranks of the communicator (llo/sharded.hpp)",
			py::arg("loss"), py::arg("params"),
			py::arg("passes") = (unsigned) llo::OPT_ALL, py::arg("ordered") = false,
			py::arg("overlap") = true, py::arg("solo") = false, py::arg("graph") = true,
			py::arg("deep_first") = true)
		.def("replays", [](llo::ShardedEvaluator& self) { return self.replays_; },
			"steps replayed from the captured CUDA graph so far")
		.def("grad_eval_device", [](llo::ShardedEvaluator& self, py::dtype dtype,
			bool with_loss)
			{
				age::_GENERATED_DTYPE ctype = pyllo::eval_dtype(dtype);
				const llo::GradientPack& pk = self.pack(ctype, with_loss);
				return pyllo::DeviceBuffer(self.grad_eval_device(ctype, with_loss),
					pk.nbytes());
			}, "one gradient evaluation; the packed, summed gradients stay in HBM "
			"(the buffer is reused by the next call)",
			py::arg("dtype"), py::arg("with_loss") = false)
		.def("grad_eval", [](llo::ShardedEvaluator& self, py::dtype dtype,
			bool with_loss)
			{
				age::_GENERATED_DTYPE ctype = pyllo::eval_dtype(dtype);
				std::vector<llo::GenericData> datas = self.grad_eval(ctype, with_loss);
				std::vector<py::array> outs;
				for (llo::GenericData& d : datas)
				{
					outs.push_back(pyllo::wrap_host(dtype,
						pyllo::c2pshape(d.shape_), d.data_));
				}
				return outs;
			}, py::arg("dtype"), py::arg("with_loss") = false)
		.def("pack_layout", [](llo::ShardedEvaluator& self, py::dtype dtype,
			bool with_loss)
			{
				const llo::GradientPack& pk = self.pack(pyllo::eval_dtype(dtype), with_loss);
				py::dict out;
				out["offsets"] = pk.offsets_;
				out["sizes"] = pk.sizes_;
				out["count"] = pk.count_;
				std::vector<std::vector<py::ssize_t>> shapes;
				for (const ade::Shape& sh : pk.shapes_)
				{
					shapes.push_back(pyllo::c2pshape(sh));
				}
				out["shapes"] = shapes;
				return out;
			}, py::arg("dtype"), py::arg("with_loss") = false)
		.def("roots", [](llo::ShardedEvaluator& self) { return self.roots_; })
		.def("opt_report", [](llo::ShardedEvaluator& self)
			{
				py::dict rep;
				rep["functors_before"] = self.report_.functors_before_;
				rep["functors_after"] = self.report_.functors_after_;
				rep["one_mul"] = self.report_.one_mul_;
				rep["zero_add"] = self.report_.zero_add_;
				rep["pow_one"] = self.report_.pow_one_;
				rep["single"] = self.report_.single_;
				rep["shared"] = self.report_.shared_;
				rep["div_cancel"] = self.report_.div_cancel_;
				return rep;
			});
	llo_m.def("shard_bounds", [](size_t batch, size_t world, size_t rank)
		{
			return llo::ShardPlan(batch, world).bounds(rank);
		}, "[first, last) rows of `rank` when `batch` rows are split over `world` ranks");

	llo_m.def("comm_unique_id", []()
		{
			char id[128];
			llo::device::check(llo_cuda_comm_unique_id(id));
			return py::bytes(id, 128);
		}, "NCCL unique id (rank 0 creates it, every rank receives it)");
	llo_m.def("comm_init", [](int rank, int nranks, py::bytes id)
		{
			std::string raw = id;
			if (nranks > 1 && raw.size() != 128)
			{
				logs::fatal("a communicator id is 128 bytes");
			}
			raw.resize(128);
			py::gil_scoped_release release;
			llo::device::check(llo_cuda_comm_init(llo::device::ctx(),
				rank, nranks, raw.data()));
		}, py::arg("rank"), py::arg("nranks"), py::arg("id") = py::bytes(""));
	llo_m.def("comm_destroy", []()
		{ llo::device::check(llo_cuda_comm_destroy(llo::device::ctx())); });
	llo_m.def("comm_size", []()
		{ return llo_cuda_comm_size(llo::device::ctx()); });
	llo_m.def("comm_rank", []()
		{ return llo_cuda_comm_rank(llo::device::ctx()); });
	llo_m.def("set_stream", [](uintptr_t stream)
		{
			llo::device::check(llo_cuda_set_stream(llo::device::ctx(),
				(void*) stream));
		}, "run the evaluator on an external cudaStream_t (0 = its own)");

	llo_m.def("available", []() { return llo::device::available(); });
	llo_m.def("sync", []()
		{ llo::device::check(llo_cuda_sync(llo::device::ctx())); });
	llo_m.def("launch_count", []()
		{ return llo_cuda_launch_count(llo::device::ctx()); });
	llo_m.def("context", []()
		{ return (uintptr_t) llo::device::ctx(); });
	llo_m.def("plan_stats", []()
		{
			const llo::PlanStats& st = llo::last_plan_stats();
			py::dict out;
			out["nodes"] = st.nodes_;
			out["views"] = st.views_;
			out["consts"] = st.consts_;
			out["fused"] = st.fused_;
			out["programs"] = st.programs_;
			out["reduces"] = st.reduces_;
			out["gemms"] = st.gemms_;
			out["gathers"] = st.gathers_;
			out["generic"] = st.generic_;
			out["simplified"] = st.simplified_;
			out["shared"] = st.shared_;
			return out;
		});
	llo_m.def("plan_preview", [](std::vector<T> roots, py::dtype dtype)
		{
			llo::PlanStats st = llo::plan_preview(roots, pyllo::eval_dtype(dtype));
			py::dict out;
			out["nodes"] = st.nodes_;
			out["views"] = st.views_;
			out["consts"] = st.consts_;
			out["generic"] = st.generic_;
			out["simplified"] = st.simplified_;
			out["shared"] = st.shared_;
			return out;
		}, "how the planner classifies the graphs, without evaluating them (no GPU needed)",
		py::arg("roots"), py::arg("dtype") = py::dtype::of<double>());
	llo_m.def("set_specialize_min", [](uint64_t elements)
		{
			llo::device::check(llo_cuda_set_specialize_min(llo::device::ctx(), elements));
		}, "elementwise launches of at least this many outputs run a kernel compiled "
		"for their program (NVRTC, cached); smaller ones run the interpreted kernel");
	llo_m.def("specialize_stats", []()
		{
			uint64_t st[4];
			llo::device::check(llo_cuda_specialize_stats(llo::device::ctx(), st));
			py::dict out;
			out["launches"] = st[0];
			out["compiles"] = st[1];
			out["disk_hits"] = st[2];
			out["failures"] = st[3];
			return out;
		});
	llo_m.def("set_option", &llo::set_option);
	llo_m.def("get_option", &llo::get_option);
}
