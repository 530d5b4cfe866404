// ORACLE (test infrastructure) — python access to the CPU restatement:
// evaluate an ade graph built with the product's `age` module on the CPU.

#include "pybind11/pybind11.h"
#include "pybind11/numpy.h"
#include "pybind11/stl.h"

#include "cpu_llo.hpp"

namespace py = pybind11;

static oracle::DType dtype_of (py::dtype dtype)
{
	char kind = dtype.kind();
	py::ssize_t nbytes = dtype.itemsize();
	switch (kind)
	{
		case 'f':
			return 4 == nbytes ? age::FLOAT : age::DOUBLE;
		case 'i':
			switch (nbytes)
			{
				case 1: return age::INT8;
				case 2: return age::INT16;
				case 4: return age::INT32;
				default: return age::INT64;
			}
		case 'u':
			switch (nbytes)
			{
				case 1: return age::UINT8;
				case 2: return age::UINT16;
				case 4: return age::UINT32;
				default: return age::UINT64;
			}
		default:
			logs::fatalf("unknown dtype %c", kind);
	}
}

// numpy shape = ade shape reversed with trailing 1s dropped
// (llo/python/llo.cpp:24-34)
static std::vector<py::ssize_t> numpy_shape (const ade::Shape& shape)
{
	auto it = shape.begin();
	auto et = shape.end();
	while (it != et && *(et - 1) == 1)
	{
		--et;
	}
	std::vector<py::ssize_t> fwd(it, et);
	return std::vector<py::ssize_t>(fwd.rbegin(), fwd.rend());
}

PYBIND11_MODULE(_oracle, m)
{
	m.doc() = "CPU oracle of the llo evaluator (test infrastructure)";

	m.def("evaluate", [](ade::TensptrT tens, py::dtype dtype)
	{
		oracle::HostData out = oracle::eval(tens, dtype_of(dtype));
		auto pshape = numpy_shape(out.shape_);
		// no base object: numpy copies the buffer
		return py::array(dtype, pshape, out.data_.get());
	}, py::arg("tens"), py::arg("dtype") = py::dtype::of<double>());

	m.def("seed", [](size_t seed) { oracle::engine().seed(seed); });
}
