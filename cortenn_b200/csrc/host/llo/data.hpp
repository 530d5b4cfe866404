///
/// data.hpp
/// llo
///
/// Purpose:
/// Device-resident llo data store.  Keeps the reference's types and names
/// (llo/data.hpp:24-283): GenericData, GenericRef, Variable, DataArg, VecRef,
/// to_ref(s), get_variable, get_scalar.
///
/// What changes against the reference:
///  * GenericData owns DEVICE memory (`device_`); `data_` is a host mirror
///    filled on demand by fetch() (llo::eval fetches its result, so
///    `out.data_.get()` reads exactly as in the reference's tests)
///  * Variable keeps its host bytes (iLeaf::data() stays a host pointer for
///    pbm / serialize) plus per-dtype device mirrors with a version stamp,
///    so a leaf is uploaded and converted once, not on every visit
///    (reference: llo/eval.hpp:31-38 re-converts per visit)
///  * DataArg::data_ / VecRef::data are DEVICE pointers
///

#ifndef LLO_DATA_HPP
#define LLO_DATA_HPP

#include <algorithm>
#include <cstring>
#include <memory>

#include "ade/coord.hpp"
#include "ade/ileaf.hpp"

#include "llo/generated/codes.hpp"
#include "llo/device.hpp"

namespace llo
{

/// GenericData for holding data when passing up the tensor graph
struct GenericData final
{
	GenericData (void) = default;

	/// Allocate uninitialised device storage for shape x dtype
	GenericData (ade::Shape shape, age::_GENERATED_DTYPE dtype);

	/// Copy over host data of specified type while retaining shape:
	/// upload + on-device conversion (reference: llo/src/data.cpp:49-72)
	void copyover (const char* indata, age::_GENERATED_DTYPE intype);

	/// Make the host mirror `data_` current and return it
	char* fetch (void);

	/// Start copying the elements to a (pinned) host mirror on the download
	/// lane without waiting for it: later device work is not held up and an
	/// upload of the next input runs at the same time.  wait() completes it.
	void fetch_async (void);

	/// Complete fetch_async (no-op otherwise) and return the host mirror
	char* wait (void);

	/// Completion event of a pending fetch_async
	std::shared_ptr<void> pending_;

	size_t nbytes (void) const
	{
		return (size_t) age::type_size(dtype_) * shape_.n_elems();
	}

	/// Host mirror of the elements (null until fetch())
	std::shared_ptr<char> data_;

	/// Shape of data_
	ade::Shape shape_;

	/// Type encoding of data_
	age::_GENERATED_DTYPE dtype_ = age::BAD_TYPE;

	/// Device-resident elements
	std::shared_ptr<char> device_;
};

/// GenericRef for holding HOST data
/// Ref uses raw pointer instead of shared, so it's memory unsafe
struct GenericRef
{
	GenericRef (char* data, ade::Shape shape, age::_GENERATED_DTYPE dtype) :
		data_(data), shape_(shape), dtype_(dtype) {}

	GenericRef (GenericData& generic) :
		data_(generic.fetch()),
		shape_(generic.shape_), dtype_(generic.dtype_) {}

	char* data_;

	ade::Shape shape_;

	age::_GENERATED_DTYPE dtype_;
};

/// Leaf node: host bytes + lazily synchronised device mirrors
struct Variable final : public ade::iLeaf
{
	Variable (const char* data, age::_GENERATED_DTYPE dtype,
		ade::Shape shape, std::string label);

	/// Constant leaf: every element equals the one at `elem`.  The host
	/// bytes are only materialised if data() is asked for, and the device
	/// evaluator folds the value into kernels instead of storing a tensor
	/// (derivative rules create one full-shape constant per rule:
	/// llo/cfg/llo.json:32-123 `llo::get_scalar(1,args[0]->shape())`)
	static Variable* filled (const char* elem, age::_GENERATED_DTYPE dtype,
		ade::Shape shape, std::string label);

	Variable (const Variable& other);

	Variable (Variable&& other) = default;

	Variable& operator = (const Variable& other);

	Variable& operator = (Variable&& other) = default;

	/// Assign vectorized data to data source
	template <typename T>
	Variable& operator = (std::vector<T> data)
	{
		GenericRef ref((char*) &data[0], shape(), age::get_type<T>());
		return operator = (ref);
	}

	/// Assign generic reference to data source
	/// (error strings: llo/data.hpp:125-133)
	Variable& operator = (GenericRef data);

	/// Same checks and effect as `= data`, but the upload runs on the copy
	/// lane (llo_cuda_h2d_async) into a fresh device buffer: evaluations
	/// already queued keep reading the old one, evaluations queued afterwards
	/// see the new values, and the caller does not wait.  `data` must be
	/// pinned (llo.pinned_empty) and stay untouched until a later result has
	/// been waited for.
	void assign_async (GenericRef data);

	/// Same checks and effect as `= data` for elements that already live in
	/// DEVICE memory (another library's tensor handed over through
	/// __cuda_array_interface__ / DLPack): one device-to-device copy into a
	/// fresh mirror, no host bounce.  The copy is ordered on the evaluator's
	/// stream; the caller orders the producer before it (llo_cuda_wait_stream).
	void assign_device (const void* device_data, ade::Shape shape,
		age::_GENERATED_DTYPE dtype);

	/// Implementation of iTensor
	const ade::Shape& shape (void) const override
	{
		return shape_;
	}

	/// Implementation of iTensor
	std::string to_string (void) const override
	{
		return label_ + "(" + shape_.to_string() + ")";
	}

	/// Implementation of iLeaf.  The mutable overload may be written
	/// through, so it invalidates the device mirrors.
	void* data (void) override
	{
		materialize();
		fill_.clear();
		++version_;
		return host_.get();
	}

	/// Implementation of iLeaf
	const void* data (void) const override
	{
		materialize();
		return host_.get();
	}

	/// True while every element is known to equal fill_value()
	bool is_filled (void) const
	{
		return false == fill_.empty();
	}

	/// The constant of a filled leaf, converted to double
	double fill_value (void) const;

	/// Implementation of iLeaf
	size_t type_code (void) const override
	{
		return dtype_;
	}

	/// Return number of bytes in data source
	size_t nbytes (void) const
	{
		return (size_t) age::type_size(dtype_) * shape_.n_elems();
	}

	/// Counter that moves whenever the elements may have changed (anything
	/// that cached work derived from this leaf compares it)
	uint64_t version (void) const
	{
		return version_;
	}

	/// Device copy of the elements converted to dtype (uploaded / converted
	/// only when the host bytes changed since the last call)
	std::shared_ptr<char> device_data (age::_GENERATED_DTYPE dtype) const;

	/// Label for distinguishing variable nodes
	std::string label_;

private:
	struct Mirror
	{
		age::_GENERATED_DTYPE dtype_;
		uint64_t version_;
		std::shared_ptr<char> data_;
	};

	void materialize (void) const;

	mutable std::shared_ptr<char> host_;

	/// True while the newest bytes only exist in the device mirror (an
	/// assignment uploads straight from the caller's buffer); the host copy
	/// is refreshed on demand
	mutable bool host_stale_ = false;

	/// One element's bytes while the leaf is a known constant
	std::vector<char> fill_;

	ade::Shape shape_;

	age::_GENERATED_DTYPE dtype_;

	uint64_t version_ = 1;

	mutable std::vector<Mirror> mirrors_;
};

/// Smart pointer for variable nodes
using VarptrT = std::shared_ptr<llo::Variable>;

/// Return new variable containing input vector data according to
/// specified shape and labelled according to input label
/// Throw error if the input vector size differs from shape.n_elems()
template <typename T>
VarptrT get_variable (std::vector<T> data, ade::Shape shape,
	std::string label = "")
{
	if (data.size() != shape.n_elems())
	{
		logs::fatalf("cannot create variable with data size %d "
			"against shape %s", (int) data.size(), shape.to_string().c_str());
	}
	return VarptrT(new Variable((char*) &data[0],
		age::get_type<T>(), shape, label));
}

/// Return new variable containing 0s
template <typename T>
VarptrT get_variable (ade::Shape shape, std::string label = "")
{
	return VarptrT(new Variable(nullptr, age::get_type<T>(), shape, label));
}

/// Return new variable filled with scalar, labelled with the scalar's
/// printed value unless a label is given ("0" / "1" labels drive
/// llo::zero_prune, llo/src/zprune.cpp:96-105)
template <typename T>
VarptrT get_scalar (T scalar, ade::Shape shape, std::string label = "")
{
	if (label.empty())
	{
		label = fmts::to_string(scalar);
	}
	return VarptrT(Variable::filled((const char*) &scalar,
		age::get_type<T>(), shape, label));
}

/// Data to pass around when evaluating (device pointers)
struct DataArg
{
	/// Smart pointer to generic device data as bytes
	std::shared_ptr<char> data_;

	/// Shape of the generic data
	ade::Shape shape_;

	/// Coordinate mapper
	ade::CoordptrT mapper_;

	/// True if the coordinate mapper accepts input coordinates,
	/// False if it accepts output coordinates
	bool fwd_;
};

/// Vector of DataArgs to hold arguments
using DataArgsT = std::vector<DataArg>;

/// Type-specific tensor data wrapper using raw DEVICE pointer
template <typename T>
struct VecRef
{
	const T* data;

	ade::Shape shape;

	ade::CoordptrT mapper;

	bool push;
};

/// Converts DataArgs to VecRef of specific type
template <typename T>
VecRef<T> to_ref (DataArg& arg)
{
	return VecRef<T>{
		(const T*) arg.data_.get(),
		arg.shape_,
		arg.mapper_,
		arg.fwd_,
	};
}

/// Converts multiple DataArgs to multiple VecRefs of the same type
template <typename T>
std::vector<VecRef<T>> to_refs (DataArgsT& args)
{
	std::vector<VecRef<T>> out;
	std::transform(args.begin(), args.end(), std::back_inserter(out),
		[](DataArg& arg)
		{
			return to_ref<T>(arg);
		});
	return out;
}

}

#endif // LLO_DATA_HPP
