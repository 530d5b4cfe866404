#include <cstdlib>
#include <map>
#include <mutex>

#include "llo/data.hpp"

#ifdef LLO_DATA_HPP

namespace llo
{

namespace device
{

static llo_cuda_ctx* global_ctx = nullptr;

static std::string init_error;

static std::once_flag init_once;

static void init_ctx (void)
{
	int dev = 0;
	if (const char* env = std::getenv("LLO_CUDA_DEVICE"))
	{
		dev = std::atoi(env);
	}
	else if (const char* env = std::getenv("LOCAL_RANK"))
	{
		dev = std::atoi(env);
	}
	if (LLO_OK != llo_cuda_init(dev, &global_ctx))
	{
		global_ctx = nullptr;
		init_error = llo_cuda_last_error();
	}
}

llo_cuda_ctx* ctx (void)
{
	std::call_once(init_once, init_ctx);
	if (nullptr == global_ctx)
	{
		logs::fatalf("llo CUDA context unavailable (%s): the B200 evaluator "
			"has no CPU fallback", init_error.c_str());
	}
	return global_ctx;
}

bool available (void)
{
	std::call_once(init_once, init_ctx);
	return nullptr != global_ctx;
}

void check (int status)
{
	if (LLO_OK == status)
	{
		return;
	}
	if (LLO_ERR_UNSUPPORTED_TYPED_OP == status)
	{
		throw std::bad_function_call();
	}
	logs::fatal(llo_cuda_last_error());
}

std::shared_ptr<char> alloc (size_t nbytes)
{
	llo_cuda_ctx* c = ctx();
	void* ptr = nullptr;
	check(llo_cuda_alloc(c, nbytes, &ptr));
	return std::shared_ptr<char>((char*) ptr,
		[c](char* p) { llo_cuda_free(c, p); });
}

static const size_t pin_threshold = 1 << 16;

/// Pinned blocks are expensive to create (cudaHostAlloc maps and locks every
/// page), so released blocks are kept for the next request of the same size
struct PinnedPool
{
	std::mutex lock_;
	std::multimap<size_t,void*> free_;
	size_t cached_ = 0;
	static constexpr size_t cap_ = (size_t) 8 << 30;

	void* take (size_t nbytes)
	{
		std::lock_guard<std::mutex> guard(lock_);
		auto it = free_.find(nbytes);
		if (free_.end() == it)
		{
			return nullptr;
		}
		void* ptr = it->second;
		free_.erase(it);
		cached_ -= nbytes;
		return ptr;
	}

	bool give (size_t nbytes, void* ptr)
	{
		std::lock_guard<std::mutex> guard(lock_);
		if (cached_ + nbytes > cap_)
		{
			return false;
		}
		free_.emplace(nbytes, ptr);
		cached_ += nbytes;
		return true;
	}
};

static PinnedPool& pinned_pool (void)
{
	static PinnedPool* pool = new PinnedPool(); // outlives every array
	return *pool;
}

std::shared_ptr<char> host_alloc (size_t nbytes)
{
	if (nbytes >= pin_threshold && available())
	{
		llo_cuda_ctx* c = ctx();
		void* ptr = pinned_pool().take(nbytes);
		if (nullptr != ptr || LLO_OK == llo_cuda_host_alloc(c, nbytes, &ptr))
		{
			return std::shared_ptr<char>((char*) ptr,
				[c, nbytes](char* p)
				{
					if (false == pinned_pool().give(nbytes, p))
					{
						llo_cuda_host_free(c, p);
					}
				});
		}
	}
	return std::shared_ptr<char>((char*) std::malloc(std::max<size_t>(nbytes, 1)),
		[](char* p) { std::free(p); });
}

}

GenericData::GenericData (ade::Shape shape, age::_GENERATED_DTYPE dtype) :
	shape_(shape), dtype_(dtype),
	device_(device::alloc(shape.n_elems() * age::type_size(dtype))) {}

void GenericData::copyover (const char* indata, age::_GENERATED_DTYPE intype)
{
	llo_cuda_ctx* c = device::ctx();
	size_t n = shape_.n_elems();
	if (dtype_ == intype)
	{
		device::check(llo_cuda_h2d(c, device_.get(), indata, nbytes()));
	}
	else
	{
		size_t inbytes = n * age::type_size(intype);
		std::shared_ptr<char> staged = device::alloc(inbytes);
		device::check(llo_cuda_h2d(c, staged.get(), indata, inbytes));
		device::check(llo_cuda_convert(c, device_.get(), dtype_,
			staged.get(), intype, n));
	}
	data_ = nullptr;
}

char* GenericData::fetch (void)
{
	if (nullptr != pending_)
	{
		return wait(); // never hand out a half-written mirror
	}
	if (nullptr == data_ && nullptr != device_)
	{
		data_ = device::host_alloc(nbytes());
		device::check(llo_cuda_d2h(device::ctx(),
			data_.get(), device_.get(), nbytes()));
	}
	return data_.get();
}

void GenericData::fetch_async (void)
{
	if (nullptr != data_ || nullptr == device_)
	{
		return;
	}
	llo_cuda_ctx* c = device::ctx();
	data_ = device::host_alloc(nbytes());
	void* event = nullptr;
	device::check(llo_cuda_d2h_async(c, data_.get(), device_.get(), nbytes(), &event));
	// The download runs on its own lane: neither buffer may go back to its
	// pool (stream-ordered device free, pinned block reuse) before the DMA has
	// finished.  The event's deleter therefore holds both and waits first, so
	// dropping every copy of this GenericData without wait() stays safe.
	std::shared_ptr<char> dev = device_;
	std::shared_ptr<char> host = data_;
	pending_ = std::shared_ptr<void>(event,
		[c, dev, host](void* ev)
		{
			llo_cuda_event_sync(c, ev);
			llo_cuda_event_destroy(c, ev);
		});
}

char* GenericData::wait (void)
{
	if (nullptr != pending_)
	{
		std::shared_ptr<void> event = pending_;
		pending_ = nullptr;
		device::check(llo_cuda_event_sync(device::ctx(), event.get()));
	}
	return fetch();
}

Variable::Variable (const char* data, age::_GENERATED_DTYPE dtype,
	ade::Shape shape, std::string label) :
	label_(label), shape_(shape), dtype_(dtype)
{
	if (nullptr != data)
	{
		host_ = device::host_alloc(nbytes());
		std::memcpy(host_.get(), data, nbytes());
	}
	else
	{
		// zeros (llo/data.hpp:72-79), kept lazy
		fill_.assign(age::type_size(dtype), 0);
	}
}

Variable* Variable::filled (const char* elem, age::_GENERATED_DTYPE dtype,
	ade::Shape shape, std::string label)
{
	Variable* out = new Variable(nullptr, dtype, shape, label);
	out->fill_.assign(elem, elem + age::type_size(dtype));
	return out;
}

Variable::Variable (const Variable& other) :
	label_(other.label_), fill_(other.fill_),
	shape_(other.shape_), dtype_(other.dtype_)
{
	if (nullptr != other.host_ || other.host_stale_)
	{
		const void* src = other.data(); // refreshes a stale host copy
		host_ = device::host_alloc(nbytes());
		std::memcpy(host_.get(), src, nbytes());
	}
}

Variable& Variable::operator = (const Variable& other)
{
	if (this != &other)
	{
		label_ = other.label_;
		fill_ = other.fill_;
		shape_ = other.shape_;
		dtype_ = other.dtype_;
		host_ = nullptr;
		host_stale_ = false;
		if (nullptr != other.host_ || other.host_stale_)
		{
			const void* src = other.data();
			host_ = device::host_alloc(nbytes());
			std::memcpy(host_.get(), src, nbytes());
		}
		++version_;
		mirrors_.clear();
	}
	return *this;
}

Variable& Variable::operator = (GenericRef data)
{
	if (false == data.shape_.compatible_after(shape(), 0))
	{
		logs::fatalf("cannot assign data of incompatible shaped %s to "
			"internal data of shape %s", data.shape_.to_string().c_str(),
			shape().to_string().c_str());
	}
	if (data.dtype_ != dtype_)
	{
		logs::fatalf("cannot assign data of incompatible types %s "
			"(external) and %s (internal)",
			age::name_type(data.dtype_).c_str(), age::name_type(dtype_).c_str());
	}
	fill_.clear();
	++version_;
	if (device::available())
	{
		// upload straight from the caller's buffer (one pass; DMA when the
		// buffer is pinned); the host copy is refreshed only if asked for.
		// Always into a FRESH buffer: results of earlier evaluations may share
		// the old mirror (a root that is a leaf, or collapses to a view of
		// one) and must keep their values; the old block returns to the pool
		// in stream order once nobody holds it.
		std::shared_ptr<char> raw = device::alloc(nbytes());
		device::check(llo_cuda_h2d(device::ctx(), raw.get(), data.data_,
			nbytes()));
		mirrors_.clear(); // converted mirrors are out of date
		mirrors_.push_back(Mirror{dtype_, version_, raw});
		host_stale_ = true;
		return *this;
	}
	if (nullptr == host_)
	{
		host_ = device::host_alloc(nbytes());
	}
	std::memcpy(host_.get(), data.data_, nbytes());
	return *this;
}

void Variable::assign_async (GenericRef data)
{
	if (false == data.shape_.compatible_after(shape(), 0))
	{
		logs::fatalf("cannot assign data of incompatible shaped %s to "
			"internal data of shape %s", data.shape_.to_string().c_str(),
			shape().to_string().c_str());
	}
	if (data.dtype_ != dtype_)
	{
		logs::fatalf("cannot assign data of incompatible types %s "
			"(external) and %s (internal)",
			age::name_type(data.dtype_).c_str(), age::name_type(dtype_).c_str());
	}
	fill_.clear();
	++version_;
	// a fresh buffer: queued evaluations still read the previous mirror,
	// which returns to the pool in stream order once they are done
	std::shared_ptr<char> raw = device::alloc(nbytes());
	device::check(llo_cuda_h2d_async(device::ctx(), raw.get(), data.data_,
		nbytes()));
	mirrors_.clear();
	mirrors_.push_back(Mirror{dtype_, version_, raw});
	host_stale_ = true;
}

void Variable::assign_device (const void* device_data, ade::Shape shape,
	age::_GENERATED_DTYPE dtype)
{
	if (false == shape.compatible_after(shape_, 0))
	{
		logs::fatalf("cannot assign data of incompatible shaped %s to "
			"internal data of shape %s", shape.to_string().c_str(),
			shape_.to_string().c_str());
	}
	if (dtype != dtype_)
	{
		logs::fatalf("cannot assign data of incompatible types %s "
			"(external) and %s (internal)",
			age::name_type(dtype).c_str(), age::name_type(dtype_).c_str());
	}
	fill_.clear();
	++version_;
	std::shared_ptr<char> raw = device::alloc(nbytes());
	device::check(llo_cuda_d2d(device::ctx(), raw.get(), device_data, nbytes()));
	mirrors_.clear();
	mirrors_.push_back(Mirror{dtype_, version_, raw});
	host_stale_ = true;
}

void Variable::materialize (void) const
{
	if (host_stale_)
	{
		if (nullptr == host_)
		{
			host_ = device::host_alloc(nbytes());
		}
		for (Mirror& m : mirrors_)
		{
			if (m.dtype_ == dtype_ && m.version_ == version_)
			{
				device::check(llo_cuda_d2h(device::ctx(), host_.get(),
					m.data_.get(), nbytes()));
				host_stale_ = false;
				return;
			}
		}
		logs::fatal("variable lost its device mirror while the host copy was stale");
	}
	if (nullptr != host_)
	{
		return;
	}
	size_t esize = age::type_size(dtype_);
	size_t n = shape_.n_elems();
	host_ = device::host_alloc(nbytes());
	char* dst = host_.get();
	bool allzero = std::all_of(fill_.begin(), fill_.end(),
		[](char c) { return 0 == c; });
	if (allzero)
	{
		std::memset(dst, 0, n * esize);
		return;
	}
	for (size_t i = 0; i < n; ++i)
	{
		std::memcpy(dst + i * esize, fill_.data(), esize);
	}
}

double Variable::fill_value (void) const
{
	if (fill_.empty())
	{
		logs::fatal("cannot get fill value of a non-constant variable");
	}
	const char* p = fill_.data();
	switch (dtype_)
	{
#define AGE_X(NAME, CTYPE) case age::NAME:\
		{ CTYPE v; std::memcpy(&v, p, sizeof(CTYPE)); return (double) v; }
		AGE_DTYPES(AGE_X)
#undef AGE_X
		default: break;
	}
	logs::fatal("cannot get fill value of bad type");
}

std::shared_ptr<char> Variable::device_data (
	age::_GENERATED_DTYPE dtype) const
{
	for (Mirror& m : mirrors_)
	{
		if (m.dtype_ == dtype)
		{
			if (m.version_ == version_)
			{
				return m.data_;
			}
			break;
		}
	}
	llo_cuda_ctx* c = device::ctx();
	size_t n = shape_.n_elems();

	// raw mirror in the leaf's own type first
	std::shared_ptr<char> raw;
	for (Mirror& m : mirrors_)
	{
		if (m.dtype_ == dtype_ && m.version_ == version_)
		{
			raw = m.data_;
		}
	}
	if (nullptr == raw)
	{
		raw = device::alloc(nbytes());
		device::check(llo_cuda_h2d(c, raw.get(), data(), nbytes()));
		mirrors_.erase(std::remove_if(mirrors_.begin(), mirrors_.end(),
			[&](Mirror& m) { return m.dtype_ == dtype_; }), mirrors_.end());
		mirrors_.push_back(Mirror{dtype_, version_, raw});
	}
	if (dtype == dtype_)
	{
		return raw;
	}
	std::shared_ptr<char> conv = device::alloc(n * age::type_size(dtype));
	device::check(llo_cuda_convert(c, conv.get(), dtype, raw.get(), dtype_, n));
	mirrors_.erase(std::remove_if(mirrors_.begin(), mirrors_.end(),
		[&](Mirror& m) { return m.dtype_ == dtype; }), mirrors_.end());
	mirrors_.push_back(Mirror{dtype, version_, conv});
	return conv;
}

}

#endif
