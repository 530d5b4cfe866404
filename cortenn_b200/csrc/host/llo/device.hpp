///
/// device.hpp
/// llo
///
/// Purpose:
/// Thin C++ access to the C-ABI CUDA library (include/llo_cuda.h): the
/// process-wide context, status-code -> exception translation, and
/// shared_ptr deleters that return device / pinned memory to the library.
/// Mirrors the reference's ownership model (std::shared_ptr<char> with a
/// free() deleter, llo/src/data.cpp:8-18) with device memory instead.
///

#ifndef LLO_DEVICE_HPP
#define LLO_DEVICE_HPP

#include <functional>
#include <memory>
#include <stdexcept>

#include "llo_cuda.h"

namespace llo
{

namespace device
{

/// Process-wide context (created on first use on device LLO_CUDA_DEVICE, else
/// LOCAL_RANK, else 0).  Throws std::runtime_error when no B200 is usable:
/// there is no CPU fallback.
llo_cuda_ctx* ctx (void);

/// True if a context exists or can be created (never throws)
bool available (void);

/// Translate a C-ABI status into the reference's exception types:
/// std::bad_function_call for unsupported typed ops, std::runtime_error
/// (logs::fatal) otherwise
void check (int status);

/// Device buffer of nbytes owned by the returned pointer
std::shared_ptr<char> alloc (size_t nbytes);

/// Host buffer of nbytes: pinned when a context is available and the buffer
/// is large enough to matter for transfers, plain malloc otherwise
std::shared_ptr<char> host_alloc (size_t nbytes);

}

}

#endif // LLO_DEVICE_HPP
