///
/// specialize.hpp — run-time specialisation of the elementwise engine.
///
/// The interpreted kernels (vm_kernels.cuh compiled by nvcc) spend two thirds
/// of their issue slots on dispatch, accumulator copies and operand-kind
/// branches (ncu, profiles/r1_chain_hot.txt).  For large launches the same
/// kernel source is compiled once per distinct (kernel, dtype, program, input
/// kinds, rank) with NVRTC for sm_100a, the program supplied as constexpr
/// tables (LLO_VM_STATIC), and cached in memory and on disk; the launch then
/// runs straight-line code.  Results are bit-identical by construction: both
/// builds apply the same operator functions (ops.cuh) in the same order.
///
/// Nothing here is a fallback for a missing GPU; if NVRTC cannot be loaded
/// the interpreted kernels run and llo_cuda_specialize_stats says so.
///

#ifndef LLO_CUDA_SPECIALIZE_HPP
#define LLO_CUDA_SPECIALIZE_HPP

#include "vm.cuh"

namespace llo_cuda
{

enum SpecKernel
{
	SPEC_FLAT = 0,
	SPEC_TILE,           // padded tiles (element sizes other than 4 bytes)
	SPEC_TILE4_ALIGNED,
	SPEC_TILE4_UNALIGNED,
	SPEC_PAIR4,
	SPEC_PAIR4_TMA,      // TMA-staged, warp-specialised form of SPEC_PAIR4
};

/// Launch the specialised build of `kernel` for program `p` if the launch is
/// large enough and the build is (or can be made) available.  Returns true
/// when the kernel was launched (status in `rc`), false when the caller
/// should launch the interpreted kernel.
bool spec_launch (llo_cuda_ctx* ctx, SpecKernel kernel, int dtype,
	const VmParams& p, unsigned blocks, size_t smem, int& rc);

/// SPEC_PAIR4_TMA: builds the tensor map of the pair program's staged input
/// (the 4-D / 5-D view described at tile4_thread_tma) and launches the
/// TMA-staged kernel.  False when the shape does not fit the view (the caller
/// then tries SPEC_PAIR4) or the build is unavailable.
bool spec_launch_pair_tma (llo_cuda_ctx* ctx, int dtype, const VmParams& p,
	uint64_t units, size_t spill, int& rc);

/// Shape of the specialised pair kernel: unit buffers in its ring and CTAs per
/// SM (the launch side sizes shared memory and the grid from it)
void spec_pair_shape (int& nbuf, int& ctas);

/// True when launches of `n` outputs take the specialised path
bool spec_wanted (llo_cuda_ctx* ctx, uint64_t n);

/// The NVRTC translation unit for a key, for tests and offline inspection
std::string spec_source (SpecKernel kernel, int dtype, const VmParams& p);

}

#endif // LLO_CUDA_SPECIALIZE_HPP
