/*
 * scan_compat.cuh -- the handful of CUDA spellings the scan kernels use.
 *
 * Product builds (nvcc, sm_100a) take the first branch.  tests/emu/ compiles the
 * very same kernel source with g++ -DSCAN_EMU against tests/emu/cuda_emu.h (one
 * OS thread per CUDA thread, barriers for __syncthreads) so that the index math
 * of every kernel can be parity-checked on the GPU-less build box.  The emulator
 * is test infrastructure only; nothing in the shipped library can reach it.
 */
#pragma once
#include <stdint.h>

#ifdef SCAN_EMU
#include "cuda_emu.h"
#define SCAN_DYN_SMEM(name) unsigned char *name = ::cuda_emu::dyn_smem()
#define SCAN_GRID_CONSTANT
#else
#include <cuda_runtime.h>
#define SCAN_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SCAN_GRID_CONSTANT __grid_constant__
#endif

#define SCAN_DEV __device__ __forceinline__

namespace rscan {

/* 16-byte asynchronous global->shared copy (LDGSTS), commit and drain. */
SCAN_DEV void cp_async16(void *smem_dst, const void *gmem_src)
{
#ifdef SCAN_EMU
	::cuda_emu::copy16(smem_dst, gmem_src);
#else
	unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
#endif
}

SCAN_DEV void cp_async_commit()
{
#ifndef SCAN_EMU
	asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
}

SCAN_DEV void cp_async_wait_all()
{
#ifndef SCAN_EMU
	asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}

/* wait until at most `pending` of the most recently committed copy groups are still in flight */
SCAN_DEV void cp_async_wait_pending(int pending)
{
#ifndef SCAN_EMU
	switch (pending) {
	case 0: asm volatile("cp.async.wait_group 0;\n" ::: "memory"); break;
	case 1: asm volatile("cp.async.wait_group 1;\n" ::: "memory"); break;
	case 2: asm volatile("cp.async.wait_group 2;\n" ::: "memory"); break;
	default: asm volatile("cp.async.wait_group 3;\n" ::: "memory"); break;
	}
#else
	(void)pending;
#endif
}

template <int PENDING>
SCAN_DEV void cp_async_wait_group()
{
#ifndef SCAN_EMU
	asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory");
#endif
}

/*
 * Programmatic dependent launch (sm_90+): a kernel launched with the
 * programmatic-stream-serialization attribute may start while the previous
 * kernel in the stream is still running; pdl_wait() blocks until that kernel
 * has completed and its writes are visible, pdl_launch_dependents() lets the
 * next kernel start early.  Both are no-ops for ordinary launches.
 */
SCAN_DEV void pdl_wait()
{
#ifndef SCAN_EMU
	asm volatile("griddepcontrol.wait;\n" ::: "memory");
#endif
}

SCAN_DEV void pdl_launch_dependents()
{
#ifndef SCAN_EMU
	asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
#endif
}

} // namespace rscan
