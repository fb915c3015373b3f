// Thread-private stash in Tensor Memory (sm_100a).
//
// The epilogue of fftconv_kernel needs the segment's own input again (the direct
// branch x_d of output i is input i - Lh/2, which for Lh/2 a multiple of 512 is
// an input of the same thread).  Registers cannot hold it across the five
// passes, shared memory is full (128 KB segment), and re-reading it from L2
// costs a second pass over the interleaved input.  TMEM (256 KB per SM, idle in
// a kernel without tensor-core work) is laid out as 128 lanes x 512 columns of
// 32 bits; with the 32x32b access shape lane l of warp w owns TMEM lane
// 32 (w % 4) + l, so it is exactly a per-thread scratch.  The four warps that
// share a lane quarter take 128 columns each: columns 0..63 hold the thread's 32
// complex inputs, columns 64..127 the filter spectrum of the two MID rows the
// thread owns (constant for the whole launch; loaded once instead of streaming
// 128 KB per segment through L1).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace prk {

constexpr int kTmemCols = 512;
constexpr int kTmemGCol = 64; // first column of the filter spectrum

// warp 0 allocates; the base address lands in *slot (shared memory); call with all threads
__device__ __forceinline__ uint32_t tmem_alloc_all (uint32_t* slot, int tid)
{
	if (tid < 32) {
		const uint32_t sa = (uint32_t)__cvta_generic_to_shared (slot);
		asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa), "n"(kTmemCols) : "memory");
		asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads ();
	asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory");
	return *slot;
}
__device__ __forceinline__ void tmem_free_all (uint32_t base, int tid)
{
	asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads ();
	if (tid < 32) {
		asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kTmemCols) : "memory");
	}
}
// address of column 0 of this thread's stash
__device__ __forceinline__ uint32_t tmem_thread_base (uint32_t base, int tid)
{
	const int warp = tid >> 5;
	return base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
}
// four complex values <-> eight consecutive columns
__device__ __forceinline__ void tmem_st4 (uint32_t taddr, float2 a, float2 b, float2 c, float2 d)
{
	asm volatile ("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint (a.x)),
	              "r"(__float_as_uint (a.y)), "r"(__float_as_uint (b.x)), "r"(__float_as_uint (b.y)), "r"(__float_as_uint (c.x)),
	              "r"(__float_as_uint (c.y)), "r"(__float_as_uint (d.x)), "r"(__float_as_uint (d.y))
	              : "memory");
}
__device__ __forceinline__ void tmem_wait_st () { asm volatile ("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4 (uint32_t taddr, float2 (&v)[4])
{
	uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
	asm volatile ("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
	              : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
	              : "r"(taddr)
	              : "memory");
	asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory");
	v[0] = make_float2 (__uint_as_float (r0), __uint_as_float (r1));
	v[1] = make_float2 (__uint_as_float (r2), __uint_as_float (r3));
	v[2] = make_float2 (__uint_as_float (r4), __uint_as_float (r5));
	v[3] = make_float2 (__uint_as_float (r6), __uint_as_float (r7));
}

// sixteen complex values = 32 consecutive columns; tmem_wait_ld16() makes them usable
__device__ __forceinline__ void tmem_ld16_issue (uint32_t taddr, uint32_t (&r)[32])
{
	asm volatile ("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
	                "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
	                "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
	                "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
	              : "r"(taddr)
	              : "memory");
}
// the registers pass through the wait so that no use can be scheduled before it
__device__ __forceinline__ void tmem_wait_ld16 (uint32_t (&r)[32])
{
	asm volatile ("tcgen05.wait::ld.sync.aligned;"
	              : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
	                "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
	                "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
	                "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
	              :
	              : "memory");
}

} // namespace prk
