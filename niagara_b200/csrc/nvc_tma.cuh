// nvc_tma.cuh — bulk asynchronous global -> shared copy (TMA engine, SASS UBLKCP) with mbarrier completion.
// Used to stage the coarse Hi-Z mips of the depth pyramid into shared memory once per CTA.
#pragma once

#include <stdint.h>

namespace nvc
{

#if defined(NVC_EMU) // CPU emulation of the kernels (tests/cuda_emu): the bulk copy is a memcpy that is complete on return
__device__ __forceinline__ void mbar_init(uint64_t*, uint32_t) {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t*) { memcpy(dst_smem, src_gmem, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t*, uint32_t) {}
#else

__device__ __forceinline__ uint32_t smem_addr(const void* p)
{
	return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
	// make the initialised barrier visible to the async (TMA) proxy
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

// bytes: multiple of 16; dst / src: 16-byte aligned
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
	             : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "WAIT_%=:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra DONE_%=;\n"
	    "bra WAIT_%=;\n"
	    "DONE_%=:\n"
	    "}\n" ::"r"(smem_addr(bar)),
	    "r"(parity)
	    : "memory");
}

#endif // NVC_EMU

} // namespace nvc
