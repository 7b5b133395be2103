// mbarrier / 1-D bulk copy ("TMA 1-D") / cp.async PTX wrappers shared by the FIR, DDC and FFT kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace gr4b200 {

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void     mbarInit(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void     mbarExpectTx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void     mbarWait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "WAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE;\n\t"
                 "bra WAIT_LOOP;\n\t"
                 "DONE:\n\t"
                 "}" ::"r"(smemAddr(bar)),
                 "r"(parity)
                 : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`; all of dst/src/bytes must be multiples of 16
__device__ __forceinline__ void bulkLoad(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory"); }
__device__ __forceinline__ void fenceBarrierInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template<int Bytes>
__device__ __forceinline__ void cpAsync(void* dstSmem, const void* srcGlobal) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal), "n"(Bytes) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int Pending>
__device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(Pending) : "memory"); }

} // namespace gr4b200
