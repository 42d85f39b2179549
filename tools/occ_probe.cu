// Which feature makes cudaOccupancyMaxActiveBlocksPerMultiprocessor report 1 CTA/SM?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
extern __shared__ uint8_t sm[];
__global__ void __launch_bounds__(384, 2) k_plain(float* o) { o[threadIdx.x] = sm[threadIdx.x]; }
__global__ void __launch_bounds__(384, 2) k_setmaxnreg(float* o) {
    if (threadIdx.x < 128) { asm volatile("setmaxnreg.dec.sync.aligned.u32 24;"); }
    else { asm volatile("setmaxnreg.inc.sync.aligned.u32 104;"); }
    o[threadIdx.x] = sm[threadIdx.x];
}
__global__ void __launch_bounds__(384, 2) k_tmem(float* o) {
    __shared__ uint32_t base;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    __syncthreads();
    o[threadIdx.x] = sm[threadIdx.x] + base;
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base));
}
__global__ void __launch_bounds__(384, 2) k_namedbar(float* o) {
    if (threadIdx.x >= 128) asm volatile("bar.sync 1, 256;");
    o[threadIdx.x] = sm[threadIdx.x];
}
__global__ void __launch_bounds__(256, 2) k_setmaxnreg256(float* o) {
    if (threadIdx.x < 128) { asm volatile("setmaxnreg.dec.sync.aligned.u32 32;"); }
    else { asm volatile("setmaxnreg.inc.sync.aligned.u32 224;"); }
    o[threadIdx.x] = sm[threadIdx.x];
}
template <typename K> void probe(const char* name, K k, int threads) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    for (int smem : {32768, 107776, 110000}) {
        int occ = -1;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, threads, smem);
        printf("%-18s threads %d regs %d smem %6d -> occupancy %d (%s)\n", name, threads, fa.numRegs, smem, occ, cudaGetErrorString(e));
    }
}
int main() {
    probe("plain", k_plain, 384);
    probe("setmaxnreg", k_setmaxnreg, 384);
    probe("tmem", k_tmem, 384);
    probe("namedbar", k_namedbar, 384);
    probe("setmaxnreg256", k_setmaxnreg256, 256);
    return 0;
}
