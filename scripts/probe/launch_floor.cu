// Fixed cost of a launch shaped like kl_rows_cluster_kernel: 120 CTAs x 576 threads, 207 KB dynamic shared memory,
// clusters of 8, TMEM alloc/dealloc, one cluster barrier - and nothing else.  Back-to-back launches, CUDA events.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(576, 1) k(float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem);
    const int warp = threadIdx.x >> 5;
    if (MODE & 1) {
        if (warp == 16) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (MODE & 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (MODE & 1) {
        __syncthreads();
        if (warp == 16) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*slot) : "memory");
    }
    if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1.f;
}
template <int MODE>
float run(int nc, int smem, int threads) {
    auto kern = k<MODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(threads); cfg.gridDim = dim3(120); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = nc; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = nc > 1 ? 1 : 0;
    float* out = nullptr;
    for (int i = 0; i < 5; ++i) cudaLaunchKernelEx(&cfg, kern, out);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 50; ++i) cudaLaunchKernelEx(&cfg, kern, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    return ms / 50 * 1e3f;
}
int main() {
    printf("empty kernel, no cluster, 1 KB smem, 576 thr      : %6.2f us\n", run<0>(1, 1024, 576));
    printf("empty kernel, no cluster, 207 KB smem             : %6.2f us\n", run<0>(1, 207 * 1024, 576));
    printf("empty kernel, cluster 8, 207 KB smem              : %6.2f us\n", run<0>(8, 207 * 1024, 576));
    printf("+ cluster barrier                                 : %6.2f us\n", run<2>(8, 207 * 1024, 576));
    printf("+ TMEM alloc/dealloc                              : %6.2f us\n", run<1>(8, 207 * 1024, 576));
    printf("+ both                                            : %6.2f us\n", run<3>(8, 207 * 1024, 576));
    return 0;
}
