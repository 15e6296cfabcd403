// How many clusters of each size can be co-resident on this GPU (1 CTA/SM: 200 KB dynamic smem)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(544, 1) k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    for (int nc = 1; nc <= 16; ++nc) {
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(544); cfg.dynamicSmemBytes = 200 * 1024;
        cfg.gridDim = dim3(nc * 64);
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = nc; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster %2d: max active clusters %3d -> %3d SMs (%s)\n", nc, n, n * nc, cudaGetErrorString(e));
        cudaGetLastError();
    }
    return 0;
}
