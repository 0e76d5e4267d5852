// How many clusters of size C (232 KB dynamic smem, 320 threads per CTA) can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int *p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("SMs %d  L2 %d MB\n", pr.multiProcessorCount, pr.l2CacheSize >> 20);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int c : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(c * 64); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = 227 * 1024;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = c; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
        printf("cluster %2d: max active clusters %d (%d SMs) %s\n", c, n, n * c, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
