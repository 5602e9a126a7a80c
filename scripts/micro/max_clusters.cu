#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
    for (int smem : {230656, 180000, 100000}) for (int cs : {16, 8, 4, 2}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("smem %d cluster %2d: max active clusters %d (%s)\n", smem, cs, n, cudaGetErrorString(e));
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); printf("SMs %d\n", p.multiProcessorCount);
}
