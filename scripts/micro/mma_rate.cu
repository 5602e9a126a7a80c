// Microbenchmark: back-to-back tcgen05.mma (kind::f16, SS operands resident in smem) rate vs N.
#include <cstdio>
#include "common.cuh"
#include "sm100.cuh"
using namespace rcnn;
using namespace rcnn::sm100;

template <int N, int DISTINCT, int COMMIT_EVERY>
__global__ void __launch_bounds__(128) k(int iters, long long *out) {
    extern __shared__ unsigned char raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bar = (uint64_t *)(smem + 160 * 1024);
    uint64_t *dummy = bar + 1;
    uint32_t *slot = (uint32_t *)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (warp == 0) {
        if (lane == 0) { mbar_init(bar, 1); mbar_init(dummy, 1); fence_barrier_init(); }
        __syncwarp();
        tmem_alloc<256>(slot);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            // DISTINCT tiles: walk through 4 A tiles (16 KB each) and 4 B tiles to defeat any operand reuse
            const int a_off = (i % DISTINCT) * 16384, b_off = 65536 + (i % DISTINCT) * 16384;
            const uint64_t ad = make_smem_desc_sw128(smem_u32(smem + a_off), 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(smem_u32(smem + b_off), 16, 1024);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_bf16(tm, ad + 2 * kk, bd + 2 * kk, idesc, 1);
            if (COMMIT_EVERY && (i % COMMIT_EVERY) == COMMIT_EVERY - 1) umma_commit(dummy);
        }
        umma_commit(bar);
        long long t1 = clock64();
        mbar_wait(bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tm); }
}

template <int N, int D, int CE>
void run(long long *d) {
    const int iters = 256;
    size_t smem = 1024 + 160 * 1024 + 64;
    cudaFuncSetAttribute(k<N, D, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<N, D, CE><<<1, 128, smem>>>(iters, d);
    k<N, D, CE><<<1, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("M=128 N=%3d distinct=%d commit_every=%d chunks: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (%s)\n", N, D, CE, h[0] / (iters * 4.0), h[1] / (iters * 4.0),
           cudaGetErrorString(e));
}

int main() {
    long long *d; cudaMalloc(&d, 64);
    run<32, 4, 0>(d); run<32, 4, 1>(d); run<32, 4, 2>(d); run<32, 4, 4>(d); run<128, 4, 0>(d); run<128, 4, 1>(d); run<128, 4, 2>(d); run<256, 4, 0>(d); run<256, 4, 1>(d);
    return 0;
}
