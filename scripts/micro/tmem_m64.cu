// Probe: where do the rows of an M=64 (cta_group::1) tcgen05.mma accumulator land in TMEM?
// D[m][n] = (m+1) + 128*(n+1); every warp dumps its 32 lanes x 32 columns.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rcnn-ocr_b200/csrc -I include -o scripts/micro/tmem_m64 scripts/micro/tmem_m64.cu
#include <cstdio>
#include <cuda_bf16.h>
#include "sm100.cuh"
using namespace rcnn::sm100;

__global__ void __launch_bounds__(128) probe(float *out, int M) {
    __shared__ __align__(1024) unsigned char a_s[16384];
    __shared__ __align__(1024) unsigned char b_s[4096];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16384 / 2; i += 128) reinterpret_cast<uint16_t *>(a_s)[i] = 0;
    for (int i = threadIdx.x; i < 4096 / 2; i += 128) reinterpret_cast<uint16_t *>(b_s)[i] = 0;
    __syncthreads();
    auto put = [](unsigned char *tile, int r, int k, float v) {
        const int off = r * 128 + ((((k * 2) >> 4) ^ (r & 7)) << 4) + ((k * 2) & 15);
        *reinterpret_cast<__nv_bfloat16 *>(tile + off) = __float2bfloat16(v);
    };
    if (threadIdx.x < M) { put(a_s, threadIdx.x, 0, (float)(threadIdx.x + 1)); put(a_s, threadIdx.x, 1, 128.f); }
    if (threadIdx.x < 32) { put(b_s, threadIdx.x, 0, 1.f); put(b_s, threadIdx.x, 1, (float)(threadIdx.x + 1)); }
    fence_proxy_async_smem();
    if (warp == 0) {
        if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
        __syncwarp();
        tmem_alloc<32>(&slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    // clear the accumulator columns first (all 128 lanes) with a zero M=128 MMA... simpler: tcgen05.st zeros
    {
        uint32_t z = 0;
        for (int c = 0; c < 32; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tm + ((uint32_t)(warp * 32) << 16) + c), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = make_idesc_bf16(M, 32);
        const uint64_t ad = make_smem_desc_sw128(smem_u32(a_s), 16, 1024);
        const uint64_t bd = make_smem_desc_sw128(smem_u32(b_s), 16, 1024);
        umma_bf16(tm, ad, bd, idesc, 0);
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 32 + c] = __uint_as_float(r[c]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<32>(tm); }
}

int main() {
    float *d; cudaMalloc(&d, 128 * 32 * 4);
    static float h[128 * 32];
    for (int M : {128, 64}) {
        cudaMemset(d, 0, 128 * 32 * 4);
        probe<<<1, 128>>>(d, M);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("M=%d (%s): lane -> (row m, first col n) decoded from column 0 and column 5\n", M, cudaGetErrorString(e));
        for (int l = 0; l < 128; ++l) {
            const int v0 = (int)h[l * 32 + 0], v5 = (int)h[l * 32 + 5];
            printf("  lane %3d: c0 -> m=%3d n=%2d | c5 -> m=%3d n=%2d%s", l, v0 % 128 - 1, v0 / 128 - 1, v5 % 128 - 1, v5 / 128 - 1, (l % 2) ? "\n" : "   ");
        }
    }
    return 0;
}
