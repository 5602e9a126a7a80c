// Probe: weights as the A operand FROM TENSOR MEMORY.  tcgen05.cp.128x256b copies one K=16 slice of a
// K-major SW128 shared-memory tile [128 rows x 64 k] into 8 TMEM columns; tcgen05.mma (A in TMEM, B in
// shared memory) then must reproduce A:  with B[n][k] = (k == 16 s + n),  D_s[m][n] = A[m][16 s + n].
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rcnn-ocr_b200/csrc -I include -o scripts/micro/tmem_a_operand scripts/micro/tmem_a_operand.cu
#include <cstdio>
#include <cuda_bf16.h>
#include "sm100.cuh"
using namespace rcnn::sm100;

// (tmem_cp_128x256b / umma_bf16_ts moved into sm100.cuh after this probe)

__global__ void __launch_bounds__(128) probe(float *out) {
    __shared__ __align__(1024) unsigned char a_s[16384];   // [128 x 64] bf16 SW128
    __shared__ __align__(1024) unsigned char b_s[4 * 2048]; // 4 tiles [16 x 64] bf16 SW128 (one per K slice)
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4 * 2048 / 2; i += 128) reinterpret_cast<uint16_t *>(b_s)[i] = 0;
    __syncthreads();
    auto put = [](unsigned char *tile, int r, int k, float v) {
        const int off = r * 128 + ((((k * 2) >> 4) ^ (r & 7)) << 4) + ((k * 2) & 15);
        *reinterpret_cast<__nv_bfloat16 *>(tile + off) = __float2bfloat16(v);
    };
    for (int k = 0; k < 64; ++k) put(a_s, threadIdx.x, k, (float)((threadIdx.x * 3 + k * 7) % 251));
    if (threadIdx.x < 64) { const int s = threadIdx.x >> 4, n = threadIdx.x & 15; put(b_s + s * 2048, n, 16 * s + n, 1.f); }
    fence_proxy_async_smem();
    if (warp == 0) {
        if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
        __syncwarp();
        tmem_alloc<128>(&slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;          // columns [0,32): A (4 slices x 8 columns); [64, 64+64): D_s (16 columns each)
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, 16);
        const uint64_t ad = make_smem_desc_sw128(smem_u32(a_s), 16, 1024);
        for (int s = 0; s < 4; ++s) tmem_cp_128x256b(tm + 8 * s, ad + 2 * s);
        for (int s = 0; s < 4; ++s) {
            const uint64_t bd = make_smem_desc_sw128(smem_u32(b_s + s * 2048), 16, 1024) + 2 * s;
            umma_bf16_ts(tm + 64 + 16 * s, tm + 8 * s, bd, idesc, 0);
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[32];
    for (int h = 0; h < 2; ++h) {
        tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + 64 + 32 * h, r);
        tmem_ld_wait();
        for (int c = 0; c < 32; ++c) out[(warp * 32 + lane) * 64 + 32 * h + c] = __uint_as_float(r[c]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<128>(tm); }
}

int main() {
    float *d; cudaMalloc(&d, 128 * 64 * 4);
    static float h[128 * 64];
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int k = 0; k < 64; ++k) {
            const float want = (float)((m * 3 + k * 7) % 251);
            if (h[m * 64 + k] != want && bad++ < 10) printf("mismatch m=%d k=%d got %g want %g\n", m, k, h[m * 64 + k], want);
        }
    printf("A-from-TMEM probe (%s): %d mismatches of %d\n", cudaGetErrorString(e), bad, 128 * 64);
    return 0;
}
