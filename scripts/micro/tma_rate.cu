// Microbenchmark: TMA 2-D box [128 rows x 64 bf16] load rate per SM vs. the global row pitch.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rcnn-ocr_b200/csrc -o /tmp/tma_rate scripts/micro/tma_rate.cu rcnn-ocr_b200/csrc/gemm.cu rcnn-ocr_b200/csrc/abi_common.cu
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "sm100.cuh"
using namespace rcnn;
using namespace rcnn::sm100;

template <int RING>
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tm, int iters, int ncol_tiles, int nrow_tiles,
                                         long long *cycles) {
    extern __shared__ unsigned char raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int nw = blockDim.x / 32, w = threadIdx.x / 32;   // each warp's lane 0 drives its own ring
    uint64_t *full = (uint64_t *)(smem + 4 * RING * 16384) + w * RING;
    smem += w * RING * 16384;
    if ((threadIdx.x & 31) == 0) {
        for (int i = 0; i < RING; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        long long t0 = clock64();
        int issued = 0, done = 0;
        // keep RING loads in flight
        for (; issued < RING && issued < iters; ++issued) {
            int id = (blockIdx.x * 7 + w * 3 + issued) % (ncol_tiles * nrow_tiles);
            mbar_arrive_expect_tx(&full[issued % RING], 16384);
            tma_load_2d(smem + (issued % RING) * 16384, &tm, &full[issued % RING], (id % ncol_tiles) * 64, (id / ncol_tiles) * 128);
        }
        for (; done < iters; ++done) {
            mbar_wait(&full[done % RING], (done / RING) & 1);
            if (issued < iters) {
                int id = (blockIdx.x * 7 + w * 3 + issued) % (ncol_tiles * nrow_tiles);
                mbar_arrive_expect_tx(&full[issued % RING], 16384);
                tma_load_2d(smem + (issued % RING) * 16384, &tm, &full[issued % RING], (id % ncol_tiles) * 64, (id / ncol_tiles) * 128);
                ++issued;
            }
        }
        if (w == 0) cycles[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    const int iters = 512;
    long long *dc; cudaMalloc(&dc, 256 * 8);
    const int rows = 256;
    for (long long pitch_elems : {4096LL, 262144LL}) {
        size_t bytes = (size_t)rows * pitch_elems * 2;
        void *buf; if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc fail\n"); continue; }
        cudaMemset(buf, 0, bytes);
        CUtensorMap tm;
        int ncol = (int)(pitch_elems / 64 < 32 ? pitch_elems / 64 : 32);
        if (make_tmap_2d(&tm, buf, 2, rows, (uint64_t)ncol * 64, pitch_elems * 2, 128, 64, 1)) { printf("tmap fail %s\n", rcnn_last_error()); return 1; }
        for (int ctas : {1, 148}) for (int warps : {1, 2, 4}) {
            for (int ring : {3}) {
                size_t smem = 1024 + 4 * ring * 16384 + 256;
                auto launch = [&](auto kern) {
                    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    kern<<<ctas, 32 * warps, smem>>>(tm, iters, ncol, rows / 128, dc);
                    kern<<<ctas, 32 * warps, smem>>>(tm, iters, ncol, rows / 128, dc);
                };
                if (ring == 3) launch(k<3>); else if (ring == 6) launch(k<6>); else launch(k<12>);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[256]; cudaMemcpy(h, dc, ctas * 8, cudaMemcpyDeviceToHost);
                double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
                printf("pitch %8lld B  ctas %3d warps %d ring %2d: %7.1f cyc/box/warp  %6.1f B/clk/SM  (%s)\n", pitch_elems * 2, ctas, warps, ring,
                       avg / iters, warps * 16384.0 * iters / avg, cudaGetErrorString(e));
            }
        }
        cudaFree(buf);
    }
    return 0;
}
