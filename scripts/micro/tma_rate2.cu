// Microbenchmark: how fast can ONE SM ingest L2-resident data?  Sweeps the TMA box height (rows of
// 128 B), the number of issuing warps, the ring depth, and compares with 1-D bulk copies
// (cp.async.bulk) and plain 128-bit LDG.  The data set is small (L2 resident).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I rcnn-ocr_b200/csrc -I include -o scripts/micro/tma_rate2 \
//      scripts/micro/tma_rate2.cu rcnn-ocr_b200/csrc/gemm.cu rcnn-ocr_b200/csrc/abi_common.cu
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "sm100.cuh"
using namespace rcnn;
using namespace rcnn::sm100;

constexpr int kMaxRingBytes = 196608;

__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// MODE 0: TMA 2-D boxes of BOX_ROWS x 128 B;  MODE 1: 1-D bulk copies of the same size.
// All loop indices are powers of two known at compile time: no integer division in the issue loop.
template <int MODE, int BOX_ROWS, int RING>
__global__ void __launch_bounds__(256) k(const __grid_constant__ CUtensorMap tm, const unsigned char *src, int iters,
                                         long long *cycles) {
    extern __shared__ unsigned char raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int w = threadIdx.x / 32;
    constexpr uint32_t box_bytes = BOX_ROWS * 128;
    uint64_t *bars = (uint64_t *)(smem + kMaxRingBytes);
    uint64_t *full = bars + w * RING;
    smem += (size_t)w * RING * box_bytes;
    if ((threadIdx.x & 31) == 0) {
        for (int i = 0; i < RING; ++i) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        long long t0 = clock64();
        int issued = 0, done = 0;
        auto issue = [&]() {
            const int id = (blockIdx.x * 7 + w * 3 + issued) & 255;       // 32 column tiles x 8 row tiles
            uint64_t *b = &full[issued & (RING - 1)];
            unsigned char *dst = smem + (size_t)(issued & (RING - 1)) * box_bytes;
            mbar_arrive_expect_tx(b, box_bytes);
            if (MODE == 0) tma_load_2d(dst, &tm, b, (id & 31) * 64, (id >> 5) * BOX_ROWS);
            else bulk_load_1d(dst, src + (size_t)id * box_bytes, box_bytes, b);
            ++issued;
        };
        while (issued < RING && issued < iters) issue();
        for (; done < iters; ++done) {
            mbar_wait(&full[done & (RING - 1)], (done / RING) & 1);
            if (issued < iters) issue();
        }
        if (w == 0) cycles[blockIdx.x] = clock64() - t0;
    }
}

template <int MODE, int BOX_ROWS, int RING>
void run(const CUtensorMap &tm, const void *buf, int warps, int ctas, long long *dc) {
    const int iters = 512;
    if ((size_t)warps * RING * BOX_ROWS * 128 > kMaxRingBytes) return;
    size_t smem = 1024 + kMaxRingBytes + 1024;
    cudaFuncSetAttribute(k<MODE, BOX_ROWS, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) k<MODE, BOX_ROWS, RING><<<ctas, 32 * warps, smem>>>(tm, (const unsigned char *)buf, iters, dc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, dc, ctas * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
    printf("%s box %3d rows (%5d B) warps %d ring %2d ctas %3d: %7.1f cyc/box/warp  %6.1f B/clk/SM  (%s)\n",
           MODE ? "bulk1d" : "tma2d ", BOX_ROWS, BOX_ROWS * 128, warps, RING, ctas, avg / iters,
           (double)warps * BOX_ROWS * 128.0 * iters / avg, cudaGetErrorString(e));
}

template <int MODE, int BOX_ROWS>
void sweep(const void *buf, long long *dc) {
    CUtensorMap tm;
    if (make_tmap_2d(&tm, buf, 2, 2048, 2048, 4096, BOX_ROWS, 64, 1)) { printf("tmap fail %s\n", rcnn_last_error()); exit(1); }
    for (int warps : {1, 2, 4})
        for (int ctas : {1, 128}) {
            run<MODE, BOX_ROWS, 1>(tm, buf, warps, ctas, dc);
            run<MODE, BOX_ROWS, 2>(tm, buf, warps, ctas, dc);
            run<MODE, BOX_ROWS, 4>(tm, buf, warps, ctas, dc);
            run<MODE, BOX_ROWS, 8>(tm, buf, warps, ctas, dc);
        }
}

// plain loads: every thread pulls 16 B per iteration, `unroll` independent loads in flight
__global__ void __launch_bounds__(1024) kldg(const uint4 *src, int n16, int iters, long long *cycles, uint4 *sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    __syncthreads();
    long long t0 = clock64();
    int idx = (blockIdx.x * 977 + threadIdx.x) % n16;
    for (int it = 0; it < iters; ++it) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v[u] = ld_nc_v4(src + idx);
            idx += blockDim.x; if (idx >= n16) idx -= n16;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    if (acc.x == 0x12345678u) sink[threadIdx.x] = acc;
}

int main() {
    long long *dc; cudaMalloc(&dc, 256 * 8);
    const int rows = 2048, cols = 2048;            // bf16: 8 MB, L2 resident
    size_t bytes = (size_t)rows * cols * 2;
    void *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    sweep<0, 32>(buf, dc); sweep<0, 64>(buf, dc); sweep<0, 128>(buf, dc); sweep<0, 256>(buf, dc);
    sweep<1, 32>(buf, dc); sweep<1, 128>(buf, dc); sweep<1, 256>(buf, dc);
    uint4 *sink; cudaMalloc(&sink, 1024 * 16);
    for (int threads : {128, 256, 512, 1024})
        for (int ctas : {1, 128}) {
            for (int rep = 0; rep < 2; ++rep) kldg<<<ctas, threads>>>((const uint4 *)buf, (int)(bytes / 16), 64, dc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[256]; cudaMemcpy(h, dc, ctas * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < ctas; ++i) avg += h[i]; avg /= ctas;
            printf("ldg128 threads %4d ctas %3d: %6.1f B/clk/SM  (%s)\n", threads, ctas, threads * 16.0 * 8 * 64 / avg, cudaGetErrorString(e));
        }
    return 0;
}
