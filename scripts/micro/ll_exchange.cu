// Probe for the next step of the recurrent kernels' exchange: what does one all-gather step between G CTAs on G SMs
// cost with (A) plain stores + ONE gpu-scope release per CTA + acquire polling of a counter + loads (what K2 does
// now), against flag-in-data protocols where every store carries the step number and readers poll the data itself
// (no fence, no counter): (B) 16-byte packets = 12 bytes of payload + 4-byte flag, (C) 8-byte packets = 4 + 4 (NCCL's
// LL), both polling one source after the other, (D) = B with the packets of up to 16 sources polled together.  Every CTA publishes 256 packets per step and reads the 256 packets of every CTA (G x 4 KB for A/B).
// Torn packets (flag of the new step, payload of the old one) are counted.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/ll_exchange scripts/micro/ll_exchange.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kThreads = 256;
constexpr int kMaxSpins = 20000000;   // a protocol bug ends the run (bit 30 of the torn count) instead of hanging the GPU

__device__ __forceinline__ void red_release(unsigned int *p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void *p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint2 ld_volatile_v2(const void *p) {
    uint2 r;
    asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_volatile_v4(void *p, const uint4 &v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_volatile_v2(void *p, const uint2 &v) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint4 ld_cg_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ unsigned int payload(unsigned int step, int src, int idx, int w) {
    return step * 2654435761u + (unsigned)(src * 4099 + idx * 17 + w);
}

// buf: [2 parity][G src][256 packets] of 16 bytes (mode 0, 1) or 8 bytes (mode 2)
__global__ void __launch_bounds__(kThreads, 1)
exchange(int mode, int G, int steps, unsigned char *buf, unsigned int *counter, long long *cycles, unsigned int *torn,
         unsigned int *sink) {
    extern __shared__ unsigned char pad[];   // 1 CTA per SM
    const int cta = blockIdx.x, tid = threadIdx.x;
    __shared__ unsigned int go;
    unsigned int acc = 0, bad = 0;
    long long t0 = 0;
    for (int s = 0; s < steps; ++s) {
        if (s == 16 && cta == 0 && tid == 0) t0 = clock64();
        const unsigned int tag = (unsigned)s + 1u;
        const size_t slot = (size_t)(s & 1) * G;
        if (mode == 3) {
            // flag-in-data, 16-byte packets, all (up to 16) sources polled together: the loads of one pass are independent
            uint4 v = make_uint4(payload(tag, cta, tid, 0), payload(tag, cta, tid, 1), payload(tag, cta, tid, 2), tag);
            st_volatile_v4(buf + ((slot + cta) * kThreads + tid) * 16, v);
            for (int base = 0; base < G; base += 16) {
                const int n = G - base < 16 ? G - base : 16;
                unsigned int pend = n == 32 ? 0xffffffffu : ((1u << n) - 1u);
                uint4 r[16];
                int spins = 0;
                while (pend) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if ((pend >> j) & 1u) r[j] = ld_volatile_v4(buf + ((slot + base + j) * kThreads + tid) * 16);
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (((pend >> j) & 1u) && r[j].w == tag) {
                            pend &= ~(1u << j);
                            bad += (r[j].x != payload(tag, base + j, tid, 0)) | (r[j].y != payload(tag, base + j, tid, 1)) |
                                   (r[j].z != payload(tag, base + j, tid, 2));
                            acc += r[j].x + r[j].y + r[j].z;
                        }
                    if (++spins >= kMaxSpins) { bad |= 0x40000000u; pend = 0; s = steps; }
                }
            }
        } else if (mode == 2) {
            uint2 v = make_uint2(payload(tag, cta, tid, 0), tag);
            st_volatile_v2(buf + ((slot + cta) * kThreads + tid) * 8, v);
            for (int src = 0; src < G; ++src) {
                uint2 r;
                const unsigned char *p = buf + ((slot + src) * kThreads + tid) * 8;
                int spins = 0;
                do { r = ld_volatile_v2(p); } while (r.y != tag && ++spins < kMaxSpins);
                if (spins >= kMaxSpins) { bad |= 0x40000000u; s = steps; break; }
                bad += r.x != payload(tag, src, tid, 0);
                acc += r.x;
            }
        } else if (mode == 1) {
            uint4 v = make_uint4(payload(tag, cta, tid, 0), payload(tag, cta, tid, 1), payload(tag, cta, tid, 2), tag);
            st_volatile_v4(buf + ((slot + cta) * kThreads + tid) * 16, v);
            for (int src = 0; src < G; ++src) {
                uint4 r;
                const unsigned char *p = buf + ((slot + src) * kThreads + tid) * 16;
                int spins = 0;
                do { r = ld_volatile_v4(p); } while (r.w != tag && ++spins < kMaxSpins);
                if (spins >= kMaxSpins) { bad |= 0x40000000u; s = steps; break; }
                bad += (r.x != payload(tag, src, tid, 0)) | (r.y != payload(tag, src, tid, 1)) | (r.z != payload(tag, src, tid, 2));
                acc += r.x + r.y + r.z;
            }
        } else {
            uint4 v = make_uint4(payload(tag, cta, tid, 0), payload(tag, cta, tid, 1), payload(tag, cta, tid, 2), payload(tag, cta, tid, 3));
            *reinterpret_cast<uint4 *>(buf + ((slot + cta) * kThreads + tid) * 16) = v;
            __syncthreads();
            if (tid == 0) {
                red_release(counter);
                const unsigned int target = tag * (unsigned)G;
                int spins = 0;
                while (ld_acquire(counter) < target && ++spins < kMaxSpins) {}
                go = spins >= kMaxSpins ? 0u : tag;
            }
            __syncthreads();
            if (go == 0u) { bad |= 0x40000000u; break; }
            for (int src = 0; src < G; ++src) {
                const uint4 r = ld_cg_v4(buf + ((slot + src) * kThreads + tid) * 16);
                bad += (r.x != payload(tag, src, tid, 0)) | (r.w != payload(tag, src, tid, 3));
                acc += r.x + r.y + r.z + r.w;
            }
        }
    }
    if (cta == 0 && tid == 0) cycles[0] = clock64() - t0;
    atomicAdd(torn, bad);
    if (acc == 0x12345678u) sink[0] = acc + go;
}

int main() {
    const int steps = 4016;
    const char *names[4] = {"A  stores + release + counter + acquire poll + loads (16 B / thread)",
                            "B  flag-in-data, 16-byte packets (12 B payload + flag)",
                            "C  flag-in-data,  8-byte packets ( 4 B payload + flag)",
                            "D  flag-in-data, 16-byte packets, 16 sources polled together"};
    unsigned char *buf;
    unsigned int *counter, *torn, *sink;
    long long *cycles;
    cudaMalloc(&buf, 2 * 64 * kThreads * 16);
    cudaMalloc(&counter, 4); cudaMalloc(&torn, 4); cudaMalloc(&sink, 4); cudaMalloc(&cycles, 8);
    cudaFuncSetAttribute(exchange, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int G : {2, 16, 64}) {
        for (int mode = 0; mode < 4; ++mode) {
            cudaMemset(buf, 0, 2 * 64 * kThreads * 16);
            cudaMemset(counter, 0, 4); cudaMemset(torn, 0, 4);
            void *args[] = {(void *)&mode, (void *)&G, (void *)&steps, (void *)&buf, (void *)&counter, (void *)&cycles, (void *)&torn, (void *)&sink};
            cudaError_t e = cudaLaunchCooperativeKernel((void *)exchange, dim3(G), dim3(kThreads), args, 200 * 1024, 0);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            long long c = 0; unsigned int t = 0;
            cudaMemcpy(&c, cycles, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(&t, torn, 4, cudaMemcpyDeviceToHost);
            printf("G=%2d  %s: %7.0f cycles/step, torn packets %u  (%s)\n", G, names[mode], (double)c / (steps - 16), t, cudaGetErrorString(e));
        }
    }
    return 0;
}
