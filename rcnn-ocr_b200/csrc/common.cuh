// Shared helpers for the sm_100a kernels (error plumbing, warp primitives, loads).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/rcnn_ocr_b200.h"

namespace rcnn {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define RCNN_CHECK_ARG(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            ::rcnn::set_error(__VA_ARGS__);       \
            return RCNN_ERR_ARG;                  \
        }                                         \
    } while (0)

#define RCNN_CUDA(call)                                             \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return ::rcnn::cuda_fail(e__, #call); \
    } while (0)

#define RCNN_LAUNCH_CHECK(name)                                      \
    do {                                                             \
        cudaError_t e__ = cudaGetLastError();                        \
        if (e__ != cudaSuccess) return ::rcnn::cuda_fail(e__, name); \
        ::rcnn::count_launch();                                      \
    } while (0)

int num_sms();
bool chain_launches();   // rcnn_chain_launches(1): launch with cudaLaunchAttributeProgrammaticStreamSerialization
inline void chain_config(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, unsigned grid, unsigned block, size_t smem,
                         cudaStream_t s) {
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cfg.attrs = attr;
    cfg.numAttrs = 0;
    if (chain_launches()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 1;
    }
}
int gemm_sms();   // num_sms() minus the SMs reserved for concurrent collectives (rcnn_reserve_sms)
// optional per-step clock64 timeline of cluster 0 / CTA 0 of the recurrent kernels (debug aid)
long long *debug_timeline();
// optional device counter of exchange packets re-fetched after the optimistic TMA fetch (debug aid)
unsigned int *debug_refetch();
void count_launch();
// `n` zeroed 32-bit counters (zeroed on `s`, in stream order) for the recurrent kernels' group barriers.
// The storage is a per-device ring of regions allocated once; a region is reused only after 64 later calls.
unsigned int *group_counters(int n, cudaStream_t s);

// Optional per-kernel event timing (see rcnn_prof_* in the header).
int prof_begin(int kernel, cudaStream_t s);
void prof_end(int slot, cudaStream_t s);
struct ProfScope {
    int slot;
    cudaStream_t s;
    ProfScope(int kernel, cudaStream_t st) : slot(prof_begin(kernel, st)), s(st) {}
    ~ProfScope() { prof_end(slot, s); }
};

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int warp_min_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// streaming 128-bit load that does not pollute L1
__device__ __forceinline__ uint4 ld_nc_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// read-only 128-bit load that may allocate in L1 (neighbouring 16-byte pieces of a sector are re-used)
__device__ __forceinline__ uint4 ld_ro_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// 256-bit (one full 32-byte sector per lane) global accesses, sm_100+
struct U8 { uint32_t v[8]; };
__device__ __forceinline__ void st_v8(void *p, const U8 &a) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]),
                 "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]) : "memory");
}
__device__ __forceinline__ U8 ld_ro_v8(const void *p) {
    U8 a;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.v[0]), "=r"(a.v[1]), "=r"(a.v[2]), "=r"(a.v[3]), "=r"(a.v[4]), "=r"(a.v[5]), "=r"(a.v[6]), "=r"(a.v[7])
                 : "l"(p));
    return a;
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ float ld_nc_f32(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}


// Programmatic dependent launch: griddep_launch() lets the next kernel of the stream start its CTAs (they run up to their own
// griddep_wait()); griddep_wait() returns once the previous kernel has completed and its writes are visible.  Both are no-ops
// for a kernel launched without the attribute.  Every kernel of a chain must wait before it reads or overwrites anything a
// predecessor touches -- a kernel that skipped the wait could finish before its predecessor and release its successor early.
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace rcnn
