// Backbone, inference ("fused SE-scale epilogues"): the squeeze-and-excitation tail of an SE-ResNet block
// (model/seresnet31.py: SELayer + the residual add + ReLU of the block),
//     gate = sigmoid(W2 relu(W1 mean_hw(y)));   out = relu(y * gate + skip)
// as two launches instead of the eight elementwise / reduction / tiny-GEMM launches torch issues for it:
//   se_gate_kernel   one CTA per image: channel means over the H*W pixels, the two small products, the sigmoid -> gate [B, C]
//   se_apply_kernel  all elements, 16 bytes per thread: out = relu(y * gate + skip)
// Tensors are channels_last ([B, H*W, C] in memory, C contiguous), bf16 or f32; W1 [Cr, C] and W2 [C, Cr] f32, no biases.
#include "common.cuh"

namespace rcnn {
namespace {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void add(const uint4 &q, float (&acc)[8]) {
        acc[0] += __uint_as_float(q.x); acc[1] += __uint_as_float(q.y); acc[2] += __uint_as_float(q.z); acc[3] += __uint_as_float(q.w);
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void add(const uint4 &q, float (&acc)[8]) {
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            acc[2 * k] += __uint_as_float(w[k] << 16);
            acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
        }
    }
};

// One CTA of kGateThreads per image.  VEC: the pixel rows are read 16 bytes per thread, C / N threads per row and
// kGateThreads / (C / N) rows per pass, four passes in flight (a thread per channel walking the rows one by one is a chain of
// H*W dependent L2 round trips: 38 us for a 256 x 8 x 32 image); partial sums meet in shared memory.
constexpr int kGateThreads = 1024;
template <typename T, bool VEC>
__global__ void __launch_bounds__(kGateThreads) se_gate_kernel(const T *__restrict__ y, int HW, int C, const float *__restrict__ w1,
                                                               const float *__restrict__ w2, int Cr, float *__restrict__ gate) {
    extern __shared__ __align__(16) float sm[];              // mean [C] | hidden [Cr] | partial [rows per pass][C] (VEC)
    float *mean = sm, *hid = sm + C, *part = sm + C + ((Cr + 3) & ~3);
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T *yb = y + (size_t)b * HW * C;
    const float inv = 1.f / (float)HW;
    if (VEC) {
        constexpr int N = Vec<T>::N;
        const int CV = C / N, rpp = kGateThreads / CV;         // threads per row, rows per pass (host: CV <= kGateThreads)
        const int cg = tid % CV, r0 = tid / CV;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        if (r0 < rpp) {
            const uint4 *base = reinterpret_cast<const uint4 *>(yb) + cg;
            int p = r0;
            for (; p + 3 * rpp < HW; p += 4 * rpp) {
                const uint4 q0 = base[(size_t)p * CV], q1 = base[(size_t)(p + rpp) * CV], q2 = base[(size_t)(p + 2 * rpp) * CV],
                            q3 = base[(size_t)(p + 3 * rpp) * CV];
                Vec<T>::add(q0, acc); Vec<T>::add(q1, acc); Vec<T>::add(q2, acc); Vec<T>::add(q3, acc);
            }
            for (; p < HW; p += rpp) Vec<T>::add(base[(size_t)p * CV], acc);
#pragma unroll
            for (int i = 0; i < N; ++i) part[(size_t)r0 * C + cg * N + i] = acc[i];
        }
        __syncthreads();
        for (int c = tid; c < C; c += kGateThreads) {
            float s = 0.f;
            for (int r = 0; r < rpp; ++r) s += part[(size_t)r * C + c];
            mean[c] = s * inv;
        }
    } else {
        for (int c = tid; c < C; c += kGateThreads) {         // any C: a thread owns a channel
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int p = 0;
            for (; p + 3 < HW; p += 4) {
                s0 += to_f<T>(yb[(size_t)p * C + c]);
                s1 += to_f<T>(yb[(size_t)(p + 1) * C + c]);
                s2 += to_f<T>(yb[(size_t)(p + 2) * C + c]);
                s3 += to_f<T>(yb[(size_t)(p + 3) * C + c]);
            }
            for (; p < HW; ++p) s0 += to_f<T>(yb[(size_t)p * C + c]);
            mean[c] = ((s0 + s1) + (s2 + s3)) * inv;
        }
    }
    __syncthreads();
    for (int o = warp; o < Cr; o += kGateThreads / 32) {      // hidden = relu(W1 mean): a warp per output
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(w1[(size_t)o * C + c], mean[c], s);
        s = warp_sum(s);
        if (lane == 0) hid[o] = fmaxf(s, 0.f);
    }
    __syncthreads();
    for (int c = tid; c < C; c += kGateThreads) {             // gate = sigmoid(W2 hidden)
        float s = 0.f;
        for (int o = 0; o < Cr; ++o) s = fmaf(w2[(size_t)c * Cr + o], hid[o], s);
        gate[(size_t)b * C + c] = 1.f / (1.f + __expf(-s));
    }
}

// elements in groups of 8 (bf16) / 4 (f32): one 16-byte load of y and of skip, one 16-byte store
__global__ void se_apply_bf16_kernel(const uint4 *__restrict__ y, const uint4 *__restrict__ skip, const float *__restrict__ gate,
                                     long long n8, int HWC8, int C8, uint4 *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / HWC8), c8 = (int)(i % C8);
        const uint4 a = y[i], s = skip[i];
        const float4 g0 = *reinterpret_cast<const float4 *>(gate + ((size_t)b * C8 + c8) * 8);
        const float4 g1 = *reinterpret_cast<const float4 *>(gate + ((size_t)b * C8 + c8) * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, sw[4] = {s.x, s.y, s.z, s.w};
        uint32_t ow[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float y0 = __uint_as_float(aw[k] << 16), y1 = __uint_as_float(aw[k] & 0xffff0000u);
            const float k0 = __uint_as_float(sw[k] << 16), k1 = __uint_as_float(sw[k] & 0xffff0000u);
            const __nv_bfloat162 r = __floats2bfloat162_rn(fmaxf(fmaf(y0, g[2 * k], k0), 0.f), fmaxf(fmaf(y1, g[2 * k + 1], k1), 0.f));
            ow[k] = *reinterpret_cast<const uint32_t *>(&r);
        }
        out[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

__global__ void se_apply_f32_kernel(const float4 *__restrict__ y, const float4 *__restrict__ skip, const float *__restrict__ gate,
                                    long long n4, int HWC4, int C4, float4 *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / HWC4), c4 = (int)(i % C4);
        const float4 a = y[i], s = skip[i];
        const float4 g = *reinterpret_cast<const float4 *>(gate + ((size_t)b * C4 + c4) * 4);
        out[i] = make_float4(fmaxf(fmaf(a.x, g.x, s.x), 0.f), fmaxf(fmaf(a.y, g.y, s.y), 0.f), fmaxf(fmaf(a.z, g.z, s.z), 0.f),
                             fmaxf(fmaf(a.w, g.w, s.w), 0.f));
    }
}

}  // namespace
}  // namespace rcnn

extern "C" int rcnn_se_gate(const void *y, int dtype, int B, int HW, int C, const float *w1, const float *w2, int Cr, float *gate,
                            rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && HW >= 1 && C >= 1 && Cr >= 1, "se_gate: bad shape B=%d HW=%d C=%d Cr=%d", B, HW, C, Cr);
    RCNN_CHECK_ARG(dtype == RCNN_F32 || dtype == RCNN_BF16, "se_gate: unsupported dtype %d", dtype);
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(y && w1 && w2 && gate, "se_gate: null pointer");
    const int N = dtype == RCNN_BF16 ? 8 : 4;
    const bool vec = C % N == 0 && C / N <= kGateThreads && ((uintptr_t)y & 15) == 0;
    const size_t smem = sizeof(float) * ((size_t)C + ((Cr + 3) & ~3) + (vec ? (size_t)(kGateThreads / (C / N)) * C : 0));
    RCNN_CHECK_ARG(smem <= 200 * 1024, "se_gate: C=%d too large", C);
    cudaStream_t st = (cudaStream_t)stream;
#define RCNN_SE_GATE(T, V)                                                                                                     \
    do {                                                                                                                       \
        if (smem > 48 * 1024)                                                                                                  \
            RCNN_CUDA(cudaFuncSetAttribute(se_gate_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        se_gate_kernel<T, V><<<B, kGateThreads, smem, st>>>((const T *)y, HW, C, w1, w2, Cr, gate);                            \
    } while (0)
    if (dtype == RCNN_F32) { if (vec) RCNN_SE_GATE(float, true); else RCNN_SE_GATE(float, false); }
    else { if (vec) RCNN_SE_GATE(__nv_bfloat16, true); else RCNN_SE_GATE(__nv_bfloat16, false); }
#undef RCNN_SE_GATE
    RCNN_LAUNCH_CHECK("se_gate_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_se_apply(const void *y, const void *skip, const float *gate, int dtype, int B, int HW, int C, void *out,
                             rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && HW >= 1 && C >= 1, "se_apply: bad shape");
    RCNN_CHECK_ARG(dtype == RCNN_F32 || dtype == RCNN_BF16, "se_apply: unsupported dtype %d", dtype);
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(y && skip && gate && out, "se_apply: null pointer");
    const int V = dtype == RCNN_BF16 ? 8 : 4;
    RCNN_CHECK_ARG(C % V == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)skip & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                       ((uintptr_t)gate & 15) == 0,
                   "se_apply: C must be a multiple of %d and the arrays 16-byte aligned", V);
    const long long n = (long long)B * HW * (C / V);
    const int blocks = (int)((n + 255) / 256 < 148LL * 8 ? (n + 255) / 256 : 148LL * 8);
    if (dtype == RCNN_BF16)
        se_apply_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4 *)y, (const uint4 *)skip, gate, n, HW * (C / V),
                                                                       C / V, (uint4 *)out);
    else
        se_apply_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)y, (const float4 *)skip, gate, n, HW * (C / V),
                                                                      C / V, (float4 *)out);
    RCNN_LAUNCH_CHECK("se_apply_kernel");
    return RCNN_OK;
}
