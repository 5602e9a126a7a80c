// Backbone, inference ("fused SE-scale epilogues"): the squeeze-and-excitation tail of an SE-ResNet block
// (model/seresnet31.py: SELayer + the residual add + ReLU of the block),
//     gate = sigmoid(W2 relu(W1 mean_hw(y)));   out = relu(y * gate + skip)
// as two launches instead of the eight elementwise / reduction / tiny-GEMM launches torch issues for it:
//   se_gate_kernel   S CTAs per image sum their share of the H*W pixel rows per channel; the last one to finish (a counter per image)
//                    adds the partial sums up, forms the two small products and the sigmoid -> gate [B, C]
//   se_apply_kernel  all elements, 16 bytes per thread: out = relu((y + ybias) * gate + skip + sbias)
// ybias / sbias (optional, [C] f32) are the biases of the convolutions that produced y / skip: folded BatchNorm shifts, which
// cuDNN would otherwise add in a separate pass over the tensor.  Also here: the stem's 2 x 2 max pooling (torch's channels_last
// kernel ran at 0.9 TB/s).  Tensors are channels_last ([B, H*W, C] in memory, C contiguous), bf16 or f32; W1 [Cr, C] and W2 TRANSPOSED [Cr, C], f32.
#include <stdlib.h>
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

// grid (S, B), kGateThreads threads.  VEC: the pixel rows are read 16 bytes per thread, C / N threads per row and
// kGateThreads / (C / N) rows per pass, four passes in flight (a thread per channel walking the rows one by one is a chain of
// H*W dependent L2 round trips); partial sums meet in shared memory, the S slices of an image in `part_g` [B, S, C]; `counter` [B]
// is zero on entry and on exit.
constexpr int kGateThreads = 256;
template <typename T, bool VEC>
__global__ void __launch_bounds__(kGateThreads) se_gate_kernel(const T *__restrict__ y, int HW, int C, const float *__restrict__ w1,
                                                               const float *__restrict__ w2, int Cr, const float *__restrict__ ybias,
                                                               float *__restrict__ gate, float *__restrict__ part_g,
                                                               unsigned int *__restrict__ counter) {
    extern __shared__ __align__(16) float sm[];              // mean [C] | hidden [Cr] | partial [rows per pass][C] (VEC)
    float *mean = sm, *hid = sm + C, *part = sm + C + ((Cr + 3) & ~3);
    __shared__ int is_last;
    const int S = gridDim.x, sidx = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T *yb = y + (size_t)b * HW * C;
    const int p_lo = (int)((long long)HW * sidx / S), p_hi = (int)((long long)HW * (sidx + 1) / S);
    float *mine = part_g + ((size_t)b * S + sidx) * C;
    if (VEC) {
        constexpr int N = Vec<T>::N;
        const int CV = C / N, rpp = kGateThreads / CV;         // threads per row, rows per pass (host: CV <= kGateThreads)
        const int cg = tid % CV, r0 = tid / CV;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        if (r0 < rpp) {
            const uint4 *base = reinterpret_cast<const uint4 *>(yb) + cg;
            int p = p_lo + r0;
            for (; p + 3 * rpp < p_hi; p += 4 * rpp) {
                const uint4 q0 = base[(size_t)p * CV], q1 = base[(size_t)(p + rpp) * CV], q2 = base[(size_t)(p + 2 * rpp) * CV],
                            q3 = base[(size_t)(p + 3 * rpp) * CV];
                Vec<T>::add(q0, acc); Vec<T>::add(q1, acc); Vec<T>::add(q2, acc); Vec<T>::add(q3, acc);
            }
            for (; p < p_hi; p += rpp) Vec<T>::add(base[(size_t)p * CV], acc);
#pragma unroll
            for (int i = 0; i < N; ++i) part[(size_t)r0 * C + cg * N + i] = acc[i];
        }
        __syncthreads();
        for (int c = tid; c < C; c += kGateThreads) {
            float s = 0.f;
            for (int r = 0; r < rpp; ++r) s += part[(size_t)r * C + c];
            mine[c] = s;
        }
    } else {
        for (int c = tid; c < C; c += kGateThreads) {         // any C: a thread owns a channel
            float s0 = 0.f, s1 = 0.f;
            int p = p_lo;
            for (; p + 1 < p_hi; p += 2) {
                s0 += to_f<T>(yb[(size_t)p * C + c]);
                s1 += to_f<T>(yb[(size_t)(p + 1) * C + c]);
            }
            if (p < p_hi) s0 += to_f<T>(yb[(size_t)p * C + c]);
            mine[c] = s0 + s1;
        }
    }
    // the last CTA of this image to arrive finishes the job (the partial sums of the others are visible to it: fence, then count)
    if (S > 1) {
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = atomicAdd(&counter[b], 1u) == (unsigned)(S - 1);
        __syncthreads();
        if (!is_last) return;
        __threadfence();
    } else {
        __syncthreads();                                      // one slice: this CTA's own sums, no hand-over
    }
    const float inv = 1.f / (float)HW;
    // every loop below requests its (up to eight) loads before the first use: unrolled, they cost one L2 round trip each
    // instead of one per element (the finish was 2/3 of this kernel's 24 us)
    for (int c = tid; c < C; c += kGateThreads) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = k < S ? __ldcg(part_g + ((size_t)b * S + k) * C + c) : 0.f;   // (host: S <= 8)
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k];                // fixed order: deterministic
        mean[c] = s * inv + (ybias ? ybias[c] : 0.f);
    }
    if (tid == 0 && S > 1) counter[b] = 0u;                   // ready for the next launch
    __syncthreads();
    // hidden = relu(W1 mean): a warp forms FOUR outputs at a time, 32 loads per lane requested before the first use (with one
    // output per warp and eight loads in flight the finish of a C = 512, Cr = 32 block was 13 of the kernel's 16 us)
    for (int ob = warp * 4; ob < Cr; ob += (kGateThreads / 32) * 4) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c0 = lane; c0 < C; c0 += 256) {
            float v[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    v[q][u] = (ob + q < Cr && c0 + 32 * u < C) ? __ldg(w1 + (size_t)(ob + q) * C + c0 + 32 * u) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float m = c0 + 32 * u < C ? mean[c0 + 32 * u] : 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) s4[q] = fmaf(v[q][u], m, s4[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float r = warp_sum(s4[q]);
            if (lane == 0 && ob + q < Cr) hid[ob + q] = fmaxf(r, 0.f);
        }
    }
    __syncthreads();
    // gate = sigmoid(W2 hidden).  W2 arrives TRANSPOSED ([Cr, C]): neighbouring threads read neighbouring channels of one row
    // (a thread walking its own row of W2 [C, Cr] made every load 32 separate lines: 16 k L1 wavefronts, 8 of the kernel's 15 us
    // at C = 512, Cr = 32)
    for (int c = tid; c < C; c += kGateThreads) {
        float s = 0.f;
        for (int o0 = 0; o0 < Cr; o0 += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = o0 + u < Cr ? __ldg(w2 + (size_t)(o0 + u) * C + c) : 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u) s = fmaf(v[u], o0 + u < Cr ? hid[o0 + u] : 0.f, s);
        }
        gate[(size_t)b * C + c] = 1.f / (1.f + __expf(-s));
    }
}

// elements in groups of 8 (bf16) / 4 (f32): one 16-byte load of y and of skip, one 16-byte store
__global__ void se_apply_bf16_kernel(const uint4 *__restrict__ y, const uint4 *__restrict__ skip, const float *__restrict__ gate,
                                     const float *__restrict__ ybias, const float *__restrict__ sbias, long long n8, int HWC8, int C8,
                                     uint4 *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / HWC8), c8 = (int)(i % C8);
        const uint4 a = y[i], s = skip[i];
        const float4 g0 = *reinterpret_cast<const float4 *>(gate + ((size_t)b * C8 + c8) * 8);
        const float4 g1 = *reinterpret_cast<const float4 *>(gate + ((size_t)b * C8 + c8) * 8 + 4);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float yb[8], sb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { yb[k] = ybias ? ybias[c8 * 8 + k] : 0.f; sb[k] = sbias ? sbias[c8 * 8 + k] : 0.f; }
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, sw[4] = {s.x, s.y, s.z, s.w};
        uint32_t ow[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float y0 = __uint_as_float(aw[k] << 16) + yb[2 * k], y1 = __uint_as_float(aw[k] & 0xffff0000u) + yb[2 * k + 1];
            const float k0 = __uint_as_float(sw[k] << 16) + sb[2 * k], k1 = __uint_as_float(sw[k] & 0xffff0000u) + sb[2 * k + 1];
            const __nv_bfloat162 r = __floats2bfloat162_rn(fmaxf(fmaf(y0, g[2 * k], k0), 0.f), fmaxf(fmaf(y1, g[2 * k + 1], k1), 0.f));
            ow[k] = *reinterpret_cast<const uint32_t *>(&r);
        }
        out[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

__global__ void se_apply_f32_kernel(const float4 *__restrict__ y, const float4 *__restrict__ skip, const float *__restrict__ gate,
                                    const float *__restrict__ ybias, const float *__restrict__ sbias, long long n4, int HWC4, int C4,
                                    float4 *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / HWC4), c4 = (int)(i % C4);
        float4 a = y[i], s = skip[i];
        const float4 g = *reinterpret_cast<const float4 *>(gate + ((size_t)b * C4 + c4) * 4);
        if (ybias) { const float4 t = *reinterpret_cast<const float4 *>(ybias + c4 * 4); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
        if (sbias) { const float4 t = *reinterpret_cast<const float4 *>(sbias + c4 * 4); s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        out[i] = make_float4(fmaxf(fmaf(a.x, g.x, s.x), 0.f), fmaxf(fmaf(a.y, g.y, s.y), 0.f), fmaxf(fmaf(a.z, g.z, s.z), 0.f),
                             fmaxf(fmaf(a.w, g.w, s.w), 0.f));
    }
}

// 2 x 2 max pooling, stride 2, channels_last: a thread per output pixel and 16-byte channel group
template <typename T>
__global__ void maxpool2x2_kernel(const uint4 *__restrict__ x, int H, int W, int CV, long long n, uint4 *__restrict__ out) {
    const int Ho = H / 2, Wo = W / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long long r = i / CV;
        const int wo = (int)(r % Wo); r /= Wo;
        const int ho = (int)(r % Ho);
        const long long b = r / Ho;
        const uint4 *p = x + ((b * H + 2 * ho) * W + 2 * wo) * CV + cv;
        const uint4 q[4] = {p[0], p[CV], p[(size_t)W * CV], p[(size_t)W * CV + CV]};
        uint4 o;
        if (sizeof(T) == 2) {
            uint32_t w[4];
            const uint32_t *a0 = reinterpret_cast<const uint32_t *>(&q[0]), *a1 = reinterpret_cast<const uint32_t *>(&q[1]),
                           *a2 = reinterpret_cast<const uint32_t *>(&q[2]), *a3 = reinterpret_cast<const uint32_t *>(&q[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t m01, m23;
                asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(m01) : "r"(a0[k]), "r"(a1[k]));
                asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(m23) : "r"(a2[k]), "r"(a3[k]));
                asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(w[k]) : "r"(m01), "r"(m23));
            }
            o = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            const float *a0 = reinterpret_cast<const float *>(&q[0]), *a1 = reinterpret_cast<const float *>(&q[1]),
                        *a2 = reinterpret_cast<const float *>(&q[2]), *a3 = reinterpret_cast<const float *>(&q[3]);
            float w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float m01, m23;
                asm("max.NaN.f32 %0, %1, %2;" : "=f"(m01) : "f"(a0[k]), "f"(a1[k]));
                asm("max.NaN.f32 %0, %1, %2;" : "=f"(m23) : "f"(a2[k]), "f"(a3[k]));
                asm("max.NaN.f32 %0, %1, %2;" : "=f"(w[k]) : "f"(m01), "f"(m23));
            }
            o = make_uint4(__float_as_uint(w[0]), __float_as_uint(w[1]), __float_as_uint(w[2]), __float_as_uint(w[3]));
        }
        out[i] = o;
    }
}

}  // namespace
}  // namespace rcnn

extern "C" size_t rcnn_se_gate_workspace_bytes(int B, int C) {
    if (B <= 0 || C <= 0) return 0;
    return ((size_t)B * 8 * C + (size_t)B) * 4;              // partial sums of up to 8 slices per image + a counter per image
}

extern "C" int rcnn_se_gate(const void *y, int dtype, int B, int HW, int C, const float *w1, const float *w2, int Cr,
                            const float *ybias, float *gate, void *workspace, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && HW >= 1 && C >= 1 && Cr >= 1 && B <= 65535, "se_gate: bad shape B=%d HW=%d C=%d Cr=%d", B, HW, C, Cr);
    RCNN_CHECK_ARG(dtype == RCNN_F32 || dtype == RCNN_BF16, "se_gate: unsupported dtype %d", dtype);
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(y && w1 && w2 && gate && workspace, "se_gate: null pointer");
    const int N = dtype == RCNN_BF16 ? 8 : 4;
    const bool vec = C % N == 0 && C / N <= kGateThreads && ((uintptr_t)y & 15) == 0;
    int S = (2 * num_sms() + B - 1) / B;                      // about two CTAs per SM over the batch
    S = S > 4 ? 4 : S;                                        // (eight slices: slower, more hand-overs than work)
    S = S > HW ? HW : S;
    S = S < 1 ? 1 : S;
    static const int force_s = getenv("RCNN_SE_SLICES") ? atoi(getenv("RCNN_SE_SLICES")) : 0;
    if (force_s >= 1 && force_s <= 8) S = force_s > HW ? HW : force_s;
    const size_t smem = sizeof(float) * ((size_t)C + ((Cr + 3) & ~3) + (vec ? (size_t)(kGateThreads / (C / N)) * C : 0));
    RCNN_CHECK_ARG(smem <= 200 * 1024, "se_gate: C=%d too large", C);
    float *part = (float *)workspace;
    unsigned int *counter = (unsigned int *)(part + (size_t)B * 8 * C);
    cudaStream_t st = (cudaStream_t)stream;
#define RCNN_SE_GATE(T, V)                                                                                                     \
    do {                                                                                                                       \
        if (smem > 48 * 1024)                                                                                                  \
            RCNN_CUDA(cudaFuncSetAttribute(se_gate_kernel<T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        se_gate_kernel<T, V><<<dim3((unsigned)S, (unsigned)B), kGateThreads, smem, st>>>((const T *)y, HW, C, w1, w2, Cr, ybias,  \
                                                                                         gate, part, counter);                  \
    } while (0)
    if (dtype == RCNN_F32) { if (vec) RCNN_SE_GATE(float, true); else RCNN_SE_GATE(float, false); }
    else { if (vec) RCNN_SE_GATE(__nv_bfloat16, true); else RCNN_SE_GATE(__nv_bfloat16, false); }
#undef RCNN_SE_GATE
    RCNN_LAUNCH_CHECK("se_gate_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_se_apply(const void *y, const void *skip, const float *gate, const float *ybias, const float *sbias, int dtype,
                             int B, int HW, int C, void *out, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && HW >= 1 && C >= 1, "se_apply: bad shape");
    RCNN_CHECK_ARG(dtype == RCNN_F32 || dtype == RCNN_BF16, "se_apply: unsupported dtype %d", dtype);
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(y && skip && gate && out, "se_apply: null pointer");
    const int V = dtype == RCNN_BF16 ? 8 : 4;
    RCNN_CHECK_ARG(C % V == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)skip & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                       ((uintptr_t)gate & 15) == 0 && ((uintptr_t)ybias & 15) == 0 && ((uintptr_t)sbias & 15) == 0,
                   "se_apply: C must be a multiple of %d and the arrays 16-byte aligned", V);
    const long long n = (long long)B * HW * (C / V);
    const int blocks = (int)((n + 255) / 256 < 148LL * 8 ? (n + 255) / 256 : 148LL * 8);
    if (dtype == RCNN_BF16)
        se_apply_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4 *)y, (const uint4 *)skip, gate, ybias, sbias, n,
                                                                       HW * (C / V), C / V, (uint4 *)out);
    else
        se_apply_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)y, (const float4 *)skip, gate, ybias, sbias, n,
                                                                      HW * (C / V), C / V, (float4 *)out);
    RCNN_LAUNCH_CHECK("se_apply_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_maxpool2x2_nhwc(const void *x, int dtype, int B, int H, int W, int C, void *out, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && H >= 2 && W >= 2 && C >= 1 && H % 2 == 0 && W % 2 == 0, "maxpool2x2: bad shape (H, W even)");
    RCNN_CHECK_ARG(dtype == RCNN_F32 || dtype == RCNN_BF16, "maxpool2x2: unsupported dtype %d", dtype);
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(x && out, "maxpool2x2: null pointer");
    const int V = dtype == RCNN_BF16 ? 8 : 4;
    RCNN_CHECK_ARG(C % V == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0, "maxpool2x2: C %% %d and 16-byte alignment", V);
    const long long n = (long long)B * (H / 2) * (W / 2) * (C / V);
    const int blocks = (int)((n + 255) / 256 < 148LL * 16 ? (n + 255) / 256 : 148LL * 16);
    if (dtype == RCNN_BF16)
        maxpool2x2_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4 *)x, H, W, C / V, n, (uint4 *)out);
    else
        maxpool2x2_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4 *)x, H, W, C / V, n, (uint4 *)out);
    RCNN_LAUNCH_CHECK("maxpool2x2_kernel");
    return RCNN_OK;
}
