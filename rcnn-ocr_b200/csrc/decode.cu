// K4: greedy CTC decode (argmax over classes, collapse repeats, strip blank).
//
// Replaces the reference's training/utils.py:122-150: `logits.argmax(dim=2)` followed by
// a host loop with one .item() per frame (B*T device round trips).  Here one launch does
// both; only ids[B,T] + len[B] ever leave the device.
//
// Mapping: a group of G warps owns one sequence (G = 8 for small batches so that the
// 148 SMs are covered, G = 1 -- warp per sequence -- once the batch alone fills the chip).
// A warp works on 4 frames at once, 8 lanes per frame: the C logits of a frame are read with a
// peeled head (scalar, up to the first 16-byte boundary), a 128-bit vector body (all loads
// issued before the first compare) and a scalar tail, because a row of C=195 floats is 780 B
// and only 4-byte aligned; the argmax reduction is 3 shuffle steps shared by the 4 frames.  The per-frame argmax goes to
// shared memory; the first warp of the group then collapses the T predictions with
// ballot/popc compaction.  HBM-bound: T*C*sizeof(logit) bytes in, (T+1)*4 bytes out per
// sequence.
#include <limits.h>
#include <math.h>
#include "common.cuh"

namespace rcnn {

namespace {

constexpr int kBlockWarps = 8;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void consider(float v, int i, float &bv, int &bi, int &ni) {
    if (v > bv) { bv = v; bi = i; }
    if (v != v) ni = min(ni, i);
}

template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int VEC = 4;
    __device__ static __forceinline__ float ld(const float *p) { return __ldg(p); }
    __device__ static __forceinline__ void unpack(const uint4 &q, float (&f)[4]) {
        f[0] = __uint_as_float(q.x); f[1] = __uint_as_float(q.y);
        f[2] = __uint_as_float(q.z); f[3] = __uint_as_float(q.w);
    }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ static __forceinline__ float ld(const __nv_bfloat16 *p) {
        return __uint_as_float(((unsigned)__ldg(reinterpret_cast<const unsigned short *>(p))) << 16);
    }
    __device__ static __forceinline__ void unpack(const uint4 &q, float (&f)[8]) {
        const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
};

// torch.argmax semantics over one frame, computed by 8 consecutive lanes (a warp works on 4
// frames at once): the first maximal index wins; NaN is maximal (first NaN wins).  `row` may be
// nullptr for a lane group without a frame; all 32 lanes must call (shuffles use the full mask).
// All loads of the frame are issued before the first compare (7 x 128-bit per lane for C = 195).
template <typename T>
__device__ __forceinline__ int frame_argmax8(const T *__restrict__ row, int C, int sub, float *best_val) {
    constexpr int VEC = Elem<T>::VEC;
    constexpr int MAXV = 8;
    float bv = -INFINITY;
    int bi = INT_MAX, ni = INT_MAX;
    if (row != nullptr) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(row);
        int head = (int)(((16u - (unsigned)(a & 15u)) & 15u) / sizeof(T));
        head = min(head, C);
        const int nvec = (C - head) / VEC;
        const int tail0 = head + nvec * VEC;
        const uint4 *vp = reinterpret_cast<const uint4 *>(row + head);
        // head (< VEC elements) and tail (< VEC elements): one scalar element per lane each
        float hvl = -INFINITY, tvl = -INFINITY;
        const bool has_h = sub < head, has_t = tail0 + sub < C;
        if (has_h) hvl = Elem<T>::ld(row + sub);
        if (has_t) tvl = Elem<T>::ld(row + tail0 + sub);
        if (has_h) consider(hvl, sub, bv, bi, ni);
        for (int base = 0; base < nvec; base += 8 * MAXV) {
            uint4 q[MAXV];
#pragma unroll
            for (int j = 0; j < MAXV; ++j) {
                const int i = base + sub + 8 * j;
                if (i < nvec) q[j] = ld_nc_v4(vp + i);
            }
#pragma unroll
            for (int j = 0; j < MAXV; ++j) {
                const int i = base + sub + 8 * j;
                if (i < nvec) {
                    float f[VEC];
                    Elem<T>::unpack(q[j], f);
                    const int e0 = head + i * VEC;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) consider(f[k], e0 + k, bv, bi, ni);
                }
            }
        }
        if (has_t) consider(tvl, tail0 + sub, bv, bi, ni);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, bv, o);
        const int oi = __shfl_xor_sync(FULL, bi, o);
        const int on = __shfl_xor_sync(FULL, ni, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        ni = min(ni, on);
    }
    if (best_val) *best_val = (ni != INT_MAX) ? NAN : bv;
    return (ni != INT_MAX) ? ni : (bi == INT_MAX ? 0 : bi);
}

// sum_c exp(x_c - m) over one frame by 8 lanes (second pass, only when a confidence is requested)
template <typename T>
__device__ __forceinline__ float frame_sumexp8(const T *__restrict__ row, int C, int sub, float m) {
    float s = 0.f;
    if (row != nullptr)
        for (int c = sub; c < C; c += 8) s += ex2((Elem<T>::ld(row + c) - m) * kLog2e);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    return s;
}

__device__ __forceinline__ void group_sync(int G, int group) {
    if (G == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(G * 32) : "memory");
}

template <typename T, int G, bool CONF>
__global__ void __launch_bounds__(kBlockWarps * 32)
ctc_greedy_kernel(const T *__restrict__ logits, int B, int Tn, int C, long long sb, long long st,
                  int blank, int *__restrict__ ids, int *__restrict__ lens, float *__restrict__ conf) {
    extern __shared__ int smem_i[];
    constexpr int SEQ_PER_CTA = kBlockWarps / G;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp / G, wig = warp % G;
    const int b = blockIdx.x * SEQ_PER_CTA + group;
    int *preds = smem_i + group * Tn * (CONF ? 2 : 1);
    float *pmax = reinterpret_cast<float *>(preds + Tn);
    if (b < B) {
        const T *seq = logits + (long long)b * sb;
        const int sub = lane & 7, fr = lane >> 3;        // 8 lanes per frame, 4 frames per warp pass
        for (int t0 = wig * 4; t0 < Tn; t0 += G * 4) {
            const int t = t0 + fr;
            const T *row = t < Tn ? seq + (long long)t * st : nullptr;
            float bv;
            const int p = frame_argmax8<T>(row, C, sub, &bv);
            float pm = 0.f;
            if (CONF) pm = 1.f / frame_sumexp8<T>(row, C, sub, bv);
            if (sub == 0 && t < Tn) {
                preds[t] = p;
                if (CONF) pmax[t] = pm;
            }
        }
    }
    group_sync(G, group);
    if (b >= B || wig != 0) return;
    // collapse: keep frame t iff pred != blank and pred != pred[t-1] (prev starts as blank)
    int *out = ids + (long long)b * Tn;
    int n = 0, nvalid = 0;
    float csum = 0.f;
    for (int t0 = 0; t0 < Tn; t0 += 32) {
        const int t = t0 + lane;
        const int p = t < Tn ? preds[t] : blank;
        const int prev = (t == 0 || t >= Tn) ? blank : preds[t - 1];
        const bool nonblank = t < Tn && p != blank;
        const bool keep = nonblank && p != prev;
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) out[n + __popc(m & ((1u << lane) - 1u))] = p;
        n += __popc(m);
        if (CONF) {
            csum += nonblank ? pmax[t] : 0.f;
            nvalid += __popc(__ballot_sync(FULL, nonblank));
        }
    }
    for (int t = n + lane; t < Tn; t += 32) out[t] = -1;
    if (CONF) csum = warp_sum(csum);
    if (lane == 0) {
        lens[b] = n;
        if (CONF) conf[b] = nvalid > 0 ? csum / (float)nvalid : 0.f;
    }
}

template <typename T, int G>
int launch_greedy(const void *logits, int B, int Tn, int C, long long sb, long long st, int blank,
                  int *ids, int *lens, float *conf, cudaStream_t stream) {
    constexpr int SEQ_PER_CTA = kBlockWarps / G;
    const int grid = (B + SEQ_PER_CTA - 1) / SEQ_PER_CTA;
    const size_t smem = (size_t)SEQ_PER_CTA * Tn * sizeof(int) * (conf ? 2 : 1);
    RCNN_CHECK_ARG(smem <= 48 * 1024, "ctc_greedy: T=%d too long for the prediction buffer", Tn);
    ProfScope prof(RCNN_K_DECODE, stream);
    if (conf)
        ctc_greedy_kernel<T, G, true><<<grid, kBlockWarps * 32, smem, stream>>>(
            (const T *)logits, B, Tn, C, sb, st, blank, ids, lens, conf);
    else
        ctc_greedy_kernel<T, G, false><<<grid, kBlockWarps * 32, smem, stream>>>(
            (const T *)logits, B, Tn, C, sb, st, blank, ids, lens, conf);
    RCNN_LAUNCH_CHECK("ctc_greedy_kernel");
    return RCNN_OK;
}

template <typename T>
int dispatch_greedy(const void *logits, int B, int Tn, int C, long long sb, long long st, int blank,
                    int *ids, int *lens, float *conf, cudaStream_t stream) {
    // warps per sequence: keep >= ~16 warps per SM in flight; a sequence never gets more
    // warps than it has frames.
    const long long warps_wanted = 16LL * num_sms();
    int G = 8;
    while (G > 1 && ((long long)B * G > 2 * warps_wanted || G * 4 > Tn + 3)) G >>= 1;
    switch (G) {
        case 8: return launch_greedy<T, 8>(logits, B, Tn, C, sb, st, blank, ids, lens, conf, stream);
        case 4: return launch_greedy<T, 4>(logits, B, Tn, C, sb, st, blank, ids, lens, conf, stream);
        case 2: return launch_greedy<T, 2>(logits, B, Tn, C, sb, st, blank, ids, lens, conf, stream);
        default: return launch_greedy<T, 1>(logits, B, Tn, C, sb, st, blank, ids, lens, conf, stream);
    }
}

}  // namespace
}  // namespace rcnn

extern "C" int rcnn_ctc_greedy(const void *logits, int dtype, int B, int T, int C,
                               int64_t stride_b, int64_t stride_t, int blank,
                               int32_t *ids_out, int32_t *len_out, float *conf_out,
                               rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 0 && C > 0, "ctc_greedy: bad shape B=%d T=%d C=%d", B, T, C);
    RCNN_CHECK_ARG(dtype == RCNN_F32 || dtype == RCNN_BF16, "ctc_greedy: unsupported dtype %d", dtype);
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(len_out && ((logits && ids_out) || T == 0), "ctc_greedy: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    if (T == 0) {
        RCNN_CUDA(cudaMemsetAsync(len_out, 0, sizeof(int32_t) * (size_t)B, s));
        if (conf_out) RCNN_CUDA(cudaMemsetAsync(conf_out, 0, sizeof(float) * (size_t)B, s));
        return RCNN_OK;
    }
    if (dtype == RCNN_F32)
        return dispatch_greedy<float>(logits, B, T, C, stride_b, stride_t, blank, ids_out, len_out, conf_out, s);
    return dispatch_greedy<__nv_bfloat16>(logits, B, T, C, stride_b, stride_t, blank, ids_out, len_out, conf_out, s);
}
