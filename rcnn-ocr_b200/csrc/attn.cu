// K6: per-step kernels of the reference's attention decoder (model/model.py:23-148), inference path.
//
// One decoding step of AttentionCell.forward + Attention._greedy_decode (model/model.py:33-47, 100-108):
//     proj_h  = h2h(h_{t-1})                                   -> tcgen05 GEMM (gemm.cu)
//     e       = score(tanh(proj_H + proj_h)),  alpha = softmax_T(e),  context = alpha^T batch_H   (K6a)
//     gates   = [context, h_{t-1}] [W_ih[:, :C] | W_hh]^T + b_ih + b_hh   -> tcgen05 GEMM
//               + W_ih[:, C + y_{t-1}]  (the one-hot half of the LSTMCell input is a column gather)
//     c, h    = LSTMCell pointwise                                                                 (K6b)
//     logits  = generator(h)                                   -> tcgen05 GEMM
//     logits[:, blank] = -1e4;  probs[:, t] = logits;  y_t = argmax                                (K6c)
// proj_H = i2h(batch_H) does not depend on the step and is hoisted out of the loop (the reference
// recomputes it every step, model/model.py:35).  Sequences are independent: no exchange between CTAs.
#include "common.cuh"
#include "sm100.cuh"

namespace rcnn {
namespace {

using namespace sm100;

__device__ __forceinline__ float tanh_fast_a(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sigmoid_fast_a(float x) { return fmaf(tanh_fast_a(0.5f * x), 0.5f, 0.5f); }

// K6a: one CTA per sequence.  proj_H [B,T,H] f32, proj_h [B,H] f32, v [H] f32, enc [B,T,C] f32 (strided rows)
// -> alpha [B,T] f32 (optional), context as bf16 into xcat[b, 0:C] (row pitch ldx elements)
__global__ void __launch_bounds__(256) attn_score_context_kernel(
    const float *__restrict__ projH, const float *__restrict__ projh, const float *__restrict__ v,
    const float *__restrict__ enc, long long enc_sb, long long enc_st, int T, int H, int C,
    float *__restrict__ alpha_out, __nv_bfloat16 *__restrict__ xcat, long long ldx, long long projh_ld) {
    extern __shared__ __align__(16) float sm[];
    float *ph = sm, *vs = sm + H, *e = sm + 2 * H;          // [H], [H], [T]
    __shared__ float red[2];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = threadIdx.x; j < H; j += blockDim.x) { ph[j] = projh[(size_t)b * projh_ld + j]; vs[j] = v[j]; }
    __syncthreads();
    const bool vec4 = (H & 3) == 0 && (C & 3) == 0 && (enc_st & 3) == 0 && (enc_sb & 3) == 0 &&
                      ((uintptr_t)projH & 15) == 0 && ((uintptr_t)enc & 15) == 0;
    for (int t0 = warp; t0 < T; t0 += 2 * nw) {                // two frames per warp pass: their rows are requested together
        const int t1 = t0 + nw;
        const float *row0 = projH + ((size_t)b * T + t0) * H;
        const float *row1 = projH + ((size_t)b * T + (t1 < T ? t1 : t0)) * H;
        float s0 = 0.f, s1 = 0.f;
        if (vec4) {
            // 128-bit loads, eight of them requested before the first value is used (the kernel is bound by L2 latency: 16
            // KB in flight per SM gave 4 TB/s over all SMs)
            for (int j0 = 4 * lane; j0 < H; j0 += 512) {
                float4 a[4], c4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 128 * u;
                    a[u] = j < H ? *reinterpret_cast<const float4 *>(row0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    c4[u] = j < H ? *reinterpret_cast<const float4 *>(row1 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 128 * u;
                    if (j < H) {
                        const float4 p4 = *reinterpret_cast<const float4 *>(ph + j), v4 = *reinterpret_cast<const float4 *>(vs + j);
                        s0 = fmaf(v4.x, tanh_fast_a(a[u].x + p4.x), s0); s0 = fmaf(v4.y, tanh_fast_a(a[u].y + p4.y), s0);
                        s0 = fmaf(v4.z, tanh_fast_a(a[u].z + p4.z), s0); s0 = fmaf(v4.w, tanh_fast_a(a[u].w + p4.w), s0);
                        s1 = fmaf(v4.x, tanh_fast_a(c4[u].x + p4.x), s1); s1 = fmaf(v4.y, tanh_fast_a(c4[u].y + p4.y), s1);
                        s1 = fmaf(v4.z, tanh_fast_a(c4[u].z + p4.z), s1); s1 = fmaf(v4.w, tanh_fast_a(c4[u].w + p4.w), s1);
                    }
                }
            }
        } else {
            for (int j = lane; j < H; j += 32) {
                s0 = fmaf(vs[j], tanh_fast_a(row0[j] + ph[j]), s0);
                s1 = fmaf(vs[j], tanh_fast_a(row1[j] + ph[j]), s1);
            }
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (lane == 0) { e[t0] = s0; if (t1 < T) e[t1] = s1; }
    }
    __syncthreads();
    if (warp == 0) {                                          // softmax over the T encoder frames
        float m = -INFINITY;
        for (int t = lane; t < T; t += 32) m = fmaxf(m, e[t]);
        m = warp_max(m);
        float z = 0.f;
        for (int t = lane; t < T; t += 32) z += __expf(e[t] - m);
        z = warp_sum(z);
        if (lane == 0) { red[0] = m; red[1] = 1.f / z; }
    }
    __syncthreads();
    const float m = red[0], iz = red[1];
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const float a = __expf(e[t] - m) * iz;
        e[t] = a;
        if (alpha_out) alpha_out[(size_t)b * T + t] = a;
    }
    __syncthreads();
    const float *eb = enc + (size_t)b * enc_sb;
    if (vec4 && (ldx & 3) == 0 && ((uintptr_t)xcat & 7) == 0) {
        for (int c = 4 * threadIdx.x; c < C; c += 4 * blockDim.x) {       // four columns per thread, four frames in flight
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int t = 0;
            for (; t + 7 < T; t += 8) {
                float4 r[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) r[k] = *reinterpret_cast<const float4 *>(eb + (size_t)(t + k) * enc_st + c);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float a = e[t + k];
                    acc.x = fmaf(a, r[k].x, acc.x); acc.y = fmaf(a, r[k].y, acc.y);
                    acc.z = fmaf(a, r[k].z, acc.z); acc.w = fmaf(a, r[k].w, acc.w);
                }
            }
            for (; t < T; ++t) {
                const float4 r = *reinterpret_cast<const float4 *>(eb + (size_t)t * enc_st + c);
                const float a = e[t];
                acc.x = fmaf(a, r.x, acc.x); acc.y = fmaf(a, r.y, acc.y); acc.z = fmaf(a, r.z, acc.z); acc.w = fmaf(a, r.w, acc.w);
            }
            const __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t *>(&lo);
            o.y = *reinterpret_cast<const uint32_t *>(&hi);
            *reinterpret_cast<uint2 *>(xcat + (size_t)b * ldx + c) = o;
        }
    } else {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float acc = 0.f;
            for (int t = 0; t < T; ++t) acc = fmaf(e[t], eb[(size_t)t * enc_st + c], acc);
            xcat[(size_t)b * ldx + c] = __float2bfloat16_rn(acc);
        }
    }
}

// mask + copy + argmax of one logits row by one warp (K6c's body; also run by the bf16 step kernel for the previous step)
__device__ __forceinline__ int warp_argmax_row(const float *__restrict__ row, int V, int blank, float *__restrict__ probs_row,
                                               int lane) {
    float best = -INFINITY;
    int arg = 0x7fffffff;
    bool nan = false;
    for (int k0 = lane; k0 < V; k0 += 256) {                 // eight loads requested before the first compare
        float xs[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xs[u] = k0 + 32 * u < V ? row[k0 + 32 * u] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k0 + 32 * u;
            if (k < V) {
                float x = xs[u];
                if (k == blank) x = -1e4f;
                if (probs_row) probs_row[k] = x;
                const bool xn = x != x;
                if (!nan && (xn || x > best || arg == 0x7fffffff)) { best = x; arg = k; nan = xn; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(FULL, best, o);
        const int oa = __shfl_xor_sync(FULL, arg, o);
        const bool on = __shfl_xor_sync(FULL, (int)nan, o) != 0;
        bool take;
        if (nan != on) take = on;                                 // a NaN beats any number
        else if (nan) take = oa < arg;                            // two NaNs: the first index
        else take = ob > best || (ob == best && oa < arg);
        if (take) { best = ob; arg = oa; nan = on; }
    }
    return arg == 0x7fffffff ? 0 : arg;
}

// K6a with bf16 operands: proj_H [B,T,H] and enc [B,T,C] as bf16 (half the bytes of the step's only large reads; both
// were rounded to bf16 on their way through the tensor cores already: proj_H is a bf16-operand product, the context is
// the bf16 A operand of the gates GEMM).  Eight warps; warp w owns the frames t = w (mod 8) in both passes, a lane reads 16
// bytes (eight values) of up to eight frames before it uses the first.  The context pass leaves per-warp partial sums
// in shared memory, added up in frame order 0..7 by the thread that owns the column (deterministic).
__device__ __forceinline__ void unpack8(const uint4 &q, float (&f)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// One work item of the score pass: eight frames (t0 + 8k) x eight hidden units (j ..) of proj_H, 16 bytes per frame
__device__ __forceinline__ void score_item_load(const __nv_bfloat16 *__restrict__ pb, int T, int H, int t0, int j, uint4 (&q)[8]) {
    if (j >= H) return;                                       // (score_item_fma skips the item as well)
    const __nv_bfloat16 *p = pb + j;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int t = min(t0 + 8 * k, T - 1);                 // past the end: a duplicate row whose score is not stored
        q[k] = *reinterpret_cast<const uint4 *>(p + (size_t)t * H);
    }
}
__device__ __forceinline__ void score_item_fma(const float *ph, const float *vs, int H, int j, const uint4 (&q)[8], float (&s)[8]) {
    if (j >= H) return;
    float p8[8], v8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { p8[i] = ph[j + i]; v8[i] = vs[j + i]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float f[8];
        unpack8(q[k], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[k] = fmaf(v8[i], tanh_fast_a(f[i] + p8[i]), s[k]);
    }
}

// STAGE: the sequence's encoder rows [T, C] are copied to shared memory with cp.async.bulk while the score pass runs (they do
// not depend on it), so that the context pass starts from on-chip data; the kernel is a chain of L2 latencies, not bytes.
template <bool STAGE>
__global__ void __launch_bounds__(256) attn_score_context_bf16_kernel(
    const __nv_bfloat16 *__restrict__ projH, const float *__restrict__ projh, const float *__restrict__ v,
    const __nv_bfloat16 *__restrict__ enc, long long enc_sb, long long enc_st, int T, int H, int C,
    float *__restrict__ alpha_out, __nv_bfloat16 *__restrict__ xcat, long long ldx, long long projh_ld,
    const float *__restrict__ prev_logits, long long prev_ld, int V, int blank, float *__restrict__ prev_probs,
    long long probs_ld, long long *__restrict__ y, const float *__restrict__ alpha_scale) {
    extern __shared__ __align__(16) float sm[];
    float *ph = sm, *vs = sm + H, *e = sm + 2 * H, *part = e + ((T + 3) & ~3);      // [H], [H], [T], [8][C]
    __nv_bfloat16 *enc_s = reinterpret_cast<__nv_bfloat16 *>(part + 8 * (size_t)C);  // [T][C] (STAGE)
    __shared__ float red[2];
    __shared__ __align__(8) uint64_t stage_bar;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __nv_bfloat16 *pb = projH + (size_t)b * T * H;
    const __nv_bfloat16 *eb = enc + (size_t)b * enc_sb;

    griddep_launch();
    // score pass, work items (frame block, unit block) double-buffered: item i+1 is requested before item i is used
    const int nJ = (H + 255) / 256, nTb = warp < T ? (T - warp + 63) / 64 : 0, nI = nTb * nJ;
    uint4 qa[8], qb[8];
    if (nI > 0) score_item_load(pb, T, H, warp, 8 * lane, qa);
    if (STAGE && threadIdx.x == 0) {                          // bulk copies (one if the rows are contiguous), one barrier
        mbar_init(&stage_bar, 1);
        fence_barrier_init();
        mbar_arrive_expect_tx(&stage_bar, (uint32_t)T * (uint32_t)C * 2u);
        const uint32_t dst0 = smem_u32(enc_s), bar = smem_u32(&stage_bar);
        const int rows = enc_st == C ? 1 : T;
        const uint32_t bytes = enc_st == C ? (uint32_t)T * (uint32_t)C * 2u : (uint32_t)C * 2u;
        for (int t = 0; t < rows; ++t)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             dst0 + (uint32_t)t * bytes), "l"(eb + (size_t)t * enc_st), "r"(bytes), "r"(bar) : "memory");
    }
    // chained launches: proj_H, enc and v above are the loop's constants; proj_h and the logits come from the predecessor
    for (int j = threadIdx.x; j < H; j += 256) vs[j] = v[j];
    griddep_wait();
    for (int j = threadIdx.x; j < H; j += 256) ph[j] = projh[(size_t)b * projh_ld + j];
    if (prev_logits != nullptr && warp == 7) {                // K6c of the previous step for this sequence (saves a launch)
        const int arg = warp_argmax_row(prev_logits + (size_t)b * prev_ld, V, blank,
                                        prev_probs ? prev_probs + (size_t)b * probs_ld : nullptr, lane);
        if (lane == 0 && y) y[b] = arg;
    }
    __syncthreads();
    {
        float s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = 0.f;
        auto finish = [&](int it) {                           // last unit block of a frame block: reduce, store, reset
            if (it % nJ != nJ - 1) return;
            const int t0 = warp + 64 * (it / nJ);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float r = warp_sum(s[k]);
                if (lane == 0 && t0 + 8 * k < T) e[t0 + 8 * k] = r;
                s[k] = 0.f;
            }
        };
        for (int it = 0; it < nI; it += 2) {
            if (it + 1 < nI) score_item_load(pb, T, H, warp + 64 * ((it + 1) / nJ), 8 * lane + 256 * ((it + 1) % nJ), qb);
            score_item_fma(ph, vs, H, 8 * lane + 256 * (it % nJ), qa, s);
            finish(it);
            if (it + 2 < nI) score_item_load(pb, T, H, warp + 64 * ((it + 2) / nJ), 8 * lane + 256 * ((it + 2) % nJ), qa);
            if (it + 1 < nI) {
                score_item_fma(ph, vs, H, 8 * lane + 256 * ((it + 1) % nJ), qb, s);
                finish(it + 1);
            }
        }
    }
    __syncthreads();
    if (warp == 0) {                                          // softmax over the T encoder frames
        float m = -INFINITY;
        for (int t = lane; t < T; t += 32) m = fmaxf(m, e[t]);
        m = warp_max(m);
        float z = 0.f;
        for (int t = lane; t < T; t += 32) z += __expf(e[t] - m);
        z = warp_sum(z);
        if (lane == 0) { red[0] = m; red[1] = 1.f / z; }
    }
    __syncthreads();
    if (STAGE) mbar_wait(&stage_bar, 0);
    const float m = red[0], iz = red[1];
    for (int t = threadIdx.x; t < T; t += 256) {
        const float a = __expf(e[t] - m) * iz;
        if (alpha_out) alpha_out[(size_t)b * T + t] = a;
        // training: F.dropout(alpha) (model/model.py:40) as a per-element multiplier 0 or 1 / (1 - p) drawn by the caller
        e[t] = alpha_scale ? a * alpha_scale[(size_t)b * T + t] : a;
    }
    __syncthreads();
    for (int c = 8 * lane; c < C; c += 256) {
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        for (int t0 = warp; t0 < T; t0 += 64) {
            uint4 q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int t = min(t0 + 8 * k, T - 1);         // past the end: a duplicate row with weight 0
                if (STAGE) q[k] = *reinterpret_cast<const uint4 *>(enc_s + (size_t)t * C + c);
                else q[k] = *reinterpret_cast<const uint4 *>(eb + (size_t)t * enc_st + c);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int t = t0 + 8 * k;
                const float a = t < T ? e[t] : 0.f;
                float f[8];
                unpack8(q[k], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(a, f[i], acc[i]);
            }
        }
        float4 *dst = reinterpret_cast<float4 *>(part + (size_t)warp * C + c);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    for (int c = 2 * threadIdx.x; c < C; c += 512) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a0 += part[(size_t)w * C + c]; a1 += part[(size_t)w * C + c + 1]; }
        *reinterpret_cast<__nv_bfloat162 *>(xcat + (size_t)b * ldx + c) = __floats2bfloat162_rn(a0, a1);
    }
}

// K6b: LSTMCell pointwise.  gates [B,4H] f32 (torch order i,f,g,o), embT [V,4H] f32 (column C+v of W_ih),
// y [B] int64 previous tokens; c [B,H] f32 in place; h -> bf16 xcat[b, C + j], f32 hid_out[b*hid_ld + j] (optional)
__global__ void attn_cell_kernel(const float *__restrict__ gates, const float *__restrict__ embT,
                                 const long long *__restrict__ y, int B, int H, int V, float *__restrict__ c,
                                 __nv_bfloat16 *__restrict__ xcat, long long ldx, int C,
                                 float *__restrict__ hid_out, long long hid_ld) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * H) return;
    const int b = (int)(idx / H), j = (int)(idx % H);
    long long tok = y[b];
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    const float *g = gates + (size_t)b * 4 * H, *em = embT + (size_t)tok * 4 * H;
    const float ig = sigmoid_fast_a(g[j] + em[j]);
    const float fg = sigmoid_fast_a(g[H + j] + em[H + j]);
    const float gg = tanh_fast_a(g[2 * H + j] + em[2 * H + j]);
    const float og = sigmoid_fast_a(g[3 * H + j] + em[3 * H + j]);
    const float cn = fmaf(fg, c[idx], ig * gg);
    c[idx] = cn;
    const float hn = og * tanh_fast_a(cn);
    xcat[(size_t)b * ldx + C + j] = __float2bfloat16_rn(hn);
    if (hid_out) hid_out[(size_t)b * hid_ld + j] = hn;
}

// K6c: one warp per sequence: mask the blank class, copy the row into probs[:, t, :], argmax (first maximum,
// NaN counts as maximal: torch.argmax)
__global__ void __launch_bounds__(128) attn_argmax_kernel(const float *__restrict__ logits, int B, int V, int blank,
                                                          float *__restrict__ probs, long long probs_ld,
                                                          long long *__restrict__ y, long long logits_ld) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + warp;
    if (b >= B) return;
    const int arg = warp_argmax_row(logits + (size_t)b * logits_ld, V, blank, probs ? probs + (size_t)b * probs_ld : nullptr, lane);
    if (lane == 0 && y) y[b] = arg;
}

}  // namespace
}  // namespace rcnn

extern "C" int rcnn_attn_score_context_ld(const float *projH, const float *projh, int64_t projh_ld, const float *v, const float *enc,
                                          int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C,
                                          float *alpha_out, void *xcat, int64_t ldx, rcnn_stream_t stream);
extern "C" int rcnn_attn_score_context(const float *projH, const float *projh, const float *v, const float *enc,
                                       int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C,
                                       float *alpha_out, void *xcat, int64_t ldx, rcnn_stream_t stream) {
    return rcnn_attn_score_context_ld(projH, projh, H, v, enc, enc_stride_b, enc_stride_t, B, T, H, C, alpha_out, xcat, ldx, stream);
}

extern "C" int rcnn_attn_score_context_ld(const float *projH, const float *projh, int64_t projh_ld, const float *v, const float *enc,
                                          int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C,
                                          float *alpha_out, void *xcat, int64_t ldx, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && C >= 1 && ldx >= C && projh_ld >= H, "attn_score_context: bad shape");
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(projH && projh && v && enc && xcat, "attn_score_context: null pointer");
    const size_t smem = sizeof(float) * (2 * (size_t)H + T);
    RCNN_CHECK_ARG(smem <= 200 * 1024, "attn_score_context: T=%d, H=%d exceed shared memory", T, H);
    if (smem > 48 * 1024)
        RCNN_CUDA(cudaFuncSetAttribute(attn_score_context_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_score_context_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(projH, projh, v, enc, enc_stride_b, enc_stride_t, T, H, C,
                                                                    alpha_out, (__nv_bfloat16 *)xcat, ldx, projh_ld);
    RCNN_LAUNCH_CHECK("attn_score_context_kernel");
    return RCNN_OK;
}

namespace {
int attn_step_launch(const void *projH, const float *projh, int64_t projh_ld, const float *v, const void *enc,
                     int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C, float *alpha_out, void *xcat, int64_t ldx,
                     const float *prev_logits, int64_t prev_ld, int V, int blank, float *prev_probs, int64_t probs_ld, int64_t *y,
                     const float *alpha_scale, rcnn_stream_t stream);
}
extern "C" int rcnn_attn_step_bf16(const void *projH, const float *projh, int64_t projh_ld, const float *v, const void *enc,
                                   int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C, float *alpha_out,
                                   void *xcat, int64_t ldx, const float *prev_logits, int64_t prev_ld, int V, int blank,
                                   float *prev_probs, int64_t probs_ld, int64_t *y, rcnn_stream_t stream) {
    return attn_step_launch(projH, projh, projh_ld, v, enc, enc_stride_b, enc_stride_t, B, T, H, C, alpha_out, xcat, ldx,
                            prev_logits, prev_ld, V, blank, prev_probs, probs_ld, y, nullptr, stream);
}
extern "C" int rcnn_attn_step_train(const void *projH, const float *projh, int64_t projh_ld, const float *v, const void *enc,
                                    int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C, float *alpha_out,
                                    const float *alpha_scale, void *xcat, int64_t ldx, rcnn_stream_t stream) {
    if (alpha_out == nullptr) {
        rcnn::set_error("attn_step_train: alpha_out is required (the backward pass reads it)");
        return RCNN_ERR_ARG;
    }
    return attn_step_launch(projH, projh, projh_ld, v, enc, enc_stride_b, enc_stride_t, B, T, H, C, alpha_out, xcat, ldx, nullptr, 0,
                            0, -1, nullptr, 0, nullptr, alpha_scale, stream);
}
namespace {
int attn_step_launch(const void *projH, const float *projh, int64_t projh_ld, const float *v, const void *enc,
                     int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H, int C, float *alpha_out, void *xcat, int64_t ldx,
                     const float *prev_logits, int64_t prev_ld, int V, int blank, float *prev_probs, int64_t probs_ld, int64_t *y,
                     const float *alpha_scale, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 1 && H >= 1 && C >= 1 && ldx >= C && projh_ld >= H, "attn_score_context_bf16: bad shape");
    RCNN_CHECK_ARG(prev_logits == nullptr || (V >= 1 && prev_ld >= V && (prev_probs == nullptr || probs_ld >= V)),
                   "attn_step_bf16: bad logits shape");
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(projH && projh && v && enc && xcat, "attn_score_context_bf16: null pointer");
    RCNN_CHECK_ARG(H % 8 == 0 && C % 8 == 0 && enc_stride_b % 8 == 0 && enc_stride_t % 8 == 0 && ldx % 2 == 0 &&
                       ((uintptr_t)projH & 15) == 0 && ((uintptr_t)enc & 15) == 0 && ((uintptr_t)xcat & 3) == 0,
                   "attn_score_context_bf16: H, C and the enc strides must be multiples of 8, the bf16 arrays 16-byte aligned");
    const size_t base = sizeof(float) * (2 * (size_t)H + ((T + 3) & ~3) + 8 * (size_t)C);
    const size_t stage_bytes = (size_t)T * C * 2;
    const bool stage = base + stage_bytes <= 100 * 1024;      // two CTAs per SM stay resident
    const size_t smem = base + (stage ? stage_bytes : 0);
    RCNN_CHECK_ARG(smem <= 200 * 1024, "attn_score_context_bf16: T=%d, H=%d, C=%d exceed shared memory", T, H, C);
    auto kern = stage ? attn_score_context_bf16_kernel<true> : attn_score_context_bf16_kernel<false>;
    if (smem > 48 * 1024) RCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    chain_config(cfg, attr, (unsigned)B, 256, smem, (cudaStream_t)stream);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16 *)projH, projh, v, (const __nv_bfloat16 *)enc,
                                 (long long)enc_stride_b, (long long)enc_stride_t, T, H, C, alpha_out, (__nv_bfloat16 *)xcat,
                                 (long long)ldx, (long long)projh_ld, prev_logits, (long long)prev_ld, V, blank, prev_probs,
                                 (long long)probs_ld, (long long *)y, alpha_scale));
    RCNN_LAUNCH_CHECK("attn_score_context_bf16_kernel");
    return RCNN_OK;
}
}  // namespace

extern "C" int rcnn_attn_score_context_bf16(const void *projH, const float *projh, int64_t projh_ld, const float *v,
                                            const void *enc, int64_t enc_stride_b, int64_t enc_stride_t, int B, int T, int H,
                                            int C, float *alpha_out, void *xcat, int64_t ldx, rcnn_stream_t stream) {
    return rcnn_attn_step_bf16(projH, projh, projh_ld, v, enc, enc_stride_b, enc_stride_t, B, T, H, C, alpha_out, xcat, ldx,
                               nullptr, 0, 0, -1, nullptr, 0, nullptr, stream);
}

extern "C" int rcnn_attn_cell(const float *gates, const float *embT, const int64_t *y, int B, int H, int V, float *c,
                              void *xcat, int64_t ldx, int C, float *hid_out, int64_t hid_ld, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && H >= 1 && V >= 1 && ldx >= (int64_t)C + H, "attn_cell: bad shape");
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(gates && embT && y && c && xcat, "attn_cell: null pointer");
    const long long total = (long long)B * H;
    attn_cell_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        gates, embT, (const long long *)y, B, H, V, c, (__nv_bfloat16 *)xcat, ldx, C, hid_out, hid_ld);
    RCNN_LAUNCH_CHECK("attn_cell_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_attn_argmax_ld(const float *logits, int64_t logits_ld, int B, int V, int blank, float *probs, int64_t probs_ld,
                                   int64_t *y, rcnn_stream_t stream);
extern "C" int rcnn_attn_argmax(const float *logits, int B, int V, int blank, float *probs, int64_t probs_ld,
                                int64_t *y, rcnn_stream_t stream) {
    return rcnn_attn_argmax_ld(logits, V, B, V, blank, probs, probs_ld, y, stream);
}

extern "C" int rcnn_attn_argmax_ld(const float *logits, int64_t logits_ld, int B, int V, int blank, float *probs, int64_t probs_ld,
                                   int64_t *y, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && V >= 1 && logits_ld >= V, "attn_argmax: bad shape");
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(logits, "attn_argmax: null pointer");
    attn_argmax_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(logits, B, V, blank, probs, probs_ld, (long long *)y, logits_ld);
    RCNN_LAUNCH_CHECK("attn_argmax_kernel");
    return RCNN_OK;
}
