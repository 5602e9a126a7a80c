// K6 backward: the per-step kernels of the attention decoder's teacher-forced training pass (model/model.py:33-47,
// 110-148 under autograd).  One backward step t, given dh_t and dc_t:
//     LSTMCell pointwise backward                       -> d(gate pre-activations) as bf16, dc_{t-1}         (K6d)
//     dcontext = dgates @ W_ih[:, :C]                    -> tcgen05 GEMM (gemm.cu)
//     attention backward for the step                   -> de_t [B,T], d proj_h (bf16), dv partial sums      (K6e)
//     dh_{t-1} = [dgates | dproj_h] @ [W_hh ; W_h2h]      -> tcgen05 GEMM
// and after the loop, once:
//     d proj_H[b,t,j] = v[j] * sum_s de_s[b,t] * (1 - tanh^2(proj_H[b,t,j] + proj_h_s[b,j]))                  (K6f)
// (the sum over the decoding steps s is formed in registers: accumulating it step by step would read and write the
// 33 MB gradient 26 times).  The weight gradients are products over all (step, sequence) rows at once (gemm.cu).
// Layouts: gate arrays are gate-interleaved along 4H (element 4u + g = gate g of unit u, torch order i, f, g, o), as the
// forward's rcnn_attn_gates_cell_train leaves them.
#include "common.cuh"

namespace rcnn {
namespace {

__device__ __forceinline__ float tanh_fast_b(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void unpack8b(const uint4 &q, float (&f)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// K6d: one thread per (sequence, unit).  gates_act [B,4H] (interleaved, post-activation), c_prev / c_t [B,H],
// dh = dh_a (+ dh_b), dc [B,H] in/out (dc_t in, dc_{t-1} out) -> dg [B, ldg] bf16 (interleaved pre-activation gradients)
__global__ void attn_cell_bwd_kernel(const float *__restrict__ gates_act, const float *__restrict__ c_prev,
                                     const float *__restrict__ c_t, const float *__restrict__ dh_a, long long dh_a_ld,
                                     const float *__restrict__ dh_b, long long dh_b_ld, float *__restrict__ dc, int B, int H,
                                     __nv_bfloat16 *__restrict__ dg, long long ldg) {
    griddep_launch();
    griddep_wait();                                           // (chained launches: dh / dc come from the predecessor)
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * H) return;
    const int b = (int)(idx / H), u = (int)(idx % H);
    const float4 ga = *reinterpret_cast<const float4 *>(gates_act + ((size_t)b * H + u) * 4);   // i, f, g, o
    const float cp = c_prev ? c_prev[idx] : 0.f;
    const float tc = tanh_fast_b(c_t[idx]);
    const float dh = dh_a[(size_t)b * dh_a_ld + u] + (dh_b ? dh_b[(size_t)b * dh_b_ld + u] : 0.f);
    const float dct = dc[idx] + dh * ga.w * (1.f - tc * tc);
    const float d_i = dct * ga.z * ga.x * (1.f - ga.x);
    const float d_f = dct * cp * ga.y * (1.f - ga.y);
    const float d_g = dct * ga.x * (1.f - ga.z * ga.z);
    const float d_o = dh * tc * ga.w * (1.f - ga.w);
    dc[idx] = dct * ga.y;
    const __nv_bfloat162 lo = __floats2bfloat162_rn(d_i, d_f), hi = __floats2bfloat162_rn(d_g, d_o);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t *>(&lo);
    o.y = *reinterpret_cast<const uint32_t *>(&hi);
    *reinterpret_cast<uint2 *>(dg + (size_t)b * ldg + 4 * (size_t)u) = o;
}

// K6e: one CTA (256 threads) per sequence.  dctx [B, ld] f32, alpha [B,T] (before dropout), alpha_scale [B,T] or NULL,
// enc / proj_H bf16, proj_h [B, ld] f32, v [H]  ->  de [B,T] f32, dprojh bf16 [B, ld] (H values), dv_acc [B,H] += .
// Warp w owns the frames t = w (mod 8) in both passes; partial sums over frames meet in shared memory.
__global__ void __launch_bounds__(256) attn_step_bwd_kernel(
    const float *__restrict__ dctx, long long dctx_ld, const float *__restrict__ alpha, const float *__restrict__ alpha_scale,
    const __nv_bfloat16 *__restrict__ enc, long long enc_sb, long long enc_st, const __nv_bfloat16 *__restrict__ projH,
    const float *__restrict__ projh, long long projh_ld, const float *__restrict__ v, int T, int H, int C,
    float *__restrict__ de_out, __nv_bfloat16 *__restrict__ dprojh, long long dprojh_ld, float *__restrict__ dv_acc) {
    extern __shared__ __align__(16) float sm[];
    float *ph = sm, *dcs = sm + H, *de = dcs + C, *part = de + ((T + 3) & ~3);      // [H], [C], [T], [8][2][H]
    __shared__ float red;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    griddep_launch();
    for (int j = threadIdx.x; j < H; j += 256) ph[j] = projh[(size_t)b * projh_ld + j];     // (kept from the forward pass)
    griddep_wait();                                           // chained launches: dctx is the predecessor's output
    for (int c = threadIdx.x; c < C; c += 256) dcs[c] = dctx[(size_t)b * dctx_ld + c];
    __syncthreads();
    // pass A: d alpha_dropped[t] = <dctx, enc[t]>; through the dropout multiplier
    const __nv_bfloat16 *eb = enc + (size_t)b * enc_sb;
    for (int t0 = warp; t0 < T; t0 += 64) {
        float s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = 0.f;
        for (int c = 8 * lane; c < C; c += 256) {
            uint4 q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = *reinterpret_cast<const uint4 *>(eb + (size_t)min(t0 + 8 * k, T - 1) * enc_st + c);
            float d8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) d8[i] = dcs[c + i];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float f[8];
                unpack8b(q[k], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) s[k] = fmaf(d8[i], f[i], s[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float r = warp_sum(s[k]);
            const int t = t0 + 8 * k;
            if (lane == 0 && t < T) de[t] = alpha_scale ? r * alpha_scale[(size_t)b * T + t] : r;
        }
    }
    __syncthreads();
    // softmax backward: de[t] = alpha[t] * (dalpha[t] - sum_t' alpha[t'] dalpha[t'])
    if (warp == 0) {
        float z = 0.f;
        for (int t = lane; t < T; t += 32) z = fmaf(alpha[(size_t)b * T + t], de[t], z);
        z = warp_sum(z);
        if (lane == 0) red = z;
    }
    __syncthreads();
    const float dot = red;
    for (int t = threadIdx.x; t < T; t += 256) {
        const float a = alpha[(size_t)b * T + t];
        const float d = a * (de[t] - dot);
        de[t] = d;
        de_out[(size_t)b * T + t] = d;
    }
    __syncthreads();
    // pass B: u = tanh(proj_H + proj_h);  dproj_h[j] = v[j] sum_t de[t] (1 - u^2),  dv[j] = sum_t de[t] u
    const __nv_bfloat16 *pb = projH + (size_t)b * T * H;
    for (int j = 8 * lane; j < H; j += 256) {
        float a1[8], a2[8], p8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { a1[i] = 0.f; a2[i] = 0.f; p8[i] = ph[j + i]; }
        for (int t0 = warp; t0 < T; t0 += 64) {
            uint4 q[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) q[k] = *reinterpret_cast<const uint4 *>(pb + (size_t)min(t0 + 8 * k, T - 1) * H + j);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int t = t0 + 8 * k;
                const float d = t < T ? de[t] : 0.f;
                float f[8];
                unpack8b(q[k], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float u = tanh_fast_b(f[i] + p8[i]);
                    a1[i] = fmaf(d, fmaf(-u, u, 1.f), a1[i]);
                    a2[i] = fmaf(d, u, a2[i]);
                }
            }
        }
        float4 *d1 = reinterpret_cast<float4 *>(part + ((size_t)warp * 2 + 0) * H + j);
        float4 *d2 = reinterpret_cast<float4 *>(part + ((size_t)warp * 2 + 1) * H + j);
        d1[0] = make_float4(a1[0], a1[1], a1[2], a1[3]); d1[1] = make_float4(a1[4], a1[5], a1[6], a1[7]);
        d2[0] = make_float4(a2[0], a2[1], a2[2], a2[3]); d2[1] = make_float4(a2[4], a2[5], a2[6], a2[7]);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += 256) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { s1 += part[((size_t)w * 2 + 0) * H + j]; s2 += part[((size_t)w * 2 + 1) * H + j]; }
        dprojh[(size_t)b * dprojh_ld + j] = __float2bfloat16_rn(v[j] * s1);
        dv_acc[(size_t)b * H + j] += s2;                       // the row belongs to this CTA: deterministic order over the steps
    }
}

// K6f: d proj_H, all steps at once.  grid (T, B), 128 threads: thread owns unit pairs; de_all [S,B,T], projh_all [S,B,H] f32
// -> dprojH bf16 [B,T,H]
__global__ void __launch_bounds__(128) attn_dprojH_kernel(const float *__restrict__ de_all, const float *__restrict__ projh_all,
                                                          const __nv_bfloat16 *__restrict__ projH, const float *__restrict__ v,
                                                          int S, int B, int T, int H, __nv_bfloat16 *__restrict__ dprojH) {
    extern __shared__ float des[];                            // [S]
    const int t = blockIdx.x, b = blockIdx.y;
    for (int s = threadIdx.x; s < S; s += 128) des[s] = de_all[((size_t)s * B + b) * T + t];
    __syncthreads();
    const size_t row = ((size_t)b * T + t) * H;
    for (int j = 2 * threadIdx.x; j < H; j += 256) {
        const uint32_t w = *reinterpret_cast<const uint32_t *>(projH + row + j);
        const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
        float a0 = 0.f, a1 = 0.f;
        for (int s = 0; s < S; ++s) {
            const float2 p = *reinterpret_cast<const float2 *>(projh_all + ((size_t)s * B + b) * H + j);
            const float u0 = tanh_fast_b(x0 + p.x), u1 = tanh_fast_b(x1 + p.y);
            a0 = fmaf(des[s], fmaf(-u0, u0, 1.f), a0);
            a1 = fmaf(des[s], fmaf(-u1, u1, 1.f), a1);
        }
        *reinterpret_cast<__nv_bfloat162 *>(dprojH + row + j) = __floats2bfloat162_rn(v[j] * a0, v[j + 1] * a1);
    }
}

// K6f, the usual case: a CTA handles eight frames of one sequence and keeps proj_h of all S steps in shared memory
// ([S][H] f32: 53 KB at S = 26, H = 512), so the steps' rows are read from L2 once per eight frames instead of once per frame
constexpr int kDpFrames = 8;
__global__ void __launch_bounds__(256) attn_dprojH_smem_kernel(const float *__restrict__ de_all, const float *__restrict__ projh_all,
                                                               const __nv_bfloat16 *__restrict__ projH, const float *__restrict__ v,
                                                               int S, int B, int T, int H, __nv_bfloat16 *__restrict__ dprojH) {
    extern __shared__ __align__(16) float dsm[];
    float *phs = dsm, *des = dsm + (size_t)S * H;             // [S][H], [S][kDpFrames]
    const int t0 = blockIdx.x * kDpFrames, b = blockIdx.y;
    for (int i = 4 * threadIdx.x; i < S * H; i += 4 * 256) {
        const int sidx = i / H, j = i - sidx * H;
        *reinterpret_cast<float4 *>(phs + i) = *reinterpret_cast<const float4 *>(projh_all + ((size_t)sidx * B + b) * H + j);
    }
    for (int i = threadIdx.x; i < S * kDpFrames; i += 256) {
        const int sidx = i / kDpFrames, f = i - sidx * kDpFrames;
        des[i] = t0 + f < T ? de_all[((size_t)sidx * B + b) * T + t0 + f] : 0.f;
    }
    __syncthreads();
    for (int j = 2 * threadIdx.x; j < H; j += 512) {
        float x0[kDpFrames], x1[kDpFrames], a0[kDpFrames], a1[kDpFrames];
#pragma unroll
        for (int f = 0; f < kDpFrames; ++f) {
            const uint32_t w = *reinterpret_cast<const uint32_t *>(projH + ((size_t)b * T + min(t0 + f, T - 1)) * H + j);
            x0[f] = __uint_as_float(w << 16); x1[f] = __uint_as_float(w & 0xffff0000u);
            a0[f] = 0.f; a1[f] = 0.f;
        }
        for (int sidx = 0; sidx < S; ++sidx) {
            const float2 p = *reinterpret_cast<const float2 *>(phs + (size_t)sidx * H + j);
#pragma unroll
            for (int f = 0; f < kDpFrames; ++f) {
                const float d = des[sidx * kDpFrames + f];
                const float u0 = tanh_fast_b(x0[f] + p.x), u1 = tanh_fast_b(x1[f] + p.y);
                a0[f] = fmaf(d, fmaf(-u0, u0, 1.f), a0[f]);
                a1[f] = fmaf(d, fmaf(-u1, u1, 1.f), a1[f]);
            }
        }
        const float v0 = v[j], v1 = v[j + 1];
#pragma unroll
        for (int f = 0; f < kDpFrames; ++f)
            if (t0 + f < T)
                *reinterpret_cast<__nv_bfloat162 *>(dprojH + ((size_t)b * T + t0 + f) * H + j) = __floats2bfloat162_rn(v0 * a0[f], v1 * a1[f]);
    }
}

}  // namespace
}  // namespace rcnn

extern "C" int rcnn_attn_cell_bwd(const float *gates_act, const float *c_prev, const float *c_t, const float *dh_a, int64_t dh_a_ld,
                                  const float *dh_b, int64_t dh_b_ld, float *dc, int B, int H, void *dg, int64_t ldg,
                                  rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && H >= 1 && ldg >= 4 * (int64_t)H && dh_a_ld >= H && (dh_b == nullptr || dh_b_ld >= H),
                   "attn_cell_bwd: bad shape");
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(gates_act && c_t && dh_a && dc && dg, "attn_cell_bwd: null pointer");
    RCNN_CHECK_ARG(((uintptr_t)gates_act & 15) == 0 && ((uintptr_t)dg & 7) == 0 && (ldg & 3) == 0, "attn_cell_bwd: alignment");
    const long long total = (long long)B * H;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    chain_config(cfg, attr, (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, attn_cell_bwd_kernel, gates_act, c_prev, c_t, dh_a, (long long)dh_a_ld, dh_b,
                                 (long long)dh_b_ld, dc, B, H, (__nv_bfloat16 *)dg, (long long)ldg));
    RCNN_LAUNCH_CHECK("attn_cell_bwd_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_attn_step_bwd(const float *dctx, int64_t dctx_ld, const float *alpha, const float *alpha_scale, const void *enc,
                                  int64_t enc_stride_b, int64_t enc_stride_t, const void *projH, const float *projh,
                                  int64_t projh_ld, const float *v, int B, int T, int H, int C, float *de_out, void *dprojh,
                                  int64_t dprojh_ld, float *dv_acc, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 1 && H >= 8 && C >= 8 && dctx_ld >= C && projh_ld >= H && dprojh_ld >= H, "attn_step_bwd: bad shape");
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(dctx && alpha && enc && projH && projh && v && de_out && dprojh && dv_acc, "attn_step_bwd: null pointer");
    RCNN_CHECK_ARG(H % 8 == 0 && C % 8 == 0 && enc_stride_b % 8 == 0 && enc_stride_t % 8 == 0 && ((uintptr_t)projH & 15) == 0 &&
                       ((uintptr_t)enc & 15) == 0,
                   "attn_step_bwd: H, C and the enc strides must be multiples of 8, the bf16 arrays 16-byte aligned");
    const size_t smem = sizeof(float) * ((size_t)H + C + ((T + 3) & ~3) + 16 * (size_t)H);
    RCNN_CHECK_ARG(smem <= 200 * 1024, "attn_step_bwd: T=%d, H=%d, C=%d exceed shared memory", T, H, C);
    if (smem > 48 * 1024)
        RCNN_CUDA(cudaFuncSetAttribute(attn_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    chain_config(cfg, attr, (unsigned)B, 256, smem, (cudaStream_t)stream);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, attn_step_bwd_kernel, dctx, (long long)dctx_ld, alpha, alpha_scale,
                                 (const __nv_bfloat16 *)enc, (long long)enc_stride_b, (long long)enc_stride_t,
                                 (const __nv_bfloat16 *)projH, projh, (long long)projh_ld, v, T, H, C, de_out,
                                 (__nv_bfloat16 *)dprojh, (long long)dprojh_ld, dv_acc));
    RCNN_LAUNCH_CHECK("attn_step_bwd_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_attn_dprojH(const float *de_all, const float *projh_all, const void *projH, const float *v, int S, int B, int T,
                                int H, void *dprojH, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(S >= 1 && B >= 0 && T >= 1 && H >= 2 && H % 2 == 0 && B <= 65535, "attn_dprojH: bad shape");
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(de_all && projh_all && projH && v && dprojH, "attn_dprojH: null pointer");
    RCNN_CHECK_ARG(S * sizeof(float) <= 48 * 1024, "attn_dprojH: too many steps");
    const size_t smem_all = sizeof(float) * ((size_t)S * H + (size_t)S * kDpFrames);
    if (smem_all <= 100 * 1024 && H % 4 == 0 && ((uintptr_t)projh_all & 15) == 0) {
        if (smem_all > 48 * 1024)
            RCNN_CUDA(cudaFuncSetAttribute(attn_dprojH_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_all));
        attn_dprojH_smem_kernel<<<dim3((unsigned)((T + kDpFrames - 1) / kDpFrames), (unsigned)B), 256, smem_all, (cudaStream_t)stream>>>(
            de_all, projh_all, (const __nv_bfloat16 *)projH, v, S, B, T, H, (__nv_bfloat16 *)dprojH);
        RCNN_LAUNCH_CHECK("attn_dprojH_smem_kernel");
        return RCNN_OK;
    }
    attn_dprojH_kernel<<<dim3((unsigned)T, (unsigned)B), 128, S * sizeof(float), (cudaStream_t)stream>>>(
        de_all, projh_all, (const __nv_bfloat16 *)projH, v, S, B, T, H, (__nv_bfloat16 *)dprojH);
    RCNN_LAUNCH_CHECK("attn_dprojH_kernel");
    return RCNN_OK;
}
