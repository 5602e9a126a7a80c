// Cell-update arithmetic shared by the forward recurrent kernels (lstm_fwd.cu, lstm_fwdx.cu).
//
// Thread layout (transposed product D^T[128 gate rows, 64 seq]): a thread owns ONE gate row (lane l of TMEM
// quadrant qd: unit 8*qd + l/4, gate l%4 in torch's i,f,g,o order) and 32 consecutive sequences.
#pragma once
#include "common.cuh"

namespace rcnn {

__device__ __forceinline__ float tanh_fast(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct CellLane {          // per-thread constants derived from the lane index
    bool g0, g1, o0, o1, o2;
    float a_scale, a_shift;
    __device__ __forceinline__ explicit CellLane(int lane) {
        const int g = lane & 3, ul = lane >> 2;
        g0 = g & 1; g1 = (g >> 1) & 1;
        a_scale = (g == 2) ? 1.f : 0.5f;     // sigmoid(x) = 0.5 tanh(0.5 x) + 0.5
        a_shift = (g == 2) ? 0.f : 0.5f;
        o0 = ul & 1; o1 = (ul >> 1) & 1; o2 = (ul >> 2) & 1;
    }
};

// one MUFU.TANH per gate pre-activation (sigmoid through tanh with the lane's constants)
__device__ __forceinline__ void cell_activate(float (&pre)[32], const CellLane &L) {
#pragma unroll
    for (int i = 0; i < 32; ++i) pre[i] = fmaf(tanh_fast(pre[i] * L.a_scale), L.a_scale, L.a_shift);
}

// activated gates -> cell update (cst: c of unit 8*qd + lane/4 for sequences 4k + lane%4, k = 0..7) -> h_t as bf16,
// transposed so that the lane returns units 8*qd .. 8*qd+7 (16 bytes) of sequence `lane` of its 32-sequence half
__device__ __forceinline__ uint4 cell_update(const float (&pre)[32], float (&cst)[8], const CellLane &L) {
    const bool g0 = L.g0, g1 = L.g1, o0 = L.o0, o1 = L.o1, o2 = L.o2;
    uint32_t hb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // 4x4 transpose across the four lanes of a unit: in = my gate for sequences 4k..4k+3,
        // out = gates i,f,g,o for sequence 4k + g
        const float v0 = pre[4 * k], v1 = pre[4 * k + 1], v2 = pre[4 * k + 2], v3 = pre[4 * k + 3];
        const float ra0 = __shfl_xor_sync(FULL, g0 ? v0 : v1, 1);
        const float ra1 = __shfl_xor_sync(FULL, g0 ? v2 : v3, 1);
        const float w0 = g0 ? ra0 : v0, w1 = g0 ? v1 : ra0, w2 = g0 ? ra1 : v2, w3 = g0 ? v3 : ra1;
        const float rb0 = __shfl_xor_sync(FULL, g1 ? w0 : w2, 2);
        const float rb1 = __shfl_xor_sync(FULL, g1 ? w1 : w3, 2);
        const float ig = g1 ? rb0 : w0, fg = g1 ? rb1 : w1, gg = g1 ? w2 : rb0, og = g1 ? w3 : rb1;
        const float cn = fmaf(fg, cst[k], ig * gg);
        cst[k] = cn;
        const __nv_bfloat16 hv = __float2bfloat16_rn(og * tanh_fast(cn));
        hb[k] = (uint32_t)__bfloat16_as_ushort(hv);
    }
    // 8x8 transpose across the eight lanes with the same g: in = my unit's h for sequences
    // 4k + g (k = 0..7), out = units 8qd .. 8qd+7 for sequence 4*ul + g = lane
    uint32_t p1[4];
    {
        uint32_t keep[4], send[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { keep[j] = o0 ? hb[2 * j + 1] : hb[2 * j]; send[j] = o0 ? hb[2 * j] : hb[2 * j + 1]; }
        const uint32_t r0 = __shfl_xor_sync(FULL, send[0] | (send[1] << 16), 4);
        const uint32_t r1 = __shfl_xor_sync(FULL, send[2] | (send[3] << 16), 4);
        const uint32_t recv[4] = {r0 & 0xffffu, r0 >> 16, r1 & 0xffffu, r1 >> 16};
#pragma unroll
        for (int j = 0; j < 4; ++j) p1[j] = o0 ? (recv[j] | (keep[j] << 16)) : (keep[j] | (recv[j] << 16));
    }
    uint32_t q2[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const uint32_t keep = o1 ? p1[2 * i + 1] : p1[2 * i], send = o1 ? p1[2 * i] : p1[2 * i + 1];
        const uint32_t rr = __shfl_xor_sync(FULL, send, 8);
        q2[i][0] = o1 ? rr : keep;
        q2[i][1] = o1 ? keep : rr;
    }
    uint4 hq;
    {
        const uint32_t k0 = o2 ? q2[1][0] : q2[0][0], k1 = o2 ? q2[1][1] : q2[0][1];
        const uint32_t s0 = o2 ? q2[0][0] : q2[1][0], s1 = o2 ? q2[0][1] : q2[1][1];
        const uint32_t r0 = __shfl_xor_sync(FULL, s0, 16), r1 = __shfl_xor_sync(FULL, s1, 16);
        hq.x = o2 ? r0 : k0; hq.y = o2 ? r1 : k1; hq.z = o2 ? k0 : r0; hq.w = o2 ? k1 : r1;
    }
    return hq;
}

}  // namespace rcnn
