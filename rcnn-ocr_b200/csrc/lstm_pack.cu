// Layout preparation around the recurrent kernels: weight packing, casts, transposes.
//
// The recurrent kernels partition one LSTM direction by HIDDEN UNIT: CTA c of a cluster owns
// units [32c, 32c+32) and all four gates of those units, so the cell update is CTA-local.
// Its 128 gate rows are ordered n = j*4 + g (unit j of the slice, gate g in torch's i,f,g,o
// order), which puts the four gates of a unit in adjacent TMEM columns of one accumulator row.
// "Packed" order of the 4H gate rows of one direction is therefore
//     p(c, j, g) = c*128 + j*4 + g   <->   torch row  g*H + 32c + j.
// The pack kernel writes every weight view the forward and backward kernels need in one launch.
#include "common.cuh"

namespace rcnn {
namespace {

struct PackArgs {
    const float *w_ih[2], *w_hh[2], *b_ih[2], *b_hh[2];
    int I, H;
    __nv_bfloat16 *wih_p;   // [2*4H, I]      packed rows (dir-major)
    float *bias_p;          // [2*4H]         b_ih + b_hh, packed
    __nv_bfloat16 *whh_p;   // [2*4H, H]      packed rows
    __nv_bfloat16 *whh_pt;  // [2, H, 4H]     transpose of whh_p per direction (k = packed column)
    __nv_bfloat16 *wih_pt;  // [I, 2*4H]      transpose of wih_p
};

__device__ __forceinline__ int torch_row(int p, int H) {
    const int c = p >> 7, j = (p >> 2) & 31, g = p & 3;
    return g * H + 32 * c + j;
}

// row-major views: one block per packed row (dir, p); reads and writes are both coalesced
__global__ void lstm_pack_rows_kernel(const PackArgs a) {
    const int H = a.H, I = a.I;
    const int dir = blockIdx.x / (4 * H), p = blockIdx.x % (4 * H);
    const int r = torch_row(p, H);
    const float *wi = a.w_ih[dir] + (size_t)r * I;
    const float *wh = a.w_hh[dir] + (size_t)r * H;
    const size_t prow = (size_t)dir * 4 * H + p;
    if ((I & 3) == 0 && (H & 3) == 0) {   // rows are 16-byte aligned: 128-bit loads, 64-bit stores
        for (int k = 4 * threadIdx.x; k < I; k += 4 * blockDim.x) {
            const float4 v = *reinterpret_cast<const float4 *>(wi + k);
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            *reinterpret_cast<uint2 *>(a.wih_p + prow * I + k) =
                make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
        }
        for (int k = 4 * threadIdx.x; k < H; k += 4 * blockDim.x) {
            const float4 v = *reinterpret_cast<const float4 *>(wh + k);
            __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            *reinterpret_cast<uint2 *>(a.whh_p + prow * H + k) =
                make_uint2(*reinterpret_cast<uint32_t *>(&lo), *reinterpret_cast<uint32_t *>(&hi));
        }
    } else {
        for (int k = threadIdx.x; k < I; k += blockDim.x) a.wih_p[prow * I + k] = __float2bfloat16_rn(wi[k]);
        for (int k = threadIdx.x; k < H; k += blockDim.x) a.whh_p[prow * H + k] = __float2bfloat16_rn(wh[k]);
    }
    if (threadIdx.x == 0) a.bias_p[prow] = a.b_ih[dir][r] + a.b_hh[dir][r];
}

// transposed views through a [32 k][64 p] shared-memory tile: which = 0 -> wih_pt [I, 8H], 1 -> whh_pt [2, H, 4H].
// Reads are 128-byte row pieces of the fp32 source, writes 128-byte pieces (64 bf16) of the transposed views.
__global__ void lstm_pack_transposed_kernel(const PackArgs a) {
    __shared__ float tile[64][33];
    const int H = a.H;
    const int which = blockIdx.z >> 1, dir = blockIdx.z & 1;
    const int K = which ? H : a.I;
    const int p0 = blockIdx.y * 64, k0 = blockIdx.x * 32;
    if (k0 >= K) return;
    const float *src = which ? a.w_hh[dir] : a.w_ih[dir];
    for (int i = threadIdx.y; i < 64; i += blockDim.y) {
        const int k = k0 + threadIdx.x;
        tile[i][threadIdx.x] = k < K ? src[(size_t)torch_row(p0 + i, H) * K + k] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int k = k0 + i, p = p0 + 2 * threadIdx.x;
        if (k < K) {
            const __nv_bfloat162 v = __floats2bfloat162_rn(tile[2 * threadIdx.x][i], tile[2 * threadIdx.x + 1][i]);
            if (which) *reinterpret_cast<__nv_bfloat162 *>(a.whh_pt + ((size_t)dir * H + k) * 4 * H + p) = v;
            else *reinterpret_cast<__nv_bfloat162 *>(a.wih_pt + (size_t)k * 8 * H + (size_t)dir * 4 * H + p) = v;
        }
    }
}

// contiguous-channel fast path of the 3-D cast: 8 elements per thread (two 128-bit loads, one 128-bit store)
__global__ void cast3_bf16_vec_kernel(const float *__restrict__ src, long long sb, long long st,
                                      __nv_bfloat16 *__restrict__ dst, int B, int T, int C8) {
    const long long total = (long long)B * T * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const long long bt = i / C8;
        const int t = (int)(bt % T);
        const long long b = bt / T;
        const float4 *s4 = reinterpret_cast<const float4 *>(src + b * sb + (long long)t * st + 8 * c8);
        const uint4 u0 = ld_nc_v4(s4), u1 = ld_nc_v4(s4 + 1);
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(u0.x), __uint_as_float(u0.y));
        const __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(u0.z), __uint_as_float(u0.w));
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(u1.x), __uint_as_float(u1.y));
        const __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(u1.z), __uint_as_float(u1.w));
        uint4 o;
        o.x = *reinterpret_cast<const uint32_t *>(&h0); o.y = *reinterpret_cast<const uint32_t *>(&h1);
        o.z = *reinterpret_cast<const uint32_t *>(&h2); o.w = *reinterpret_cast<const uint32_t *>(&h3);
        *reinterpret_cast<uint4 *>(dst + (bt * C8 + c8) * 8) = o;
    }
}

// fp32 [R, C] (row stride ld) -> bf16 [R, ldd] (columns [C, ldd) zero-filled)
__global__ void cast_bf16_kernel(const float *__restrict__ src, long long ld, __nv_bfloat16 *__restrict__ dst,
                                 long long ldd, long long rows, int cols) {
    const long long total = rows * ldd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ldd;
        const int c = (int)(i - r * ldd);
        dst[i] = __float2bfloat16_rn(c < cols ? src[r * ld + c] : 0.f);
    }
}

// 3-D strided fp32 [B, T, C] (strides sb, st, sc) -> contiguous bf16 [B, T, C]
// (the encoder input is a permuted view of [B, C, T], model/model.py:218)
__global__ void cast3_bf16_kernel(const float *__restrict__ src, long long sb, long long st, long long sc,
                                  __nv_bfloat16 *__restrict__ dst, int B, int T, int C) {
    __shared__ float tile[32][33];
    // tile over (t, c) for one b; handles both c-contiguous and t-contiguous sources coalesced
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const float *s = src + (long long)b * sb;
    const bool c_fast = (sc == 1) || (st != 1);
    if (c_fast) {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int t = t0 + i, c = c0 + threadIdx.x;
            if (t < T && c < C) tile[i][threadIdx.x] = s[t * st + c * sc];
        }
    } else {
        for (int i = threadIdx.y; i < 32; i += blockDim.y) {
            const int c = c0 + i, t = t0 + threadIdx.x;
            if (t < T && c < C) tile[threadIdx.x][i] = s[t * st + c * sc];
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        if (t < T && c < C) dst[((long long)b * T + t) * C + c] = __float2bfloat16_rn(tile[i][threadIdx.x]);
    }
}

// bf16 [R, C] (row stride ld) -> bf16 [C, R] with output row stride ldo >= R
__global__ void transpose_bf16_kernel(const __nv_bfloat16 *__restrict__ src, long long ld,
                                      __nv_bfloat16 *__restrict__ dst, long long ldo, int R, int C) {
    __shared__ __nv_bfloat16 tile[64][66];
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    for (int i = threadIdx.y; i < 64; i += blockDim.y) {
        const int r = r0 + i;
        for (int j = threadIdx.x; j < 64; j += blockDim.x) {
            const int c = c0 + j;
            if (r < R && c < C) tile[i][j] = src[(long long)r * ld + c];
        }
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 64; i += blockDim.y) {
        const int c = c0 + i;
        for (int j = threadIdx.x; j < 64; j += blockDim.x) {
            const int r = r0 + j;
            if (r < R && c < C) dst[(long long)c * ldo + r] = tile[j][i];
        }
    }
}

}  // namespace
}  // namespace rcnn

extern "C" size_t rcnn_lstm_packed_bytes(int I, int H) {
    if (I <= 0 || H <= 0) return 0;
    const size_t H8 = 8 * (size_t)H;
    size_t b = 0;
    b += H8 * I * 2;          // wih_p
    b += H8 * 4;              // bias_p
    b += H8 * H * 2;          // whh_p
    b += H8 * H * 2;          // whh_pt
    b += H8 * I * 2;          // wih_pt
    return (b + 255) & ~(size_t)255;
}

extern "C" int rcnn_lstm_pack_weights(const float *w_ih_f, const float *w_hh_f, const float *b_ih_f, const float *b_hh_f,
                                      const float *w_ih_r, const float *w_hh_r, const float *b_ih_r, const float *b_hh_r,
                                      int I, int H, void *packed, rcnn_stream_t stream) {
    return rcnn_lstm_pack_weights_parts(w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_r, w_hh_r, b_ih_r, b_hh_r, I, H, packed, 3, stream);
}

extern "C" int rcnn_lstm_pack_weights_parts(const float *w_ih_f, const float *w_hh_f, const float *b_ih_f, const float *b_hh_f,
                                            const float *w_ih_r, const float *w_hh_r, const float *b_ih_r, const float *b_hh_r,
                                            int I, int H, void *packed, int parts, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(I > 0 && H > 0 && H % 32 == 0, "lstm_pack: bad sizes I=%d H=%d", I, H);
    RCNN_CHECK_ARG(parts >= 1 && parts <= 3, "lstm_pack: parts must be 1 (forward views), 2 (transposed views) or 3");
    RCNN_CHECK_ARG(w_ih_f && w_hh_f && b_ih_f && b_hh_f && w_ih_r && w_hh_r && b_ih_r && b_hh_r && packed,
                   "lstm_pack: null pointer");
    PackArgs a;
    a.w_ih[0] = w_ih_f; a.w_hh[0] = w_hh_f; a.b_ih[0] = b_ih_f; a.b_hh[0] = b_hh_f;
    a.w_ih[1] = w_ih_r; a.w_hh[1] = w_hh_r; a.b_ih[1] = b_ih_r; a.b_hh[1] = b_hh_r;
    a.I = I; a.H = H;
    const size_t H8 = 8 * (size_t)H;
    char *p = (char *)packed;
    a.wih_p = (__nv_bfloat16 *)p;  p += H8 * I * 2;
    a.bias_p = (float *)p;         p += H8 * 4;
    a.whh_p = (__nv_bfloat16 *)p;  p += H8 * H * 2;
    a.whh_pt = (__nv_bfloat16 *)p; p += H8 * H * 2;
    a.wih_pt = (__nv_bfloat16 *)p;
    if (parts & 1) {
        lstm_pack_rows_kernel<<<(unsigned)H8, 128, 0, (cudaStream_t)stream>>>(a);
        RCNN_LAUNCH_CHECK("lstm_pack_rows_kernel");
    }
    if (parts & 2) {
        const int kmax = I > H ? I : H;
        dim3 grid((kmax + 31) / 32, 4 * H / 64, 4), block(32, 8);
        lstm_pack_transposed_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a);
        RCNN_LAUNCH_CHECK("lstm_pack_transposed_kernel");
    }
    return RCNN_OK;
}

extern "C" int rcnn_cast_bf16_3d(const float *src, int64_t sb, int64_t st, int64_t sc, void *dst, int B, int T, int C,
                                 rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 0 && C >= 0, "cast_bf16_3d: bad shape");
    if (B == 0 || T == 0 || C == 0) return RCNN_OK;
    RCNN_CHECK_ARG(src && dst, "cast_bf16_3d: null pointer");
    if (sc == 1 && (C & 7) == 0 && (sb & 3) == 0 && (st & 3) == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0) {
        const long long total = (long long)B * T * (C / 8);
        const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        cast3_bf16_vec_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, sb, st, (__nv_bfloat16 *)dst, B, T, C / 8);
        RCNN_LAUNCH_CHECK("cast3_bf16_vec_kernel");
        return RCNN_OK;
    }
    RCNN_CHECK_ARG(B <= 65535, "cast_bf16_3d: batch %d exceeds the grid limit", B);
    dim3 grid((C + 31) / 32, (T + 31) / 32, B), block(32, 8);
    cast3_bf16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, sb, st, sc, (__nv_bfloat16 *)dst, B, T, C);
    RCNN_LAUNCH_CHECK("cast3_bf16_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_cast_bf16_2d(const float *src, int64_t ld_src, void *dst, int64_t ld_dst, int64_t rows, int cols,
                                 rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(rows >= 0 && cols >= 0 && ld_dst >= cols, "cast_bf16_2d: bad shape");
    if (rows == 0 || ld_dst == 0) return RCNN_OK;
    RCNN_CHECK_ARG(src && dst, "cast_bf16_2d: null pointer");
    const long long total = rows * ld_dst;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    cast_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__nv_bfloat16 *)dst, ld_dst, rows, cols);
    RCNN_LAUNCH_CHECK("cast_bf16_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_transpose_bf16(const void *src, int64_t ld, void *dst, int64_t ldo, int R, int C,
                                   rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(R >= 0 && C >= 0, "transpose_bf16: bad shape");
    if (R == 0 || C == 0) return RCNN_OK;
    RCNN_CHECK_ARG(src && dst && ldo >= R, "transpose_bf16: null pointer or ldo < R");
    dim3 grid((C + 63) / 64, (R + 63) / 64), block(32, 8);
    RCNN_CHECK_ARG(grid.y <= 65535, "transpose_bf16: too many rows");
    transpose_bf16_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)src, ld,
                                                                   (__nv_bfloat16 *)dst, ldo, R, C);
    RCNN_LAUNCH_CHECK("transpose_bf16_kernel");
    return RCNN_OK;
}
