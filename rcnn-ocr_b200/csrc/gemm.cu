// K1: dense bf16 GEMM on the 5th-gen tensor cores: tcgen05.mma with the accumulator in TMEM,
// operands staged in shared memory by TMA (128-byte swizzle), mbarrier pipeline.
//
//     D[M,N] = A[M,K] * B[N,K]^T (+ bias[N])        A, B bf16 row-major; D fp32 or bf16
//
// This is the LSTM input projection for all timesteps at once (the `W_ih x_t + b` half of
// nn.LSTM at model/model.py:154-156,161, M = T*B rows), the block's nn.Linear(2H -> H)
// (model/model.py:157,162) and the CTC head; the backward pass reuses it for dX and dW.
//
// 128 x 128 output tiles, warp-specialised: warps 0 and 6 = TMA producers (one elected lane each,
// one operand each: a single thread gets one 16 KB box per ~410 cycles out of the TMA unit,
// independent issuers run in parallel -- see profiles/tma_rate_r01.txt), warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2-5 = epilogue (one warp per TMEM lane quadrant:
// tcgen05.ld -> +bias -> convert -> global).  The main kernel is persistent with two TMEM
// accumulators (see gemm_tn_kernel); the weight-gradient kernel (gemm_atb_kernel) is split-K.
// M / N / K tails: TMA zero-fills out-of-bounds rows and columns; stores are masked.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "sm100.cuh"

namespace rcnn {

// ---- tensor maps (host) -----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int encode(CUtensorMap *out, const void *base, int elem_bytes, int rank, const cuuint64_t *gdim,
                  const cuuint64_t *gstride, const cuuint32_t *box, int swizzle128) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return RCNN_ERR_DEVICE;
    }
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle128 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                    : (swizzle128 == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): base=%p rank=%d dims=(%llu,%llu) pitch=%llu box=(%u,%u)", (int)r,
                  base, rank, (unsigned long long)gdim[0], (unsigned long long)gdim[1],
                  (unsigned long long)gstride[0], box[0], box[1]);
        return RCNN_ERR_ARG;
    }
    return RCNN_OK;
}

int make_tmap_2d(CUtensorMap *out, const void *base, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols, int swizzle128) {
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {row_pitch_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    return encode(out, base, elem_bytes, 2, gdim, gstride, box, swizzle128);
}

int make_tmap_3d(CUtensorMap *out, const void *base, int elem_bytes, uint64_t d2, uint64_t rows, uint64_t cols,
                 uint64_t pitch2_bytes, uint64_t row_pitch_bytes, uint32_t box2, uint32_t box_rows,
                 uint32_t box_cols, int swizzle128) {
    const cuuint64_t gdim[3] = {cols, rows, d2};
    const cuuint64_t gstride[2] = {row_pitch_bytes, pitch2_bytes};
    const cuuint32_t box[3] = {box_cols, box_rows, box2};
    return encode(out, base, elem_bytes, 3, gdim, gstride, box, swizzle128);
}

// N-D map (rank <= 5) with caller-given dimension order (dims[0] innermost, strides[i] = byte stride of dims[i+1])
int make_tmap_nd(CUtensorMap *out, const void *base, int elem_bytes, int rank, const uint64_t *dims, const uint64_t *strides,
                 const uint32_t *box, int swizzle128) {
    cuuint64_t gdim[5], gstride[4];
    cuuint32_t b[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) gstride[i] = strides[i];
    return encode(out, base, elem_bytes, rank, gdim, gstride, b, swizzle128);
}

// 4-D map with caller-given dimension order (dims[0] innermost, strides[i] = byte stride of dims[i+1]):
// lets one TMA operation gather several K chunks of a row tile into consecutive UMMA-shaped boxes.
int make_tmap_4d(CUtensorMap *out, const void *base, int elem_bytes, const uint64_t dims[4], const uint64_t strides[3],
                 const uint32_t box[4], int swizzle128) {
    const cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
    const cuuint64_t gstride[3] = {strides[0], strides[1], strides[2]};
    const cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
    return encode(out, base, elem_bytes, 4, gdim, gstride, b, swizzle128);
}

namespace {

using namespace sm100;

constexpr int BM = 128, BN = 128, BK = 64, UK = 16;
constexpr int kStages = 3;
constexpr int kThreads = 288;   // warp 0,6,7,8 TMA producers, warp 1 MMA, warps 2-5 epilogue
constexpr uint32_t kABytes = BM * BK * 2, kBBytes = BN * BK * 2;
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr size_t kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + 256 /*barriers*/ + BN * sizeof(float);
constexpr size_t kAtbSmemBytes = kSmemBytes;   // the epilogue stages its chunks in the (by then idle) operand ring: 2 CTAs per SM

// vec_ok: 32-byte aligned row chunks -> 256-bit stores (full sectors per lane)
__device__ __forceinline__ void store_row_chunk(float *dst, const float (&v)[32], int ncols, bool vec_ok) {
    if (vec_ok && ncols == 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            U8 q;
#pragma unroll
            for (int k = 0; k < 8; ++k) q.v[k] = __float_as_uint(v[j + k]);
            st_v8(dst + j, q);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < ncols) dst[j] = v[j];
    }
}
__device__ __forceinline__ void store_row_chunk(__nv_bfloat16 *dst, const float (&v)[32], int ncols, bool vec_ok) {
    if (vec_ok && ncols == 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 16) {
            U8 q;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[j + 2 * k], v[j + 2 * k + 1]);
                q.v[k] = *reinterpret_cast<uint32_t *>(&h);
            }
            st_v8(dst + j, q);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < ncols) dst[j] = __float2bfloat16_rn(v[j]);
    }
}

__device__ __forceinline__ void store_row_chunk(__half *dst, const float (&v)[32], int ncols, bool vec_ok) {
    if (vec_ok && ncols == 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 16) {
            U8 q;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                __half2 h = __floats2half2_rn(v[j + 2 * k], v[j + 2 * k + 1]);
                q.v[k] = *reinterpret_cast<uint32_t *>(&h);
            }
            st_v8(dst + j, q);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < ncols) dst[j] = __float2half_rn(v[j]);
    }
}

// Persistent main kernel.  One CTA per SM walks a static list of 128 x 128 output tiles (tile_n
// fastest, so neighbouring CTAs share A row-blocks in L2).  A pipeline stage holds K = 128 (two TMA
// boxes per operand, 64 KB): one barrier round trip then feeds eight tcgen05.mma -- with K = 64 per
// round the issuing thread spent as long on wait/commit bookkeeping as the tensor pipe spent on
// the four MMAs (profiles/timeline_r01.txt).  Two 128-column TMEM accumulators alternate between
// tiles, so the epilogue warps drain tile i while the MMA warp is already on tile i+1.
// Tile width TN: 128 (K = 128 per stage, 3 stages) or 256 (K = 64 per stage, 4 stages); 64 and 32 (K = 128 per stage, 4
// stages) exist for products with few rows (the attention decoder's per-step GEMMs, M = batch): there the time is what ONE SM
// can pull through its L2 port (~85 GB/s: a CTA of a 16-tile launch reads 768 KB), so narrow tiles spread the B operand over
// 4-8x as many SMs.  With 128 x 128 tiles
// every tile re-reads (128 + 128) x K operand elements for 128 x 128 x K MACs = 64 FLOP per L2 byte, and the
// large GEMMs of the step then sit exactly at the L2 -> SM bandwidth (~12.5 TB/s: 830 TFLOP/s); 128 x 256
// tiles raise that to 85 FLOP per byte and need half as many tcgen05.mma instructions (N = 256 each).
constexpr int kPThreads = 352;                               // warps: 0 A-TMA, 1 MMA, 2-5 and 7-10 epilogue, 6 B-TMA
template <int TN> struct PCfg {
    static constexpr int kBoxes = TN <= 128 ? 2 : 1;                       // k-boxes of 64 per stage
    static constexpr int kStages = TN == 128 ? 3 : 4;
    static constexpr uint32_t kAOp = kBoxes * kABytes;                     // A bytes per stage
    static constexpr uint32_t kBBox = TN * BK * 2;                         // one B box [TN x 64] bf16
    static constexpr uint32_t kBOp = kBoxes * kBBox;
    static constexpr uint32_t kStage = kAOp + kBOp;
    static constexpr size_t kSmem = 1024 + kStages * kStage + 256 + 2 * TN * sizeof(float);
};

template <typename OutT> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<float>(float a, float) { return __float_as_uint(a); }

// CELL epilogue (TN = 32, the attention decoder's gate product, attn.cu): the B rows are gate-interleaved (row 4u + g = gate g
// of hidden unit u, torch order i, f, g, o), so the 32 accumulator columns a thread holds are all four gates of eight units and
// the LSTMCell pointwise step (K6b) runs on them in registers: nothing is written to D.
struct CellEpi {
    const float *embT;           // [V, 4H] gate-interleaved: the one-hot half of the LSTMCell input, one row per token
    const long long *y;          // [M] previous tokens
    float *c;                    // [M, H] cell state c_t out (in place when c_in == c)
    const float *c_in;           // [M, H] c_{t-1}
    float *gates_act;            // optional [M, 4H] gate-interleaved: sigmoid(i), sigmoid(f), tanh(g), sigmoid(o), kept for backward
    __nv_bfloat16 *h_out;        // h_t as bf16, row pitch h_ld (NOT the A operand's buffer: other CTAs still read that)
    float *hid_out;              // optional f32 copy, row pitch hid_ld
    long long h_ld, hid_ld;
    int V, H;
};
__device__ __forceinline__ float tanh_fast_g(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sigmoid_fast_g(float x) { return fmaf(tanh_fast_g(0.5f * x), 0.5f, 0.5f); }

template <typename OutT, int TN, bool CELL = false>
__global__ void __launch_bounds__(kPThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               OutT *__restrict__ D, long long ldd, const float *__restrict__ bias, int M, int N, int K, const CellEpi ce) {
    static_assert(!CELL || TN == 32, "the cell epilogue works on 32-column tiles");
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using C = PCfg<TN>;
    constexpr int kPStages = C::kStages, kPBoxes = C::kBoxes;
    constexpr uint32_t kPStage = C::kStage, kPAOp = C::kAOp, kPBBox = C::kBBox;
    unsigned char *tiles = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kPStages * kPStage);
    uint64_t *empty = full + kPStages;
    uint64_t *tmem_full = empty + kPStages;      // [2]
    uint64_t *tmem_empty = tmem_full + 2;        // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *bias_s = reinterpret_cast<float *>(smem + kPStages * kPStage + 256);   // [2][TN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntn = (N + TN - 1) / TN, ntm = (M + BM - 1) / BM;
    const int num_tiles = ntn * ntm;
    const int rounds = (K + PCfg<TN>::kBoxes * BK - 1) / (PCfg<TN>::kBoxes * BK);

    griddep_launch();                                         // (chained launches: the successor may set up meanwhile)
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kPStages; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<2 * TN>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();                                           // the predecessor's output is complete from here on

    if (warp == 0 || warp == 6) {
        // ===== TMA producers: warp 0 streams A, warp 6 streams B (independent issuers) =============
        if (elect_one()) {
            const bool isA = warp == 0;
            int st = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int row0 = isA ? (tile / ntn) * BM : (tile % ntn) * TN;
                for (int r = 0; r < rounds; ++r) {
                    mbar_wait(&empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&full[st], isA ? kPAOp : C::kBOp);
                    unsigned char *dst = tiles + st * kPStage + (isA ? 0 : kPAOp);
#pragma unroll
                    for (int j = 0; j < kPBoxes; ++j)
                        tma_load_2d(dst + j * (isA ? kABytes : kPBBox), isA ? &tmA : &tmB, &full[st], (r * kPBoxes + j) * BK, row0);
                    if (++st == kPStages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer ==========================================================================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, TN);
            int st = 0, it = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);     // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TN);
                for (int r = 0; r < rounds; ++r) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < kPBoxes; ++j) {
                        const uint64_t adesc = make_smem_desc_sw128(smem_u32(tiles + st * kPStage + j * kABytes), 16, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(tiles + st * kPStage + kPAOp + j * kPBBox), 16, 1024);
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (r | j | k) != 0);
                    }
                    umma_commit(&empty[st]);
                    if (++st == kPStages) { st = 0; ph ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp != 6) {
        // ===== epilogue: TMEM -> registers -> (+bias, convert) -> global.  Eight warps: two per TMEM lane
        // quadrant (warp % 4), each draining half of the tile's columns -- with K = 512 a tile is only
        // 32 MMAs (~2,000 cycles) and four warps took about as long to drain it. =====================
        const int q = warp & 3;
        const int chalf = warp >= 7 ? 1 : 0;
        constexpr int kCHalf = TN >= 64 ? TN / 2 : TN;                        // (a 32-column tile is drained by four warps)
        const int etid = warp < 6 ? threadIdx.x - 64 : threadIdx.x - 96;    // 0..255 over the epilogue threads
        const bool vec_ok = (ldd % (32 / (long long)sizeof(OutT)) == 0) && ((reinterpret_cast<uintptr_t>(D) & 31) == 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int tile_m = tile / ntn, tile_n = tile % ntn;
            {   // this tile's 128 bias values -> shared memory (broadcast reads below)
                const int j = etid, col = tile_n * TN + j;
                if (j < TN) bias_s[acc * TN + j] = (bias != nullptr && col < N) ? __ldg(bias + col) : 0.f;
            }
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            asm volatile("bar.sync 2, 256;" ::: "memory");   // bias_s visible to the eight epilogue warps
            tc_fence_after();
            const int row = tile_m * BM + q * 32 + lane;
#pragma unroll 1
            for (int c0 = chalf * kCHalf; c0 < (chalf + 1) * kCHalf && c0 < TN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TN + c0), r);
                tmem_ld_wait();
                const int col0 = tile_n * TN + c0;
                if (CELL) {
                    if (row < M) {                                          // (N = 4H is a multiple of 32)
                        long long tok = ce.y[row];
                        tok = tok < 0 ? 0 : (tok >= ce.V ? ce.V - 1 : tok);
                        const float4 *em4 = reinterpret_cast<const float4 *>(ce.embT + (size_t)tok * 4 * ce.H + col0);
                        const int unit0 = col0 >> 2;
                        float4 *c4 = reinterpret_cast<float4 *>(ce.c + (size_t)row * ce.H + unit0);
                        const float4 *ci4 = reinterpret_cast<const float4 *>(ce.c_in + (size_t)row * ce.H + unit0);
                        float4 *ga4 = ce.gates_act ? reinterpret_cast<float4 *>(ce.gates_act + (size_t)row * 4 * ce.H + col0) : nullptr;
                        const float4 cA = ci4[0], cB = ci4[1];
                        const float cp[8] = {cA.x, cA.y, cA.z, cA.w, cB.x, cB.y, cB.z, cB.w};
                        float cn[8], hn[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float4 em = __ldg(em4 + u);
                            const float ig = sigmoid_fast_g(__uint_as_float(r[4 * u]) + bias_s[acc * TN + 4 * u] + em.x);
                            const float fg = sigmoid_fast_g(__uint_as_float(r[4 * u + 1]) + bias_s[acc * TN + 4 * u + 1] + em.y);
                            const float gg = tanh_fast_g(__uint_as_float(r[4 * u + 2]) + bias_s[acc * TN + 4 * u + 2] + em.z);
                            const float og = sigmoid_fast_g(__uint_as_float(r[4 * u + 3]) + bias_s[acc * TN + 4 * u + 3] + em.w);
                            cn[u] = fmaf(fg, cp[u], ig * gg);
                            hn[u] = og * tanh_fast_g(cn[u]);
                            if (ga4) ga4[u] = make_float4(ig, fg, gg, og);
                        }
                        c4[0] = make_float4(cn[0], cn[1], cn[2], cn[3]);
                        c4[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
                        uint4 hb;
                        hb.x = pack2<__nv_bfloat16>(hn[0], hn[1]); hb.y = pack2<__nv_bfloat16>(hn[2], hn[3]);
                        hb.z = pack2<__nv_bfloat16>(hn[4], hn[5]); hb.w = pack2<__nv_bfloat16>(hn[6], hn[7]);
                        *reinterpret_cast<uint4 *>(ce.h_out + (size_t)row * ce.h_ld + unit0) = hb;
                        if (ce.hid_out) {
                            float4 *o4 = reinterpret_cast<float4 *>(ce.hid_out + (size_t)row * ce.hid_ld + unit0);
                            o4[0] = make_float4(hn[0], hn[1], hn[2], hn[3]);
                            o4[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
                        }
                    }
                } else if (row < M && col0 < N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias_s[acc * TN + c0 + j];
                    store_row_chunk(D + (long long)row * ldd + col0, v, min(32, N - col0), vec_ok);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * TN>(tmem_base);
    }
}


// ---- CTA-pair variant: tcgen05.mma.cta_group::2, 256 x 256 output tile per pair ----------------------------
// The two CTAs of a cluster compute one 256 x 256 tile: CTA r holds A rows [128r, 128r+128) and HALF of the B
// tile (N rows [128r, 128r+128)); the leader (even CTA) issues M = 256, N = 256 MMAs that read both CTAs'
// shared memory, each CTA accumulates its own 128 rows x 256 columns in its own TMEM.  Per stage (K = 64) a
// CTA ingests 32 KB for 128 x 256 x 64 MACs: 128 FLOP per L2 byte, twice the single-CTA 128 x 256 tile --
// the large GEMMs of the step are bound by L2 -> SM bandwidth, not by the tensor pipe.
// Epilogue: a thread owns one accumulator row, so direct stores touch 32 different rows per instruction and the
// LSU, not HBM, bounded the kernel (7,100 cycles to drain a tile against 4,400 cycles of MMAs, measured with
// scripts/timeline_gemm.py).  Each epilogue warp therefore stages its [32 rows x 32 columns] chunks in shared
// memory (swizzled, conflict-free) and hands them to TMA stores, which also clip the M / N tails.
constexpr int kQStages = 4;
constexpr uint32_t kQStage = 2 * kABytes;                          // A box + B-half box, 16 KB each
constexpr uint32_t kQOutBuf = 4096;                                // one staged chunk: 32 rows x 32 columns x <= 4 bytes
constexpr size_t kQSmem = 1024 + kQStages * kQStage + 8 * 2 * kQOutBuf + 256 + 2 * 256 * sizeof(float);

template <typename OutT>
__global__ void __launch_bounds__(kPThreads, 1)
gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    OutT *__restrict__ D, long long ldd, const float *__restrict__ bias, int M, int N, int K,
                    long long *tlbuf, const __grid_constant__ CUtensorMap tmD, int tma_out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int TN = 256;
    unsigned char *tiles = smem;
    unsigned char *outbuf = smem + kQStages * kQStage;                            // [8 warps][2] staged output chunks
    uint64_t *full = reinterpret_cast<uint64_t *>(outbuf + 8 * 2 * kQOutBuf);    // used in the leader only
    uint64_t *empty = full + kQStages;
    uint64_t *tmem_full = empty + kQStages;      // [2]
    uint64_t *tmem_empty = tmem_full + 2;        // [2], used in the leader only (16 arrivals: both CTAs' epilogue warps)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *bias_s = reinterpret_cast<float *>(outbuf + 8 * 2 * kQOutBuf + 256);  // [2][TN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int pid = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    long long *tl = blockIdx.x == 0 ? tlbuf : nullptr;       // debug: per-tile clock64 marks of CTA 0 (see rcnn_debug_timeline)
#define GT_MARK(k) do { if (tl && it < 64) tl[it * 8 + (k)] = clock64(); } while (0)
    const int ntn = (N + TN - 1) / TN, ntm = (M + 255) / 256;
    const int num_tiles = ntn * ntm;
    const int rounds = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kQStages; ++s) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 16); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc_2sm<2 * TN>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    // both CTAs' barriers and TMEM allocations exist before either signals the other
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 || warp == 6) {
        // ===== TMA producers (both CTAs): warp 0 its A rows, warp 6 its half of the B tile; every load
        // completes on the LEADER's full barrier, which the leader's two producers arm for both CTAs =====
        if (elect_one()) {
            const bool isA = warp == 0;
            int st = 0;
            uint32_t ph = 0;
            for (int tile = pid; tile < num_tiles; tile += npairs) {
                const int row0 = isA ? (tile / ntn) * 256 + (int)rank * 128 : (tile % ntn) * TN + (int)rank * 128;
                for (int r = 0; r < rounds; ++r) {
                    mbar_wait(&empty[st], ph ^ 1);
                    if (leader) mbar_arrive_expect_tx(&full[st], 2 * kABytes);      // this operand: own box + the peer's
                    tma_load_2d_2sm(tiles + st * kQStage + (isA ? 0 : kABytes), isA ? &tmA : &tmB, &full[st], r * BK, row0);
                    if (++st == kQStages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader's elected thread drives both tensor cores =======================
        if (leader && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(256, TN);
            int st = 0, it = 0;
            uint32_t ph = 0;
            for (int tile = pid; tile < num_tiles; tile += npairs, ++it) {
                const int acc = it & 1;
                GT_MARK(0);
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);     // both epilogues drained this accumulator
                GT_MARK(1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TN);
                for (int r = 0; r < rounds; ++r) {
                    mbar_wait(&full[st], ph);
                    if (r == 0) GT_MARK(2);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc_sw128(smem_u32(tiles + st * kQStage), 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(tiles + st * kQStage + kABytes), 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k)
                        umma_bf16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (r | k) != 0);
                    umma_commit_2sm(&empty[st], (uint16_t)3);          // frees the stage in both CTAs
                    if (++st == kQStages) { st = 0; ph ^= 1; }
                }
                umma_commit_2sm(&tmem_full[acc], (uint16_t)3);
                GT_MARK(3);
            }
        }
    } else {
        // ===== epilogue (both CTAs): own 128 rows x 256 columns ==========================================
        const int q = warp & 3;
        const int chalf = warp >= 7 ? 1 : 0;
        const int etid = warp < 6 ? threadIdx.x - 64 : threadIdx.x - 96;    // 0..255 over the epilogue threads
        const bool vec_ok = (ldd % (32 / (long long)sizeof(OutT)) == 0) && ((reinterpret_cast<uintptr_t>(D) & 31) == 0);
        int it = 0;
        for (int tile = pid; tile < num_tiles; tile += npairs, ++it) {
            const int acc = it & 1;
            const int tile_m = tile / ntn, tile_n = tile % ntn;
            {
                const int col = tile_n * TN + etid;
                bias_s[acc * TN + etid] = (bias != nullptr && col < N) ? __ldg(bias + col) : 0.f;
            }
            mbar_wait(&tmem_full[acc], (it >> 1) & 1);
            if (threadIdx.x == 64) GT_MARK(4);
            asm volatile("bar.sync 2, 256;" ::: "memory");
            tc_fence_after();
            const int row = tile_m * 256 + (int)rank * 128 + q * 32 + lane;
            const int ew = etid >> 5;                                   // epilogue warp 0..7
            int nchunk = 0;
#pragma unroll 1
            for (int c0 = chalf * (TN / 2); c0 < (chalf + 1) * (TN / 2); c0 += 32, ++nchunk) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TN + c0), r);
                tmem_ld_wait();
                if (c0 + 32 == (chalf + 1) * (TN / 2)) {                // accumulator fully read: hand it back before the stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
                }
                const int col0 = tile_n * TN + c0;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias_s[acc * TN + c0 + j];
                if (tma_out) {
                    // 8 KB of staging per warp = two blocks [32 rows x 128 bytes] (SWIZZLE_128B): one TMA store (3-D
                    // map: column-in-block, row, block) writes both -- the store engine, not the LSU, is the limiter
                    // once the rows are coalesced, so fewer and larger operations matter
                    constexpr int kCPB = sizeof(OutT) == 2 ? 2 : 1;                 // 32-column chunks per 128-byte row block
                    unsigned char *buf = outbuf + (size_t)ew * 2 * kQOutBuf;
                    const int blk = (nchunk / kCPB) & 1, sub = nchunk % kCPB;
                    if ((nchunk % (2 * kCPB)) == 0) {                               // first chunk of a store: buffer free?
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
                    }
                    unsigned char *rowp = buf + blk * kQOutBuf + lane * 128;
                    const int sw = lane & 7;
                    if (sizeof(OutT) == 2) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 o;
                            uint32_t *ow = reinterpret_cast<uint32_t *>(&o);
#pragma unroll
                            for (int k = 0; k < 4; ++k) ow[k] = pack2<OutT>(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
                            *reinterpret_cast<uint4 *>(rowp + (((sub * 4 + j) ^ sw) << 4)) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4 *>(rowp + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    if ((nchunk % (2 * kCPB)) == 2 * kCPB - 1) {                    // both blocks staged: store
                        fence_proxy_async_smem();
                        __syncwarp();
                        const int colb = (col0 + 32 - 2 * kCPB * 32) / (32 * kCPB);   // first 128-byte column block of the store
                        if (lane == 0 && tile_m * 256 + (int)rank * 128 + q * 32 < M) {
                            tma_store_3d(&tmD, buf, 0, tile_m * 256 + (int)rank * 128 + q * 32, colb);
                            tma_store_commit();
                        }
                    }
                } else if (row < M && col0 < N) {
                    store_row_chunk(D + (long long)row * ldd + col0, v, min(32, N - col0), vec_ok);
                }
            }
            if (threadIdx.x == 64) GT_MARK(5);
        }
    }
#undef GT_MARK
    if (tma_out && warp >= 2 && warp != 6 && lane == 0) tma_store_wait<0>();
    // no CTA leaves while its peer may still read its shared memory / signal its barriers
    tc_fence_before();
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<2 * TN>(tmem_base);
    }
}

// ---- D[M,N] += A[K,M]^T * B[K,N]  (weight-gradient shape) -------------------------------------
// Both operands are "MN-major": the contraction index k is the ROW of the row-major global
// arrays, so no transposed copies are needed for dW = dY^T X.  A k-block of 64 rows is staged as
// two TMA boxes per operand (64 k-rows x 64 contiguous M/N elements = 128-byte swizzled rows);
// the UMMA descriptors use the MN-major canonical layout (LBO = distance between the two
// 64-element halves, SBO = 8 k-rows) and idesc.a_major = idesc.b_major = 1.  Split-K over
// blockIdx.z (K = T*B is 16384 while M*N gives only 32-128 tiles): fp32 partial tiles are
// added with red.global.add.v4.f32 into a zero-initialised D.
constexpr uint32_t kHalf = 64 * BK * 2;   // one 64-wide box: 64 k-rows x 128 B = 8 KB

__device__ __forceinline__ void red_add_v4(float *dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads)
gemm_atb_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                float *__restrict__ D, long long ldd, int M, int N, int K, int kb_per_split, int splits,
                int a_gcols, int b_gcols, long long d_goff, const __grid_constant__ CUtensorMap tmD, int tma_out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *tiles = smem;
    unsigned char *outbuf = smem;          // [4 warps] staged 32 x 32 fp32 chunks: the operand ring, idle once tmem_full fired
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
    uint64_t *empty = full + kStages;
    uint64_t *tmem_full = empty + kStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_n = blockIdx.x, tile_m = blockIdx.y;
    const int total_kb = (K + BK - 1) / BK;
    const int grp = blockIdx.z / splits;              // independent problems sharing one launch
    const int kb0 = (blockIdx.z % splits) * kb_per_split;
    const int num_kb = min(kb_per_split, total_kb - kb0);
    if (num_kb <= 0) return;
    const int a_c0 = grp * a_gcols + tile_m * BM, b_c0 = grp * b_gcols + tile_n * BN;
    D += grp * d_goff;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 4); mbar_init(&empty[s], 1); }
            mbar_init(tmem_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<BN>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 || warp >= 6) {
        if (elect_one()) {
            const int which = warp == 0 ? 0 : warp - 5;      // 0,1: the two A halves; 2,3: the two B halves
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], kHalf);
                unsigned char *st = tiles + s * kStageBytes;
                const int k0 = (kb0 + kb) * BK;
                if (which == 0) tma_load_2d(st, &tmA, &full[s], a_c0, k0);
                else if (which == 1) tma_load_2d(st + kHalf, &tmA, &full[s], a_c0 + 64, k0);
                else if (which == 2) tma_load_2d(st + kABytes, &tmB, &full[s], b_c0, k0);
                else tma_load_2d(st + kABytes + kHalf, &tmB, &full[s], b_c0 + 64, k0);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t adesc = make_smem_desc_sw128(smem_u32(tiles + s * kStageBytes), kHalf, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(tiles + s * kStageBytes + kABytes), kHalf, 1024);
#pragma unroll
                for (int k = 0; k < BK / UK; ++k)   // 16 k-rows = 2048 bytes further down each half
                    umma_bf16(tmem_base, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, (kb | k) != 0);
                umma_commit(&empty[s]);
            }
            umma_commit(tmem_full);
        }
    } else if (warp < 6) {
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = tile_m * BM + q * 32 + lane;
        const bool vec_ok = (ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(D) & 15) == 0);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            const int col0 = tile_n * BN + c0;
            if (tma_out) {
                // a thread owns one accumulator row: direct red.add touches 32 rows per instruction (LSU-bound);
                // stage the chunk (128-byte rows, SWIZZLE_128B) and let TMA do a coalesced reduce-add that also
                // clips the M / N tails
                if (tile_m * BM + q * 32 < M && col0 < N) {
                    unsigned char *buf = outbuf + (warp - 2) * 4096;
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                    unsigned char *rowp = buf + lane * 128;
                    const int sw = lane & 7;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4 *>(rowp + ((j ^ sw) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_3d(&tmD, buf, col0, tile_m * BM + q * 32, grp);
                        tma_store_commit();
                    }
                }
            } else if (row < M && col0 < N) {
                float *dst = D + (long long)row * ldd + col0;
                const int ncols = min(32, N - col0);
                if (vec_ok && ncols == 32) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        red_add_v4(dst + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) atomicAdd(dst + j, __uint_as_float(r[j]));
                }
            }
        }
        if (tma_out && lane == 0) tma_store_wait<0>();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<BN>(tmem_base);
    }
}

// ---- CTA-pair weight-gradient GEMM: D[256 x 256 tile] += A^T B over a K range, tcgen05.mma.cta_group::2 ---------
// Same operand layout as gemm_atb_kernel (MN-major boxes of 64 k-rows x 64 columns) and the same split-K +
// TMA reduce-add output; the pair mechanics are those of gemm_tn_pair_kernel: CTA r stages its own 128 columns of
// A (rows of D) and 128 of the tile's 256 B columns, every load completes on the leader's barrier, the leader
// issues M = 256, N = 256 MMAs for both tensor cores.  A CTA ingests 32 KB per 128 x 256 x 64 MACs instead of
// 32 KB per 128 x 128 x 64: the single-CTA kernel sat at the per-SM L2 ingest limit (two CTAs x 32 KB per 256
// cycles of MMA = 125 B/clk).
//
// The operands are read through 3-D maps (column, inner k, outer k): a flat [K, cols] matrix is (cols, K, 1); the
// LSTM's dW_hh uses (cols, t, b) with the B operand's t shifted by -1 / +1 per group -- h_{t-1} of the forward and
// h_{t+1} of the reverse direction straight from hcat, out-of-range rows zero-filled by TMA (x_seq.shift[]).
// x_seq.perm: the 32 packed gate rows of an output piece (unit-major, gate-minor) are scattered to nn.LSTM's
// gate-major row order by a 5-D reduce-add map, so the gradients land in torch layout without an unpack pass.
struct AtbSeq {
    int cps;        // 64-row K chunks per outer index (flat: all of them)
    int shift[2];   // B operand: offset of the inner k coordinate for group 0 / group >= 1
    int perm;       // output rows in packed gate order -> 5-D map (col, gate, unit, column block, group)
};
constexpr int kRStages = 5;
constexpr uint32_t kRStage = 2 * kABytes;                                 // A: two halves, B-half: two halves (8 KB each)
constexpr size_t kRSmem = 1024 + kRStages * kRStage + 4 * 8192 + 256;     // + one 8 KB staging buffer per epilogue warp

__global__ void __launch_bounds__(kThreads, 1)
gemm_atb_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmD, int M, int N, int total_kb, int kb_per_split, int splits,
                     int a_gcols, int b_gcols, const AtbSeq x_seq) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *tiles = smem;
    unsigned char *outbuf = smem + kRStages * kRStage;       // [4 warps][2 blocks of 32 rows x 128 bytes]
    uint64_t *full = reinterpret_cast<uint64_t *>(outbuf + 4 * 8192);       // used in the leader only
    uint64_t *empty = full + kRStages;
    uint64_t *tmem_full = empty + kRStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int tile_n = blockIdx.x >> 1, tile_m = blockIdx.y;        // gridDim.x = 2 * column tiles (cluster along x)
    const int grp = blockIdx.z / splits;
    const int kb0 = (blockIdx.z % splits) * kb_per_split;
    const int num_kb = min(kb_per_split, total_kb - kb0);            // identical in both CTAs of the pair
    if (num_kb <= 0) return;
    const int a_c0 = grp * a_gcols + tile_m * 256 + (int)rank * 128, b_c0 = grp * b_gcols + tile_n * 256 + (int)rank * 128;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmD); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kRStages; ++s) { mbar_init(&full[s], 4); mbar_init(&empty[s], 1); }
            mbar_init(tmem_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc_2sm<256>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 || warp >= 6) {
        if (elect_one()) {
            const int which = warp == 0 ? 0 : warp - 5;      // 0,1: the two A halves; 2,3: the two halves of this CTA's B columns
            const int bshift = x_seq.shift[grp > 0 ? 1 : 0];
            int ko = kb0 / x_seq.cps, ki = (kb0 % x_seq.cps) * BK;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kRStages;
                const uint32_t ph = (kb / kRStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (leader) mbar_arrive_expect_tx(&full[s], 2 * kHalf);          // own box + the peer's twin
                unsigned char *st = tiles + s * kRStage;
                if (which == 0) tma_load_3d_2sm(st, &tmA, &full[s], a_c0, ki, ko);
                else if (which == 1) tma_load_3d_2sm(st + kHalf, &tmA, &full[s], a_c0 + 64, ki, ko);
                else if (which == 2) tma_load_3d_2sm(st + kABytes, &tmB, &full[s], b_c0, ki + bshift, ko);
                else tma_load_3d_2sm(st + kABytes + kHalf, &tmB, &full[s], b_c0 + 64, ki + bshift, ko);
                ki += BK;
                if (ki >= x_seq.cps * BK) { ki = 0; ++ko; }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(256, 256, 1, 1);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kRStages;
                const uint32_t ph = (kb / kRStages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t adesc = make_smem_desc_sw128(smem_u32(tiles + s * kRStage), kHalf, 1024);
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(tiles + s * kRStage + kABytes), kHalf, 1024);
#pragma unroll
                for (int k = 0; k < BK / UK; ++k)
                    umma_bf16_2sm(tmem_base, adesc + (uint64_t)(128 * k), bdesc + (uint64_t)(128 * k), idesc, (kb | k) != 0);
                umma_commit_2sm(&empty[s], (uint16_t)3);
            }
            umma_commit_2sm(tmem_full, (uint16_t)3);
        }
    } else if (warp < 6) {
        // epilogue: this CTA's 128 rows x 256 columns, 64 columns (two 128-byte blocks) per TMA reduce-add
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        unsigned char *buf = outbuf + (warp - 2) * 8192;
        const int row0 = tile_m * 256 + (int)rank * 128 + q * 32;
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            const int blk = (c0 >> 5) & 1;
            if (blk == 0) {
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
            }
            unsigned char *rowp = buf + blk * 4096 + lane * 128;
            const int sw = lane & 7;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4 *>(rowp + ((j ^ sw) << 4)) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            if (blk == 1) {
                fence_proxy_async_smem();
                __syncwarp();
                const int col0 = tile_n * 256 + c0 - 32;
                if (lane == 0 && row0 < M && col0 < N) {
                    if (x_seq.perm) tma_reduce_add_5d(&tmD, buf, 0, 0, row0 >> 2, col0 / 32, grp);
                    else tma_reduce_add_4d(&tmD, buf, 0, row0, col0 / 32, grp);
                    tma_store_commit();
                }
            }
        }
        if (lane == 0) tma_store_wait<0>();
        tc_fence_before();
    }
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<256>(tmem_base);
    }
}

template <typename OutT, int TN>
int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, void *D, long long ldd, const float *bias, int M,
                int N, int K, cudaStream_t s) {
    constexpr size_t smem = PCfg<TN>::kSmem;
    RCNN_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<OutT, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ((N + TN - 1) / TN) * ((M + BM - 1) / BM);
    const int grid = tiles < gemm_sms() ? tiles : gemm_sms();
    ProfScope prof(RCNN_K_GEMM, s);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    chain_config(cfg, attr, grid, kPThreads, smem, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, gemm_tn_kernel<OutT, TN>, ta, tb, (OutT *)D, (long long)ldd, bias, M, N, K, CellEpi{}));
    count_launch();
    return RCNN_OK;
}

int launch_gemm_cell(const CUtensorMap &ta, const CUtensorMap &tb, const float *bias, int M, int N, int K, const CellEpi &ce,
                     cudaStream_t s) {
    constexpr size_t smem = PCfg<32>::kSmem;
    RCNN_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<float, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (N / 32) * ((M + BM - 1) / BM);
    const int grid = tiles < gemm_sms() ? tiles : gemm_sms();
    ProfScope prof(RCNN_K_GEMM, s);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    chain_config(cfg, attr, grid, kPThreads, smem, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, gemm_tn_kernel<float, 32, true>, ta, tb, (float *)nullptr, (long long)0, bias, M, N, K, ce));
    count_launch();
    return RCNN_OK;
}

template <typename OutT>
int launch_gemm_pair(const CUtensorMap &ta, const CUtensorMap &tb, void *D, long long ldd, const float *bias, int M,
                     int N, int K, cudaStream_t s) {
    RCNN_CUDA(cudaFuncSetAttribute(gemm_tn_pair_kernel<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQSmem));
    const int tiles = ((N + 255) / 256) * ((M + 255) / 256);
    const int pairs = tiles < gemm_sms() / 2 ? tiles : gemm_sms() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(kPThreads);
    cfg.dynamicSmemBytes = kQSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // TMA-store epilogue when the output rows are 16-byte aligned (else direct stores)
    CUtensorMap td = ta;
    int tma_out = 0;
    constexpr int kCW = 128 / (int)sizeof(OutT);         // columns per 128-byte block
    if (((uintptr_t)D & 15) == 0 && (ldd * (long long)sizeof(OutT)) % 16 == 0 && N % kCW == 0) {
        const uint64_t dims[3] = {(uint64_t)kCW, (uint64_t)M, (uint64_t)(N / kCW)};
        const uint64_t strides[2] = {(uint64_t)ldd * sizeof(OutT), 128};
        const uint32_t box[3] = {(uint32_t)kCW, 32, 2};
        const int rc = make_tmap_nd(&td, D, (int)sizeof(OutT), 3, dims, strides, box, 1);
        if (rc) return rc;
        tma_out = 1;
    }
    ProfScope prof(RCNN_K_GEMM, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, gemm_tn_pair_kernel<OutT>, ta, tb, (OutT *)D, ldd, bias, M, N, K, debug_timeline(), td,
                                 tma_out));
    count_launch();
    return RCNN_OK;
}

int launch_atb_pair(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &td, int M, int N, int tkb, int groups,
                    int a_gcols, int b_gcols, const AtbSeq &seq, cudaStream_t s) {
    const int ptiles = ((M + 255) / 256) * ((N + 255) / 256) * groups;
    int sp = (gemm_sms() / 2) / ptiles;                      // fill the 74 CTA pairs once
    sp = sp < 1 ? 1 : (sp > tkb ? tkb : sp);
    const int kbs = (tkb + sp - 1) / sp;
    sp = (tkb + kbs - 1) / kbs;
    RCNN_CUDA(cudaFuncSetAttribute(gemm_atb_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRSmem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * ((N + 255) / 256)), (unsigned)((M + 255) / 256), (unsigned)(sp * groups));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kRSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ProfScope prof(RCNN_K_GEMM_ATB, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, gemm_atb_pair_kernel, ta, tb, td, M, N, tkb, kbs, sp, a_gcols, b_gcols, seq));
    count_launch();
    return RCNN_OK;
}

}  // namespace
}  // namespace rcnn

extern "C" int rcnn_gemm_bf16(const void *A, int64_t lda, const void *B, int64_t ldb, void *D, int64_t ldd,
                              int out_dtype, const float *bias, int M, int N, int K, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(M >= 0 && N >= 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
    RCNN_CHECK_ARG(out_dtype == RCNN_F32 || out_dtype == RCNN_BF16 || out_dtype == RCNN_F16, "gemm: bad out dtype %d", out_dtype);
    if (M == 0 || N == 0) return RCNN_OK;
    RCNN_CHECK_ARG(A && B && D, "gemm: null pointer");
    RCNN_CHECK_ARG(lda >= K && ldb >= K && ldd >= N, "gemm: leading dimension smaller than the row");
    RCNN_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0,
                   "gemm: A/B rows must be 16-byte aligned (lda=%lld ldb=%lld)", (long long)lda, (long long)ldb);
    CUtensorMap ta, tb;
    int rc = make_tmap_2d(&ta, A, 2, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, BM, BK, 1);
    if (rc) return rc;
    static const int force_tn = getenv("RCNN_GEMM_TN") ? atoi(getenv("RCNN_GEMM_TN")) : 0;
    static const int use_pair = getenv("RCNN_GEMM_PAIR") ? atoi(getenv("RCNN_GEMM_PAIR")) : 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (use_pair && !force_tn && M >= 512 && N >= 256) {      // CTA-pair kernel: 256 x 256 tiles, each CTA loads 128 B rows
        rc = make_tmap_2d(&tb, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 128, BK, 1);
        if (rc) return rc;
        if (out_dtype == RCNN_F32) return launch_gemm_pair<float>(ta, tb, D, ldd, bias, M, N, K, s);
        if (out_dtype == RCNN_F16) return launch_gemm_pair<__half>(ta, tb, D, ldd, bias, M, N, K, s);
        return launch_gemm_pair<__nv_bfloat16>(ta, tb, D, ldd, bias, M, N, K, s);
    }
    int tn = N > 128 ? 256 : 128;
    {   // few row blocks: narrow the tiles until about half of the SMs have one (see the note above gemm_tn_kernel)
        const int ntm = (M + BM - 1) / BM;
        while (tn > 32 && ntm * ((N + tn - 1) / tn) < gemm_sms() / 2) tn >>= 1;
    }
    if (force_tn == 32 || force_tn == 64 || force_tn == 128 || force_tn == 256) tn = force_tn;
    rc = make_tmap_2d(&tb, B, 2, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, (uint32_t)tn, BK, 1);
    if (rc) return rc;
    if (tn == 64) {
        if (out_dtype == RCNN_F32) return launch_gemm<float, 64>(ta, tb, D, ldd, bias, M, N, K, s);
        if (out_dtype == RCNN_F16) return launch_gemm<__half, 64>(ta, tb, D, ldd, bias, M, N, K, s);
        return launch_gemm<__nv_bfloat16, 64>(ta, tb, D, ldd, bias, M, N, K, s);
    }
    if (tn == 32) {
        if (out_dtype == RCNN_F32) return launch_gemm<float, 32>(ta, tb, D, ldd, bias, M, N, K, s);
        if (out_dtype == RCNN_F16) return launch_gemm<__half, 32>(ta, tb, D, ldd, bias, M, N, K, s);
        return launch_gemm<__nv_bfloat16, 32>(ta, tb, D, ldd, bias, M, N, K, s);
    }
    if (tn == 256) {
        if (out_dtype == RCNN_F32) return launch_gemm<float, 256>(ta, tb, D, ldd, bias, M, N, K, s);
        if (out_dtype == RCNN_F16) return launch_gemm<__half, 256>(ta, tb, D, ldd, bias, M, N, K, s);
        return launch_gemm<__nv_bfloat16, 256>(ta, tb, D, ldd, bias, M, N, K, s);
    }
    if (out_dtype == RCNN_F32) return launch_gemm<float, 128>(ta, tb, D, ldd, bias, M, N, K, s);
    if (out_dtype == RCNN_F16) return launch_gemm<__half, 128>(ta, tb, D, ldd, bias, M, N, K, s);
    return launch_gemm<__nv_bfloat16, 128>(ta, tb, D, ldd, bias, M, N, K, s);
}

extern "C" int rcnn_attn_gates_cell_train(const void *xcat, int64_t ldx, const void *wcat_il, int64_t ldw, const float *bias_il,
                                          const float *embT_il, const int64_t *y, int B, int H, int K, int V, const float *c_in,
                                          float *c, void *h_out, int64_t h_ld, float *hid_out, int64_t hid_ld, float *gates_act,
                                          rcnn_stream_t stream);
extern "C" int rcnn_attn_gates_cell(const void *xcat, int64_t ldx, const void *wcat_il, int64_t ldw, const float *bias_il,
                                    const float *embT_il, const int64_t *y, int B, int H, int K, int V, float *c, void *h_out,
                                    int64_t h_ld, float *hid_out, int64_t hid_ld, rcnn_stream_t stream) {
    return rcnn_attn_gates_cell_train(xcat, ldx, wcat_il, ldw, bias_il, embT_il, y, B, H, K, V, c, c, h_out, h_ld, hid_out, hid_ld,
                                      nullptr, stream);
}

extern "C" int rcnn_attn_gates_cell_train(const void *xcat, int64_t ldx, const void *wcat_il, int64_t ldw, const float *bias_il,
                                          const float *embT_il, const int64_t *y, int B, int H, int K, int V, const float *c_in,
                                          float *c, void *h_out, int64_t h_ld, float *hid_out, int64_t hid_ld, float *gates_act,
                                          rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && H >= 8 && H % 8 == 0 && K > 0 && V >= 1, "attn_gates_cell: bad shape B=%d H=%d K=%d V=%d (H %% 8 == 0)",
                   B, H, K, V);
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(xcat && wcat_il && bias_il && embT_il && y && c && c_in && h_out, "attn_gates_cell: null pointer");
    RCNN_CHECK_ARG(((uintptr_t)c_in % 16) == 0 && ((uintptr_t)gates_act % 16) == 0, "attn_gates_cell: c_in / gates_act alignment");
    RCNN_CHECK_ARG(h_out != xcat, "attn_gates_cell: h_out must not alias the A operand (other CTAs still read it)");
    RCNN_CHECK_ARG(ldx >= K && ldw >= K && (ldx % 8) == 0 && (ldw % 8) == 0 && ((uintptr_t)xcat % 16) == 0 &&
                       ((uintptr_t)wcat_il % 16) == 0,
                   "attn_gates_cell: operand rows must be 16-byte aligned");
    RCNN_CHECK_ARG(h_ld >= H && (h_ld % 8) == 0 && ((uintptr_t)h_out % 16) == 0 && ((uintptr_t)c % 16) == 0 &&
                       ((uintptr_t)embT_il % 16) == 0 && (hid_out == nullptr || (hid_ld >= H && (hid_ld % 4) == 0 &&
                                                                                 ((uintptr_t)hid_out % 16) == 0)),
                   "attn_gates_cell: h_out / c / embT / hid_out must allow 16-byte accesses");
    CUtensorMap ta, tb;
    int rc = make_tmap_2d(&ta, xcat, 2, (uint64_t)B, (uint64_t)K, (uint64_t)ldx * 2, BM, BK, 1);
    if (rc) return rc;
    rc = make_tmap_2d(&tb, wcat_il, 2, (uint64_t)4 * H, (uint64_t)K, (uint64_t)ldw * 2, 32, BK, 1);
    if (rc) return rc;
    CellEpi ce;
    ce.embT = embT_il; ce.y = (const long long *)y; ce.c = c; ce.c_in = c_in; ce.gates_act = gates_act;
    ce.h_out = (__nv_bfloat16 *)h_out; ce.hid_out = hid_out;
    ce.h_ld = h_ld; ce.hid_ld = hid_ld; ce.V = V; ce.H = H;
    return launch_gemm_cell(ta, tb, bias_il, B, 4 * H, K, ce, (cudaStream_t)stream);
}

extern "C" int rcnn_gemm_bf16_atb_grouped(const void *A, int64_t lda, int a_gcols, const void *B, int64_t ldb, int b_gcols,
                                          float *D, int64_t ldd, int64_t d_goff, int groups, int M, int N, int K,
                                          int accumulate, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && groups >= 1, "gemm_atb: bad shape M=%d N=%d K=%d groups=%d", M, N, K, groups);
    if (M == 0 || N == 0) return RCNN_OK;
    RCNN_CHECK_ARG(D && ldd >= N, "gemm_atb: bad output");
    cudaStream_t s = (cudaStream_t)stream;
    if (!accumulate)
        for (int g = 0; g < groups; ++g)
            RCNN_CUDA(cudaMemset2DAsync(D + g * d_goff, (size_t)ldd * 4, 0, (size_t)N * 4, (size_t)M, s));
    if (K == 0) return RCNN_OK;
    RCNN_CHECK_ARG(A && B, "gemm_atb: null pointer");
    const long long a_cols = (long long)(groups - 1) * a_gcols + M, b_cols = (long long)(groups - 1) * b_gcols + N;
    RCNN_CHECK_ARG(lda >= a_cols && ldb >= b_cols, "gemm_atb: leading dimension smaller than the row");
    RCNN_CHECK_ARG((lda % 8) == 0 && (ldb % 8) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0,
                   "gemm_atb: A/B rows must be 16-byte aligned (lda=%lld ldb=%lld)", (long long)lda, (long long)ldb);
    CUtensorMap ta, tb;
    static const int use_pair = getenv("RCNN_GEMM_PAIR") ? atoi(getenv("RCNN_GEMM_PAIR")) : 1;
    if (use_pair && M >= 256 && N >= 256 && (N % 32) == 0 && ((uintptr_t)D & 15) == 0 && (ldd % 4) == 0 &&
        (groups == 1 || (d_goff % 4) == 0)) {
        // CTA-pair kernel: 256 x 256 tiles, TMA reduce-add of [32 rows x 64 columns] pieces (4-D map: column in
        // block, row, 32-column block, group); flat operands as (cols, K, 1)
        int rc = make_tmap_3d(&ta, A, 2, 1, (uint64_t)K, (uint64_t)a_cols, (uint64_t)lda * 2 * (uint64_t)K, (uint64_t)lda * 2, 1, BK, 64, 1);
        if (rc) return rc;
        rc = make_tmap_3d(&tb, B, 2, 1, (uint64_t)K, (uint64_t)b_cols, (uint64_t)ldb * 2 * (uint64_t)K, (uint64_t)ldb * 2, 1, BK, 64, 1);
        if (rc) return rc;
        CUtensorMap td;
        const uint64_t dims[4] = {32, (uint64_t)M, (uint64_t)(N / 32), (uint64_t)groups};
        const uint64_t strides[3] = {(uint64_t)ldd * 4, 128, (uint64_t)(groups > 1 ? d_goff : (int64_t)M * ldd) * 4};
        const uint32_t box[4] = {32, 32, 2, 1};
        rc = make_tmap_nd(&td, D, 4, 4, dims, strides, box, 1);
        if (rc) return rc;
        const int tkb = (K + BK - 1) / BK;
        const AtbSeq flat = {tkb, {0, 0}, 0};
        return launch_atb_pair(ta, tb, td, M, N, tkb, groups, a_gcols, b_gcols, flat, s);
    }
    int rc = make_tmap_2d(&ta, A, 2, (uint64_t)K, (uint64_t)a_cols, (uint64_t)lda * 2, BK, 64, 1);
    if (rc) return rc;
    rc = make_tmap_2d(&tb, B, 2, (uint64_t)K, (uint64_t)b_cols, (uint64_t)ldb * 2, BK, 64, 1);
    if (rc) return rc;
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * groups;
    const int total_kb = (K + BK - 1) / BK;
    int splits = (2 * gemm_sms() + tiles - 1) / tiles;          // aim at ~2 CTAs per SM
    splits = splits < 1 ? 1 : (splits > total_kb ? total_kb : splits);
    const int kb_per_split = (total_kb + splits - 1) / splits;
    splits = (total_kb + kb_per_split - 1) / kb_per_split;
    RCNN_CUDA(cudaFuncSetAttribute(gemm_atb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAtbSmemBytes));
    // TMA reduce-add epilogue when the output rows (and the group stride) are 16-byte aligned
    CUtensorMap td = ta;
    int tma_out = 0;
    if (((uintptr_t)D & 15) == 0 && (ldd % 4) == 0 && (groups == 1 || (d_goff % 4) == 0)) {
        rc = make_tmap_3d(&td, D, 4, (uint64_t)groups, (uint64_t)M, (uint64_t)N, (uint64_t)(groups > 1 ? d_goff : (int64_t)M * ldd) * 4,
                          (uint64_t)ldd * 4, 1, 32, 32, 1);
        if (rc) return rc;
        tma_out = 1;
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits * groups);
    ProfScope prof(RCNN_K_GEMM_ATB, s);
    gemm_atb_kernel<<<grid, kThreads, kAtbSmemBytes, s>>>(ta, tb, D, ldd, M, N, K, kb_per_split, splits, a_gcols, b_gcols,
                                                        (long long)d_goff, td, tma_out);
    RCNN_LAUNCH_CHECK("gemm_atb_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_gemm_bf16_atb(const void *A, int64_t lda, const void *B, int64_t ldb, float *D, int64_t ldd,
                                  int M, int N, int K, int accumulate, rcnn_stream_t stream) {
    return rcnn_gemm_bf16_atb_grouped(A, lda, 0, B, ldb, 0, D, ldd, 0, 1, M, N, K, accumulate, stream);
}

// Weight gradients of one bidirectional LSTM block, written in nn.LSTM's row order (see include/rcnn_ocr_b200.h)
extern "C" int rcnn_lstm_weight_grads(const void *dG, const void *x, const void *hcat, int B, int T, int I, int H,
                                      float *dwih, float *dwhh, int accumulate, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 1 && T >= 1 && (H == 256 || H == 512) && I >= 256 && (I % 32) == 0,
                   "lstm_weight_grads: needs H in {256, 512} and I >= 256, a multiple of 32 (B=%d T=%d I=%d H=%d)", B, T, I, H);
    RCNN_CHECK_ARG(dG && x && hcat && dwih && dwhh, "lstm_weight_grads: null pointer");
    RCNN_CHECK_ARG(((uintptr_t)dG % 16) == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)hcat % 16) == 0 &&
                   ((uintptr_t)dwih % 16) == 0 && ((uintptr_t)dwhh % 16) == 0, "lstm_weight_grads: 16-byte aligned buffers");
    cudaStream_t s = (cudaStream_t)stream;
    const int H4 = 4 * H;
    if (!accumulate) {
        RCNN_CUDA(cudaMemsetAsync(dwih, 0, sizeof(float) * 2 * (size_t)H4 * I, s));
        RCNN_CUDA(cudaMemsetAsync(dwhh, 0, sizeof(float) * 2 * (size_t)H4 * H, s));
    }
    auto out_map = [&](CUtensorMap *td, float *D, int N) {   // (col in block, gate, unit, 32-column block, direction)
        const uint64_t dims[5] = {32, 4, (uint64_t)H, (uint64_t)(N / 32), 2};
        const uint64_t strides[4] = {(uint64_t)H * N * 4, (uint64_t)N * 4, 128, (uint64_t)H4 * N * 4};
        const uint32_t box[5] = {32, 4, 8, 2, 1};
        return make_tmap_nd(td, D, 4, 5, dims, strides, box, 1);
    };
    CUtensorMap ta, tb, td;
    // dW_ih[d] = dG_d^T x: flat K = B*T rows, both directions read the same x (b_gcols = 0)
    const uint64_t K = (uint64_t)B * T;
    int rc = make_tmap_3d(&ta, dG, 2, 1, K, 8ull * H, K * 16ull * H, 16ull * H, 1, BK, 64, 1);
    if (rc) return rc;
    rc = make_tmap_3d(&tb, x, 2, 1, K, (uint64_t)I, K * 2ull * I, 2ull * I, 1, BK, 64, 1);
    if (rc) return rc;
    rc = out_map(&td, dwih, I);
    if (rc) return rc;
    const int tkb = (int)((K + BK - 1) / BK);
    const AtbSeq flat = {tkb, {0, 0}, 1};
    rc = launch_atb_pair(ta, tb, td, H4, I, tkb, 2, H4, 0, flat, s);
    if (rc) return rc;
    // dW_hh[d] = dG_d^T h_d(t -/+ 1): K runs over (b, t) with the h rows shifted inside each sequence
    rc = make_tmap_3d(&ta, dG, 2, (uint64_t)B, (uint64_t)T, 8ull * H, (uint64_t)T * 16ull * H, 16ull * H, 1, BK, 64, 1);
    if (rc) return rc;
    rc = make_tmap_3d(&tb, hcat, 2, (uint64_t)B, (uint64_t)T, 2ull * H, (uint64_t)T * 4ull * H, 4ull * H, 1, BK, 64, 1);
    if (rc) return rc;
    rc = out_map(&td, dwhh, H);
    if (rc) return rc;
    const int cps = (T + BK - 1) / BK;
    const AtbSeq seq = {cps, {-1, 1}, 1};
    return launch_atb_pair(ta, tb, td, H4, H, B * cps, 2, H4, H, seq, s);
}
