// K2 (forward): persistent recurrent LSTM kernel, both directions of a BidirectionalLSTM block.
//
// Replaces the T dependent timesteps inside nn.LSTM(bidirectional=True, batch_first=True) at
// model/model.py:154-156,161 (cuDNN RNN in the reference): gates = xp_t + h_{t-1} W_hh^T,
// c = f*c + i*g, h = o*tanh(c), h_0 = c_0 = 0, gate order i,f,g,o.  The input projection
// xp = x W_ih^T + b_ih + b_hh for all timesteps comes from the tcgen05 GEMM (gemm.cu).
//
// Decomposition.  One thread-block CLUSTER owns (direction, tile of 128 sequences).  CTA c of
// the cluster owns hidden units [32c, 32c+32) with all four gates, so cluster size = H/32
// (16 CTAs for H=512 -- a non-portable cluster -- 8 for H=256).  Its 128 x H slice of W_hh
// (bf16, 128 KB for H=512) is loaded ONCE by TMA and stays resident in shared memory for all T
// steps.  Per timestep:
//   warp 0  (TMA)      streams h_{t-1}[128 seq, H] (bf16, straight out of the block's output
//                      tensor, two 64-column boxes per ring slot) and prefetches the next step's
//                      xp tile (fp16, four 32-column chunks, all resident);
//   warp 1  (MMA)      tcgen05.mma  D[128 seq, 128 gate cols] += h_chunk * W_chunk^T, fp32 in TMEM;
//   warps 2-9 (cell)   two warps per TMEM lane quadrant; a thread owns one sequence and 16 of the
//                      CTA's 32 units: tcgen05.ld its accumulator columns, add xp, sigmoid/tanh,
//                      update c (fp32, in registers for the whole sequence), write h_t (bf16) and,
//                      for training, the activated gates (fp16) and c_t (fp32);
//   all               barrier.cluster (release/acquire): h_t of every unit slice is visible to
//                      the other CTAs' TMA loads of the next step.
// Only h crosses CTAs, through L2; nothing is exchanged between clusters.
#include <cuda_fp16.h>
#include "common.cuh"
#include "sm100.cuh"

namespace rcnn {
namespace {

using namespace sm100;

constexpr int LB = 128;   // sequences per cluster tile (UMMA M)
constexpr int LN = 128;   // gate columns per CTA: 32 units x 4 gates (UMMA N)
constexpr int LK = 64;    // K chunk (one 128-byte swizzle row of bf16)
constexpr uint32_t kTile = 16384;   // one operand TMA box: 128 x 64 bf16 (h chunk or W chunk) = 16 KB
constexpr int kBoxes = 2;           // h boxes per ring slot: one barrier round trip per K = 128
constexpr int kARing = 2;           // ring slots of kBoxes * 16 KB
constexpr uint32_t kASlot = kBoxes * kTile;
constexpr int kXRing = 4;           // one slot per 32-column xp chunk: the whole step's tile is resident
constexpr uint32_t kXTile = LB * 32 * 2;   // 128 seq x 32 cols fp16 = 8 KB, 64-byte swizzle
constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 cell update

struct FwdParams {
    int B, T, H;
    __nv_bfloat16 *hcat;  // [B, T, 2H]
    __half *gates;        // [2, T, B, 4H] packed column order, activated (training only)
    float *csave;         // [2, T, B, H]  (training only)
    long long *tl;        // debug timeline or nullptr
};
#define TL_MARK(k) do { if (tl) tl[s * 8 + (k)] = clock64(); } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory"); }
// Publish the cell warps' global stores with ONE gpu-scope release per CTA (the pattern of a
// cooperative-groups grid sync): CTA barrier among the 256 cell threads, then one thread arrives with
// release semantics (its MEMBAR is cumulative over the stores it observed through the barrier) while the
// others arrive relaxed.
__device__ __forceinline__ void cell_publish_arrive() {
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 64) cluster_arrive(); else cluster_arrive_relaxed();
}
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ float tanh_fast(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

template <bool SAVE>
__global__ void __launch_bounds__(kThreads, 1)
lstm_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
                const __grid_constant__ CUtensorMap tmX, const FwdParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int H = p.H, T = p.T, B = p.B;
    const int nkc = H / LK;
    unsigned char *w_s = smem;                          // nkc tiles [128 n x 64 k] bf16, SW128
    unsigned char *a_s = w_s + (size_t)nkc * kTile;     // kARing slots of kBoxes tiles [128 b x 64 k] bf16, SW128
    unsigned char *x_s = a_s + kARing * kASlot;         // kXRing tiles [128 b x 32 col] fp16, SW64
    uint64_t *bars = reinterpret_cast<uint64_t *>(x_s + kXRing * kXTile);
    uint64_t *w_full = bars;
    uint64_t *a_full = bars + 1, *a_empty = a_full + kARing;
    uint64_t *x_full = a_empty + kARing, *x_empty = x_full + kXRing;
    uint64_t *tmem_full = x_empty + kXRing;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = (int)cluster_ctarank();        // unit slice
    const int cid = (int)cluster_id_x();
    const int dir = cid & 1, tile = cid >> 1;
    const int b0 = tile * LB;
    long long *tl = (p.tl && blockIdx.x == 0) ? p.tl : nullptr;

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(w_full, 1);
            for (int i = 0; i < kARing; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
            for (int i = 0; i < kXRing; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 4); }
            mbar_init(tmem_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<LN>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer ======================================================================
        if (lane == 0) {
            tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmX);
            mbar_arrive_expect_tx(w_full, (uint32_t)nkc * kTile);
            for (int kc = 0; kc < nkc; ++kc)
                tma_load_2d(w_s + (size_t)kc * kTile, &tmW, w_full, kc * LK, dir * 4 * H + c * LN);
        }
        int pslot = 0;
        uint32_t pphase = 0;
        // xp chunk q of step s lives in slot q; phase parity = s & 1.  Step 0 is loaded up front, step
        // s+1 while step s runs (each slot is refilled as soon as the cell warps released it).
        auto load_x = [&](int xs) {
            const int xt = dir ? T - 1 - xs : xs;
            for (int q = 0; q < kXRing; ++q) {
                mbar_wait(&x_empty[q], (xs & 1) ^ 1);
                mbar_arrive_expect_tx(&x_full[q], kXTile);
                tma_load_3d(x_s + q * kXTile, &tmX, &x_full[q], dir * 4 * H + c * LN + q * 32, xt, b0);
            }
        };
        if (lane == 0 && T > 0) load_x(0);
        for (int s = 0; s < T; ++s) {
            if (lane == 0) {
                const int t = dir ? T - 1 - s : s;
                if (s > 0) {
                    const int tprev = dir ? t + 1 : t - 1;
                    TL_MARK(0);
                    fence_proxy_async_global();  // h_{t-1} was written through the generic proxy
                    for (int g = 0; g * kBoxes < nkc; ++g) {
                        const int nb = min(kBoxes, nkc - g * kBoxes);     // H = 64 has a single box
                        mbar_wait(&a_empty[pslot], pphase ^ 1);
                        mbar_arrive_expect_tx(&a_full[pslot], (uint32_t)nb * kTile);
                        for (int j = 0; j < nb; ++j)
                            tma_load_3d(a_s + pslot * kASlot + j * kTile, &tmH, &a_full[pslot],
                                        dir * H + (g * kBoxes + j) * LK, tprev, b0);
                        if (++pslot == kARing) { pslot = 0; pphase ^= 1; }
                    }
                    TL_MARK(1);
                }
                if (s + 1 < T) load_x(s + 1);
            }
            __syncwarp();
            cluster_arrive_relaxed();
            cluster_wait();
        }
    } else if (warp == 1) {
        // ===== MMA issuer ========================================================================
        constexpr uint32_t idesc = make_idesc_bf16(LB, LN);
        int mslot = 0;
        uint32_t mphase = 0;
        if (lane == 0) mbar_wait(w_full, 0);
        __syncwarp();
        for (int s = 0; s < T; ++s) {
            if (lane == 0 && s > 0) {
                for (int g = 0; g * kBoxes < nkc; ++g) {
                    const int nb = min(kBoxes, nkc - g * kBoxes);
                    mbar_wait(&a_full[mslot], mphase);
                    if (g == 0) TL_MARK(2);
                    tc_fence_after();
                    for (int j = 0; j < nb; ++j) {
                        const int kc = g * kBoxes + j;
                        const uint64_t adesc = make_smem_desc_sw128(smem_u32(a_s + mslot * kASlot + j * kTile), 16, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(w_s + (size_t)kc * kTile), 16, 1024);
#pragma unroll
                        for (int k = 0; k < LK / 16; ++k)
                            umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (g | j | k) != 0);
                    }
                    umma_commit(&a_empty[mslot]);
                    if (++mslot == kARing) { mslot = 0; mphase ^= 1; }
                }
                umma_commit(tmem_full);
                TL_MARK(3);
            }
            __syncwarp();
            cluster_arrive_relaxed();
            cluster_wait();
        }
    } else {
        // ===== cell update (one thread per sequence of the tile) ================================
        const int qd = warp & 3;            // TMEM lane quadrant of this warp
        const int hf = (warp - 2) >> 2;     // which half of the CTA's gate columns (units hf*16 .. +16)
        const int row = qd * 32 + lane;     // sequence within the tile == accumulator row
        const int b = b0 + row;
        const bool valid = b < B;
        float cst[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cst[i] = 0.f;
        for (int s = 0; s < T; ++s) {
            const int t = dir ? T - 1 - s : s;
            if (s > 0) {
                mbar_wait(tmem_full, (s - 1) & 1);
                if (threadIdx.x == 64) TL_MARK(4);
                tc_fence_after();
            }
            __nv_bfloat16 *hrow = p.hcat + ((size_t)b * T + t) * 2 * H + (size_t)dir * H + 32 * c;
            __half *grow = SAVE ? p.gates + (((size_t)dir * T + t) * B + b) * 4 * H + (size_t)c * LN : nullptr;
            float *crow = SAVE ? p.csave + (((size_t)dir * T + t) * B + b) * H + 32 * c : nullptr;
            U8 gsave[2][2], csv[2];   // training: activated gates / c of both chunks, stored AFTER the barrier arrive
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
                const int q = hf * 2 + qq;          // 32-column chunk of the accumulator / xp tile
                uint32_t acc[32];
                if (s > 0) {
                    tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(q * 32), acc);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0u;
                }
                mbar_wait(&x_full[q], s & 1);
                // fp16 tile, 64-byte rows, TMA SWIZZLE_64B: 16-byte piece j of row r sits at j ^ ((r >> 1) & 3)
                const unsigned char *xrow = x_s + q * kXTile + row * 64;
                float pre[32];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(xrow + ((j ^ ((row >> 1) & 3)) << 4));
                    const __half2 *h2 = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = __half22float2(h2[k]);
                        pre[8 * j + 2 * k] = f.x;
                        pre[8 * j + 2 * k + 1] = f.y;
                    }
                }
                // Free the slot only once every lane's loads have RETURNED: fold one word of each
                // 16-byte load into a value the arrive depends on, and vote it across the warp.
                uint32_t dep = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) dep |= __float_as_uint(pre[8 * j]);
                dep = __any_sync(FULL, dep != 0u) ? 1u : 0u;
                if (lane == 0) mbar_arrive_after(&x_empty[q], dep);
                if (s > 0) tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) pre[i] += __uint_as_float(acc[i]);
                float hv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float ig = sigmoid_fast(pre[4 * j]);
                    const float fg = sigmoid_fast(pre[4 * j + 1]);
                    const float gg = tanh_fast(pre[4 * j + 2]);
                    const float og = sigmoid_fast(pre[4 * j + 3]);
                    const float cn = fmaf(fg, cst[qq * 8 + j], ig * gg);
                    cst[qq * 8 + j] = cn;
                    hv[j] = og * tanh_fast(cn);
                    pre[4 * j] = ig; pre[4 * j + 1] = fg; pre[4 * j + 2] = gg; pre[4 * j + 3] = og;
                }
                if (valid) {
                    uint4 hq;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(hv[0], hv[1]), h1 = __floats2bfloat162_rn(hv[2], hv[3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(hv[4], hv[5]), h3 = __floats2bfloat162_rn(hv[6], hv[7]);
                    hq.x = *reinterpret_cast<uint32_t *>(&h0); hq.y = *reinterpret_cast<uint32_t *>(&h1);
                    hq.z = *reinterpret_cast<uint32_t *>(&h2); hq.w = *reinterpret_cast<uint32_t *>(&h3);
                    *reinterpret_cast<uint4 *>(hrow + q * 8) = hq;
                }
                if (SAVE) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            __half2 g2 = __floats2half2_rn(pre[16 * j + 2 * k], pre[16 * j + 2 * k + 1]);
                            gsave[qq][j].v[k] = *reinterpret_cast<uint32_t *>(&g2);
                        }
#pragma unroll
                    for (int k = 0; k < 8; ++k) csv[qq].v[k] = __float_as_uint(cst[qq * 8 + k]);
                }
            }
            if (threadIdx.x == 64) TL_MARK(5);
            tc_fence_before();
            fence_proxy_async_global();  // order the h_t stores before the other CTAs' TMA reads
            cell_publish_arrive();       // one release per CTA; its MEMBAR only has the h stores to wait for
            if (SAVE && valid) {
                // The saved tensors are not needed until the backward pass: issue their stores after
                // the arrive so they drain while this CTA waits and during the next step's MMA phase.
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    const int q = hf * 2 + qq;
                    st_v8(grow + q * 32, gsave[qq][0]);
                    st_v8(grow + q * 32 + 16, gsave[qq][1]);
                    st_v8(crow + q * 8, csv[qq]);
                }
            }
            cluster_wait();
            if (threadIdx.x == 64) TL_MARK(6);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<LN>(tmem_base);
    }
}

size_t fwd_smem_bytes(int H) { return 1024 + (size_t)(H / LK) * kTile + kARing * kASlot + kXRing * kXTile + 256; }

template <bool SAVE>
int launch_fwd(const CUtensorMap &tw, const CUtensorMap &th, const CUtensorMap &tx, const FwdParams &p,
               cudaStream_t s) {
    const int csize = p.H / 32;
    const int ntiles = (p.B + LB - 1) / LB;
    const size_t smem = fwd_smem_bytes(p.H);
    RCNN_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8)
        RCNN_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<SAVE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(csize * ntiles * 2));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ProfScope prof(RCNN_K_LSTM_FWD, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, lstm_fwd_kernel<SAVE>, tw, th, tx, p));
    count_launch();
    return RCNN_OK;
}

}  // namespace
}  // namespace rcnn

extern "C" int rcnn_lstm_forward(const void *xp, const void *whh_packed, int B, int T, int H, void *hcat,
                                 void *gates_save, float *c_save, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 0, "lstm_forward: bad shape B=%d T=%d", B, T);
    RCNN_CHECK_ARG(H == 64 || H == 128 || H == 256 || H == 512,
                   "lstm_forward: hidden size %d unsupported (64, 128, 256 or 512)", H);
    if (B == 0 || T == 0) return RCNN_OK;
    RCNN_CHECK_ARG(xp && whh_packed && hcat, "lstm_forward: null pointer");
    RCNN_CHECK_ARG((gates_save == nullptr) == (c_save == nullptr), "lstm_forward: gates_save and c_save go together");
    CUtensorMap tw, th, tx;
    int rc = make_tmap_2d(&tw, whh_packed, 2, 8ull * H, (uint64_t)H, (uint64_t)H * 2, LN, LK, 1);
    if (rc) return rc;
    rc = make_tmap_3d(&th, hcat, 2, (uint64_t)B, (uint64_t)T, 2ull * H, (uint64_t)T * 2 * H * 2, 2ull * H * 2, LB, 1, LK, 1);
    if (rc) return rc;
    rc = make_tmap_3d(&tx, xp, 2, (uint64_t)B, (uint64_t)T, 8ull * H, (uint64_t)T * 8 * H * 2, 8ull * H * 2, LB, 1, 32, 2);
    if (rc) return rc;
    FwdParams p;
    p.B = B; p.T = T; p.H = H;
    p.hcat = (__nv_bfloat16 *)hcat;
    p.gates = (__half *)gates_save;
    p.csave = c_save;
    p.tl = debug_timeline();
    cudaStream_t s = (cudaStream_t)stream;
    return gates_save ? launch_fwd<true>(tw, th, tx, p, s) : launch_fwd<false>(tw, th, tx, p, s);
}
