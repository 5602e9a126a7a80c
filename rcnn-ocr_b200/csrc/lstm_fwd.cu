// K2 (forward), UNFUSED variant: persistent recurrent LSTM kernel, both directions of a BidirectionalLSTM block,
// on a precomputed input projection.  The default path is lstm_fwdx.cu (input projection fused, two half-items
// per group, dedicated publisher warp); this kernel serves the input sizes that one does not take (I not a
// multiple of 64, or > 512).
//
// Replaces the T dependent timesteps inside nn.LSTM(bidirectional=True, batch_first=True) at
// model/model.py:154-156,161 (cuDNN RNN in the reference): gates = xp_t + h_{t-1} W_hh^T,
// c = f*c + i*g, h = o*tanh(c), h_0 = c_0 = 0, gate order i,f,g,o.  The input projection
// xp = x W_ih^T + b_ih + b_hh for all timesteps comes from the tcgen05 GEMM (gemm.cu).
//
// Decomposition.  A GROUP of H/32 CTAs (16 for H=512) owns one work item = (direction, tile of 64
// sequences) at a time; CTA c of the group owns hidden units [32c, 32c+32) with all four gates.
// Its 128 x H slice of W_hh (bf16, 128 KB for H=512) is loaded ONCE by TMA and stays resident in
// shared memory.  The product is computed TRANSPOSED, D^T[128 gate rows, 64 seq] = W_slice h^T, so
// the resident weights are the M = 128 operand (full tensor-pipe rate) and the sequence tile is the
// N operand: a 64-sequence tile keeps the per-step ingest of a CTA at 64 KB of h (all of it in
// flight at once: the whole tile has its own shared-memory slots) and lets a B = 256 batch spread
// over 8 groups = 128 SMs.  Per timestep:
//   warp 0  (TMA)      waits until all CTAs of the group have published h_{t-1} (a counter in global
//                      memory, red.release / ld.acquire at gpu scope: groups are not clusters, so any
//                      number of them fits the machine), then streams h_{t-1}[64 seq, H] (bf16,
//                      straight out of the block's output tensor); prefetches the next step's xp
//                      tile (fp16) and, for training, TMA-stores the previous step's activated gates;
//   warp 1  (MMA)      tcgen05.mma  D^T[128, 64] += W_chunk * h_chunk^T, fp32 in TMEM;
//   warps 2-9 (cell)   a thread owns one gate row (TMEM lane) and 32 sequences: tcgen05.ld, add xp,
//                      one MUFU.TANH per gate value (sigmoid through tanh with per-lane constants),
//                      4x4 register transposes across the 4 lanes of a unit (warp shuffles) so that each
//                      lane has i,f,g,o of 8 (unit, sequence) cells, cell update (c stays in registers
//                      for the whole sequence), 8x8 shuffle transpose of the bf16 h values so that each
//                      lane stores 16 contiguous bytes of h_t; then ONE gpu-scope release per CTA.
// Only h crosses CTAs, through L2; nothing is exchanged between groups.  The kernel is launched
// cooperatively (all CTAs co-resident), groups loop over work items when there are more items than
// groups.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "sm100.cuh"
#include "lstm_cell.cuh"

namespace rcnn {
namespace {

using namespace sm100;

constexpr int NS = 64;    // sequences per work item (UMMA N)
constexpr int GR = 128;   // gate rows per CTA: 32 units x 4 gates (UMMA M)
constexpr int LK = 64;    // K chunk (one 128-byte swizzle row of bf16)
constexpr uint32_t kWTile = GR * LK * 2;   // 16 KB: [128 gate rows x 64 k] bf16, SW128
constexpr uint32_t kHBox = NS * LK * 2;    //  8 KB: [64 seq x 64 k] bf16, SW128
constexpr uint32_t kXTile = NS * GR * 2;   // 16 KB: [64 seq x 128 gate cols] fp16, no swizzle
constexpr int kMaxHBars = 4;               // h_{t-1} arrives as (at most) four TMA operations of H/256 chunks each
constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 cell update

struct FwdParams {
    int B, T, H;
    int nitems, ngroups;
    int csize;            // CTAs per cluster (1, 2 or 4): the h tile is fetched once per cluster and multicast
    __nv_bfloat16 *hcat;  // [B, T, 2H]
    float *csave;         // [2, T, B, H]  (training only)
    unsigned int *sync;   // [ngroups] zeroed before the launch
    long long *tl;        // debug timeline or nullptr
};
#define TL_MARK(k) do { if (tl) tl[(s) * 8 + (k)] = clock64(); } while (0)

// Publishing h_t: written by a TMA store of this thread (complete: cp.async.bulk.wait_group 0 + fence.proxy.async);
// the counter increment is a gpu-scope RELEASE and the poll an ACQUIRE.  A relaxed increment is not enough:
// completion makes the writes visible to the issuing thread only (see lstm_bwd.cu, where readers were observed to
// fetch the previous contents of the exchange slot after seeing the counter).
__device__ __forceinline__ void red_release_gpu_inc(unsigned int *p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Bounded spin on the group counter: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void wait_counter(const unsigned int *p, unsigned int target) {
    if (ld_acquire_gpu(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu(p) < target) {
        if (clock64() - t0 > 4000000000LL) {
            printf("rcnn-ocr_b200: lstm_fwd group counter timed out (block %d)\n", blockIdx.x);
            __trap();
        }
    }
}

template <bool SAVE>
__global__ void __launch_bounds__(kThreads, 1)
lstm_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH,
                const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                const __grid_constant__ CUtensorMap tmHs, const FwdParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int H = p.H, T = p.T, B = p.B;
    const int nkc = H / LK;
    const int cpb = nkc >= 4 ? nkc / 4 : 1;        // K chunks per barrier / TMA operation
    const int nhb = nkc / cpb;                     // h barriers per step (4; 2 for H = 128, 1 for H = 64)
    const int cpc = nkc / p.csize;                 // K chunks this CTA fetches for its whole cluster
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int gsize = H / 32;
    unsigned char *w_s = smem;                          // nkc tiles [128 gate rows x 64 k] bf16, SW128
    unsigned char *h_s = w_s + (size_t)nkc * kWTile;    // nkc boxes [64 seq x 64 k] bf16, SW128
    unsigned char *x_s = h_s + (size_t)nkc * kHBox;     // 2 tiles [64 seq x 128 cols] fp16
    uint64_t *bars = reinterpret_cast<uint64_t *>(x_s + 2 * kXTile);
    uint64_t *w_full = bars;
    uint64_t *h_full = bars + 1;                        // kMaxHBars
    uint64_t *x_full = h_full + kMaxHBars;              // 2
    uint64_t *x_free = x_full + 2;                      // 2: slot may be overwritten by the next xp load
    uint64_t *g_ready = x_free + 2;                     // 2: (training) activated gates are in the slot
    uint64_t *tmem_full = g_ready + 2;
    uint64_t *h_staged = tmem_full + 1;                 // the 8 cell warps put h_t into shared memory
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(h_staged + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x / gsize;
    const int c = blockIdx.x % gsize;            // unit slice
    long long *tl = (p.tl && blockIdx.x == 0) ? p.tl : nullptr;
    unsigned int *counter = p.sync + group;

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(w_full, 1);
            for (int i = 0; i < kMaxHBars; ++i) mbar_init(&h_full[i], 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&x_full[i], 1);
                mbar_init(&x_free[i], SAVE ? 1 : 8);
                mbar_init(&g_ready[i], 8);
            }
            mbar_init(tmem_full, 1);
            mbar_init(h_staged, 8);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);   // [0, 256): the W_hh slice as the A operand (H/2 columns); [256, 320): accumulator
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (one elected thread: elect.sync keeps operands in uniform registers) =
        if (elect_one()) {
            tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmX);
            if (SAVE) tma_prefetch_desc(&tmG);
            int cur_dir = -1;
            unsigned int xidx = 0;       // xp tiles requested so far (slot = xidx & 1)
            unsigned int gidx = 0;       // gate tiles stored so far
            unsigned int published = 0;  // steps this group has completed before the current one
            for (int item = group; item < p.nitems; item += p.ngroups) {
                const int dir = item & 1, b0 = (item >> 1) * NS;
                if (dir != cur_dir) {
                    // every MMA that read the old slice has completed: its step was published
                    mbar_arrive_expect_tx(w_full, (uint32_t)nkc * kWTile);
                    for (int kc = 0; kc < nkc; ++kc)
                        tma_load_2d(w_s + (size_t)kc * kWTile, &tmW, w_full, kc * LK, dir * 4 * H + c * GR);
                    cur_dir = dir;
                }
                auto load_x = [&](int xs) {
                    const int slot = xidx & 1;
                    mbar_wait(&x_free[slot], ((xidx >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&x_full[slot], kXTile);
                    tma_load_3d(x_s + slot * kXTile, &tmX, &x_full[slot], dir * 4 * H + c * GR, dir ? T - 1 - xs : xs, b0);
                    ++xidx;
                };
                auto store_gates = [&](int gs) {   // activated gates of step gs: shared memory -> gates_save
                    const int slot = gidx & 1;
                    mbar_wait(&g_ready[slot], (gidx >> 1) & 1);
                    tma_store_3d(&tmG, x_s + slot * kXTile, c * GR, b0, dir * T + (dir ? T - 1 - gs : gs));
                    tma_store_commit();
                    tma_store_wait_read<0>();
                    mbar_arrive(&x_free[slot]);
                    ++gidx;
                };
                load_x(0);
                for (int s = 0; s < T; ++s) {
                    if (s > 0) {
                        const int t = dir ? T - 1 - s : s;
                        const int tprev = dir ? t + 1 : t - 1;
                        TL_MARK(7);
                        wait_counter(counter, (published + (unsigned)s) * (unsigned)gsize);
                        TL_MARK(0);
                        fence_proxy_async_global();  // h_{t-1} was written through the generic proxy
                        // Every CTA of the cluster expects the whole tile and fetches 1/csize of it (cpc boxes
                        // [64 seq x 64 k], chunk-major) for all of them: L2 is read once per cluster.  The peers'
                        // slots are free: they published step s-1 after their MMAs had consumed the old tile.
                        for (int g = 0; g < nhb; ++g) mbar_arrive_expect_tx(&h_full[g], (uint32_t)cpb * kHBox);
                        const int bxc = cpc < cpb ? cpc : cpb;      // chunks per TMA operation (= the map's box)
                        for (int kc0 = (int)crank * cpc; kc0 < ((int)crank + 1) * cpc; kc0 += bxc) {
                            if (p.csize > 1)
                                tma_load_4d_mcast(h_s + (size_t)kc0 * kHBox, &tmH, &h_full[kc0 / cpb], 0, b0, dir * nkc + kc0,
                                                  tprev, (uint16_t)((1u << p.csize) - 1u));
                            else
                                tma_load_4d(h_s + (size_t)kc0 * kHBox, &tmH, &h_full[kc0 / cpb], 0, b0, dir * nkc + kc0, tprev);
                        }
                        TL_MARK(1);
                        if (SAVE) store_gates(s - 1);
                    }
                    if (s + 1 < T) load_x(s + 1);
                }
                if (SAVE) store_gates(T - 1);
                published += (unsigned)T;
            }
            if (SAVE) tma_store_wait<0>();
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) ===================================================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(GR, NS);
            int cur_dir = -1;
            uint32_t wphase = 0, hphase = 0, sphase = 0;
            tma_prefetch_desc(&tmHs);
            // h_t of this CTA's 32 units leaves by ONE TMA store (staged by the cell warps in the idle h tile);
            // its completion is awaited by this thread alone and followed by ONE release increment of the counter
            // (instead of a gpu-scope fence after 256 threads' stores).
            auto publish = [&](int dir, int b0, int s) {
                const int t = dir ? T - 1 - s : s;
                mbar_wait(h_staged, sphase);
                sphase ^= 1;
                tma_store_3d(&tmHs, h_s, dir * H + 32 * c, t, b0);
                tma_store_commit();
                tma_store_wait<0>();
                fence_proxy_async_global();
                red_release_gpu_inc(counter);
                TL_MARK(6);
            };
            for (int item = group; item < p.nitems; item += p.ngroups) {
                const int dir = item & 1, b0 = (item >> 1) * NS;
                if (dir != cur_dir) {
                    // weights: shared memory -> tensor memory, one K = 16 slice per copy (in issue order with the MMAs)
                    mbar_wait(w_full, wphase);
                    wphase ^= 1;
                    cur_dir = dir;
                    tc_fence_after();
                    for (int kc = 0; kc < nkc; ++kc) {
                        const uint64_t wdesc = make_smem_desc_sw128(smem_u32(w_s + (size_t)kc * kWTile), 16, 1024);
#pragma unroll
                        for (int k = 0; k < LK / 16; ++k)
                            tmem_cp_128x256b(tmem_base + (uint32_t)(8 * (kc * (LK / 16) + k)), wdesc + (uint64_t)(2 * k));
                    }
                }
                if (T > 0) publish(dir, b0, 0);
                for (int s = 1; s < T; ++s) {
                    for (int g = 0; g < nhb; ++g) {
                        mbar_wait(&h_full[g], hphase);
                        if (g == 0) TL_MARK(2);
                        tc_fence_after();
                        for (int j = 0; j < cpb; ++j) {
                            const int kc = g * cpb + j;
                            const uint64_t bdesc = make_smem_desc_sw128(smem_u32(h_s + (size_t)kc * kHBox), 16, 1024);
#pragma unroll
                            for (int k = 0; k < LK / 16; ++k)
                                umma_bf16_ts(tmem_base + 256u, tmem_base + (uint32_t)(8 * (kc * (LK / 16) + k)),
                                             bdesc + (uint64_t)(2 * k), idesc, (g | j | k) != 0);
                        }
                    }
                    umma_commit(tmem_full);
                    TL_MARK(3);
                    hphase ^= 1;
                    publish(dir, b0, s);
                }
            }
        }
    } else {
        // ===== cell update ========================================================================
        const int qd = warp & 3;            // TMEM lane quadrant of this warp
        const int ch = (warp - 2) >> 2;     // which 32 of the tile's 64 sequences
        const int r = qd * 32 + lane;       // gate row within the CTA == accumulator lane: unit r>>2, gate r&3
        const int g = lane & 3, ul = lane >> 2;
        const CellLane CL(lane);
        unsigned int xidx = 0, mcount = 0;
        for (int item = group; item < p.nitems; item += p.ngroups) {
            const int dir = item & 1, b0 = (item >> 1) * NS;
            float cst[8];   // cell state of unit 8*qd+ul for sequences 32ch + 4k + g
#pragma unroll
            for (int k = 0; k < 8; ++k) cst[k] = 0.f;
            for (int s = 0; s < T; ++s) {
                const int t = dir ? T - 1 - s : s;
                const int slot = xidx & 1;
                unsigned char *xs = x_s + slot * kXTile + (size_t)(32 * ch) * (GR * 2) + 2 * r;
                float pre[32];
                mbar_wait(&x_full[slot], (xidx >> 1) & 1);
#pragma unroll
                for (int i = 0; i < 32; ++i) pre[i] = __half2float(*reinterpret_cast<const __half *>(xs + i * (GR * 2)));
                if (!SAVE) {
                    // Free the slot only once every lane's loads have RETURNED (an mbarrier arrive is not
                    // ordered behind shared-memory loads still in flight): make the arrive data-dependent.
                    uint32_t dep = 0;
#pragma unroll
                    for (int i = 0; i < 32; i += 8) dep |= __float_as_uint(pre[i]);
                    dep = __any_sync(FULL, dep != 0u) ? 1u : 0u;
                    if (lane == 0) mbar_arrive_after(&x_free[slot], dep);
                }
                if (s > 0) {
                    uint32_t acc[32];
                    mbar_wait(tmem_full, mcount & 1);
                    ++mcount;
                    if (threadIdx.x == 64) TL_MARK(4);
                    tc_fence_after();
                    tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + 256u + (uint32_t)(ch * 32), acc);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) pre[i] += __uint_as_float(acc[i]);
                }
                cell_activate(pre, CL);
                uint32_t gsv[16];   // (training) activated gates as fp16 pairs, stashed after the publish
                if (SAVE) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const __half2 h2 = __floats2half2_rn(pre[2 * i], pre[2 * i + 1]);
                        gsv[i] = *reinterpret_cast<const uint32_t *>(&h2);
                    }
                }
                const uint4 hq = cell_update(pre, cst, CL);
                // stage h_t [64 seq x 32 units] in the (now idle) h tile: row = sequence, 64 bytes per row; the MMA
                // warp's thread stores it with TMA (rows >= B are clipped by the tensor map) and publishes the step
                *reinterpret_cast<uint4 *>(h_s + (size_t)(32 * ch + lane) * 64 + qd * 16) = hq;
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(h_staged);
                if (threadIdx.x == 64) TL_MARK(5);
                if (SAVE) {
                    // activated gates back into the xp slot (same addresses this thread read), then out by TMA
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        *reinterpret_cast<uint16_t *>(xs + (2 * i) * (GR * 2)) = (uint16_t)(gsv[i] & 0xffffu);
                        *reinterpret_cast<uint16_t *>(xs + (2 * i + 1) * (GR * 2)) = (uint16_t)(gsv[i] >> 16);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&g_ready[slot]);
                    // c_t is not needed until the backward pass: stored after the publish
                    float *crow = p.csave + (((size_t)dir * T + t) * B) * H + 32 * c + 8 * qd + ul;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int bc = b0 + 32 * ch + 4 * k + g;
                        if (bc < B) crow[(size_t)bc * H] = cst[k];
                    }
                }
                ++xidx;
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

int fwd_cluster_size(int nkc) {
    static const int want = getenv("RCNN_FWD_CLUSTER") ? atoi(getenv("RCNN_FWD_CLUSTER")) : 1;
    return want >= 4 && nkc >= 4 ? 4 : (want >= 2 && nkc >= 2 ? 2 : 1);
}

size_t fwd_smem_bytes(int H) { return 1024 + (size_t)(H / LK) * (kWTile + kHBox) + 2 * kXTile + 256; }

template <bool SAVE>
int launch_fwd(const CUtensorMap &tw, const CUtensorMap &th, const CUtensorMap &tx, const CUtensorMap &tg,
               const CUtensorMap &ths, FwdParams &p, cudaStream_t s) {
    const int gsize = p.H / 32;
    const int nkc = p.H / LK;
    // cluster = multicast domain for the h tile; it divides the group.  Measured on B200 (B=256, H=512): with
    // multicast the step is ~7% SLOWER (the tile's first half then waits for two CTAs' polls), L2 read
    // bandwidth is not the limiter, so the default is no cluster; RCNN_FWD_CLUSTER=4 turns it on.
    p.csize = fwd_cluster_size(nkc);
    p.nitems = 2 * ((p.B + NS - 1) / NS);
    const size_t smem = fwd_smem_bytes(p.H);
    RCNN_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel<SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the groups spin on each other
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)p.csize;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    static thread_local int max_clusters[32][2][4] = {};   // [device][SAVE][log2 H/64]
    int dev_id = 0;
    RCNN_CUDA(cudaGetDevice(&dev_id));
    int &mc = max_clusters[dev_id & 31][SAVE ? 1 : 0][p.H == 64 ? 0 : p.H == 128 ? 1 : p.H == 256 ? 2 : 3];
    if (mc == 0) {
        cfg.gridDim = dim3((unsigned)(p.csize * 1024));
        RCNN_CUDA(cudaOccupancyMaxActiveClusters(&mc, lstm_fwd_kernel<SAVE>, &cfg));
        if (mc * p.csize < gsize) {
            set_error("lstm_forward: only %d clusters of %d CTAs fit the device, a group needs %d CTAs", mc, p.csize, gsize);
            mc = 0;
            return RCNN_ERR_DEVICE;
        }
    }
    const int max_groups = mc * p.csize / gsize;
    p.ngroups = p.nitems < max_groups ? p.nitems : max_groups;
    if (p.ngroups > 1 && (p.ngroups & 1) && p.nitems > p.ngroups) --p.ngroups;   // even: a group keeps its direction (and W slice)
    p.sync = group_counters(p.ngroups, s);
    if (!p.sync) return RCNN_ERR_CUDA_BASE;
    cfg.gridDim = dim3((unsigned)(gsize * p.ngroups));
    ProfScope prof(RCNN_K_LSTM_FWD, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, lstm_fwd_kernel<SAVE>, tw, th, tx, tg, ths, p));
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
    CUtensorMap tw, th, tx, tg, ths;
    int rc = make_tmap_2d(&tw, whh_packed, 2, 8ull * H, (uint64_t)H, (uint64_t)H * 2, GR, LK, 1);
    if (rc) return rc;
    {   // hcat [B, T, 2H] seen as (k within chunk, b, chunk, t): one box = cpc chunks of [64 seq x 64 k],
        // the share one CTA of a cluster fetches (see launch_fwd: csize = min(4, nkc))
        const int nkc = H / LK, cpb = nkc >= 4 ? nkc / 4 : 1;
        const int cpc = nkc / fwd_cluster_size(nkc) < cpb ? nkc / fwd_cluster_size(nkc) : cpb;
        const uint64_t dims[4] = {(uint64_t)LK, (uint64_t)B, 2ull * nkc, (uint64_t)T};
        const uint64_t strides[3] = {(uint64_t)T * 2 * H * 2, (uint64_t)LK * 2, 2ull * H * 2};
        const uint32_t box[4] = {(uint32_t)LK, (uint32_t)NS, (uint32_t)cpc, 1u};
        rc = make_tmap_4d(&th, hcat, 2, dims, strides, box, 1);
        if (rc) return rc;
    }
    // store view of hcat: box = [64 seq x 1 t x 32 units] (this CTA's slice of h_t), plain rows of 64 bytes
    rc = make_tmap_3d(&ths, hcat, 2, (uint64_t)B, (uint64_t)T, 2ull * H, (uint64_t)T * 2 * H * 2, 2ull * H * 2, NS, 1, 32, 0);
    if (rc) return rc;
    rc = make_tmap_3d(&tx, xp, 2, (uint64_t)B, (uint64_t)T, 8ull * H, (uint64_t)T * 8 * H * 2, 8ull * H * 2, NS, 1, GR, 0);
    if (rc) return rc;
    tg = tx;
    if (gates_save) {
        rc = make_tmap_3d(&tg, gates_save, 2, 2ull * T, (uint64_t)B, 4ull * H, (uint64_t)B * 4 * H * 2, 4ull * H * 2, 1, NS, GR, 0);
        if (rc) return rc;
    }
    FwdParams p;
    p.B = B; p.T = T; p.H = H;
    p.hcat = (__nv_bfloat16 *)hcat;
    p.csave = c_save;
    p.tl = debug_timeline();
    cudaStream_t s = (cudaStream_t)stream;
    return gates_save ? launch_fwd<true>(tw, th, tx, tg, ths, p, s) : launch_fwd<false>(tw, th, tx, tg, ths, p, s);
}
