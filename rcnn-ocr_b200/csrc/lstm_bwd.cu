// K2 (backward): persistent BPTT kernel for both directions of a BidirectionalLSTM block.
//
// Backward of the recurrence in lstm_fwd.cu (the reference gets it from autograd through
// nn.LSTM / cuDNN RNN backward, model/model.py:161):
//     dh_t   = dhcat[:, t] + dG_{t'} W_hh            (t' = the step processed just before)
//     do     = dh_t * tanh(c_t)            dc += dh_t * o * (1 - tanh(c_t)^2)
//     di = dc*g   dg = dc*i   df = dc*c_prev   dc <- dc*f
//     dG_t   = [di*i(1-i), df*f(1-f), dg*(1-g^2), do*o(1-o)]   (pre-activation gradients)
// dW_ih, dW_hh and dX are GEMMs over dG after the loop (host side); db is accumulated here.
//
// Same grouping as the forward kernel: a GROUP of H/32 CTAs owns one work item = (direction, tile
// of 64 sequences) at a time, CTA c owns hidden units [32c, 32c+32); CTAs of a group synchronise per
// step through a counter in global memory and the kernel is launched cooperatively.
//
// The contraction dG_{t'} W_hh is K-LOCAL: a CTA multiplies only the 128 gate columns IT produced
// (its own dG slice, written by its cell warps straight into shared memory as the UMMA B operand:
// no ingest from L2) with the matching 128 rows of W_hh, for ALL H outputs:
//     partial_c^T [H units, 64 seq] = W_hh[rows of c]^T (H x 128, resident in shared memory, bf16)
//                                     x dG_c^T (128 x 64)
// i.e. H/128 x 8 tcgen05.mma of M = 128, N = 64 into H/128 TMEM accumulators (an all-gather of dG
// would need 4H/16 = 128 MMAs of N = 32 per step and 256 KB of ingest per CTA).  The partial sums are
// exchanged through L2 as bf16 (64 KB out, 64 KB in per CTA and step), both ways by ONE TMA operation
// on a 5-D tensor map of the exchange buffer [parity x group][src CTA][dst CTA][4][1 KB]: the drained
// accumulators are staged in shared memory and stored, the 16 partials of the CTA's own 32 units are
// loaded back into the same shared-memory area the next step and summed by the cell warps.
// With NH = 2 (default) the 64 sequences are two halves of 32 with independent chains (own half of every
// exchange block, accumulator columns, barriers, counter, publisher), see the kernel's comment.
//   warp 0      loads W once; per step and half: polls the group counter (acquire gpu-scope loads) and
//               TMA-loads the partials
//   warp 1      MMA issuer (one elected thread)
//   warps 2-9   cell update: warp w owns sequences 8w..8w+7 of the tile, lane = hidden unit of the
//               CTA, so every global access of a warp is one contiguous 128/256-byte row; dc and the
//               bias-gradient sums stay in registers for the whole sequence; after the MMA the same
//               warps drain TMEM (lane = output unit) to the exchange area.
//   warps 10-11 publisher of half 0 / 1 (one elected thread each): TMA-stores this CTA's partials, waits
//               for their completion and RELEASES the counter (the only gpu-scope fence of the step)
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "sm100.cuh"

namespace rcnn {
namespace {

using namespace sm100;

constexpr int NS = 64;     // sequences per work item (UMMA N)
constexpr int LK = 64;
constexpr uint32_t kWTile = 128 * LK * 2;   // [128 output units x 64 k] bf16, SW128 = 16 KB
constexpr uint32_t kBChunk = NS * LK * 2;   // [64 seq x 64 k] bf16, SW128 = 8 KB
constexpr uint32_t kStageWarp = 3 * 8 * 32 * 4;   // bytes of activation staging per cell warp (two-slot kernel)
constexpr int kThreads = 384;
constexpr int kThreads2 = 640;  // two slots: warps 2-9 cell update of slot 0, 10-17 of slot 1, 18-19 publishers   // warp 0 poll/TMA loads, warp 1 MMA, warps 2-9 cell update, warps 10-11 publishers (one per half)

struct BwdParams {
    int B, T, H;
    int nitems, ngroups;
    const __half *gates;       // [2, T, B, 4H] activated gates, packed order
    const float *csave;        // [2, T, B, H]
    const float *dhcat;        // [B, T, 2H] upstream gradient of the block's LSTM output
    __nv_bfloat16 *dG;         // [B, T, 2*4H] out: pre-activation gate gradients, packed order
    float *db;                 // [2*4H] out or nullptr: column sums of dG (bias gradient), zeroed before the launch
    __nv_bfloat16 *xbuf;       // exchange buffer [2][ngroups][src CTA][dst CTA][8 warps][32 units][8 seq]
    unsigned int *sync;        // [ngroups][slot][2 halves] zeroed before the launch
    long long *tl;             // debug timeline or nullptr
};
#define TL_MARK(k) do { if (tl) tl[(s) * 8 + (k)] = clock64(); } while (0)

__device__ __forceinline__ float tanh_fast_b(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// Publishing a step: the partials were written by TMA stores of this thread (complete: cp.async.bulk.wait_group 0,
// then fence.proxy.async); the counter increment must be a gpu-scope RELEASE and the poll an ACQUIRE.  A relaxed
// increment after the completed stores is NOT enough on this part: completion makes the writes visible to the
// issuing thread only, and readers on other SMs were observed (scripts/stress_block.py, B=4736 T=3 H=128 under
// RCNN_POISON=1) to fetch the previous contents of the slot after seeing the counter.
__device__ __forceinline__ void red_release_gpu_inc_b(unsigned int *p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_b(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// the four chain counters of a two-slot group ([slot][half], 16 bytes) in ONE acquire load: one round trip polls them all
__device__ __forceinline__ uint4 ld_acquire_gpu_v4_b(const unsigned int *p) {
    uint4 v;
    asm volatile("ld.acquire.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// Bounded spin on the group counter: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void wait_counter_b(const unsigned int *p, unsigned int target) {
    if (ld_acquire_gpu_b(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu_b(p) < target) {
        if (clock64() - t0 > 4000000000LL) {
            printf("rcnn-ocr_b200: lstm_bwd group counter timed out (block %d)\n", blockIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ uint2 ld_ro_v2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// NH = 2: the item's 64 sequences are two HALVES of 32 with independent dependency chains (the cell warps are split
// that way already: warps 0-3 / 4-7 own and drain sequences 0-31 / 32-63): own half of every 4 KB exchange block, own
// accumulator columns, barriers, counter and publisher thread.  The halves run half a step apart, so the exchange
// latency of one (TMA stores, release, counter propagation, TMA load) hides behind the MMAs and cell phase of the other.
//
// NSLOT = 2: the group works on TWO ITEMS at once (slot 0: item i, slot 1: item i + ngroups; same direction, so the
// same W_hh slice in tensor memory) -- four independent chains q = 2 slot + half per CTA.  Chosen by the host when a
// batch has more items than the GPU has groups (B > 256 at H = 512): the items of a group used to run back to back,
// each of them 55 % exchange latency (profiles/timeline_bwd_r01_final3.txt); now a step of one item runs inside the
// exchange of the other.  Per slot: own dG tile, own exchange area in shared and in global memory, own barriers and
// counters.  Shared between the slots: the cell warps (they alternate: step s of slot 0, step s of slot 1), the
// publisher threads, and the ACCUMULATOR columns (tensor memory is full: 256 columns of W_hh, 256 of accumulators) --
// a warp drains slot 0's accumulators before it hands slot 1's dG tile to the MMA thread, and the other way round in
// the next step, so a chain's b_ready barrier already orders the reuse.  Shared memory: the W_hh staging tiles
// (needed only until tcgen05.cp has copied them, once per launch: groups keep their direction) and the two exchange
// areas occupy the same bytes.
template <int NH, int NSLOT>
__global__ void __launch_bounds__(NSLOT == 2 ? kThreads2 : kThreads, 1)
lstm_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmXin,
                const __grid_constant__ CUtensorMap tmXout, const BwdParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int H = p.H, T = p.T, B = p.B;
    const int nmb = (H + 127) / 128;           // accumulators: blocks of 128 output units (H = 64: one, half used)
    const int gsize = H / 32;
    const size_t wbytes = (size_t)nmb * 2 * kWTile, xslot = (size_t)gsize * 4096;
    // NSLOT = 1: [W tiles][dG tile][exchange area];  NSLOT = 2: [W tiles | 2 exchange areas][2 dG tiles]
    unsigned char *w_s = smem;                                  // nmb x 2 tiles [128 units x 64 k] bf16, SW128
    unsigned char *b_s = NSLOT == 1 ? w_s + wbytes : smem + (wbytes > NSLOT * xslot ? wbytes : NSLOT * xslot);
                                                                // per slot 2 chunks [64 seq x 64 k] bf16, SW128: this CTA's dG_t
    unsigned char *x_s = NSLOT == 1 ? b_s + 2 * kBChunk : smem; // per slot gsize x 4 KB: partials in (by source) / out (by destination)
    // (two slots) per cell warp 3 KB: c_t, c_{t-1} and dh of the warp's 8 sequences x 32 units, fetched a step ahead with
    // cp.async instead of into registers (the sixteen cell warps have 112 registers each)
    unsigned char *stage_s = b_s + (size_t)NSLOT * 2 * kBChunk;
    uint64_t *bars = reinterpret_cast<uint64_t *>((NSLOT == 1 ? x_s + xslot : stage_s + 16 * kStageWarp));
    // barriers of chain q = 2 slot + half
    uint64_t *w_full = bars, *b_ready = bars + 1, *x_ready = bars + 5;
    uint64_t *d_full = bars + 9;     // [4 chains][4] accumulator block mb complete
    uint64_t *staged = bars + 25;    // [4 chains][4] block mb drained into x_s (the chain's cell warps)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 41);
    constexpr int HS = NS / NH;                 // sequences per half
    constexpr uint32_t XB = 4096 / NH;          // bytes of a half's share of one (source, destination) exchange block
    constexpr int XR = 4 / NH;                  // ... in 1 KB rows of the exchange tensor map

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warp roles.  One slot: 0 poller, 1 MMA, 2-9 cell update, 10-11 publishers.  Two slots: warps 0-3 are the four
    // single-thread roles (poller, MMA, two publishers) and give most of their registers to the sixteen cell warps
    // (setmaxnreg: 640 threads leave 96 registers each, the cell code wants ~150 -- spilled, with the poller's acquire
    // loads invalidating L1 all the time, its phases ran 5-10x slower)
    constexpr int kPubWarp = NSLOT == 2 ? 2 : 10;    // first publisher warp
    constexpr int kCell0 = NSLOT == 2 ? 4 : 2;       // first cell warp
    constexpr bool kIsCellFirst = true;
    (void)kIsCellFirst;
    const int group = blockIdx.x / gsize;
    const int c = blockIdx.x % gsize;
    long long *tl = (p.tl && blockIdx.x == 0) ? p.tl : nullptr;
    unsigned int *counter = p.sync + (size_t)group * NSLOT * 2;   // [slot][half]
    const int xgroups = p.ngroups * NSLOT;        // exchange areas per parity in global memory
    const int rstride = NSLOT * p.ngroups;        // items a group advances by per round
    // exchange buffer: [parity][group][slot][src CTA][dst CTA][warp 8][unit 32][seq 8] bf16 (4 KB per (src, dst))

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(w_full, 1);
            for (int i = 0; i < 4; ++i) { mbar_init(&b_ready[i], 8 / NH); mbar_init(&x_ready[i], 1); }
            for (int i = 0; i < 16; ++i) { mbar_init(&d_full[i], 1); mbar_init(&staged[i], 8 / NH); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);   // [0, 256): W_hh^T blocks as A operands (64 columns each); [256, 512): accumulators
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // processing order: the forward direction is back-propagated from t = T-1 down to 0, the
    // reverse direction from t = 0 up to T-1.  Every step but an item's last publishes partial sums.
    const bool single_thread_role = NSLOT == 2 ? warp < 4 : (warp < 2 || warp >= kPubWarp);
    if (single_thread_role) {
    if (warp == 0) {
        // ===== W loader + counter poller (one elected thread) ====================================
        if (elect_one()) {
            tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmXin); tma_prefetch_desc(&tmXout);
            int cur_dir = -1;
            unsigned int published = 0, npub = 0;
            for (int it0 = group; it0 < p.nitems; it0 += rstride) {
                const int dir = it0 & 1;
                const int nsl = (NSLOT == 2 && it0 + p.ngroups < p.nitems) ? 2 : 1;
                if (dir != cur_dir) {
                    if (NSLOT == 2 && cur_dir != -1) {   // the staging tiles share their bytes with the exchange areas
                        printf("rcnn-ocr_b200: lstm_bwd two-slot groups must keep their direction (block %d)\n", blockIdx.x);
                        __trap();
                    }
                    // rows = output units j of this direction, columns = the 128 packed gate indices of CTA c
                    mbar_arrive_expect_tx(w_full, (uint32_t)nmb * 2 * kWTile);
                    for (int mb = 0; mb < nmb; ++mb)
                        for (int kc = 0; kc < 2; ++kc)
                            tma_load_2d(w_s + (size_t)(mb * 2 + kc) * kWTile, &tmW, w_full, c * 128 + kc * LK, dir * H + mb * 128);
                    cur_dir = dir;
                }
                if (NSLOT == 2) {
                    // four chains, each at its own step: whichever counter has reached its target gets its partials fetched
                    unsigned int sq[4] = {1u, 1u, 1u, 1u};     // next step (1 .. T-1) whose partials chain q needs
                    int remaining = nsl * NH * (T - 1);
                    long long t0 = 0;
                    for (unsigned spins = 0; remaining > 0; ++spins) {
                        const uint4 v = ld_acquire_gpu_v4_b(counter);
                        const unsigned int val[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int slot = q >> 1, hf = q & 1;
                            if (hf >= NH || slot >= nsl || sq[q] >= (unsigned)T) continue;
                            if (val[q] < (published + sq[q]) * (unsigned)gsize) continue;
                            const int s = (int)sq[q];
                            if (q == 0) TL_MARK(0);
                            fence_proxy_async_global();
                            mbar_arrive_expect_tx(&x_ready[q], (uint32_t)gsize * XB);
                            tma_load_5d(x_s + (size_t)slot * xslot + (size_t)hf * gsize * XB, &tmXin, &x_ready[q], 0, XR * hf, c, 0,
                                        (int)((npub + sq[q] - 1u) & 1u) * xgroups + group * NSLOT + slot);
                            ++sq[q];
                            --remaining;
                            spins = 0;
                        }
                        if (spins == 4096u) t0 = clock64();
                        if (spins > 4096u && clock64() - t0 > 4000000000LL) {
                            printf("rcnn-ocr_b200: lstm_bwd group counters timed out (block %d)\n", blockIdx.x);
                            __trap();
                        }
                    }
                    npub += (unsigned)(T - 1);
                } else {
                    for (int s = 0; s < T; ++s) {
                        if (s > 0) {
                            // the partials of this CTA's units (all sources), published by the group in the previous step.
                            // (The counter includes this CTA's own increment, which followed the completion of its stores
                            // out of the same shared-memory area.)
                            for (int hf = 0; hf < NH; ++hf) {
                                wait_counter_b(counter + hf, (published + (unsigned)s) * (unsigned)gsize);
                                if (hf == 0) TL_MARK(0);
                                fence_proxy_async_global();
                                mbar_arrive_expect_tx(&x_ready[hf], (uint32_t)gsize * XB);
                                tma_load_5d(x_s + (size_t)hf * gsize * XB, &tmXin, &x_ready[hf], 0, XR * hf, c, 0,
                                            (int)((npub - 1) & 1) * xgroups + group);
                            }
                        }
                        if (s + 1 < T) ++npub;
                    }
                }
                published += (unsigned)(T - 1);
            }
        }
    } else if (warp >= kPubWarp && warp < kPubWarp + 2) {
        // ===== publisher of half (warp - kPubWarp), both slots: partial sums -> exchange buffer -> release ====
        const int hf = warp - kPubWarp;
        if (hf < NH && elect_one()) {
            tma_prefetch_desc(&tmXout);
            unsigned int npubs[2] = {0u, 0u};     // publishes so far, per slot
            auto publish = [&](const int slot, const int s) {
                const int q = slot * 2 + hf;
                unsigned char *xs = x_s + (size_t)slot * xslot + (size_t)hf * gsize * XB;
                const unsigned int np = npubs[slot];
                // block by block as the cell warps drain the accumulators: the partials for destination
                // CTAs 4mb .. 4mb+3 (contiguous in shared and global memory)
                for (int mb = 0; mb < nmb; ++mb) {
                    mbar_wait(&staged[q * 4 + mb], np & 1u);
                    if (q == 0 && mb == nmb - 1) TL_MARK(7);
                    tma_store_5d(&tmXout, xs + (size_t)mb * 4 * XB, 0, XR * hf, 4 * mb, c,
                                 (int)(np & 1u) * xgroups + group * NSLOT + slot);
                    tma_store_commit();
                }
                tma_store_wait<0>();                      // written, not merely read out of shared memory
                if (q == 0) TL_MARK(1);
                // the partials are complete for this thread; the release increment makes them visible to the
                // group (readers poll with acquire loads and fetch with TMA)
                fence_proxy_async_global();
                red_release_gpu_inc_b(counter + q);
                if (q == 0) TL_MARK(6);
                ++npubs[slot];
            };
            for (int it0 = group; it0 < p.nitems; it0 += rstride) {
                const int nsl = (NSLOT == 2 && it0 + p.ngroups < p.nitems) ? 2 : 1;
                if (NSLOT == 2) {
                    // whichever slot's first block has been drained
                    int sp[2] = {0, 0};
                    int remaining = nsl * (T - 1);
                    long long t0 = 0;
                    for (unsigned spins = 0; remaining > 0; ++spins) {
#pragma unroll
                        for (int slot = 0; slot < 2; ++slot) {
                            if (slot >= nsl || sp[slot] >= T - 1) continue;
                            if (!mbar_test(&staged[(slot * 2 + hf) * 4], npubs[slot] & 1u)) continue;
                            publish(slot, sp[slot]);
                            ++sp[slot]; --remaining; spins = 0;
                        }
                        if (spins == 65536u) t0 = clock64();
                        if (spins > 65536u && clock64() - t0 > 4000000000LL) {
                            printf("rcnn-ocr_b200: lstm_bwd publisher timed out (block %d)\n", blockIdx.x);
                            __trap();
                        }
                    }
                } else {
                    for (int s = 0; s + 1 < T; ++s) publish(0, s);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) ===================================================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, HS);
            int cur_dir = -1;
            uint32_t wphase = 0, bphase = 0;
            unsigned int ni[4] = {0u, 0u, 0u, 0u};   // (two slots) products issued per chain since the launch
            for (int it0 = group; it0 < p.nitems; it0 += rstride) {
                const int dir = it0 & 1;
                const int nsl = (NSLOT == 2 && it0 + p.ngroups < p.nitems) ? 2 : 1;
                if (dir != cur_dir) {
                    // weights: shared memory -> tensor memory, one K = 16 slice per copy (in issue order with the MMAs)
                    mbar_wait(w_full, wphase);
                    wphase ^= 1;
                    cur_dir = dir;
                    tc_fence_after();
                    for (int t2 = 0; t2 < nmb * 2; ++t2) {
                        const uint64_t wdesc = make_smem_desc_sw128(smem_u32(w_s + (size_t)t2 * kWTile), 16, 1024);
#pragma unroll
                        for (int k = 0; k < LK / 16; ++k)
                            tmem_cp_128x256b(tmem_base + (uint32_t)(8 * (t2 * (LK / 16) + k)), wdesc + (uint64_t)(2 * k));
                    }
                }
                auto product = [&](int slot, int hf) {   // partial^T blocks of chain (slot, hf): nmb x 8 MMAs
                    const int q = slot * 2 + hf;
                    tc_fence_after();
                    for (int mb = 0; mb < nmb; ++mb) {
#pragma unroll
                        for (int kc = 0; kc < 2; ++kc) {
                            // rows (sequences) [HS hf, +HS) of the slot's dG chunk
                            const uint64_t bdesc = make_smem_desc_sw128(
                                smem_u32(b_s + (size_t)slot * 2 * kBChunk + kc * kBChunk + hf * HS * 128), 16, 1024);
#pragma unroll
                            for (int k = 0; k < LK / 16; ++k)
                                umma_bf16_ts(tmem_base + 256u + (uint32_t)(mb * NS + hf * HS),
                                             tmem_base + (uint32_t)(8 * ((mb * 2 + kc) * (LK / 16) + k)),
                                             bdesc + (uint64_t)(2 * k), idesc, (kc | k) != 0);
                        }
                        umma_commit(&d_full[q * 4 + mb]);    // the cell warps drain block mb while the next one is multiplied
                    }
                };
                if (NSLOT == 2) {
                    // four chains in whatever order their dG tiles arrive.  The two slots of a half share the accumulator
                    // columns: a product is issued only when the other slot's previous one has been drained.
                    unsigned int rs[4] = {0u, 0u, 0u, 0u};     // products issued this round
                    int remaining = nsl * NH * (T - 1);
                    long long t0 = 0;
                    for (unsigned spins = 0; remaining > 0; ++spins) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int slot = q >> 1, hf = q & 1, o = q ^ 2;
                            if (hf >= NH || slot >= nsl || rs[q] >= (unsigned)(T - 1)) continue;
                            if (!mbar_test(&b_ready[q], ni[q] & 1u)) continue;
                            if (ni[o] > 0u && !mbar_test(&staged[o * 4 + nmb - 1], (ni[o] - 1u) & 1u)) continue;
                            const int s = (int)rs[q];
                            if (q == 0) TL_MARK(2);
                            product(slot, hf);
                            if (q == NH - 1) TL_MARK(3);
                            ++ni[q];
                            ++rs[q];
                            --remaining;
                            spins = 0;
                        }
                        if (spins == 65536u) t0 = clock64();
                        if (spins > 65536u && clock64() - t0 > 4000000000LL) {
                            printf("rcnn-ocr_b200: lstm_bwd MMA thread timed out (block %d)\n", blockIdx.x);
                            __trap();
                        }
                    }
                } else {
                    for (int s = 0; s + 1 < T; ++s) {
                        for (int hf = 0; hf < NH; ++hf) {
                            mbar_wait(&b_ready[hf], bphase);
                            if (hf == 0) TL_MARK(2);
                            product(0, hf);
                            if (hf == NH - 1) TL_MARK(3);
                        }
                        bphase ^= 1;
                    }
                }
            }
        }
    }
    } else {
        // ===== cell update + exchange ============================================================
        const int myslot = (warp - kCell0) >> 3;   // (two slots) warps 4-11 serve slot 0, warps 12-19 slot 1
        const int w = (warp - kCell0) & 7;    // sequences 8w .. 8w+7 of the tile; lane = hidden unit 32c + lane
        const int qd = warp & 3;              // TMEM lane quadrant (drain phase)
        const int hq = w >> 2;                // which 32 of the 64 sequence columns this warp drains
        const int hf = NH == 2 ? hq : 0;      // the half this warp belongs to (owns AND drains the same 32 sequences)
        // saved activations of one (slot, step): issued before anything of the phase is waited for.  (Not earlier: the
        // proxy fences of the dG store and of the drain wait for the thread's outstanding loads -- with the loads of the
        // next step in flight across them the step grew from 9,000 to 10,500 cycles.)
        uint2 gq[8];
        float cc[8], cp[8], dhu[8];
        const uint32_t stg = smem_u32(stage_s) + (uint32_t)((warp - kCell0) & 15) * kStageWarp + (uint32_t)lane * 4u;   // + (array * 8 + i) * 128
        auto issue_loads = [&](int dir, int b0, int s) {
            const int t = dir ? s : T - 1 - s;
            const int tfp = dir ? t + 1 : t - 1;   // forward-time predecessor: where c_{prev} lives
            // (two slots: the base pointers pass through an empty asm so that the compiler cannot keep the 32 per-sequence
            // addresses of a step alive across the whole time loop)
            const __half *gates_p = p.gates;
            const float *csave_p = p.csave, *dhcat_p = p.dhcat;
            if (NSLOT == 2) asm volatile("" : "+l"(gates_p), "+l"(csave_p), "+l"(dhcat_p));
            const bool has_prev = tfp >= 0 && tfp < T;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int b = b0 + 8 * w + i;
                if (b < B) {
                    const size_t rtb = ((size_t)dir * T + t) * B + b;
                    gq[i] = ld_ro_v2(gates_p + rtb * 4 * H + (size_t)c * 128 + 4 * lane);
                    const float *pc = csave_p + rtb * H + 32 * c + lane;
                    const float *pp = csave_p + (((size_t)dir * T + (has_prev ? tfp : t)) * B + b) * H + 32 * c + lane;
                    const float *pd = dhcat_p + ((size_t)b * T + t) * 2 * H + (size_t)dir * H + 32 * c + lane;
                    if (NSLOT == 2) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(stg + (uint32_t)(0 * 8 + i) * 128u), "l"(pc) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(stg + (uint32_t)(1 * 8 + i) * 128u), "l"(pp) : "memory");
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(stg + (uint32_t)(2 * 8 + i) * 128u), "l"(pd) : "memory");
                    } else {
                        cc[i] = ld_nc_f32(pc);
                        cp[i] = has_prev ? ld_nc_f32(pp) : 0.f;
                        dhu[i] = ld_nc_f32(pd);
                    }
                } else {
                    gq[i] = make_uint2(0u, 0u);
                    if (NSLOT == 2) {
#pragma unroll
                        for (int a = 0; a < 3; ++a) asm volatile("st.shared.f32 [%0], %1;" ::"r"(stg + (uint32_t)(a * 8 + i) * 128u), "f"(0.f) : "memory");
                    } else {
                        cc[i] = 0.f; cp[i] = 0.f; dhu[i] = 0.f;
                    }
                }
            }
            if (NSLOT == 2) asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // (two slots) the staged activations of this phase -> registers; c_{t-1} of a direction's first step is zero
        auto fetch_staged = [&](int dir, int s) {
            const int t = dir ? s : T - 1 - s;
            const int tfp = dir ? t + 1 : t - 1;
            const bool has_prev = tfp >= 0 && tfp < T;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cc[i]) : "r"(stg + (uint32_t)(0 * 8 + i) * 128u) : "memory");
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cp[i]) : "r"(stg + (uint32_t)(1 * 8 + i) * 128u) : "memory");
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dhu[i]) : "r"(stg + (uint32_t)(2 * 8 + i) * 128u) : "memory");
                if (!has_prev) cp[i] = 0.f;
            }
        };
        unsigned int xph = 0u, dph = 0u;       // x_ready / d_full phases consumed by the warp's chain
        for (int it0 = group; it0 < p.nitems; it0 += rstride) {
            const int dir = it0 & 1;
            const int nsl = (NSLOT == 2 && it0 + p.ngroups < p.nitems) ? 2 : 1;
            float dc[NSLOT][8], dbacc[4];   // (only dc[0] / dc[NSLOT-1] of the warp's own slot is live)
#pragma unroll
            for (int sl = 0; sl < NSLOT; ++sl)
#pragma unroll
                for (int i = 0; i < 8; ++i) dc[sl][i] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) dbacc[i] = 0.f;
            // ---- cell phase of (slot, step s): partial sums in -> dh -> gate gradients -> dG tile for the MMA thread
            auto cell_phase = [&](const int slot, float (&dcs)[8], const int s) {
                const int t = dir ? s : T - 1 - s;
                const bool more = s + 1 < T;
                const int q = slot * 2 + hf;
                const int b0 = ((it0 + slot * p.ngroups) >> 1) * NS;
                unsigned char *xs = x_s + (size_t)slot * xslot + (size_t)hf * gsize * XB;   // the chain's exchange area: partials in / out
                unsigned char *bs = b_s + (size_t)slot * 2 * kBChunk;
                issue_loads(dir, b0, s);
                // recurrent term: sum of the partials every CTA of the group published last step
                float rec[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) rec[i] = 0.f;
                if (s > 0) {
                    mbar_wait(&x_ready[q], xph & 1u);
                    const unsigned char *src = xs + ((size_t)(w % (8 / NH)) * 256 + (size_t)lane * 8) * 2;
                    constexpr int NV = NSLOT == 2 ? 4 : 8;      // shared-memory loads in flight per round (registers)
                    for (int sc = 0; sc < gsize; sc += NV) {
                        uint4 v[NV];
#pragma unroll
                        for (int u = 0; u < NV; ++u)
                            v[u] = sc + u < gsize ? *reinterpret_cast<const uint4 *>(src + (size_t)(sc + u) * XB) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                        for (int u = 0; u < NV; ++u) {
                            const uint32_t wd[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                rec[2 * k] += __uint_as_float(wd[k] << 16);
                                rec[2 * k + 1] += __uint_as_float(wd[k] & 0xffff0000u);
                            }
                        }
                    }
                    ++xph;
                }
                if (NSLOT == 2) fetch_staged(dir, s);
                uint2 dgq[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 if_ = __half22float2(*reinterpret_cast<const __half2 *>(&gq[i].x));   // (i, f)
                    const float2 go_ = __half22float2(*reinterpret_cast<const __half2 *>(&gq[i].y));   // (g, o)
                    const float ig = if_.x, fg = if_.y, gg = go_.x, og = go_.y;
                    const float dh = dhu[i] + rec[i];
                    const float tc = tanh_fast_b(cc[i]);
                    const float d_o = dh * tc;
                    const float dct = fmaf(dh * og, 1.f - tc * tc, dcs[i]);
                    dcs[i] = dct * fg;
                    const float o0 = dct * gg * ig * (1.f - ig);
                    const float o1 = dct * cp[i] * fg * (1.f - fg);
                    const float o2 = dct * ig * (1.f - gg * gg);
                    const float o3 = d_o * og * (1.f - og);
                    dbacc[0] += o0; dbacc[1] += o1; dbacc[2] += o2; dbacc[3] += o3;   // (rows b >= B contribute zeros)
                    const __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1), hi = __floats2bfloat162_rn(o2, o3);
                    dgq[i].x = *reinterpret_cast<const uint32_t *>(&lo);
                    dgq[i].y = *reinterpret_cast<const uint32_t *>(&hi);
                }
                if (more) {
                    // dG_t of this CTA as the K-major SW128 B operand: row = sequence, k = 4*lane + gate
                    const int chunk = lane >> 4, piece = (lane & 15) >> 1, sub = (lane & 1) * 8;
                    // (32-bit shared addresses formed from one base: eight generic pointers kept across the time loop cost
                    // the two-slot kernel's 96-register budget six spilled pairs)
                    const uint32_t bs0 = smem_u32(bs) + (uint32_t)(chunk * kBChunk + 8 * w * 128 + sub);
#pragma unroll
                    for (int i = 0; i < 8; ++i)   // row = 8 w + i, so row & 7 = i
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(bs0 + (uint32_t)(i * 128) + (uint32_t)((piece ^ i) << 4)),
                                     "r"(dgq[i].x), "r"(dgq[i].y) : "memory");
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&b_ready[q]);
                }
                __nv_bfloat16 *dG_p = p.dG;
                int tt = t;
                if (NSLOT == 2) asm volatile("" : "+l"(dG_p), "+r"(tt));   // addresses formed here, not carried through the phase
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int b = b0 + 8 * w + i;
                    if (b < B)
                        *reinterpret_cast<uint2 *>(dG_p + ((size_t)b * T + tt) * 8 * H + (size_t)dir * 4 * H + (size_t)c * 128 + 4 * lane) = dgq[i];
                }
                if (warp == kCell0 && lane == 0) TL_MARK(5);
            };
            // ---- drain of (slot, step s): accumulators partial^T [unit, seq] -> exchange area (bf16), block by block
            auto drain_phase = [&](const int slot, const int s) {
                const int q = slot * 2 + hf;
                unsigned char *xs = x_s + (size_t)slot * xslot + (size_t)hf * gsize * XB;
                for (int mb = 0; mb < nmb; ++mb) {
                    mbar_wait(&d_full[q * 4 + mb], dph & 1u);
                    if (warp == kCell0 && lane == 0 && mb == nmb - 1) TL_MARK(4);
                    tc_fence_after();
                    uint32_t acc[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + 256u + (uint32_t)(mb * NS + hq * 32), acc);
                    tmem_ld_wait();
                    const int dstc = mb * 4 + qd;                 // CTA that owns output units [128mb + 32qd, +32)
                    if (dstc < gsize) {
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) {          // 8 sequences = one reader warp's slice
                            uint4 v;
                            uint32_t *vw = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(acc[8 * g4 + 2 * k]),
                                                                                 __uint_as_float(acc[8 * g4 + 2 * k + 1]));
                                vw[k] = *reinterpret_cast<const uint32_t *>(&h2);
                            }
                            *reinterpret_cast<uint4 *>(xs + (size_t)dstc * XB + ((size_t)((NH == 2 ? 0 : hq) * 4 + g4) * 256 + (size_t)lane * 8) * 2) = v;
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&staged[q * 4 + mb]);
                }
                tc_fence_before();
                ++dph;
            };
            // every slot has its own eight warps (their own registers for the activations loaded a step ahead), so the
            // chains of the two slots depend on each other only through the tensor pipe and the accumulator columns.
            // (Measured first: ONE set of warps serving both slots, in a fixed order A0 B0 A1 B1 / A0 A1 B0 B1 -- one
            // chain's wait holds the other's publish back, 14,100 - 15,200 cycles per step pair -- and event driven, which
            // needs the activation loads issued on demand: 20,600.  Back to back the pair costs 2 x 9,150.)
            if (myslot < nsl) {
                for (int s = 0; s < T; ++s) {
                    if (myslot == 0) cell_phase(0, dc[0], s); else cell_phase(NSLOT - 1, dc[NSLOT - 1], s);
                    if (s + 1 < T) drain_phase(myslot, s);
                }
            }
            if (p.db != nullptr && myslot < nsl) {
                float *dst = p.db + (size_t)dir * 4 * H + (size_t)c * 128 + 4 * lane;
#pragma unroll
                for (int i = 0; i < 4; ++i) atomicAdd(dst + i, dbacc[i]);
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

size_t bwd_smem_bytes(int H, int nslot) {
    const size_t wbytes = (size_t)((H + 127) / 128) * 2 * kWTile, xslot = (size_t)(H / 32) * 4096;
    if (nslot == 1) return 1024 + wbytes + 2 * kBChunk + xslot + 512;
    return 1024 + (wbytes > nslot * xslot ? wbytes : nslot * xslot) + (size_t)nslot * 2 * kBChunk + 16 * kStageWarp + 512;
}

// How a batch is laid over the GPU: groups of H/32 CTAs (one CTA per SM: shared memory), each working on `nslot`
// items (64 sequences of one direction) at a time.  One slot while every item gets a group of its own; two slots
// once items would queue behind each other (RCNN_BWD_SLOTS=1 keeps the one-slot kernel for A/B measurements).
struct BwdCfg { int nslot, ngroups; };
BwdCfg bwd_config(int B, int H) {
    static const int force = getenv("RCNN_BWD_SLOTS") ? atoi(getenv("RCNN_BWD_SLOTS")) : 0;
    const int gsize = H / 32, nitems = 2 * ((B + NS - 1) / NS);
    const int max_groups = num_sms() / gsize;
    BwdCfg cfg;
    if (force != 1 && nitems > max_groups && max_groups >= 2) {
        const int even_max = max_groups & ~1;
        int need = (nitems + 1) / 2;
        need += need & 1;                            // even: slot 1 = item + ngroups has the direction of slot 0
        cfg.nslot = 2;
        cfg.ngroups = need < even_max ? need : even_max;
        return cfg;
    }
    int ng = nitems < max_groups ? nitems : max_groups;
    if (ng > 1 && (ng & 1) && nitems > ng) --ng;   // even: a group keeps its direction (and W slice)
    cfg.nslot = 1;
    cfg.ngroups = ng;
    return cfg;
}

// column sums of dG: db[col] = sum_rows dG[row, col]   (rows = B*T, cols = 8H)
__global__ void colsum_bf16_kernel(const __nv_bfloat16 *__restrict__ src, long long ld, long long rows, int cols,
                                   float *__restrict__ out) {
    // block = 32 x 8 threads: 32 consecutive columns, 8 row phases; grid.y splits the rows
    __shared__ float part[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (col < cols) {
        for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8)
            s += __bfloat162float(src[r * ld + col]);
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < cols) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += part[i][threadIdx.x];
        atomicAdd(&out[col], tot);
    }
}

// vectorised variant: 8 columns (one 128-bit load) per thread; needs cols % 8 == 0, ld % 8 == 0
__global__ void colsum_bf16_v8_kernel(const uint4 *__restrict__ src, long long ld8, long long rows, int cols8,
                                      float *__restrict__ out) {
    __shared__ float part[8][32][9];
    const int c8 = blockIdx.x * 32 + threadIdx.x;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (c8 < cols8) {
        const long long step = (long long)gridDim.y * 8;
        long long r = (long long)blockIdx.y * 8 + threadIdx.y;
        auto add = [&](const uint4 &v) {
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc[2 * k] += __uint_as_float(w[k] << 16);
                acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
            }
        };
        for (; r + 7 * step < rows; r += 8 * step) {       // eight independent 128-bit loads in flight per thread
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = ld_nc_v4(src + (r + j * step) * ld8 + c8);
#pragma unroll
            for (int j = 0; j < 8; ++j) add(v[j]);
        }
        for (; r < rows; r += step) add(ld_nc_v4(src + r * ld8 + c8));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[threadIdx.y][threadIdx.x][k] = acc[k];
    __syncthreads();
    if (threadIdx.y == 0 && c8 < cols8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float tot = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) tot += part[i][threadIdx.x][k];
            atomicAdd(&out[c8 * 8 + k], tot);
        }
    }
}

// hprev for the dW_hh GEMM: out[b, t, dir*H + u] = hcat[b, t-1 (dir 0) / t+1 (dir 1), dir*H + u], zero at
// the first step of that direction: the h that multiplied W_hh when gates_t were formed.
__global__ void hprev_shift_kernel(const uint4 *__restrict__ hcat, uint4 *__restrict__ out, int B, int T, int H) {
    const int vec_per_row = 2 * H / 8;           // uint4 = 8 bf16
    const int vec_per_dir = H / 8;
    const long long total = (long long)B * T * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec_per_row);
        const long long bt = i / vec_per_row;
        const int t = (int)(bt % T);
        const int dir = v / vec_per_dir;
        const int tp = dir ? t + 1 : t - 1;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (tp >= 0 && tp < T) val = hcat[(bt + (tp - t)) * vec_per_row + v];
        out[i] = val;
    }
}

// Scatter the packed-order weight gradients back to torch's layout (fp32):
//   dW_ih[dir][g*H+32c+j, :] = dWih_p[dir*4H + p, :], same for dW_hh; db_ih = db_hh = db_p.
struct UnpackArgs {
    const float *dwih_p, *dwhh_p, *db_p;   // [8H, I], [8H, H], [8H]
    float *dw_ih[2], *dw_hh[2], *db_ih[2], *db_hh[2];
    int I, H;
};
__global__ void lstm_unpack_grads_kernel(const UnpackArgs a) {
    const int H = a.H, I = a.I;
    const int dir = blockIdx.x / (4 * H), pidx = blockIdx.x % (4 * H);
    const int cc = pidx >> 7, j = (pidx >> 2) & 31, g = pidx & 3;
    const int r = g * H + 32 * cc + j;
    const size_t prow = (size_t)dir * 4 * H + pidx;
    if ((I & 3) == 0 && (H & 3) == 0) {
        for (int k = 4 * threadIdx.x; k < I; k += 4 * blockDim.x)
            *reinterpret_cast<float4 *>(a.dw_ih[dir] + (size_t)r * I + k) = *reinterpret_cast<const float4 *>(a.dwih_p + prow * I + k);
        for (int k = 4 * threadIdx.x; k < H; k += 4 * blockDim.x)
            *reinterpret_cast<float4 *>(a.dw_hh[dir] + (size_t)r * H + k) = *reinterpret_cast<const float4 *>(a.dwhh_p + prow * H + k);
    } else {
        for (int k = threadIdx.x; k < I; k += blockDim.x) a.dw_ih[dir][(size_t)r * I + k] = a.dwih_p[prow * I + k];
        for (int k = threadIdx.x; k < H; k += blockDim.x) a.dw_hh[dir][(size_t)r * H + k] = a.dwhh_p[prow * H + k];
    }
    if (threadIdx.x == 0) {
        const float v = a.db_p[prow];
        a.db_ih[dir][r] = v;
        a.db_hh[dir][r] = v;
    }
}

}  // namespace
}  // namespace rcnn

namespace rcnn { void fwdx_plan(int B, int H, int *nslot, int *ngroups); }

extern "C" int rcnn_lstm_plan(int which, int B, int H, int *nslot, int *ngroups) {
    using namespace rcnn;
    RCNN_CHECK_ARG(nslot && ngroups, "lstm_plan: null pointer");
    RCNN_CHECK_ARG(B > 0 && (H == 64 || H == 128 || H == 256 || H == 512), "lstm_plan: bad shape B=%d H=%d", B, H);
    RCNN_CHECK_ARG(which == 0 || which == 1, "lstm_plan: which = %d (0 forward, 1 backward)", which);
    if (which == 1) {
        const BwdCfg cfg = bwd_config(B, H);
        *nslot = cfg.nslot;
        *ngroups = cfg.ngroups;
    } else {
        fwdx_plan(B, H, nslot, ngroups);
    }
    return RCNN_OK;
}

extern "C" size_t rcnn_lstm_backward_workspace_bytes(int B, int T, int H) {
    using namespace rcnn;
    if (B <= 0 || T <= 0 || !(H == 64 || H == 128 || H == 256 || H == 512)) return 0;
    const size_t gsize = H / 32;
    const BwdCfg cfg = bwd_config(B, H);
    return 2 * (size_t)cfg.ngroups * cfg.nslot * gsize * gsize * 8 * 32 * 8 * sizeof(__nv_bfloat16);
}

extern "C" int rcnn_lstm_backward(const void *whh_pt, const void *gates_save, const float *c_save, const float *dhcat,
                                  int B, int T, int H, void *dG, float *db, void *workspace, size_t workspace_bytes,
                                  rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 0, "lstm_backward: bad shape B=%d T=%d", B, T);
    RCNN_CHECK_ARG(H == 64 || H == 128 || H == 256 || H == 512,
                   "lstm_backward: hidden size %d unsupported (64, 128, 256 or 512)", H);
    if (db) RCNN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * 8 * (size_t)H, (cudaStream_t)stream));
    if (B == 0 || T == 0) return RCNN_OK;
    RCNN_CHECK_ARG(whh_pt && gates_save && c_save && dhcat && dG, "lstm_backward: null pointer");
    if (!workspace || workspace_bytes < rcnn_lstm_backward_workspace_bytes(B, T, H)) {
        set_error("lstm_backward: workspace of %zu bytes needed, %zu given", rcnn_lstm_backward_workspace_bytes(B, T, H),
                  workspace ? workspace_bytes : (size_t)0);
        return RCNN_ERR_WORKSPACE;
    }
    CUtensorMap tw;
    // whh_pt [2, H, 4H]: rows = (direction, output unit j), columns = packed gate index; box = [128 units x 64 k]
    // (H = 64: the box reads 64 rows past the direction's block or zero fill; those accumulator lanes are unused)
    int rc = make_tmap_2d(&tw, whh_pt, 2, 2ull * H, 4ull * H, 4ull * H * 2, 128, LK, 1);
    if (rc) return rc;
    static const int halves = getenv("RCNN_BWD_HALVES") ? (atoi(getenv("RCNN_BWD_HALVES")) == 1 ? 1 : 2) : 2;
    const BwdCfg bcfg = bwd_config(B, H);
    const int gsz = H / 32, ngr = bcfg.ngroups * bcfg.nslot;   // exchange areas per parity
    CUtensorMap txin, txout;
    {   // exchange buffer as float32 [2*ngroups][src][dst][4][256]: load box = all sources of one destination,
        // store box = all destinations of one source
        const uint64_t dims[5] = {256, 4, (uint64_t)gsz, (uint64_t)gsz, 2ull * ngr};
        const uint64_t strides[4] = {1024, 4096, (uint64_t)gsz * 4096, (uint64_t)gsz * gsz * 4096};
        const uint32_t xr = 4u / (uint32_t)halves;
        const uint32_t box_in[5] = {256, xr, 1, (uint32_t)gsz, 1}, box_out[5] = {256, xr, (uint32_t)(gsz < 4 ? gsz : 4), 1, 1};
        rc = make_tmap_nd(&txin, workspace, 4, 5, dims, strides, box_in, 0);
        if (rc) return rc;
        rc = make_tmap_nd(&txout, workspace, 4, 5, dims, strides, box_out, 0);
        if (rc) return rc;
    }
    BwdParams p;
    p.B = B; p.T = T; p.H = H;
    p.gates = (const __half *)gates_save;
    p.csave = c_save;
    p.dhcat = dhcat;
    p.dG = (__nv_bfloat16 *)dG;
    p.db = db;
    p.xbuf = (__nv_bfloat16 *)workspace;
    p.tl = debug_timeline();
    const int gsize = H / 32;
    const size_t smem = bwd_smem_bytes(H, bcfg.nslot);
    cudaStream_t s = (cudaStream_t)stream;
    auto kern = bcfg.nslot == 2 ? (halves == 2 ? lstm_bwd_kernel<2, 2> : lstm_bwd_kernel<1, 2>)
                                : (halves == 2 ? lstm_bwd_kernel<2, 1> : lstm_bwd_kernel<1, 1>);
    RCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.nitems = 2 * ((B + NS - 1) / NS);
    p.ngroups = bcfg.ngroups;
    p.sync = group_counters(p.ngroups * bcfg.nslot * 2, s);
    if (!p.sync) return RCNN_ERR_CUDA_BASE;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(gsize * p.ngroups));
    cfg.blockDim = dim3(bcfg.nslot == 2 ? kThreads2 : kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the groups spin on each other
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ProfScope prof(RCNN_K_LSTM_BWD, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, txin, txout, p));
    count_launch();
    return RCNN_OK;
}

extern "C" int rcnn_colsum_bf16(const void *src, int64_t ld, int64_t rows, int cols, float *out, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(rows >= 0 && cols >= 0, "colsum: bad shape");
    if (cols == 0) return RCNN_OK;
    RCNN_CHECK_ARG(out, "colsum: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    RCNN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)cols, s));
    if (rows == 0) return RCNN_OK;
    RCNN_CHECK_ARG(src, "colsum: null pointer");
    if (cols % 8 == 0 && ld % 8 == 0 && ((uintptr_t)src % 16) == 0) {
        const int cols8 = cols / 8;
        // ~1 block of 256 threads per SM: the column sums end in same-address atomics (one per block and
        // column), which serialise in L2 -- few fat blocks with deep loads beat many thin ones
        const long long want = ((long long)num_sms() * 32) / cols8 + 1;
        const long long maxy = (rows + 7) / 8;
        dim3 block(32, 8), grid((cols8 + 31) / 32, (unsigned)(want < maxy ? want : maxy));
        colsum_bf16_v8_kernel<<<grid, block, 0, s>>>((const uint4 *)src, ld / 8, rows, cols8, out);
        RCNN_LAUNCH_CHECK("colsum_bf16_v8_kernel");
        return RCNN_OK;
    }
    dim3 block(32, 8), grid((cols + 31) / 32, (unsigned)(((rows + 7) / 8) < 64 ? ((rows + 7) / 8) : 64));
    colsum_bf16_kernel<<<grid, block, 0, s>>>((const __nv_bfloat16 *)src, ld, rows, cols, out);
    RCNN_LAUNCH_CHECK("colsum_bf16_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_lstm_hprev(const void *hcat, void *out, int B, int T, int H, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 0 && H > 0 && H % 8 == 0, "hprev: bad shape");
    if (B == 0 || T == 0) return RCNN_OK;
    RCNN_CHECK_ARG(hcat && out, "hprev: null pointer");
    const long long total = (long long)B * T * (2 * H / 8);
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    hprev_shift_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint4 *)hcat, (uint4 *)out, B, T, H);
    RCNN_LAUNCH_CHECK("hprev_shift_kernel");
    return RCNN_OK;
}

extern "C" int rcnn_lstm_unpack_grads(const float *dwih_p, const float *dwhh_p, const float *db_p, int I, int H,
                                      float *dw_ih_f, float *dw_hh_f, float *db_ih_f, float *db_hh_f,
                                      float *dw_ih_r, float *dw_hh_r, float *db_ih_r, float *db_hh_r,
                                      rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(I > 0 && H > 0 && H % 32 == 0, "unpack_grads: bad sizes");
    RCNN_CHECK_ARG(dwih_p && dwhh_p && db_p && dw_ih_f && dw_hh_f && db_ih_f && db_hh_f && dw_ih_r && dw_hh_r &&
                       db_ih_r && db_hh_r, "unpack_grads: null pointer");
    UnpackArgs a;
    a.dwih_p = dwih_p; a.dwhh_p = dwhh_p; a.db_p = db_p;
    a.dw_ih[0] = dw_ih_f; a.dw_hh[0] = dw_hh_f; a.db_ih[0] = db_ih_f; a.db_hh[0] = db_hh_f;
    a.dw_ih[1] = dw_ih_r; a.dw_hh[1] = dw_hh_r; a.db_ih[1] = db_ih_r; a.db_hh[1] = db_hh_r;
    a.I = I; a.H = H;
    lstm_unpack_grads_kernel<<<8 * H, 128, 0, (cudaStream_t)stream>>>(a);
    RCNN_LAUNCH_CHECK("lstm_unpack_grads_kernel");
    return RCNN_OK;
}
