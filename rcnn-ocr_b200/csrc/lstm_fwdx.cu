// K2 (forward, inference and training): recurrent LSTM kernel with the INPUT PROJECTION fused in.
//
// Same decomposition, exchange and cell arithmetic as lstm_fwd.cu (groups of H/32 CTAs, 64-sequence work
// items, transposed product with the weights as the M = 128 operand, W_hh in tensor memory).  In addition the
// CTA keeps its 128 x I slice of W_ih resident in SHARED memory and computes
//     acc_t  = W_ih_slice x_t^T                 (K = I, operands from shared memory: x_t streams through a
//                                                2-slot TMA ring straight out of the bf16 input tensor)
//     acc_t += W_hh_slice h_{t-1}^T             (K = H, A from tensor memory, h tile by TMA as before)
// into one of two alternating TMEM accumulators: the x half of step t+1 does not depend on the recurrence, so
// its MMAs are issued right after the h half of step t and execute while the cell warps, the publish and the
// group counter of step t are in flight -- the tensor pipe idles ~75 % of a step otherwise.  The separate xp
// GEMM (nn.LSTM's `W_ih x_t + b` for all t, model/model.py:154-156,161), its 134 MB fp16 output and the
// re-read of that output disappear from the inference path; the bias is a per-thread constant (a thread owns
// one gate row).  Threads: warp 0 TMA producer + counter poller, warp 1 MMA issuer, warps 2-9 cell update (they
// store h_t to hcat themselves), warp 10 publisher (one release per half and step: a MEMBAR issued by the MMA
// thread would wait behind its queued MMAs).  SAVE (training): the activated gates (fp16, lane pairs exchange halves so that every store is 32
// bits) and c_t leave by plain global stores after the step is published -- off the dependent chain, and shared
// memory has no room left for a staging tile.
//
// EXCHANGE, experimental variant (LL = true, RCNN_EXCHANGE=ll; measured SLOWER than the counter protocol, which stays
// the default -- profiles/ll_exchange_r02.txt): FLAG-IN-DATA, fetched optimistically by TMA and VALIDATED element by element.
// hcat is filled with the bf16 pattern 0xFFFF before the launch (a NaN encoding the cell arithmetic never produces:
// cvt.rn.bf16 returns the canonical NaN 0x7FFF), so "this element has been written" can be read off the element.
//   * the cell warps store h_t with relaxed gpu-scope 16-byte stores -- no fence, no counter;
//   * the loader warp (warp 10) polls one CANARY word per source CTA and half with relaxed gpu-scope loads (L2 is
//     the point of coherence); when the canaries of a half are all in, it TMA-loads the half's h tile (as before);
//   * the canaries prove nothing about the other packets, so the half's cell warps SCAN the tile in shared memory
//     and re-fetch any packet that still holds a sentinel element straight from global memory (relaxed gpu-scope
//     loads, polled until valid) before they release the tile to the MMA thread (generic -> async proxy fence).
// Every element the MMA consumes has therefore been seen with a written value: the protocol makes no assumption
// about store atomicity or ordering (each location of hcat is written exactly once per launch).  The step's exchange is
// store trip + canary poll + TMA round trip + scan, without the MEMBAR.ALL.GPU of a release (1,300 cycles) and the
// counter's own round trip.  Pure LSU polling of the whole tile was measured first and is 2x SLOWER than the
// counter protocol (17 B/clk per SM against TMA's 125: profiles/timeline_fwdx_r02_ll_lsu.txt).
// LL = false (default) is the release / acquire counter protocol.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "sm100.cuh"
#include "lstm_cell.cuh"

namespace rcnn {
namespace {

using namespace sm100;

constexpr int NS = 64;    // sequences per work item (UMMA N)
constexpr int GR = 128;   // gate rows per CTA
constexpr int LK = 64;
constexpr uint32_t kWTile = GR * LK * 2;   // 16 KB
constexpr uint32_t kHBox = NS * LK * 2;    //  8 KB
constexpr int kThreads = 352;              // warp 0 TMA, warp 1 MMA, warps 2-9 cell update, warp 10 publisher
#ifndef RCNN_FWD_WARP_RELEASE
#define RCNN_FWD_WARP_RELEASE 0
#endif
// Experiment (compile with -DRCNN_FWD_WARP_RELEASE=1): every cell warp releases the group counter itself after its h stores
// instead of handing off to the publisher thread (whose mbarrier wake-up is ~190 cycles of every step).  Measured SLOWER:
// 248 against 219 us per inference launch at B = 256 (step 6,400 against 5,450 cycles) -- four MEMBAR.ALL.GPU per half
// and step instead of one, each ~1,400 cycles in which the warp cannot start on the next accumulator.
constexpr bool kWarpRelease = RCNN_FWD_WARP_RELEASE != 0;

struct FxParams {
    int B, T, H, I;
    int nitems, ngroups;
    const float *bias;    // [2*4H] b_ih + b_hh, packed order
    float *csave;         // [2, T, B, H] (training only)
    __half *gsave;        // [2, T, B, 4H] activated gates, packed order (training only)
    __nv_bfloat16 *hcat;  // [B, T, 2H]
    unsigned int *sync;   // [ngroups][slot][NH] zeroed before the launch
    int ll_delay;           // (LL) cycles between the last canary and the TMA fetch (RCNN_LL_DELAY, default 0)
    unsigned int *refetch;  // (LL, optional) counts warp-level re-fetches of packets the TMA fetch overtook
    long long *tl;        // debug timeline (CTA 0, half 0) or nullptr
};
#define TLX(k) do { if (tl) tl[(s) * 8 + (k)] = clock64(); } while (0)

// release / acquire at gpu scope around the TMA-stored h_t: see lstm_fwd.cu
__device__ __forceinline__ void red_release_gpu_inc_x(unsigned int *p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_x(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_counter_x(const unsigned int *p, unsigned int target) {
    if (ld_acquire_gpu_x(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu_x(p) < target) {
        if (clock64() - t0 > 4000000000LL) {
            printf("rcnn-ocr_b200: lstm_fwdx group counter timed out (block %d)\n", blockIdx.x);
            __trap();
        }
    }
}

// flag-in-data exchange: relaxed gpu-scope 16-byte accesses (L1 is bypassed; L2 is the point of coherence)
__device__ __forceinline__ uint4 ld_relaxed_gpu_v4(const void *p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_relaxed_gpu_v4(void *p, const uint4 &v) {
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// true when none of the eight bf16 elements of the packet is the sentinel 0xFFFF
__device__ __forceinline__ bool packet_valid(const uint4 &v) {
    return (__vcmpeq2(v.x, 0xffffffffu) | __vcmpeq2(v.y, 0xffffffffu) | __vcmpeq2(v.z, 0xffffffffu) | __vcmpeq2(v.w, 0xffffffffu)) == 0u;
}

// NH = 2: the item's 64 sequences are two HALVES of 32 with independent dependency chains -- own h tile, own
// accumulator columns, own counter; the cell warps were split that way already (4 warps per 32 sequences).  The
// halves fall half a step apart, so the exchange latency of one (release, counter propagation, TMA load of h) hides
// behind the MMAs and the cell phase of the other.  The x half of a step is still one N = 64 product.
// HL = log2(H / 64): compile-time tile geometry for the LL validation pass (ignored when !LL)
//
// NSLOT = 2 (needs NH = 2, counter protocol): the group works on TWO ITEMS at once (slot 0: item i, slot 1: item
// i + ngroups -- same direction, same weights), four independent chains q = 2 slot + half per CTA.  Chosen by the
// host when a batch has more items than the GPU has groups (B > 256 at H = 512): a group's items used to run back
// to back, each waiting two thirds of a step for its exchange; now one slot's MMAs and cell phase run inside the
// other's exchange.  Per slot: own accumulator columns ([256 + 128 slot, +128): tensor memory is exactly full), own
// counters and barriers, own c state in the cell warps' registers.  Shared: W_hh / W_ih, the cell warps (step s of
// slot 0, then step s of slot 1), the x ring, and the h tile of a half -- shared memory has no room for a second
// one, so the TMA load of slot 1's h_{t-1} waits until the MMAs that read slot 0's have completed (h_free), and the
// other way round.  One publisher warp per half (a release blocks its thread for ~1,300 cycles).
template <bool SAVE, int NH, bool LL, int HL, int NSLOT>
__global__ void __launch_bounds__(NSLOT == 2 ? kThreads + 32 : kThreads, 1)
lstm_fwdx_kernel(const __grid_constant__ CUtensorMap tmWh, const __grid_constant__ CUtensorMap tmWi,
                 const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmX,
                 const FxParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int H = p.H, T = p.T, I = p.I;
    const int nkc = H / LK, nki = I / LK;
    const int nkw = nkc > nki ? nkc : nki;        // tiles of the weight area (W_hh is staged there on its way to TMEM)
    constexpr int HS = NS / NH;                   // sequences per half
    constexpr uint32_t kHalfBox = HS * LK * 2;    // [HS seq x 64 k] bf16
    const int cpb = NH == 1 ? (nkc >= 4 ? nkc / 4 : 1) : (nkc >= 2 ? nkc / 2 : 1);   // h chunks per TMA operation / barrier
    const int nhb = nkc / cpb;
    const int xcs = nki >= 2 ? 2 : 1;             // x chunks (K = 64 each) per ring slot
    const int nxs = nki / xcs;                    // ring slots consumed per step
    const int gsize = H / 32;
    unsigned char *w_s = smem;                            // nkw tiles [128 gate rows x 64 k] bf16, SW128: W_ih (resident)
    unsigned char *h_s = w_s + (size_t)nkw * kWTile;      // NH x nkc boxes [HS seq x 64 k]
    unsigned char *x_s = h_s + (size_t)nkc * kHBox;       // 2 ring slots of xcs boxes [64 seq x 64 k]
    uint64_t *bars = reinterpret_cast<uint64_t *>(x_s + 2 * (size_t)xcs * kHBox);
    uint64_t *wh_full = bars, *wcp_done = bars + 1, *wi_full = bars + 2;
    uint64_t *h_full = bars + 3;                          // [2 halves][4]
    uint64_t *x_full = h_full + 8, *x_empty = x_full + 2; // [2] each
    uint64_t *tmem_full = x_empty + 2;                    // [2 halves][2]: accumulator columns of (half, step parity)
    uint64_t *h_staged = tmem_full + 4;                   // [2 halves]: h_t of the half is in hcat (its cell warps)
    uint64_t *h_land = h_staged + 2;                      // [2 halves][4] (LL): the TMA-fetched tile landed, not yet validated
    // two-slot extras: the barriers of slot 1's chains and the "h tile of half hf may be overwritten" barriers
    uint64_t *tmem_full1 = h_land + 8;                    // [2 halves][2]
    uint64_t *h_staged1 = tmem_full1 + 4;                 // [2 halves]
    uint64_t *h_free = h_staged1 + 2;                     // [2 halves]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(h_free + 2);
    static_assert(NSLOT == 1 || (NSLOT == 2 && NH == 2 && !LL), "two slots: two halves each, counter protocol");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x / gsize;
    const int c = blockIdx.x % gsize;
    unsigned int *counter = p.sync + (size_t)group * NSLOT * NH;   // [slot][half]
    long long *tl = (p.tl && blockIdx.x == 0) ? p.tl : nullptr;
    const int rstride = NSLOT * p.ngroups;                // items a group advances by per round
    auto tfull = [&](int slot, int i) { return slot ? &tmem_full1[i] : &tmem_full[i]; };     // i = half * 2 + parity
    auto hstaged = [&](int slot, int hf) { return slot ? &h_staged1[hf] : &h_staged[hf]; };

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(wh_full, 1); mbar_init(wcp_done, 1); mbar_init(wi_full, 1);
            for (int i = 0; i < 8; ++i) { mbar_init(&h_full[i], LL ? 8 / NH : 1); mbar_init(&h_land[i], 1); }   // LL: h_full = one arrival per validating warp
            for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); mbar_init(&h_staged[i], 8 / NH); }
            for (int i = 0; i < 4; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_full1[i], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&h_staged1[i], 8 / NH); mbar_init(&h_free[i], 1); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_slot);   // [0, 256): W_hh slice (A operand); [256 + 128 slot + 64 parity, +64): the accumulators
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer ========================================================================
        if (elect_one()) {
            tma_prefetch_desc(&tmWh); tma_prefetch_desc(&tmWi); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmX);
            int cur_dir = -1;
            unsigned int xn = 0, wn = 0, published = 0;
            unsigned int nload[2] = {0u, 0u};            // (two slots) loads issued into the h tile of half hf
            for (int it0 = group; it0 < p.nitems; it0 += rstride) {
                const int dir = it0 & 1;
                const int nsl = (NSLOT == 2 && it0 + p.ngroups < p.nitems) ? 2 : 1;
                if (dir != cur_dir) {
                    if (NSLOT == 2 && cur_dir != -1) {
                        printf("rcnn-ocr_b200: lstm_fwdx two-slot groups must keep their direction (block %d)\n", blockIdx.x);
                        __trap();
                    }
                    // W_hh slice -> weight area -> (MMA thread) tensor memory; then W_ih takes the area for good.
                    // (On a direction change the previous item's last x-half MMAs must have read the old W_ih.)
                    if (xn > 0) mbar_wait(&x_empty[(xn - 1) & 1], ((xn - 1) >> 1) & 1);
                    mbar_arrive_expect_tx(wh_full, (uint32_t)nkc * kWTile);
                    for (int kc = 0; kc < nkc; ++kc)
                        tma_load_2d(w_s + (size_t)kc * kWTile, &tmWh, wh_full, kc * LK, dir * 4 * H + c * GR);
                    mbar_wait(wcp_done, wn & 1);
                    mbar_arrive_expect_tx(wi_full, (uint32_t)nki * kWTile);
                    for (int kc = 0; kc < nki; ++kc)
                        tma_load_2d(w_s + (size_t)kc * kWTile, &tmWi, wi_full, kc * LK, dir * 4 * H + c * GR);
                    ++wn;
                    cur_dir = dir;
                }
                auto load_x = [&](int slot, int xs) {       // x_t of step xs: nxs ring operations of xcs K-chunks each
                    const int t = dir ? T - 1 - xs : xs;
                    const int b0 = ((it0 + slot * p.ngroups) >> 1) * NS;
                    for (int j = 0; j < nxs; ++j, ++xn) {
                        const int rs = xn & 1;
                        mbar_wait(&x_empty[rs], ((xn >> 1) & 1) ^ 1);
                        mbar_arrive_expect_tx(&x_full[rs], (uint32_t)xcs * kHBox);
                        tma_load_4d(x_s + (size_t)rs * xcs * kHBox, &tmX, &x_full[rs], 0, b0, j * xcs, t);
                    }
                };
                for (int slot = 0; slot < nsl; ++slot) load_x(slot, 0);
                for (int s = 0; s < T; ++s) {
                    for (int slot = 0; slot < nsl; ++slot) {
                        if (!LL && s > 0) {   // (LL: the cell warps ingest h_{t-1} themselves)
                            const int t = dir ? T - 1 - s : s;
                            const int tprev = dir ? t + 1 : t - 1;
                            const int b0 = ((it0 + slot * p.ngroups) >> 1) * NS;
#pragma unroll
                            for (int hf = 0; hf < NH; ++hf) {
                                wait_counter_x(counter + slot * NH + hf,
                                               (published + (unsigned)s) * (unsigned)gsize * (kWarpRelease && !LL ? (unsigned)(8 / NH) : 1u));
                                if (slot == 0 && hf == 0) TLX(0);                     // P0: half 0's counter seen
                                if (NSLOT == 2) {
                                    // the half's tile is shared by the slots: the MMAs that read its previous contents are done
                                    if (nload[hf] > 0) mbar_wait(&h_free[hf], (nload[hf] - 1) & 1);
                                    ++nload[hf];
                                }
                                fence_proxy_async_global();
                                for (int g = 0; g < nhb; ++g) {
                                    mbar_arrive_expect_tx(&h_full[hf * 4 + g], (uint32_t)cpb * kHalfBox);
                                    tma_load_4d(h_s + (size_t)(hf * nkc + g * cpb) * kHalfBox, &tmH, &h_full[hf * 4 + g], 0, b0 + hf * HS,
                                                dir * nkc + g * cpb, tprev);
                                }
                            }
                        }
                        if (s + 1 < T) load_x(slot, s + 1);
                    }
                }
                published += (unsigned)T;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) ======================================================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(GR, NS), idesc_h = make_idesc_bf16(GR, HS);
            int cur_dir = -1;
            unsigned int xn = 0, wn = 0, nst = 0;
            uint32_t hph[2] = {0u, 0u};                  // phase of h_full[half][*] (two slots: two loads per step)
            int prev_nsl = 0;
            for (int it0 = group; it0 < p.nitems; it0 += rstride) {
                const int dir = it0 & 1;
                const int nsl = (NSLOT == 2 && it0 + p.ngroups < p.nitems) ? 2 : 1;
                if (nst > 0) {
                    // the cell warps have read the last accumulators of the previous item (they arrive on h_staged
                    // after their tcgen05.ld): the x halves below may overwrite them
                    for (int slot = 0; slot < prev_nsl; ++slot)
                        for (int hf = 0; hf < NH; ++hf) mbar_wait(hstaged(slot, hf), (nst - 1) & 1);
                    tc_fence_after();
                }
                prev_nsl = nsl;
                if (dir != cur_dir) {
                    mbar_wait(wh_full, wn & 1);
                    tc_fence_after();
                    for (int kc = 0; kc < nkc; ++kc) {
                        const uint64_t wdesc = make_smem_desc_sw128(smem_u32(w_s + (size_t)kc * kWTile), 16, 1024);
#pragma unroll
                        for (int k = 0; k < LK / 16; ++k)
                            tmem_cp_128x256b(tmem_base + (uint32_t)(8 * (kc * (LK / 16) + k)), wdesc + (uint64_t)(2 * k));
                    }
                    umma_commit(wcp_done);            // the copies have read the staging area
                    mbar_wait(wi_full, wn & 1);
                    tc_fence_after();
                    ++wn;
                    cur_dir = dir;
                }
                // x half of a step: acc[slot][par] = W_ih_slice x_t^T  (first MMA overwrites)
                auto x_part = [&](int slot, int par) {
                    const uint32_t d_tmem = tmem_base + 256u + (uint32_t)(slot * 2 * NS + par * NS);
                    for (int j = 0; j < nxs; ++j, ++xn) {
                        const int rs = xn & 1;
                        mbar_wait(&x_full[rs], (xn >> 1) & 1);
                        tc_fence_after();
                        for (int jj = 0; jj < xcs; ++jj) {
                            const int kc = j * xcs + jj;
                            const uint64_t adesc = make_smem_desc_sw128(smem_u32(w_s + (size_t)kc * kWTile), 16, 1024);
                            const uint64_t bdesc = make_smem_desc_sw128(smem_u32(x_s + (size_t)(rs * xcs + jj) * kHBox), 16, 1024);
#pragma unroll
                            for (int k = 0; k < LK / 16; ++k)
                                umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (j | jj | k) != 0);
                        }
                        umma_commit(&x_empty[rs]);
                    }
                };
                for (int slot = 0; slot < nsl; ++slot) x_part(slot, 0);
                for (int s = 0; s < T; ++s) {
                    const int par = s & 1;
                    for (int slot = 0; slot < nsl; ++slot) {
#pragma unroll
                        for (int hf = 0; hf < NH; ++hf) {
                            if (s > 0) {
                                const uint32_t d_tmem = tmem_base + 256u + (uint32_t)(slot * 2 * NS + par * NS + hf * HS);
                                for (int g = 0; g < nhb; ++g) {
                                    mbar_wait(&h_full[hf * 4 + g], hph[hf]);
                                    if (slot == 0 && hf == 0 && g == 0) TLX(1);       // M0: first h box of half 0 landed
                                    tc_fence_after();
                                    for (int j = 0; j < cpb; ++j) {
                                        const int kc = g * cpb + j;
                                        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(h_s + (size_t)(hf * nkc + kc) * kHalfBox), 16, 1024);
#pragma unroll
                                        for (int k = 0; k < LK / 16; ++k)
                                            umma_bf16_ts(d_tmem, tmem_base + (uint32_t)(8 * (kc * (LK / 16) + k)), bdesc + (uint64_t)(2 * k),
                                                         idesc_h, 1u);
                                    }
                                }
                                hph[hf] ^= 1u;
                                if (NSLOT == 2) umma_commit(&h_free[hf]);   // the half's tile may take the other slot's h
                            }
                            umma_commit(tfull(slot, hf * 2 + par));
                            if (slot == 0 && hf == 0) TLX(2);                         // M1: half 0's MMAs issued
                            if (slot == 0 && hf == NH - 1) TLX(3);                    // M2: last half's MMAs issued
                        }
                        if (s + 1 < T) x_part(slot, par ^ 1);   // runs while step s is in its cell / publish / counter phases
                    }
                    ++nst;
                }
            }
        }
    } else if (warp >= 10) {
        if (LL) {
            // ===== h loader (whole warp): canary poll -> optimistic TMA load of the half's tile ==================
            // lane l watches source CTA l % gsize of half l / gsize: the first word (units 32 src, 32 src + 1) of the
            // half's first sequence, stored by lane 0 of that CTA's first cell warp of the half
            const int src = lane % gsize, hfl = lane / gsize;
            if (elect_one()) tma_prefetch_desc(&tmH);
            for (int item = group; item < p.nitems; item += p.ngroups) {
                const int dir = item & 1, b0 = (item >> 1) * NS;
                const bool watch = hfl < NH && b0 + hfl * HS < p.B;
                for (int s = 1; s < T; ++s) {
                    const int t = dir ? T - 1 - s : s;
                    const int tprev = dir ? t + 1 : t - 1;
                    const unsigned int *canary = reinterpret_cast<const unsigned int *>(
                        p.hcat + ((size_t)(b0 + hfl * HS) * T + tprev) * 2 * H + (size_t)dir * H + 32 * src);
                    // four canary words per (source, half): word 0 of the packet each of the source's four cell warps
                    // of that half stores for the half's first sequence
                    unsigned need = watch ? 0xfu : 0u;
                    long long seen_at = 0;
                    // a half without a single live sequence is never fetched (its cell warps zero the tile once)
                    bool issued[2] = {false, NH == 1 || b0 + HS >= p.B};
                    long long t0 = 0;
                    for (unsigned spins = 0; !(issued[0] && issued[1]); ++spins) {
                        if (need) {
                            unsigned int v[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if ((need >> q) & 1u)
                                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v[q]) : "l"(canary + 4 * q) : "memory");
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (((need >> q) & 1u) && __vcmpeq2(v[q], 0xffffffffu) == 0u) need &= ~(1u << q);
                            if (tl && !need) seen_at = clock64();
                        }
                        const unsigned waiting = __ballot_sync(FULL, need != 0u);
#pragma unroll
                        for (int hf = 0; hf < NH; ++hf) {
                            const unsigned hmask = ((gsize >= 32 ? 0u : (1u << gsize)) - 1u) << (hf * gsize);
                            if (!issued[hf] && (waiting & hmask) == 0u) {
                                if (p.ll_delay > 0) { const long long d0 = clock64(); while (clock64() - d0 < p.ll_delay) {} }
                                if (elect_one()) {
                                    if (hf == 0) TLX(0);             // P0: half 0's canaries seen
                                    for (int g = 0; g < nhb; ++g) {
                                        mbar_arrive_expect_tx(&h_land[hf * 4 + g], (uint32_t)cpb * kHalfBox);
                                        tma_load_4d(h_s + (size_t)(hf * nkc + g * cpb) * kHalfBox, &tmH, &h_land[hf * 4 + g], 0,
                                                    b0 + hf * HS, dir * nkc + g * cpb, tprev);
                                    }
                                }
                                __syncwarp();
                                issued[hf] = true;
                            }
                        }
                        if (spins == 4096u) t0 = clock64();
                        if (spins > 4096u && clock64() - t0 > 4000000000LL) {
                            if (lane == 0) printf("rcnn-ocr_b200: lstm_fwdx h exchange timed out (block %d, step %d)\n", blockIdx.x, s);
                            __trap();
                        }
                    }
                    if (tl) tl[(size_t)T * 8 + (size_t)s * 32 + lane] = seen_at;   // per-source arrival times (debug timeline)
                }
            }
        }
        // ===== publisher (one elected thread): h_t of a half was stored to hcat by its cell warps, which arrived on
        // h_staged; ONE gpu-scope release (cumulative over what the barrier ordered before it) makes it visible.
        // One slot: warp 10 serves both halves.  Two slots: warp 10 + hf serves half hf of both slots in turn.
        if (!LL && !kWarpRelease && elect_one()) {
            uint32_t sphase = 0;
            const int hf_lo = NSLOT == 2 ? warp - 10 : 0, hf_hi = NSLOT == 2 ? warp - 9 : NH;
            for (int it0 = group; it0 < p.nitems; it0 += rstride) {
                const int nsl = (NSLOT == 2 && it0 + p.ngroups < p.nitems) ? 2 : 1;
                for (int s = 0; s < T; ++s) {
                    for (int slot = 0; slot < nsl; ++slot)
                        for (int hf = hf_lo; hf < hf_hi; ++hf) {
                            mbar_wait(hstaged(slot, hf), sphase);
                            if (slot == 0 && hf == 0) TLX(6);                         // R0: half 0 stored by its cell warps
                            red_release_gpu_inc_x(counter + slot * NH + hf);
                            if (slot == 0 && hf == 0) TLX(7);                         // R1: half 0 released
                        }
                    sphase ^= 1;
                }
            }
        }
    } else {
        // ===== cell update ==========================================================================
        const int qd = warp & 3;
        const int ch = (warp - 2) >> 2;
        const int hf = NH == 2 ? ch : 0;          // the half this warp's 32 sequences belong to
        const int r = qd * 32 + lane;
        const CellLane CL(lane);
        unsigned int use[2] = {0u, 0u};           // completions of tmem_full[parity] consumed so far
        uint32_t lphase = 0;                      // (LL) phase of h_land
        const int g = lane & 3, ul = lane >> 2;
        for (int it0 = group; it0 < p.nitems; it0 += rstride) {
            const int dir = it0 & 1;
            const int nsl = (NSLOT == 2 && it0 + p.ngroups < p.nitems) ? 2 : 1;
            const float bias = p.bias[(size_t)dir * 4 * H + (size_t)c * GR + r];
            float cstate[NSLOT][8];
#pragma unroll
            for (int sl = 0; sl < NSLOT; ++sl)
#pragma unroll
                for (int k = 0; k < 8; ++k) cstate[sl][k] = 0.f;
            for (int s = 0; s < T; ++s) {
                const int par = s & 1;
#pragma unroll
              for (int slot = 0; slot < NSLOT; ++slot) {
                if (slot >= nsl) continue;
                const int b0 = ((it0 + slot * p.ngroups) >> 1) * NS;
                float (&cst)[8] = cstate[slot];
                uint32_t acc[32];
                mbar_wait(tfull(slot, hf * 2 + par), use[par] & 1);
                if (threadIdx.x == 64 && slot == 0) TLX(4);          // E0: half 0's accumulator complete
                tc_fence_after();
                tmem_ld_32x32(tmem_base + ((uint32_t)(qd * 32) << 16) + 256u + (uint32_t)(slot * 2 * NS + par * NS + ch * 32), acc);
                tmem_ld_wait();
                float pre[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) pre[i] = __uint_as_float(acc[i]) + bias;
                cell_activate(pre, CL);
                uint32_t gsv[16];   // (training) activated gates as fp16 pairs, staged after the publish
                if (SAVE) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const __half2 h2 = __floats2half2_rn(pre[2 * i], pre[2 * i + 1]);
                        gsv[i] = *reinterpret_cast<const uint32_t *>(&h2);
                    }
                }
                const uint4 hq = cell_update(pre, cst, CL);
                {   // h_t: 8 units (16 bytes) of sequence 32ch + lane, straight to hcat[b, t, dir*H + 32c + 8qd ..]
                    const int t = dir ? T - 1 - s : s;
                    const int b = b0 + 32 * ch + lane;
                    if (b < p.B) {
                        __nv_bfloat16 *dst = p.hcat + ((size_t)b * T + t) * 2 * H + (size_t)dir * H + 32 * c + 8 * qd;
                        if (LL) st_relaxed_gpu_v4(dst, hq); else *reinterpret_cast<uint4 *>(dst) = hq;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(hstaged(slot, hf));                  // (the MMA thread: the accumulators have been read)
                    if (kWarpRelease && !LL) {
                        if (threadIdx.x == 64 && slot == 0) TLX(6);
                        red_release_gpu_inc_x(counter + slot * NH + hf);   // cumulative over the warp's h stores (__syncwarp)
                        if (threadIdx.x == 64 && slot == 0) TLX(7);
                    }
                }
                if (threadIdx.x == 64 && slot == 0) TLX(5);          // E1: cell phase done, h_t stored
                if (SAVE) {
                    // gates_save [2, T, B, 4H]: this thread holds gate row r for 32 sequences; a lane pair swaps halves so
                    // that the even lane writes rows (r, r+1) of sequence 2i and the odd lane those of sequence 2i+1.
                    // Neither this nor c_t is needed before the backward pass.
                    const int t = dir ? T - 1 - s : s;
                    const int odd = lane & 1;
                    __half *grow = p.gsave + (((size_t)dir * T + t) * p.B + b0 + 32 * ch + odd) * (size_t)(4 * H) + c * GR + (r & ~1);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t pv = __shfl_xor_sync(FULL, gsv[i], 1);
                        const uint32_t out = odd ? ((pv >> 16) | (gsv[i] & 0xffff0000u)) : ((gsv[i] & 0xffffu) | (pv << 16));
                        if (b0 + 32 * ch + 2 * i + odd < p.B) *reinterpret_cast<uint32_t *>(grow + (size_t)(2 * i) * (4 * H)) = out;
                    }
                    float *crow = p.csave + (((size_t)dir * T + t) * p.B) * H + 32 * c + 8 * qd + ul;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int bc = b0 + 32 * ch + 4 * k + g;
                        if (bc < p.B) crow[(size_t)bc * H] = cst[k];
                    }
                }
                if (LL && s + 1 < T) {
                    // ---- validate the TMA-fetched h_t tile of this half for step s+1 (see the header comment) ----------
                    // warp qd of the half checks sequences 8qd .. 8qd+7.  Per barrier group (CPB chunks of 64 units) a lane
                    // reads piece lane%8 (16 bytes) of rows lane/8 and lane/8 + 4 of every chunk: 2*CPB packets, all at
                    // compile-time offsets from two swizzled base addresses
                    constexpr int NKC = 1 << HL, CPB = NKC >= 2 ? NKC / 2 : 1, NHB = NKC / CPB;
                    const int t = dir ? T - 1 - s : s;
                    const int piece = lane & 7, rsub = lane >> 3;
                    unsigned char *hs = h_s + (size_t)(hf * NKC) * kHalfBox + (size_t)(8 * qd + rsub) * 128;
                    const uint32_t sbase[2] = {smem_u32(hs) + (uint32_t)((piece ^ rsub) << 4),
                                               smem_u32(hs) + 4u * 128u + (uint32_t)((piece ^ (rsub + 4)) << 4)};
                    const bool half_live = b0 + 32 * hf < p.B;
                    if (!half_live && s == 0) {      // no live sequence in this half: zero operand rows, once per item
#pragma unroll
                        for (int i = 0; i < 2 * NKC; ++i)
                            asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(sbase[i & 1] + (uint32_t)(i >> 1) * kHalfBox), "r"(0u) : "memory");
                        fence_proxy_async_smem();
                    }
#pragma unroll
                    for (int g2 = 0; g2 < NHB; ++g2) {
                        if (!half_live) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&h_full[hf * 4 + g2]);
                            continue;
                        }
                        mbar_wait(&h_land[hf * 4 + g2], lphase);
                        if (threadIdx.x == 64 && g2 == 0) TLX(6);    // R0: half 0's first TMA box landed
                        uint4 v[2 * CPB];
#pragma unroll
                        for (int i = 0; i < 2 * CPB; ++i)
                            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[i].x), "=r"(v[i].y), "=r"(v[i].z), "=r"(v[i].w)
                                         : "r"(sbase[i & 1] + (uint32_t)(g2 * CPB + (i >> 1)) * kHalfBox) : "memory");
                        unsigned pend = 0u;
#pragma unroll
                        for (int i = 0; i < 2 * CPB; ++i)
                            if (!packet_valid(v[i])) pend |= 1u << i;
                        if (__any_sync(FULL, pend != 0u)) {
                            // rare: packets the fetch overtook -- poll them from global memory until they are there
                            const unsigned char *grow = reinterpret_cast<const unsigned char *>(p.hcat + (size_t)t * 2 * H + (size_t)dir * H)
                                                        + (size_t)(b0 + 32 * hf + 8 * qd + rsub) * ((size_t)T * 2 * H * 2) + (size_t)piece * 16;
                            long long t0 = 0;
                            for (unsigned spins = 0;; ++spins) {
#pragma unroll
                                for (int i = 0; i < 2 * CPB; ++i)
                                    if ((pend >> i) & 1u) {
                                        const uint4 w = ld_relaxed_gpu_v4(grow + (size_t)(4 * (i & 1)) * ((size_t)T * 2 * H * 2)
                                                                          + (size_t)(g2 * CPB + (i >> 1)) * 128);
                                        if (packet_valid(w)) {
                                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sbase[i & 1] + (uint32_t)(g2 * CPB + (i >> 1)) * kHalfBox),
                                                         "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
                                            pend &= ~(1u << i);
                                        }
                                    }
                                if (!__any_sync(FULL, pend != 0u)) break;
                                if (spins == 1024u) t0 = clock64();
                                if (spins > 1024u && clock64() - t0 > 4000000000LL) {
                                    printf("rcnn-ocr_b200: lstm_fwdx h packet never arrived (block %d, step %d)\n", blockIdx.x, s);
                                    __trap();
                                }
                            }
                            fence_proxy_async_smem();
                            if (p.refetch && lane == 0) atomicAdd(p.refetch, 1u);
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&h_full[hf * 4 + g2]);
                        if (threadIdx.x == 64 && g2 == 0) TLX(7);    // R1: warp 2's share of half 0's first box validated
                    }
                    if (half_live) lphase ^= 1;
                }
              }
                ++use[par];
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

size_t fwdx_smem_bytes(int H, int I, bool save) {
    (void)save;
    const int nkc = H / LK, nki = I / LK, nkw = nkc > nki ? nkc : nki, xcs = nki >= 2 ? 2 : 1;
    return 1024 + (size_t)nkw * kWTile + (size_t)nkc * kHBox + 2 * (size_t)xcs * kHBox + 512;
}

}  // namespace

// Groups of H/32 CTAs and work items per group at a time.  The forward kernel stays with ONE item per group: the
// two-slot variant (RCNN_FWD_SLOTS=2; parity-tested) measures 432 us per launch at B = 512 against 439 us for two items
// back to back -- the tensor pipe is the shared resource here (128 small MMAs per item and step, 4,100 of a 5,450-cycle
// step) and the strictly ordered producer / MMA threads serialise the two slots' chains; rec-before-x ordering made it
// 452 us.  The backward kernel (11 % tensor pipe) is where two slots pay: lstm_bwd.cu.
void fwdx_plan(int B, int H, int *nslot, int *ngroups) {
    static const int force_slots = getenv("RCNN_FWD_SLOTS") ? atoi(getenv("RCNN_FWD_SLOTS")) : 1;
    static const bool one_chain = getenv("RCNN_FWD_HALVES") && atoi(getenv("RCNN_FWD_HALVES")) == 1;
    static const bool ll_env = getenv("RCNN_EXCHANGE") && strcmp(getenv("RCNN_EXCHANGE"), "ll") == 0;
    const int gsize = H / 32, nitems = 2 * ((B + NS - 1) / NS);
    const int max_groups = num_sms() / gsize;
    if (force_slots != 1 && !ll_env && !one_chain && nitems > max_groups && max_groups >= 2) {
        const int even_max = max_groups & ~1;
        int need = (nitems + 1) / 2;
        need += need & 1;                            // even: slot 1 = item + ngroups has the direction of slot 0
        *nslot = 2;
        *ngroups = need < even_max ? need : even_max;
        return;
    }
    int ng = nitems < max_groups ? nitems : max_groups;
    if (ng > 1 && (ng & 1) && nitems > ng) --ng;
    *nslot = 1;
    *ngroups = ng;
}

}  // namespace rcnn

extern "C" int rcnn_lstm_forward_fused(const void *x, const void *wih_p, const float *bias_p, const void *whh_p, int B, int T,
                                       int I, int H, void *hcat, void *gates_save, float *c_save, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 0, "lstm_forward_fused: bad shape B=%d T=%d", B, T);
    RCNN_CHECK_ARG(H == 64 || H == 128 || H == 256 || H == 512,
                   "lstm_forward_fused: hidden size %d unsupported (64, 128, 256 or 512)", H);
    RCNN_CHECK_ARG(I >= 64 && I % 64 == 0 && I <= 512, "lstm_forward_fused: input size %d unsupported (multiple of 64, <= 512)", I);
    if (B == 0 || T == 0) return RCNN_OK;
    RCNN_CHECK_ARG(x && wih_p && bias_p && whh_p && hcat, "lstm_forward_fused: null pointer");
    RCNN_CHECK_ARG((gates_save == nullptr) == (c_save == nullptr), "lstm_forward_fused: gates_save and c_save go together");
    const bool save = gates_save != nullptr;
    CUtensorMap twh, twi, th, tx;
    int rc = make_tmap_2d(&twh, whh_p, 2, 8ull * H, (uint64_t)H, (uint64_t)H * 2, GR, LK, 1);
    if (rc) return rc;
    rc = make_tmap_2d(&twi, wih_p, 2, 8ull * H, (uint64_t)I, (uint64_t)I * 2, GR, LK, 1);
    if (rc) return rc;
    static const int halves = getenv("RCNN_FWD_HALVES") ? (atoi(getenv("RCNN_FWD_HALVES")) == 1 ? 1 : 2) : 2;
    {   // hcat [B, T, 2H] as (k in chunk, b, chunk, t): box = cpb chunks of [64 / halves seq x 64 k]
        const int nkc = H / LK, cpb = halves == 1 ? (nkc >= 4 ? nkc / 4 : 1) : (nkc >= 2 ? nkc / 2 : 1);
        const uint64_t dims[4] = {(uint64_t)LK, (uint64_t)B, 2ull * nkc, (uint64_t)T};
        const uint64_t strides[3] = {(uint64_t)T * 2 * H * 2, (uint64_t)LK * 2, 2ull * H * 2};
        const uint32_t box[4] = {(uint32_t)LK, (uint32_t)(NS / halves), (uint32_t)cpb, 1u};
        rc = make_tmap_4d(&th, hcat, 2, dims, strides, box, 1);
        if (rc) return rc;
    }
    {   // x [B, T, I] bf16 the same way: box = xcs chunks
        const int nki = I / LK, xcs = nki >= 2 ? 2 : 1;
        const uint64_t dims[4] = {(uint64_t)LK, (uint64_t)B, (uint64_t)nki, (uint64_t)T};
        const uint64_t strides[3] = {(uint64_t)T * I * 2, (uint64_t)LK * 2, (uint64_t)I * 2};
        const uint32_t box[4] = {(uint32_t)LK, (uint32_t)NS, (uint32_t)xcs, 1u};
        rc = make_tmap_4d(&tx, x, 2, dims, strides, box, 1);
        if (rc) return rc;
    }
    if (rc) return rc;
    FxParams p;
    p.B = B; p.T = T; p.H = H; p.I = I;
    p.bias = bias_p;
    p.csave = c_save;
    p.hcat = (__nv_bfloat16 *)hcat;
    p.tl = debug_timeline();
    p.refetch = debug_refetch();
    static const int ll_delay = getenv("RCNN_LL_DELAY") ? atoi(getenv("RCNN_LL_DELAY")) : 0;
    p.ll_delay = ll_delay;
    p.gsave = (__half *)gates_save;
    const int gsize = H / 32;
    p.nitems = 2 * ((B + NS - 1) / NS);
    static const bool ll_env = getenv("RCNN_EXCHANGE") && strcmp(getenv("RCNN_EXCHANGE"), "ll") == 0;
    const bool ll = ll_env && halves == 2;
    int nslot = 1;
    fwdx_plan(B, H, &nslot, &p.ngroups);
    cudaStream_t s = (cudaStream_t)stream;
    // exchange protocol: flag-in-data (default; needs the two-halves layout) or the release / acquire counter
    // (measured: the counter protocol is faster -- profiles/ll_exchange_r02.txt; "ll" is kept as a tested experiment)
    if (ll) {
        // every element of hcat that the kernel will poll starts as the sentinel (bf16 0xFFFF)
        RCNN_CUDA(cudaMemsetAsync(hcat, 0xFF, (size_t)B * T * 2 * H * sizeof(__nv_bfloat16), s));
        p.sync = nullptr;
    } else {
        p.sync = group_counters(p.ngroups * nslot * halves, s);
        if (!p.sync) return RCNN_ERR_CUDA_BASE;
    }
    const size_t smem = fwdx_smem_bytes(H, I, save);
    using KernT = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const FxParams);
    KernT kern;
    if (ll) {
        static const KernT tab[2][4] = {
            {lstm_fwdx_kernel<false, 2, true, 0, 1>, lstm_fwdx_kernel<false, 2, true, 1, 1>, lstm_fwdx_kernel<false, 2, true, 2, 1>, lstm_fwdx_kernel<false, 2, true, 3, 1>},
            {lstm_fwdx_kernel<true, 2, true, 0, 1>, lstm_fwdx_kernel<true, 2, true, 1, 1>, lstm_fwdx_kernel<true, 2, true, 2, 1>, lstm_fwdx_kernel<true, 2, true, 3, 1>}};
        kern = tab[save ? 1 : 0][H == 64 ? 0 : H == 128 ? 1 : H == 256 ? 2 : 3];
    } else {
        kern = nslot == 2 ? (save ? lstm_fwdx_kernel<true, 2, false, 0, 2> : lstm_fwdx_kernel<false, 2, false, 0, 2>)
               : save     ? (halves == 2 ? lstm_fwdx_kernel<true, 2, false, 0, 1> : lstm_fwdx_kernel<true, 1, false, 0, 1>)
                          : (halves == 2 ? lstm_fwdx_kernel<false, 2, false, 0, 1> : lstm_fwdx_kernel<false, 1, false, 0, 1>);
    }
    RCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(gsize * p.ngroups));
    cfg.blockDim = dim3(nslot == 2 ? kThreads + 32 : kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ProfScope prof(RCNN_K_LSTM_FWD, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, kern, twh, twi, th, tx, p));
    count_launch();
    return RCNN_OK;
}
