// K2 (backward): persistent BPTT kernel for both directions of a BidirectionalLSTM block.
//
// Backward of the recurrence in lstm_fwd.cu (the reference gets it from autograd through
// nn.LSTM / cuDNN RNN backward, model/model.py:161):
//     dh_t   = dhcat[:, t] + dG_{t'} W_hh            (t' = the step processed just before)
//     do     = dh_t * tanh(c_t)            dc += dh_t * o * (1 - tanh(c_t)^2)
//     di = dc*g   dg = dc*i   df = dc*c_prev   dc <- dc*f
//     dG_t   = [di*i(1-i), df*f(1-f), dg*(1-g^2), do*o(1-o)]   (pre-activation gradients)
// dW_ih, dW_hh, db and dX are GEMMs / reductions over dG after the loop (host side).
//
// Same cluster decomposition as the forward kernel: CTA c owns hidden units [32c, 32c+32).
// Here the resident operand is the CTA's 32 x 4H slice of W_hh^T (bf16, 128 KB for H=512) and
// the streamed operand is the full dG_{t'} tile [128 seq, 4H] (two 64-column TMA boxes per slot of a
// 3-slot ring, so one barrier round trip covers K = 128), accumulated by tcgen05.mma into a 128 x 32 fp32 TMEM tile.  The cell
// threads (one per sequence) keep dc in registers for the whole sequence, read the saved gates
// / cell states / upstream dh directly from global memory (prefetched one 8-unit chunk ahead)
// and write dG_t in the packed column order, which is at once the next step's MMA operand and
// the operand of the dX / dW GEMMs.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "sm100.cuh"

namespace rcnn {
namespace {

using namespace sm100;

constexpr int LB = 128;   // sequences per cluster tile (UMMA M)
constexpr int LU = 32;    // hidden units per CTA (UMMA N)
constexpr int LK = 64;
constexpr int kMaxRing = 6;
constexpr uint32_t kABox = LB * LK * 2;    // one TMA box: 128 seq x 64 k bf16 = 16 KB
constexpr int kBoxes = 2;                  // boxes per ring slot: one barrier round trip per K = 128
constexpr uint32_t kATile = kBoxes * kABox; // 32 KB ring slot
constexpr uint32_t kWTile = LU * LK * 2;   // 4 KB
constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 cell update

struct BwdParams {
    int B, T, H;
    const __half *gates;       // [2, T, B, 4H] activated gates, packed order
    const float *csave;        // [2, T, B, H]
    const float *dhcat;        // [B, T, 2H] upstream gradient of the block's LSTM output
    __nv_bfloat16 *dG;         // [B, T, 2*4H] out: pre-activation gate gradients, packed order
    long long *tl;             // debug timeline or nullptr
};
#define TL_MARK(k) do { if (tl) tl[s * 8 + (k)] = clock64(); } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank_b() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x_b() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_arrive_relaxed_b() { asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_b() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
// one gpu-scope release per CTA for the cell warps' dG stores (see lstm_fwd.cu: cell_publish_arrive)
__device__ __forceinline__ void cell_publish_sync_b() {
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 64) asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    else cluster_arrive_relaxed_b();
    cluster_wait_b();
}
__device__ __forceinline__ void cluster_sync_b() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ float tanh_fast_b(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct ChunkIn {     // saved state of 8 units of one sequence at one step
    U8 g[2];         // 32 halves: (i,f,g,o) x 8 units
    U8 c, cp, dh;    // 8 floats each
};

__device__ __forceinline__ void load_chunk(ChunkIn &ci, const __half *grow, const float *crow, const float *cprow,
                                           const float *dhrow, int q, bool valid) {
    U8 z;
#pragma unroll
    for (int k = 0; k < 8; ++k) z.v[k] = 0u;
    if (valid) {
        ci.g[0] = ld_ro_v8(grow + q * 32);
        ci.g[1] = ld_ro_v8(grow + q * 32 + 16);
        ci.c = ld_ro_v8(crow + q * 8);
        ci.dh = ld_ro_v8(dhrow + q * 8);
        ci.cp = cprow ? ld_ro_v8(cprow + q * 8) : z;
    } else {
        ci.g[0] = z; ci.g[1] = z; ci.c = z; ci.cp = z; ci.dh = z;
    }
}

// TWO_SM: the cluster is 8 CTA pairs; a pair runs one cta_group::2 MMA of M = 128 sequences (64 per
// CTA) x N = 64 units (32 per CTA), so each SM streams only HALF of the dG tile (the per-SM L2 read
// port, ~40 B/clk, is what bounds this kernel).  A CTA then owns 64 sequences x the pair's 64 units.
template <int kARing, bool TWO_SM>
__global__ void __launch_bounds__(kThreads, 1)
lstm_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG, const BwdParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int H = p.H, T = p.T, B = p.B;
    const int nkc = 4 * H / LK;
    unsigned char *w_s = smem;                          // nkc tiles [32 n x 64 k] bf16, SW128
    unsigned char *a_s = w_s + (size_t)nkc * kWTile;    // kARing tiles [128 b x 64 k] bf16, SW128
    uint64_t *bars = reinterpret_cast<uint64_t *>(a_s + kARing * kATile);
    uint64_t *w_full = bars;
    uint64_t *a_full = bars + 1, *a_empty = a_full + kARing;
    uint64_t *tmem_full = a_empty + kARing;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = (int)cluster_ctarank_b();
    const int cid = (int)cluster_id_x_b();
    const int dir = cid & 1, tile = cid >> 1;
    const int b0 = tile * LB;
    long long *tl = (p.tl && blockIdx.x == 0) ? p.tl : nullptr;
    const bool leader = (c & 1) == 0;
    const uint16_t pair_mask = (uint16_t)(3u << (c & ~1));

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(w_full, 1);
            for (int i = 0; i < kARing; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
            mbar_init(tmem_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        if (TWO_SM) tmem_alloc_2sm<LU>(tmem_slot); else tmem_alloc<LU>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (TWO_SM) {
        // the resident W^T slice of BOTH CTAs must be in place, and the peer's barriers initialised,
        // before the leader issues the first pair MMA / the peer's TMA signals the leader's barrier
        if (warp == 0 && lane == 0) {
            mbar_arrive_expect_tx(w_full, (uint32_t)nkc * kWTile);
            for (int kc = 0; kc < nkc; ++kc)
                tma_load_2d(w_s + (size_t)kc * kWTile, &tmW, w_full, kc * LK, dir * H + c * LU);
        }
        if (warp == 1 && lane == 0) mbar_wait(w_full, 0);
        __syncwarp();
        cluster_sync_b();
    }

    // processing order: the forward direction is back-propagated from t = T-1 down to 0, the
    // reverse direction from t = 0 up to T-1.
    if (warp == 0) {
        if (lane == 0 && !TWO_SM) {
            tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmG);
            mbar_arrive_expect_tx(w_full, (uint32_t)nkc * kWTile);
            for (int kc = 0; kc < nkc; ++kc)
                tma_load_2d(w_s + (size_t)kc * kWTile, &tmW, w_full, kc * LK, dir * H + c * LU);
        }
        int pslot = 0;
        uint32_t pphase = 0;
        for (int s = 0; s < T; ++s) {
            if (lane == 0 && s > 0) {
                const int t = dir ? s : T - 1 - s;
                const int tsrc = dir ? t - 1 : t + 1;   // step processed just before
                TL_MARK(0);
                fence_proxy_async_global();
                for (int g = 0; g < nkc / kBoxes; ++g) {
                    mbar_wait(&a_empty[pslot], pphase ^ 1);
                    if (!TWO_SM || leader) mbar_arrive_expect_tx(&a_full[pslot], kATile);
#pragma unroll
                    for (int j = 0; j < kBoxes; ++j) {
                        const int kc = g * kBoxes + j;
                        unsigned char *dst = a_s + pslot * kATile + j * kABox;
                        if (!TWO_SM)
                            tma_load_3d(dst, &tmG, &a_full[pslot], dir * 4 * H + kc * LK, tsrc, b0);
                        else   // each CTA fetches its own 64 sequences; both halves complete on the LEADER's barrier
                            tma_load_3d_2sm(dst, &tmG, &a_full[pslot], dir * 4 * H + kc * LK, tsrc, b0 + 64 * (c & 1));
                    }
                    if (++pslot == kARing) { pslot = 0; pphase ^= 1; }
                }
                TL_MARK(1);
            }
            __syncwarp();
            cluster_arrive_relaxed_b();
            cluster_wait_b();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_bf16(LB, TWO_SM ? 2 * LU : LU);
        int mslot = 0;
        uint32_t mphase = 0;
        if (lane == 0 && !TWO_SM) mbar_wait(w_full, 0);
        __syncwarp();
        for (int s = 0; s < T; ++s) {
            if (lane == 0 && s > 0 && (!TWO_SM || leader)) {
                for (int g = 0; g < nkc / kBoxes; ++g) {
                    mbar_wait(&a_full[mslot], mphase);
                    if (g == 0) TL_MARK(2);
                    tc_fence_after();
#pragma unroll
                    for (int j = 0; j < kBoxes; ++j) {
                        const int kc = g * kBoxes + j;
                        const uint64_t adesc = make_smem_desc_sw128(smem_u32(a_s + mslot * kATile + j * kABox), 16, 1024);
                        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(w_s + (size_t)kc * kWTile), 16, 1024);
#pragma unroll
                        for (int k = 0; k < LK / 16; ++k) {
                            if (TWO_SM)
                                umma_bf16_2sm(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (g | j | k) != 0);
                            else
                                umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (g | j | k) != 0);
                        }
                    }
                    if (TWO_SM) umma_commit_2sm(&a_empty[mslot], pair_mask);   // frees the slot in both CTAs
                    else umma_commit(&a_empty[mslot]);
                    if (++mslot == kARing) { mslot = 0; mphase ^= 1; }
                }
                if (TWO_SM) umma_commit_2sm(tmem_full, pair_mask);
                else umma_commit(tmem_full);
                TL_MARK(3);
            }
            __syncwarp();
            cluster_arrive_relaxed_b();
            cluster_wait_b();
        }
    } else {
        const int qd = warp & 3;            // TMEM lane quadrant
        const int hf = (warp - 2) >> 2;     // which 16 of a 32-unit slice this warp handles
        // 1-SM: the CTA owns 128 sequences x its own 32-unit slice c.  TWO_SM: it owns 64 sequences x
        // both slices of the pair; TMEM lanes 0-63 hold slice 2p, lanes 64-127 slice 2p+1 ("2x2" layout).
        const int row = TWO_SM ? 64 * (c & 1) + (qd & 1) * 32 + lane : qd * 32 + lane;
        const int cs = TWO_SM ? (c & ~1) + (qd >> 1) : c;     // unit slice this thread works on
        const int b = b0 + row;
        const bool valid = b < B;
        float dc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) dc[i] = 0.f;
        for (int s = 0; s < T; ++s) {
            const int t = dir ? s : T - 1 - s;
            const int tfp = dir ? t + 1 : t - 1;   // forward-time predecessor: where c_{prev} lives
            const size_t rtb = ((size_t)dir * T + t) * B + b;
            const __half *grow = p.gates + rtb * 4 * H + (size_t)cs * 128;
            const float *crow = p.csave + rtb * H + 32 * cs;
            const float *cprow = (tfp >= 0 && tfp < T) ? p.csave + (((size_t)dir * T + tfp) * B + b) * H + 32 * cs : nullptr;
            const float *dhrow = p.dhcat + ((size_t)b * T + t) * 2 * H + (size_t)dir * H + 32 * cs;
            __nv_bfloat16 *dgrow = p.dG + ((size_t)b * T + t) * 8 * H + (size_t)dir * 4 * H + (size_t)cs * 128;

            ChunkIn cur, nxt;
            load_chunk(cur, grow, crow, cprow, dhrow, hf * 2, valid);
            load_chunk(nxt, grow, crow, cprow, dhrow, hf * 2 + 1, valid);
            if (valid && s + 1 < T && hf == 0) {   // pull the next step's saved rows from HBM into L2 while the MMA runs
                const int tn = dir ? t + 1 : t - 1;
                const size_t rn = ((size_t)dir * T + tn) * B + b;
                prefetch_l2(p.gates + rn * 4 * H + (size_t)cs * 128);
                prefetch_l2(p.gates + rn * 4 * H + (size_t)cs * 128 + 64);
                prefetch_l2(p.csave + rn * H + 32 * cs);
                prefetch_l2(p.dhcat + ((size_t)b * T + tn) * 2 * H + (size_t)dir * H + 32 * cs);
            }
            uint32_t acc[16];
            if (s > 0) {
                mbar_wait(tmem_full, (s - 1) & 1);
                if (threadIdx.x == 64) TL_MARK(4);
                tc_fence_after();
                tmem_ld_32x16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(hf * 16), acc);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0u;
            }
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
                const int q = hf * 2 + qq;
                const ChunkIn &ci = qq ? nxt : cur;
                const __half2 *gh = reinterpret_cast<const __half2 *>(ci.g);
                float cc[8], cp[8], dhu[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    cc[k] = __uint_as_float(ci.c.v[k]);
                    cp[k] = __uint_as_float(ci.cp.v[k]);
                    dhu[k] = __uint_as_float(ci.dh.v[k]);
                }
                float o32[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 if_ = __half22float2(gh[2 * j]);       // (i, f)
                    const float2 go_ = __half22float2(gh[2 * j + 1]);   // (g, o)
                    const float ig = if_.x, fg = if_.y, gg = go_.x, og = go_.y;
                    const float dh = dhu[j] + __uint_as_float(acc[qq * 8 + j]);
                    const float tc = tanh_fast_b(cc[j]);
                    const float d_o = dh * tc;
                    const float dct = fmaf(dh * og, 1.f - tc * tc, dc[qq * 8 + j]);
                    dc[qq * 8 + j] = dct * fg;
                    o32[4 * j] = dct * gg * ig * (1.f - ig);
                    o32[4 * j + 1] = dct * cp[j] * fg * (1.f - fg);
                    o32[4 * j + 2] = dct * ig * (1.f - gg * gg);
                    o32[4 * j + 3] = d_o * og * (1.f - og);
                }
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        U8 v;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(o32[16 * j + 2 * k], o32[16 * j + 2 * k + 1]);
                            v.v[k] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        st_v8(dgrow + q * 32 + j * 16, v);
                    }
                }
            }
            if (threadIdx.x == 64) TL_MARK(5);
            tc_fence_before();
            fence_proxy_async_global();   // dG_t stores before the other CTAs' TMA reads
            cell_publish_sync_b();
            if (threadIdx.x == 64) TL_MARK(6);
        }
    }
    if (TWO_SM) cluster_sync_b();   // no CTA leaves while its peer may still signal its barriers / use its TMEM
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (TWO_SM) tmem_dealloc_2sm<LU>(tmem_base); else tmem_dealloc<LU>(tmem_base);
    }
}

size_t bwd_smem_bytes(int H, int ring) { return 1024 + (size_t)(4 * H / LK) * kWTile + (size_t)ring * kATile + 256; }

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
        for (long long r = (long long)blockIdx.y * 8 + threadIdx.y; r < rows; r += (long long)gridDim.y * 8) {
            const uint4 v = ld_nc_v4(src + r * ld8 + c8);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc[2 * k] += __uint_as_float(w[k] << 16);
                acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
            }
        }
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
    for (int k = threadIdx.x; k < I; k += blockDim.x) a.dw_ih[dir][(size_t)r * I + k] = a.dwih_p[prow * I + k];
    for (int k = threadIdx.x; k < H; k += blockDim.x) a.dw_hh[dir][(size_t)r * H + k] = a.dwhh_p[prow * H + k];
    if (threadIdx.x == 0) {
        const float v = a.db_p[prow];
        a.db_ih[dir][r] = v;
        a.db_hh[dir][r] = v;
    }
}

}  // namespace
}  // namespace rcnn

extern "C" int rcnn_lstm_backward(const void *whh_pt, const void *gates_save, const float *c_save, const float *dhcat,
                                  int B, int T, int H, void *dG, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 0, "lstm_backward: bad shape B=%d T=%d", B, T);
    RCNN_CHECK_ARG(H == 64 || H == 128 || H == 256 || H == 512,
                   "lstm_backward: hidden size %d unsupported (64, 128, 256 or 512)", H);
    if (B == 0 || T == 0) return RCNN_OK;
    RCNN_CHECK_ARG(whh_pt && gates_save && c_save && dhcat && dG, "lstm_backward: null pointer");
    CUtensorMap tw, tg;
    int rc = make_tmap_2d(&tw, whh_pt, 2, 2ull * H, 4ull * H, 4ull * H * 2, LU, LK, 1);
    if (rc) return rc;
    static const bool two_sm = getenv("RCNN_BWD_2SM") ? atoi(getenv("RCNN_BWD_2SM")) != 0 : false;
    rc = make_tmap_3d(&tg, dG, 2, (uint64_t)B, (uint64_t)T, 8ull * H, (uint64_t)T * 8 * H * 2, 8ull * H * 2,
                      two_sm ? LB / 2 : LB, 1, LK, 1);
    if (rc) return rc;
    BwdParams p;
    p.B = B; p.T = T; p.H = H;
    p.gates = (const __half *)gates_save;
    p.csave = c_save;
    p.dhcat = dhcat;
    p.dG = (__nv_bfloat16 *)dG;
    p.tl = debug_timeline();
    const int csize = H / 32;
    const int ntiles = (B + LB - 1) / LB;
    const bool stagger = two_sm;
    const size_t smem = bwd_smem_bytes(H, 3);
    cudaStream_t s = (cudaStream_t)stream;
    auto kern = stagger ? lstm_bwd_kernel<3, true> : lstm_bwd_kernel<3, false>;
    RCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8) RCNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
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
    ProfScope prof(RCNN_K_LSTM_BWD, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, kern, tw, tg, p));
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
        const long long want = (2LL * num_sms() * 32) / cols8 + 1;   // ~2 blocks per SM in total
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
