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
// Same decomposition as the forward kernel: a GROUP of H/32 CTAs owns one work item = (direction,
// tile of 64 sequences) at a time, CTA c owns hidden units [32c, 32c+32); groups synchronise per step
// through a counter in global memory (red.release / relaxed poll) and the kernel is launched
// cooperatively.  Here the resident operand is the CTA's 32 x 4H slice of W_hh^T (bf16, 128 KB for
// H=512) and the streamed operand is the full dG_{t'} tile [64 seq, 4H]: one TMA operation brings
// four K chunks (a 32 KB slot of a 3-slot ring, 4-D tensor map), accumulated by tcgen05.mma (M = 64)
// into a 64 x 32 fp32 TMEM tile.  The cell threads (16 rows per TMEM lane quadrant, the two half-warps
// of a warp split the columns) keep dc in registers for the whole sequence, read the saved gates /
// cell states / upstream dh directly from global memory (issued before the MMA wait) and write dG_t
// in the packed column order, which is at once the next step's MMA operand and the operand of the
// dX / dW GEMMs.
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "sm100.cuh"

namespace rcnn {
namespace {

using namespace sm100;

constexpr int LB = 64;    // sequences per work item (UMMA M)
constexpr int LU = 32;    // hidden units per CTA (UMMA N)
constexpr int LK = 64;
constexpr int kARing = 3;
constexpr int kChunks = 4;                   // K chunks per TMA operation / ring slot
constexpr uint32_t kABox = LB * LK * 2;      // [64 seq x 64 k] bf16 = 8 KB
constexpr uint32_t kATile = kChunks * kABox; // 32 KB ring slot
constexpr uint32_t kWTile = LU * LK * 2;     // 4 KB
constexpr int kThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 cell update

struct BwdParams {
    int B, T, H;
    int nitems, ngroups;
    const __half *gates;       // [2, T, B, 4H] activated gates, packed order
    const float *csave;        // [2, T, B, H]
    const float *dhcat;        // [B, T, 2H] upstream gradient of the block's LSTM output
    __nv_bfloat16 *dG;         // [B, T, 2*4H] out: pre-activation gate gradients, packed order
    float *db;                 // [2*4H] out or nullptr: column sums of dG (bias gradient), zeroed before the launch
    unsigned int *sync;        // [ngroups] zeroed before the launch
    long long *tl;             // debug timeline or nullptr
};
#define TL_MARK(k) do { if (tl) tl[(s) * 8 + (k)] = clock64(); } while (0)

__device__ __forceinline__ float tanh_fast_b(float x) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void red_release_gpu_inc_b(unsigned int *p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu_b(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_counter_b(const unsigned int *p, unsigned int target) {
    if (ld_relaxed_gpu_b(p) >= target) return;
    const long long t0 = clock64();
    while (ld_relaxed_gpu_b(p) < target) {
        if (clock64() - t0 > 4000000000LL) {
            printf("rcnn-ocr_b200: lstm_bwd group counter timed out (block %d)\n", blockIdx.x);
            __trap();
        }
    }
}

struct ChunkIn {     // saved state of 8 units of one sequence at one step
    U8 g[2];         // 32 halves: (i,f,g,o) x 8 units
    U8 c, cp, dh;    // 8 floats each
};

__device__ __forceinline__ void load_chunk(ChunkIn &ci, const __half *grow, const float *crow, const float *cprow,
                                           const float *dhrow, bool valid) {
    U8 z;
#pragma unroll
    for (int k = 0; k < 8; ++k) z.v[k] = 0u;
    if (valid) {
        ci.g[0] = ld_ro_v8(grow);
        ci.g[1] = ld_ro_v8(grow + 16);
        ci.c = ld_ro_v8(crow);
        ci.dh = ld_ro_v8(dhrow);
        ci.cp = cprow ? ld_ro_v8(cprow) : z;
    } else {
        ci.g[0] = z; ci.g[1] = z; ci.c = z; ci.cp = z; ci.dh = z;
    }
}

__global__ void __launch_bounds__(kThreads, 1)
lstm_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG, const BwdParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int H = p.H, T = p.T, B = p.B;
    const int nkc = 4 * H / LK;
    const int nslots = nkc / kChunks;          // ring slots consumed per step
    const int gsize = H / 32;
    unsigned char *w_s = smem;                          // nkc tiles [32 n x 64 k] bf16, SW128
    unsigned char *a_s = w_s + (size_t)nkc * kWTile;    // kARing slots of kChunks boxes [64 seq x 64 k] bf16, SW128
    uint64_t *bars = reinterpret_cast<uint64_t *>(a_s + kARing * kATile);
    uint64_t *w_full = bars;
    uint64_t *a_full = bars + 1, *a_empty = a_full + kARing;
    uint64_t *tmem_full = a_empty + kARing;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.x / gsize;
    const int c = blockIdx.x % gsize;
    long long *tl = (p.tl && blockIdx.x == 0) ? p.tl : nullptr;
    unsigned int *counter = p.sync + group;

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(w_full, 1);
            for (int i = 0; i < kARing; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
            mbar_init(tmem_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<LU>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // processing order: the forward direction is back-propagated from t = T-1 down to 0, the
    // reverse direction from t = 0 up to T-1.
    if (warp == 0) {
        // ===== TMA producer (one elected thread) =================================================
        if (elect_one()) {
            tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmG);
            int cur_dir = -1;
            int pslot = 0;
            uint32_t pphase = 0;
            unsigned int published = 0;
            for (int item = group; item < p.nitems; item += p.ngroups) {
                const int dir = item & 1, b0 = (item >> 1) * LB;
                if (dir != cur_dir) {
                    mbar_arrive_expect_tx(w_full, (uint32_t)nkc * kWTile);
                    for (int kc = 0; kc < nkc; ++kc)
                        tma_load_2d(w_s + (size_t)kc * kWTile, &tmW, w_full, kc * LK, dir * H + c * LU);
                    cur_dir = dir;
                }
                for (int s = 1; s < T; ++s) {
                    const int t = dir ? s : T - 1 - s;
                    const int tsrc = dir ? t - 1 : t + 1;   // step processed just before
                    TL_MARK(7);
                    wait_counter_b(counter, (published + (unsigned)s) * (unsigned)gsize);
                    TL_MARK(0);
                    fence_proxy_async_global();
                    for (int g = 0; g < nslots; ++g) {
                        mbar_wait(&a_empty[pslot], pphase ^ 1);
                        mbar_arrive_expect_tx(&a_full[pslot], kATile);
                        tma_load_4d(a_s + pslot * kATile, &tmG, &a_full[pslot], 0, b0, dir * nkc + g * kChunks, tsrc);
                        if (++pslot == kARing) { pslot = 0; pphase ^= 1; }
                    }
                    TL_MARK(1);
                }
                published += (unsigned)T;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) ===================================================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(LB, LU);
            int cur_dir = -1;
            int mslot = 0;
            uint32_t mphase = 0, wphase = 0;
            for (int item = group; item < p.nitems; item += p.ngroups) {
                const int dir = item & 1;
                if (dir != cur_dir) { mbar_wait(w_full, wphase); wphase ^= 1; cur_dir = dir; }
                for (int s = 1; s < T; ++s) {
                    for (int g = 0; g < nslots; ++g) {
                        mbar_wait(&a_full[mslot], mphase);
                        if (g == 0) TL_MARK(2);
                        tc_fence_after();
#pragma unroll
                        for (int j = 0; j < kChunks; ++j) {
                            const int kc = g * kChunks + j;
                            const uint64_t adesc = make_smem_desc_sw128(smem_u32(a_s + mslot * kATile + j * kABox), 16, 1024);
                            const uint64_t bdesc = make_smem_desc_sw128(smem_u32(w_s + (size_t)kc * kWTile), 16, 1024);
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
            }
        }
    } else {
        // ===== cell update ========================================================================
        // M = 64 accumulator layout: row m sits in TMEM lane 32*(m/16) + m%16, so warp quadrant qd reads
        // rows 16qd .. 16qd+15 in its lanes 0-15; the upper half-warp takes over 8 of the 16 columns.
        const int qd = warp & 3;
        const int hf = (warp - 2) >> 2;                   // which 16 of the CTA's 32 units this warp handles
        const int row = qd * 16 + (lane & 15);            // sequence within the tile
        const int u0 = hf * 16 + (lane >> 4) * 8;         // first of this thread's 8 units (within the CTA)
        unsigned int mcount = 0;
        for (int item = group; item < p.nitems; item += p.ngroups) {
            const int dir = item & 1, b0 = (item >> 1) * LB;
            const int b = b0 + row;
            const bool valid = b < B;
            float dc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) dc[i] = 0.f;
            float dbacc[32];   // this sequence's contribution to the bias gradient of the thread's 32 gate columns
#pragma unroll
            for (int i = 0; i < 32; ++i) dbacc[i] = 0.f;
            for (int s = 0; s < T; ++s) {
                const int t = dir ? s : T - 1 - s;
                const int tfp = dir ? t + 1 : t - 1;   // forward-time predecessor: where c_{prev} lives
                const size_t rtb = ((size_t)dir * T + t) * B + b;
                const __half *grow = p.gates + rtb * 4 * H + (size_t)c * 128 + u0 * 4;
                const float *crow = p.csave + rtb * H + 32 * c + u0;
                const float *cprow = (tfp >= 0 && tfp < T) ? p.csave + (((size_t)dir * T + tfp) * B + b) * H + 32 * c + u0 : nullptr;
                const float *dhrow = p.dhcat + ((size_t)b * T + t) * 2 * H + (size_t)dir * H + 32 * c + u0;
                __nv_bfloat16 *dgrow = p.dG + ((size_t)b * T + t) * 8 * H + (size_t)dir * 4 * H + (size_t)c * 128 + u0 * 4;

                ChunkIn ci;
                load_chunk(ci, grow, crow, cprow, dhrow, valid);
                if (valid && s + 1 < T) {   // pull the next step's saved rows from HBM into L2 while the MMA runs
                    const int tn = dir ? t + 1 : t - 1;
                    const size_t rn = ((size_t)dir * T + tn) * B + b;
                    prefetch_l2(p.gates + rn * 4 * H + (size_t)c * 128 + u0 * 4);
                    prefetch_l2(p.csave + rn * H + 32 * c + u0);
                    prefetch_l2(p.dhcat + ((size_t)b * T + tn) * 2 * H + (size_t)dir * H + 32 * c + u0);
                }
                float rec[8];   // dG_{t'} W_hh for this thread's 8 units
                if (s > 0) {
                    uint32_t acc[16];
                    mbar_wait(tmem_full, mcount & 1);
                    ++mcount;
                    if (threadIdx.x == 64) TL_MARK(4);
                    tc_fence_after();
                    tmem_ld_32x16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(hf * 16), acc);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t hi = __shfl_sync(FULL, acc[8 + j], lane & 15);
                        rec[j] = __uint_as_float(lane < 16 ? acc[j] : hi);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) rec[j] = 0.f;
                }
                const __half2 *gh = reinterpret_cast<const __half2 *>(ci.g);
                float o32[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 if_ = __half22float2(gh[2 * j]);       // (i, f)
                    const float2 go_ = __half22float2(gh[2 * j + 1]);   // (g, o)
                    const float ig = if_.x, fg = if_.y, gg = go_.x, og = go_.y;
                    const float dh = __uint_as_float(ci.dh.v[j]) + rec[j];
                    const float tc = tanh_fast_b(__uint_as_float(ci.c.v[j]));
                    const float d_o = dh * tc;
                    const float dct = fmaf(dh * og, 1.f - tc * tc, dc[j]);
                    dc[j] = dct * fg;
                    o32[4 * j] = dct * gg * ig * (1.f - ig);
                    o32[4 * j + 1] = dct * __uint_as_float(ci.cp.v[j]) * fg * (1.f - fg);
                    o32[4 * j + 2] = dct * ig * (1.f - gg * gg);
                    o32[4 * j + 3] = d_o * og * (1.f - og);
                }
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) dbacc[i] += o32[i];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        U8 v;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(o32[16 * j + 2 * k], o32[16 * j + 2 * k + 1]);
                            v.v[k] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        st_v8(dgrow + j * 16, v);
                    }
                }
                if (threadIdx.x == 64) TL_MARK(5);
                tc_fence_before();
                fence_proxy_async_global();   // dG_t stores before the other CTAs' TMA reads
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (threadIdx.x == 64) {      // one gpu-scope release per CTA
                    red_release_gpu_inc_b(counter);
                    TL_MARK(6);
                }
            }
            if (p.db != nullptr) {
                // db: sum over the 16 sequences of the half-warp, then one atomic per column and half-warp
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float v = dbacc[i];
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                    dbacc[i] = v;
                }
                if ((lane & 15) == 0) {
                    float *dst = p.db + (size_t)dir * 4 * H + (size_t)c * 128 + u0 * 4;
#pragma unroll
                    for (int i = 0; i < 32; ++i) atomicAdd(dst + i, dbacc[i]);
                }
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<LU>(tmem_base);
    }
}

size_t bwd_smem_bytes(int H) { return 1024 + (size_t)(4 * H / LK) * kWTile + (size_t)kARing * kATile + 256; }

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
                                  int B, int T, int H, void *dG, float *db, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 0, "lstm_backward: bad shape B=%d T=%d", B, T);
    RCNN_CHECK_ARG(H == 64 || H == 128 || H == 256 || H == 512,
                   "lstm_backward: hidden size %d unsupported (64, 128, 256 or 512)", H);
    if (db) RCNN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * 8 * (size_t)H, (cudaStream_t)stream));
    if (B == 0 || T == 0) return RCNN_OK;
    RCNN_CHECK_ARG(whh_pt && gates_save && c_save && dhcat && dG, "lstm_backward: null pointer");
    CUtensorMap tw, tg;
    int rc = make_tmap_2d(&tw, whh_pt, 2, 2ull * H, 4ull * H, 4ull * H * 2, LU, LK, 1);
    if (rc) return rc;
    {   // dG [B, T, 8H] seen as (k within chunk, b, chunk, t): one box = kChunks chunks of [64 seq x 64 k]
        const uint64_t dims[4] = {(uint64_t)LK, (uint64_t)B, 8ull * H / LK, (uint64_t)T};
        const uint64_t strides[3] = {(uint64_t)T * 8 * H * 2, (uint64_t)LK * 2, 8ull * H * 2};
        const uint32_t box[4] = {(uint32_t)LK, (uint32_t)LB, (uint32_t)kChunks, 1u};
        rc = make_tmap_4d(&tg, dG, 2, dims, strides, box, 1);
        if (rc) return rc;
    }
    BwdParams p;
    p.B = B; p.T = T; p.H = H;
    p.gates = (const __half *)gates_save;
    p.csave = c_save;
    p.dhcat = dhcat;
    p.dG = (__nv_bfloat16 *)dG;
    p.db = db;
    p.tl = debug_timeline();
    const int gsize = H / 32;
    const size_t smem = bwd_smem_bytes(H);
    cudaStream_t s = (cudaStream_t)stream;
    RCNN_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.nitems = 2 * ((B + LB - 1) / LB);
    const int max_groups = num_sms() / gsize;    // one CTA per SM (shared memory)
    p.ngroups = p.nitems < max_groups ? p.nitems : max_groups;
    if (p.ngroups > 1 && (p.ngroups & 1) && p.nitems > p.ngroups) --p.ngroups;   // even: a group keeps its direction
    p.sync = group_counters(p.ngroups, s);
    if (!p.sync) return RCNN_ERR_CUDA_BASE;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(gsize * p.ngroups));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the groups spin on each other
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ProfScope prof(RCNN_K_LSTM_BWD, s);
    RCNN_CUDA(cudaLaunchKernelEx(&cfg, lstm_bwd_kernel, tw, tg, p));
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
