// K3: CTC loss forward-backward fused with log_softmax and its backward.
//
// The reference has no CTC code; this implements the
// nn.CTCLoss(blank, reduction, zero_infinity) call the north_star places at
// training/train.py:289,503-505.  torch runs log_softmax, ctc_loss, ctc_loss_backward and
// log_softmax_backward as four passes over [T,N,C]; here ONE launch reads the logits and
// writes the gradient w.r.t. the logits (algorithmic traffic 2*T*C*4 bytes per sequence).
//
// One CTA per sequence.  All arithmetic is float32 in the log2 domain (ex2/lg2 are single
// MUFU ops):
//   phase 1 (all warps, frame-parallel): per-frame logsumexp, then gather the
//            log-probabilities of the extended label sequence l' (blank,l1,blank,...) into a
//            [T][S] lattice in shared memory.  Each lattice row is shifted by its maximum
//            u_t so that alpha/beta stay small in magnitude (the shifts are summed exactly
//            in float64 and only enter the loss, never the gradient).
//   phase 2 (warp 0: alpha forward in time, warp 1: beta backward in time, concurrently):
//            each lane owns K consecutive lattice states in registers; the s-1 / s-2
//            neighbours of a lane's first states come from the previous lane by warp
//            shuffle, so a time step needs no block barrier.  Both tables stay in shared
//            memory (they spill to the caller's workspace only when T*S is too large).
//   phase 3 (all warps, frame-parallel): occupancy_t(s) = normalised exp2(alpha+beta-lp),
//            scattered per class into a per-warp row accumulator, then
//            grad[t,c] = scale_n * (softmax_t(c) - occupancy_t(c)), coalesced stores.
#include <math.h>
#include "common.cuh"

namespace rcnn {

namespace {

constexpr int kWarps = 16;   // frame-parallel phases 1 and 3 are latency-bound: 16 warps per sequence (8: 49 us, 16: 35 us, 32: 44 us at N = 256)
constexpr int kThreads = kWarps * 32;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr size_t kSmemBudget = 200 * 1024;

// log2(2^a + 2^b + 2^c), safe for -inf operands
__device__ __forceinline__ float lse3_2(float a, float b, float c) {
    const float m = fmaxf(fmaxf(a, b), c);
    const float mm = (m == -INFINITY) ? 0.f : m;
    return mm + lg2(ex2(a - mm) + ex2(b - mm) + ex2(c - mm));
}
__device__ __forceinline__ float lse2_2(float a, float b) {
    const float m = fmaxf(a, b);
    const float mm = (m == -INFINITY) ? 0.f : m;
    return mm + lg2(ex2(a - mm) + ex2(b - mm));
}

struct CtcParams {
    const float *x;
    int from_logits, T, N, C;
    long long st, sn;
    const long long *targets;
    long long tgt_stride;
    const long long *tgt_offsets;  // exclusive prefix sums (concatenated targets only)
    const long long *in_len, *tg_len;
    int blank, reduction, zero_inf;
    float *nll, *grad;
    long long gst, gsn;
    int SP;            // lattice row stride (floats) >= 2*max_target_len+1
    float *gtab;       // global spill for the three [T][SP] tables, or nullptr (shared memory)
    long long *tl;     // debug: clock64 at the phase boundaries of CTA 0 (rcnn_debug_timeline) or nullptr
};

// alpha recursion: lane owns states s = lane*K + k.
template <int K>
__device__ void alpha_pass(const float *__restrict__ lpe, float *__restrict__ alpha, const int *ext,
                           int S, int SP, int Tn, int lane) {
    float a[K], l[K];
    unsigned skip = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int s = lane * K + k;
        const bool ok = (s & 1) && s >= 2 && s < S && ext[s] != ext[s - 2];
        skip |= (ok ? 1u : 0u) << k;
        a[k] = (s < 2 && s < S) ? lpe[s] : -INFINITY;
        if (s < S) alpha[s] = a[k];
    }
    if (Tn > 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int s = lane * K + k;
            l[k] = (s < S) ? lpe[SP + s] : -INFINITY;
        }
    }
    for (int t = 1; t < Tn; ++t) {
        float ln[K];
        if (t + 1 < Tn) {  // prefetch the next frame's lattice row off the dependent chain
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int s = lane * K + k;
                ln[k] = (s < S) ? lpe[(size_t)(t + 1) * SP + s] : -INFINITY;
            }
        }
        float p1 = __shfl_up_sync(FULL, a[K - 1], 1);
        float p2 = (K >= 2) ? __shfl_up_sync(FULL, a[K >= 2 ? K - 2 : 0], 1) : __shfl_up_sync(FULL, a[0], 2);
        if (lane == 0) { p1 = -INFINITY; p2 = -INFINITY; }
        if (K == 1 && lane == 1) p2 = -INFINITY;
        float n[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float x0 = a[k];
            const float x1 = (k >= 1) ? a[k >= 1 ? k - 1 : 0] : p1;
            float x2 = (k >= 2) ? a[k >= 2 ? k - 2 : 0] : (k == 1 ? p1 : p2);
            x2 = ((skip >> k) & 1u) ? x2 : -INFINITY;
            n[k] = lse3_2(x0, x1, x2) + l[k];
        }
        float *arow = alpha + (size_t)t * SP;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int s = lane * K + k;
            a[k] = n[k];
            if (s < S) arow[s] = a[k];
            l[k] = ln[k];
        }
    }
}

// beta recursion, mirrored: neighbours s+1 / s+2 come from the next lane.
template <int K>
__device__ void beta_pass(const float *__restrict__ lpe, float *__restrict__ beta, const int *ext,
                          int S, int SP, int Tn, int lane) {
    float b[K], l[K];
    unsigned skip = 0;
    const float *last = lpe + (size_t)(Tn - 1) * SP;
    float *brow = beta + (size_t)(Tn - 1) * SP;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int s = lane * K + k;
        const bool ok = (s & 1) && s + 2 < S && ext[s] != ext[s + 2];
        skip |= (ok ? 1u : 0u) << k;
        b[k] = (s < S && s >= S - 2) ? last[s] : -INFINITY;
        if (s < S) brow[s] = b[k];
    }
    if (Tn > 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int s = lane * K + k;
            l[k] = (s < S) ? lpe[(size_t)(Tn - 2) * SP + s] : -INFINITY;
        }
    }
    for (int t = Tn - 2; t >= 0; --t) {
        float ln[K];
        if (t > 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int s = lane * K + k;
                ln[k] = (s < S) ? lpe[(size_t)(t - 1) * SP + s] : -INFINITY;
            }
        }
        float n1 = __shfl_down_sync(FULL, b[0], 1);
        float n2 = (K >= 2) ? __shfl_down_sync(FULL, b[K >= 2 ? 1 : 0], 1) : __shfl_down_sync(FULL, b[0], 2);
        if (lane == 31) { n1 = -INFINITY; n2 = -INFINITY; }
        if (K == 1 && lane == 30) n2 = -INFINITY;
        float n[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float x0 = b[k];
            const float x1 = (k + 1 < K) ? b[k + 1 < K ? k + 1 : 0] : n1;
            float x2 = (k + 2 < K) ? b[k + 2 < K ? k + 2 : 0] : (k + 2 == K ? n1 : n2);
            x2 = ((skip >> k) & 1u) ? x2 : -INFINITY;
            n[k] = lse3_2(x0, x1, x2) + l[k];
        }
        brow = beta + (size_t)t * SP;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int s = lane * K + k;
            b[k] = n[k];
            if (s < S) brow[s] = b[k];
            l[k] = ln[k];
        }
    }
}

// CV = ceil(C / 32) when C <= 256 (a frame's row lives in CV registers per lane and the rows of up to FB frames are
// requested before the first is used: the frame-parallel phases were bound by one L2 / HBM round trip per load, 7
// dependent ones per pass over a row of 195 classes); CV = 0: any C, rows re-read in loops.
constexpr int FB = 2;        // frames in flight per warp (rows in registers: FB x CV values; 64 registers per thread keep 2 CTAs per SM)
template <int K, int CV>
__global__ void __launch_bounds__(kThreads, 2)
ctc_kernel(const CtcParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = p.T, C = p.C, SP = p.SP;

    // shared layout: ext[SP] | link[SP] | occs[kWarps][SP] | lse[T] | ush[T] | acc[kWarps][C] | misc | tables
    int *ext = reinterpret_cast<int *>(smem_raw);
    int *link = ext + SP;      // odd s: next lattice state with the same label (0 = none); sign bit set = not its first occurrence
    float *occs = reinterpret_cast<float *>(link + SP);
    float *lse = occs + kWarps * SP;
    float *ush = lse + T;
    float *acc = ush + T;
    float *misc = acc + kWarps * C;  // [0] = nll (natural log), [1] = status flag
    float *tab = p.gtab ? p.gtab + (size_t)n * 3 * T * SP : misc + 4;
    float *lpe = tab, *alpha = tab + (size_t)T * SP, *beta = tab + 2 * (size_t)T * SP;

    if (p.tl && n == 0 && tid == 0) p.tl[0] = clock64();
    long long Tn_ll = p.in_len[n], L_ll = p.tg_len[n];
    const bool bad_len = Tn_ll < 0 || Tn_ll > T || L_ll < 0 || 2 * L_ll + 1 > SP;
    const int Tn = bad_len ? 0 : (int)Tn_ll;
    const int L = bad_len ? 0 : (int)L_ll;
    const int S = 2 * L + 1;
    const long long *tg = p.targets + (p.tgt_stride > 0 ? (long long)n * p.tgt_stride : p.tgt_offsets[n]);
    const float *xs = p.x + (long long)n * p.sn;

    // request every row of the sequence now (one 128-byte line per lane): the setup below and the first frames overlap
    // the HBM latency, and the second batch of frames finds its rows in L2
    for (int t = warp; t < Tn; t += kWarps)
        if (lane * 32 < C) prefetch_l2(xs + (long long)t * p.st + lane * 32);
    int bad_label = 0;
    for (int s = tid; s < S; s += kThreads) {
        int lab = p.blank;
        if (s & 1) {
            const long long v = tg[s >> 1];
            if (v < 0 || v >= C) bad_label = 1; else lab = (int)v;
        }
        ext[s] = lab;
    }
    for (int i = tid; i < kWarps * C; i += kThreads) acc[i] = 0.f;
    const int any_bad = __syncthreads_or(bad_label) || bad_len;
    // Occurrences of the same label are chained so that phase 3 sums a class's occupancy WITHOUT shared-memory float
    // atomics (ATOMS.CAST.SPIN compare-and-swap loops: they were 2/3 of phase 3): the first occurrence walks the chain.
    for (int s2 = 2 * tid + 1; s2 < S; s2 += 2 * kThreads) {
        const int lab = ext[s2];
        int nx = 0, notfirst = 0;
        for (int q = s2 + 2; q < S; q += 2) if (ext[q] == lab) { nx = q; break; }
        for (int q = 1; q < s2; q += 2) if (ext[q] == lab) { notfirst = 1; break; }
        link[s2] = nx | (notfirst ? (int)0x80000000 : 0);
    }

    // ---- phase 1: per-frame logsumexp and lattice gather --------------------------------
    if (CV > 0) {
        for (int tb = warp; tb < Tn; tb += kWarps * FB) {
            float xv[FB][CV > 0 ? CV : 1];
#pragma unroll
            for (int f = 0; f < FB; ++f) {
                const int t = tb + f * kWarps;
                const float *row = xs + (long long)t * p.st;
#pragma unroll
                for (int j = 0; j < CV; ++j) {
                    const int c = lane + 32 * j;
                    xv[f][j] = (t < Tn && c < C) ? __ldg(row + c) : -INFINITY;
                }
            }
#pragma unroll
            for (int f = 0; f < FB; ++f) {
                const int t = tb + f * kWarps;
                if (t >= Tn) break;
                const float *row = xs + (long long)t * p.st;
                float z = 0.f;
                if (p.from_logits) {
                    float m = -INFINITY;
#pragma unroll
                    for (int j = 0; j < CV; ++j) m = fmaxf(m, xv[f][j]);
                    m = warp_max(m);
                    float sum = 0.f;
#pragma unroll
                    for (int j = 0; j < CV; ++j) sum += (lane + 32 * j < C) ? ex2((xv[f][j] - m) * kLog2e) : 0.f;
                    sum = warp_sum(sum);
                    z = m * kLog2e + lg2(sum);
                }
                float g[K], u = -INFINITY;
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int s = lane + 32 * j;
                    g[j] = -INFINITY;
                    if (32 * j < S) {      // (uniform over the CTA: slots beyond this sequence's lattice cost no issue slots)
                        g[j] = (s < S) ? __ldg(row + ext[s]) * kLog2e - z : -INFINITY;   // (L1 hit: the row was just read)
                        u = fmaxf(u, g[j]);
                    }
                }
                u = warp_max(u);
                if (u == -INFINITY) u = 0.f;
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int s = lane + 32 * j;
                    if (s < S) lpe[(size_t)t * SP + s] = g[j] - u;
                }
                if (lane == 0) { lse[t] = z; ush[t] = u; }
            }
        }
    } else
    for (int t = warp; t < Tn; t += kWarps) {
        const float *row = xs + (long long)t * p.st;
        float z = 0.f;
        if (p.from_logits) {
            float m = -INFINITY;
            for (int c = lane; c < C; c += 32) m = fmaxf(m, __ldg(row + c));
            m = warp_max(m);
            float sum = 0.f;
            for (int c = lane; c < C; c += 32) sum += ex2((__ldg(row + c) - m) * kLog2e);
            sum = warp_sum(sum);
            z = m * kLog2e + lg2(sum);
        }
        float g[K], u = -INFINITY;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int s = lane + 32 * j;
            g[j] = -INFINITY;
            if (32 * j < S) {
                g[j] = (s < S) ? __ldg(row + ext[s]) * kLog2e - z : -INFINITY;
                u = fmaxf(u, g[j]);
            }
        }
        u = warp_max(u);
        if (u == -INFINITY) u = 0.f;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int s = lane + 32 * j;
            if (s < S) lpe[(size_t)t * SP + s] = g[j] - u;
        }
        if (lane == 0) { lse[t] = z; ush[t] = u; }
    }
    __syncthreads();
    if (p.tl && n == 0 && tid == 0) p.tl[1] = clock64();

    // ---- phase 2: alpha (warp 0) and beta (warp 1) concurrently ---------------------------
    const bool want_grad = p.grad != nullptr;
    if (Tn > 0 && !any_bad) {
        if (warp == 0) {
            // states per lane chosen per SEQUENCE (S = 2 L + 1 <= 32: one; <= 64: two; ...): a short label sequence pays for
            // one logsumexp per lane and step instead of the K the longest label of the batch needs
            if (K >= 3 && S <= 32) alpha_pass<1>(lpe, alpha, ext, S, SP, Tn, lane);
            else if (K >= 3 && S <= 64) alpha_pass<2>(lpe, alpha, ext, S, SP, Tn, lane);
            else if (K == 2 && S <= 32) alpha_pass<1>(lpe, alpha, ext, S, SP, Tn, lane);
            else alpha_pass<K>(lpe, alpha, ext, S, SP, Tn, lane);
            __syncwarp();
            // log-likelihood: the two terminal states, plus the exact sum of the row shifts
            double us = 0.0;
            for (int t = lane; t < Tn; t += 32) us += (double)ush[t];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) us += __shfl_xor_sync(FULL, us, o);
            if (lane == 0) {
                const float *last = alpha + (size_t)(Tn - 1) * SP;
                const float ll2 = (S > 1) ? lse2_2(last[S - 1], last[S - 2]) : last[0];
                misc[0] = (float)(-((double)ll2 + us) * (double)kLn2);
            }
        } else if (warp == 1 && want_grad) {
            if (K >= 3 && S <= 32) beta_pass<1>(lpe, beta, ext, S, SP, Tn, lane);
            else if (K >= 3 && S <= 64) beta_pass<2>(lpe, beta, ext, S, SP, Tn, lane);
            else if (K == 2 && S <= 32) beta_pass<1>(lpe, beta, ext, S, SP, Tn, lane);
            else beta_pass<K>(lpe, beta, ext, S, SP, Tn, lane);
        }
    } else if (tid == 0) {
        misc[0] = any_bad ? NAN : (L == 0 ? 0.f : INFINITY);
    }
    __syncthreads();
    if (p.tl && n == 0 && tid == 0) p.tl[2] = clock64();

    const float nll = misc[0];
    const bool infeasible = isinf(nll);
    if (tid == 0) p.nll[n] = (infeasible && p.zero_inf) ? 0.f : nll;
    if (!want_grad) return;

    // ---- phase 3: gradient rows ----------------------------------------------------------
    float scale = 1.f;
    if (p.reduction == RCNN_REDUCE_MEAN) scale = 1.f / ((float)p.N * (float)max(L, 1));
    float *gs = p.grad + (long long)n * p.gsn;
    float *wacc = acc + warp * C;
    float *wocc = occs + warp * SP;
    for (int tb = warp; tb < T; tb += kWarps * FB) {
        // the rows of the FB frames this warp handles next are requested before the first one is needed
        float xv[FB][CV > 0 ? CV : 1];
        if (CV > 0 && !(infeasible || any_bad)) {
#pragma unroll
            for (int f = 0; f < FB; ++f) {
                const int t = tb + f * kWarps;
                const float *row = xs + (long long)t * p.st;
#pragma unroll
                for (int j = 0; j < CV; ++j) {
                    const int c = lane + 32 * j;
                    xv[f][j] = (t < Tn && c < C) ? __ldg(row + c) : 0.f;
                }
            }
        }
#pragma unroll
        for (int f = 0; f < FB; ++f) {
            const int t = tb + f * kWarps;
            if (t >= T) break;
            float *grow = gs + (long long)t * p.gst;
            if (t >= Tn) {
                for (int c = lane; c < C; c += 32) grow[c] = 0.f;
                continue;
            }
            if (infeasible || any_bad) {
                const float v = (infeasible && p.zero_inf) ? 0.f : NAN;
                for (int c = lane; c < C; c += 32) grow[c] = v;
                continue;
            }
            // occupancy of every lattice state at frame t, normalised over s (sums to 1)
            float v[K], mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int s = lane + 32 * j;
                float w = -INFINITY;
                if (32 * j < S) {          // (uniform over the CTA)
                    if (s < S) {
                        const float lp = lpe[(size_t)t * SP + s];
                        if (lp != -INFINITY) w = alpha[(size_t)t * SP + s] + beta[(size_t)t * SP + s] - lp;
                    }
                    mx = fmaxf(mx, w);
                }
                v[j] = w;
            }
            mx = warp_max(mx);
            float zsum = 0.f, blank_sum = 0.f;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (32 * j < S) {
                    v[j] = (v[j] == -INFINITY) ? 0.f : ex2(v[j] - mx);
                    zsum += v[j];
                } else {
                    v[j] = 0.f;
                }
            }
            zsum = warp_sum(zsum);
            const float inv = 1.f / zsum;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int s = lane + 32 * j;
                if (32 * j < S) {
                    v[j] *= inv;
                    if (s < S) {
                        if (s & 1) wocc[s] = v[j];
                        else blank_sum += v[j];
                    }
                }
            }
            blank_sum = warp_sum(blank_sum);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int s = lane + 32 * j;
                if (32 * j < S && s < S && (s & 1)) {
                    const int lk = link[s];
                    if (lk >= 0) {                       // first occurrence of its label: sum the chain
                        float tot = v[j];
                        for (int q = lk; q != 0; q = link[q] & 0x7fffffff) tot += wocc[q];
                        wacc[ext[s]] = tot;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) wacc[p.blank] += blank_sum;   // (a label equal to the blank class lands here too)
            __syncwarp();
            const float z = lse[t];
            if (CV > 0) {
#pragma unroll
                for (int j = 0; j < CV; ++j) {
                    const int c = lane + 32 * j;
                    if (c < C) {
                        const float sm = ex2(xv[f][j] * kLog2e - z);
                        grow[c] = (sm - wacc[c]) * scale;
                        wacc[c] = 0.f;
                    }
                }
            } else {
                const float *row = xs + (long long)t * p.st;
                for (int c = lane; c < C; c += 32) {
                    const float sm = ex2(__ldg(row + c) * kLog2e - z);
                    grow[c] = (sm - wacc[c]) * scale;
                    wacc[c] = 0.f;
                }
            }
            __syncwarp();
        }
    }
    if (p.tl && n == 0 && tid == 0) p.tl[3] = clock64();
}

// exclusive prefix sum of target_lengths (concatenated-target form), one CTA
__global__ void ctc_offsets_kernel(const long long *tg_len, long long *off, int N) {
    __shared__ long long carry;
    __shared__ long long wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < N; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const long long v = i < N ? max(tg_len[i], 0LL) : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w += y;
            }
            wsum[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const long long before = carry + (warp ? wsum[warp - 1] : 0) + inc - v;
        if (i < N) off[i] = before;
        __syncthreads();
        if (threadIdx.x == 0) carry += wsum[(blockDim.x >> 5) - 1];
        __syncthreads();
    }
}

// loss = sum_n w_n * nll_n, fixed summation order (deterministic), one CTA
__global__ void ctc_reduce_kernel(const float *nll, const long long *tg_len, int N, int reduction,
                                  float *loss) {
    __shared__ double part[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double w = 1.0;
        if (reduction == RCNN_REDUCE_MEAN) w = 1.0 / (double)max(tg_len[i], 1LL);
        s += w * (double)nll[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
        if (reduction == RCNN_REDUCE_MEAN) tot = N > 0 ? tot / (double)N : NAN;
        *loss = (float)tot;
    }
}

__global__ void ctc_scale_kernel(float *grad, int T, int N, int C, long long gst, long long gsn,
                                 const float *scale, int per_sample) {
    const int n = blockIdx.x;
    const float f = scale[per_sample ? n : 0];
    if (f == 1.0f) return;
    for (int t = blockIdx.y; t < T; t += gridDim.y) {
        float *row = grad + (long long)t * gst + (long long)n * gsn;
        for (int c = threadIdx.x; c < C; c += blockDim.x) row[c] *= f;
    }
}

size_t table_floats(int T, int SP) { return 3 * (size_t)T * SP; }

size_t smem_fixed_bytes(int T, int C, int SP) {
    return sizeof(int) * 2 * (size_t)SP + sizeof(float) * ((size_t)kWarps * SP + 2 * (size_t)T + (size_t)kWarps * C + 4);
}

int lattice_stride(int max_target_len) { return 2 * max_target_len + 1; }

template <int K, int CV>
int launch_ctc_cv(const CtcParams &p, size_t smem, cudaStream_t s) {
    if (smem > 48 * 1024)
        RCNN_CUDA(cudaFuncSetAttribute(ctc_kernel<K, CV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    ProfScope prof(RCNN_K_CTC, s);
    ctc_kernel<K, CV><<<p.N, kThreads, smem, s>>>(p);
    RCNN_LAUNCH_CHECK("ctc_kernel");
    return RCNN_OK;
}

template <int K>
int launch_ctc(const CtcParams &p, size_t smem, cudaStream_t s) {
    // register-resident rows for the usual alphabets (C <= 256: 4 or 8 values per lane), loops otherwise
    if (K <= 4 && p.C <= 128) return launch_ctc_cv<K, 4>(p, smem, s);
    if (K <= 4 && p.C <= 256) return launch_ctc_cv<K, 8>(p, smem, s);
    return launch_ctc_cv<K, 0>(p, smem, s);
}

}  // namespace
}  // namespace rcnn

extern "C" size_t rcnn_ctc_workspace_bytes(int T, int N, int C, int max_target_len) {
    using namespace rcnn;
    if (T < 0 || N < 0 || C <= 0 || max_target_len < 0) return 0;
    const int SP = lattice_stride(max_target_len);
    size_t bytes = sizeof(long long) * (size_t)(N + 1);  // concatenated-target offsets
    bytes = (bytes + 255) & ~(size_t)255;
    if (smem_fixed_bytes(T, C, SP) + sizeof(float) * table_floats(T, SP) > kSmemBudget)
        bytes += sizeof(float) * table_floats(T, SP) * (size_t)N;
    return bytes;
}

extern "C" int rcnn_ctc_loss(const float *x, int from_logits, int T, int N, int C,
                             int64_t stride_t, int64_t stride_n,
                             const int64_t *targets, int64_t tgt_stride,
                             const int64_t *input_lengths, const int64_t *target_lengths,
                             int max_target_len, int blank, int reduction, int zero_infinity,
                             float *nll_out, float *loss_out,
                             float *grad_out, int64_t gstride_t, int64_t gstride_n,
                             void *workspace, size_t workspace_bytes, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(T >= 0 && N >= 0 && C > 0, "ctc_loss: bad shape T=%d N=%d C=%d", T, N, C);
    RCNN_CHECK_ARG(blank >= 0 && blank < C, "ctc_loss: blank=%d outside [0,%d)", blank, C);
    RCNN_CHECK_ARG(reduction >= RCNN_REDUCE_NONE && reduction <= RCNN_REDUCE_SUM, "ctc_loss: bad reduction %d", reduction);
    RCNN_CHECK_ARG(max_target_len >= 0, "ctc_loss: max_target_len=%d", max_target_len);
    RCNN_CHECK_ARG(tgt_stride >= 0, "ctc_loss: tgt_stride must be >= 0");
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0) {
        if (reduction != RCNN_REDUCE_NONE && loss_out) {
            // torch: mean over an empty batch is NaN, sum is 0
            const float v = reduction == RCNN_REDUCE_MEAN ? NAN : 0.f;
            RCNN_CUDA(cudaMemcpyAsync(loss_out, &v, sizeof(float), cudaMemcpyHostToDevice, s));
        }
        return RCNN_OK;
    }
    RCNN_CHECK_ARG(x && targets && input_lengths && target_lengths && nll_out, "ctc_loss: null pointer");
    RCNN_CHECK_ARG(reduction == RCNN_REDUCE_NONE || loss_out, "ctc_loss: loss_out is required for mean/sum");
    const int SP = lattice_stride(max_target_len);
    const int K = (SP + 31) / 32;
    RCNN_CHECK_ARG(K <= 16, "ctc_loss: max_target_len=%d exceeds the supported 255", max_target_len);
    const size_t need = rcnn_ctc_workspace_bytes(T, N, C, max_target_len);
    if (!workspace || workspace_bytes < need) {
        set_error("ctc_loss: workspace of %zu bytes required, got %zu", need, workspace_bytes);
        return RCNN_ERR_WORKSPACE;
    }
    CtcParams p;
    p.x = x; p.from_logits = from_logits; p.T = T; p.N = N; p.C = C;
    p.st = stride_t; p.sn = stride_n;
    p.targets = (const long long *)targets; p.tgt_stride = tgt_stride;
    p.in_len = (const long long *)input_lengths; p.tg_len = (const long long *)target_lengths;
    p.blank = blank; p.reduction = reduction; p.zero_inf = zero_infinity;
    p.nll = nll_out; p.grad = grad_out; p.gst = gstride_t; p.gsn = gstride_n;
    p.SP = SP;
    p.tl = debug_timeline();
    long long *offs = (long long *)workspace;
    p.tgt_offsets = offs;
    if (tgt_stride == 0) {
        ctc_offsets_kernel<<<1, 1024, 0, s>>>(p.tg_len, offs, N);
        RCNN_LAUNCH_CHECK("ctc_offsets_kernel");
    }
    size_t smem = smem_fixed_bytes(T, C, SP);
    const size_t off_bytes = (sizeof(long long) * (size_t)(N + 1) + 255) & ~(size_t)255;
    if (smem + sizeof(float) * table_floats(T, SP) > kSmemBudget) {
        p.gtab = (float *)((char *)workspace + off_bytes);
    } else {
        p.gtab = nullptr;
        smem += sizeof(float) * table_floats(T, SP);
    }
    RCNN_CHECK_ARG(smem <= kSmemBudget, "ctc_loss: T=%d C=%d do not fit shared memory", T, C);
    int rc;
    if (K <= 1) rc = launch_ctc<1>(p, smem, s);
    else if (K <= 2) rc = launch_ctc<2>(p, smem, s);
    else if (K <= 3) rc = launch_ctc<3>(p, smem, s);
    else if (K <= 4) rc = launch_ctc<4>(p, smem, s);
    else if (K <= 6) rc = launch_ctc<6>(p, smem, s);
    else if (K <= 8) rc = launch_ctc<8>(p, smem, s);
    else if (K <= 12) rc = launch_ctc<12>(p, smem, s);
    else rc = launch_ctc<16>(p, smem, s);
    if (rc) return rc;
    if (reduction != RCNN_REDUCE_NONE) {
        ctc_reduce_kernel<<<1, 256, 0, s>>>(nll_out, p.tg_len, N, reduction, loss_out);
        RCNN_LAUNCH_CHECK("ctc_reduce_kernel");
    }
    return RCNN_OK;
}

extern "C" int rcnn_ctc_scale_grad(float *grad, int T, int N, int C, int64_t gstride_t, int64_t gstride_n,
                                   const float *scale, int per_sample, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(T >= 0 && N >= 0 && C > 0, "ctc_scale_grad: bad shape");
    if (T == 0 || N == 0) return RCNN_OK;
    RCNN_CHECK_ARG(grad && scale, "ctc_scale_grad: null pointer");
    dim3 grid(N, min(T, 8));
    ctc_scale_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(grad, T, N, C, gstride_t, gstride_n, scale, per_sample);
    RCNN_LAUNCH_CHECK("ctc_scale_kernel");
    return RCNN_OK;
}
