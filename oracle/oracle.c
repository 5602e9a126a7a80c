/*
 * oracle.c -- CPU restatement of the RCNN-OCR sequence-recognition hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (rcnn-ocr_b200/) may
 * import, link or execute this file.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker
 * (or as the timed CPU baseline), never as the thing shipped.
 *
 * What it restates (citations are path:line under the upstream reference tree):
 *   - greedy CTC decode            training/utils.py:122-150 (ctc_greedy_decoder)
 *   - BidirectionalLSTM.forward    model/model.py:151-163   (nn.LSTM bidirectional,
 *                                  batch_first, gate order i,f,g,o + nn.Linear)
 *   - CTC loss + gradient          NOT in the reference tree:
 *                                  the north_star loss is torch.nn.CTCLoss, a
 *                                  third-party dependency (torch 2.11.0 here; the
 *                                  reference's requirements.txt:5-6 implies 2.0.1).
 *                                  This restates the published algorithm
 *                                  (Graves et al. 2006, eq. 6-16; ATen
 *                                  native/LossCTC.cpp conventions) in float64.
 *
 * Pinning: the reference ships no tests or golden vectors, so
 * this oracle is pinned against outputs of the reference's own Python code and of
 * torch float64, generated in the build container by tests/make_golden.py and
 * committed under tests/golden/ (see tests/test_oracle_golden.py).
 *
 * Everything is plain C, float64 arithmetic (float32 inputs are widened), single
 * threaded (the OpenMP pragmas are inert: libgomp is not in the image; the timed CPU
 * baseline fans batch slices out over host threads from Python instead).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define NEG_INF (-INFINITY)

int oracle_version(void) { return 1; }

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* greedy CTC decode                                                         */
/* ------------------------------------------------------------------------- */

/* torch.argmax tie/NaN rule: the first maximal element wins; NaN compares as
 * larger than everything, and the first NaN wins. */
static int argmax_row_f32(const float *row, int C) {
    int best = 0;
    float bv = row[0];
    if (bv != bv) return 0;
    for (int c = 1; c < C; ++c) {
        float v = row[c];
        if (v != v) return c;
        if (v > bv) { bv = v; best = c; }
    }
    return best;
}

/* training/utils.py:135-150.  logits is [B,T,C] addressed through element strides
 * (sb, st, class stride 1).  ids_out is [B,T] int32, rows padded with -1; len_out
 * is [B].  The reference's layout heuristic (utils.py:132-133) lives in the Python
 * wrapper, not here. */
int oracle_ctc_greedy_f32(const float *logits, int B, int T, int C,
                          long long sb, long long st, int blank,
                          int32_t *ids_out, int32_t *len_out) {
    if (B < 0 || T < 0 || C <= 0) return 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int b = 0; b < B; ++b) {
        int prev = blank, n = 0;
        int32_t *row = ids_out + (size_t)b * T;
        for (int t = 0; t < T; ++t) {
            int p = argmax_row_f32(logits + b * sb + t * st, C);
            if (p != blank && p != prev) row[n++] = p;
            prev = p;
        }
        for (int t = n; t < T; ++t) row[t] = -1;
        len_out[b] = n;
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* CTC loss, forward-backward, float64 log space                              */
/* ------------------------------------------------------------------------- */

static inline double lse2(double a, double b) {
    if (a == NEG_INF) return b;
    if (b == NEG_INF) return a;
    double m = a > b ? a : b;
    return m + log(exp(a - m) + exp(b - m));
}

static inline double lse3(double a, double b, double c) { return lse2(lse2(a, b), c); }

/*
 * x          : [T,N,C] through element strides (st, sn), class stride 1.  Logits when
 *              from_logits != 0 (log_softmax is applied internally), log-probs else.
 * targets    : if tgt_stride > 0: padded [N, tgt_stride]; if 0: 1-D concatenation.
 * nll_out    : [N] negative log likelihood per sample (inf when no alignment exists
 *              and zero_infinity == 0; 0 when zero_infinity != 0).
 * grad_out   : NULL or contiguous [T,N,C]: d nll_n / d x[t,n,c] per sample,
 *              unscaled by any reduction.
 *                from_logits != 0 : gradient w.r.t. the logits (softmax - occupancy)
 *                from_logits == 0 : torch's convention for log-prob inputs,
 *                                   exp(lp) - occupancy (what ATen's backward emits)
 *              Frames t >= input_len get 0.  Infeasible samples get NaN on frames
 *              t < input_len when zero_infinity == 0 (torch behaviour) and 0 when
 *              zero_infinity != 0.
 */
int oracle_ctc_loss_f64(const double *x, int from_logits, int T, int N, int C,
                        long long st, long long sn,
                        const int64_t *targets, long long tgt_stride,
                        const int64_t *in_len, const int64_t *tg_len,
                        int blank, int zero_infinity,
                        double *nll_out, double *grad_out) {
    if (T < 0 || N < 0 || C <= 0 || blank < 0 || blank >= C) return 1;
    /* offsets for concatenated targets */
    int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(N + 1));
    if (!off) return 2;
    off[0] = 0;
    for (int n = 0; n < N; ++n)
        off[n + 1] = off[n] + (tgt_stride > 0 ? tgt_stride : tg_len[n]);
    int bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int n = 0; n < N; ++n) {
        const int Tn = (int)in_len[n];
        const int L = (int)tg_len[n];
        const int S = 2 * L + 1;
        const int64_t *tg = targets + off[n];
        int lbad = 0;
        if (Tn < 0 || Tn > T || L < 0) { bad = 1; continue; }
        double *lp = (double *)malloc(sizeof(double) * (size_t)(Tn > 0 ? Tn : 1) * C);
        double *la = (double *)malloc(sizeof(double) * (size_t)(Tn > 0 ? Tn : 1) * S);
        double *lb = (double *)malloc(sizeof(double) * (size_t)(Tn > 0 ? Tn : 1) * S);
        int *ext = (int *)malloc(sizeof(int) * (size_t)S);
        for (int s = 0; s < S; ++s) ext[s] = (s & 1) ? (int)tg[s >> 1] : blank;
        for (int s = 1; s < S; s += 2)
            if (ext[s] < 0 || ext[s] >= C) { lbad = 1; ext[s] = blank; }
        /* log-probs of this sample */
        for (int t = 0; t < Tn; ++t) {
            const double *row = x + t * st + n * sn;
            double z = 0.0;
            if (from_logits) {
                double m = row[0];
                for (int c = 1; c < C; ++c) if (row[c] > m) m = row[c];
                double sum = 0.0;
                for (int c = 0; c < C; ++c) sum += exp(row[c] - m);
                z = m + log(sum);
            }
            for (int c = 0; c < C; ++c) lp[(size_t)t * C + c] = row[c] - z;
        }
        double nll = INFINITY;
        if (Tn == 0) {
            nll = (L == 0) ? 0.0 : INFINITY;
        } else if (!lbad) {
            /* alpha */
            for (int s = 0; s < S; ++s) la[s] = NEG_INF;
            la[0] = lp[blank];
            if (S > 1) la[1] = lp[ext[1]];
            for (int t = 1; t < Tn; ++t) {
                const double *p = la + (size_t)(t - 1) * S;
                double *q = la + (size_t)t * S;
                for (int s = 0; s < S; ++s) {
                    double a = p[s];
                    double b = s >= 1 ? p[s - 1] : NEG_INF;
                    double c = (s >= 2 && (s & 1) && ext[s] != ext[s - 2]) ? p[s - 2] : NEG_INF;
                    double v = lse3(a, b, c);
                    q[s] = (v == NEG_INF) ? NEG_INF : v + lp[(size_t)t * C + ext[s]];
                }
            }
            const double *last = la + (size_t)(Tn - 1) * S;
            double ll = (S > 1) ? lse2(last[S - 1], last[S - 2]) : last[0];
            nll = -ll;
            /* beta */
            double *e = lb + (size_t)(Tn - 1) * S;
            for (int s = 0; s < S; ++s) e[s] = NEG_INF;
            e[S - 1] = lp[(size_t)(Tn - 1) * C + blank];
            if (S > 1) e[S - 2] = lp[(size_t)(Tn - 1) * C + ext[S - 2]];
            for (int t = Tn - 2; t >= 0; --t) {
                const double *p = lb + (size_t)(t + 1) * S;
                double *q = lb + (size_t)t * S;
                for (int s = 0; s < S; ++s) {
                    double a = p[s];
                    double b = s + 1 < S ? p[s + 1] : NEG_INF;
                    double c = (s + 2 < S && (s & 1) && ext[s] != ext[s + 2]) ? p[s + 2] : NEG_INF;
                    double v = lse3(a, b, c);
                    q[s] = (v == NEG_INF) ? NEG_INF : v + lp[(size_t)t * C + ext[s]];
                }
            }
        }
        const int infeasible = isinf(nll);
        nll_out[n] = (infeasible && zero_infinity) ? 0.0 : nll;
        if (grad_out) {
            for (int t = 0; t < T; ++t) {
                double *g = grad_out + ((size_t)t * N + n) * C;
                if (t >= Tn) { for (int c = 0; c < C; ++c) g[c] = 0.0; continue; }
                if (infeasible) {
                    double v = zero_infinity ? 0.0 : NAN;
                    for (int c = 0; c < C; ++c) g[c] = v;
                    continue;
                }
                /* occupancy: sum_{s: ext[s]==c} exp(alpha+beta - lp + nll)        */
                /* accumulate in log space (lcab) like ATen, then convert once.   */
                for (int c = 0; c < C; ++c) g[c] = NEG_INF;
                for (int s = 0; s < S; ++s) {
                    double ab = la[(size_t)t * S + s] + lb[(size_t)t * S + s];
                    g[ext[s]] = lse2(g[ext[s]], ab);
                }
                for (int c = 0; c < C; ++c) {
                    double l = lp[(size_t)t * C + c];
                    double occ = (g[c] == NEG_INF) ? 0.0 : exp(g[c] - l + nll);
                    g[c] = exp(l) - occ;
                }
            }
        }
        free(lp); free(la); free(lb); free(ext);
        if (lbad) bad = 1;
    }
    free(off);
    return bad ? 3 : 0;
}

/* float32 front end used by the timed CPU baseline: widens per sample. */
int oracle_ctc_loss_f32(const float *x, int from_logits, int T, int N, int C,
                        long long st, long long sn,
                        const int64_t *targets, long long tgt_stride,
                        const int64_t *in_len, const int64_t *tg_len,
                        int blank, int zero_infinity,
                        double *nll_out, float *grad_out) {
    /* gather into a contiguous float64 [T,N,C] copy, run, narrow the gradient */
    size_t tot = (size_t)T * N * C;
    double *xd = (double *)calloc(tot ? tot : 1, sizeof(double));
    double *gd = grad_out ? (double *)malloc(sizeof(double) * (tot ? tot : 1)) : NULL;
    if (!xd || (grad_out && !gd)) { free(xd); free(gd); return 2; }
    for (int t = 0; t < T; ++t)
        for (int n = 0; n < N; ++n) {
            const float *row = x + t * st + n * sn;
            double *dst = xd + ((size_t)t * N + n) * C;
            for (int c = 0; c < C; ++c) dst[c] = (double)row[c];
        }
    int rc = oracle_ctc_loss_f64(xd, from_logits, T, N, C, (long long)N * C, C, targets,
                                 tgt_stride, in_len, tg_len, blank, zero_infinity,
                                 nll_out, gd);
    if (grad_out)
        for (size_t i = 0; i < tot; ++i) grad_out[i] = (float)gd[i];
    free(xd); free(gd);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* BidirectionalLSTM forward (model/model.py:151-163), float64                */
/* ------------------------------------------------------------------------- */

static inline double sigm(double v) { return 1.0 / (1.0 + exp(-v)); }

/* One direction of a 1-layer LSTM, batch_first.  Weight layout is torch's:
 * w_ih [4H, I], w_hh [4H, H], b_ih [4H], b_hh [4H], gate blocks i,f,g,o.
 * x [B,T,I]; hcat [B,T,2H] receives h_t at column offset col0.  h0 = c0 = 0. */
static void lstm_dir(const double *x, int B, int T, int I, int H,
                     const double *w_ih, const double *w_hh,
                     const double *b_ih, const double *b_hh,
                     int reverse, double *hcat, int col0) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int b = 0; b < B; ++b) {
        double *h = (double *)calloc((size_t)H, sizeof(double));
        double *c = (double *)calloc((size_t)H, sizeof(double));
        double *g = (double *)malloc(sizeof(double) * 4 * (size_t)H);
        for (int step = 0; step < T; ++step) {
            int t = reverse ? T - 1 - step : step;
            const double *xt = x + ((size_t)b * T + t) * I;
            for (int r = 0; r < 4 * H; ++r) {
                double acc = b_ih[r] + b_hh[r];
                const double *wi = w_ih + (size_t)r * I;
                for (int k = 0; k < I; ++k) acc += wi[k] * xt[k];
                const double *wh = w_hh + (size_t)r * H;
                for (int k = 0; k < H; ++k) acc += wh[k] * h[k];
                g[r] = acc;
            }
            double *out = hcat + ((size_t)b * T + t) * 2 * H + col0;
            for (int j = 0; j < H; ++j) {
                double ig = sigm(g[j]), fg = sigm(g[H + j]);
                double gg = tanh(g[2 * H + j]), og = sigm(g[3 * H + j]);
                c[j] = fg * c[j] + ig * gg;
                h[j] = og * tanh(c[j]);
                out[j] = h[j];
            }
        }
        free(h); free(c); free(g);
    }
}

/* x [B,T,I] -> out [B,T,O].  w_* / b_* are the forward direction, *_r the
 * "_reverse" twins; lin_w [O, 2H], lin_b [O].  hcat_out (optional, may be NULL)
 * receives the [B,T,2H] LSTM output before the affine map. */
int oracle_bilstm_forward_f64(const double *x, int B, int T, int I, int H, int O,
                              const double *w_ih, const double *w_hh,
                              const double *b_ih, const double *b_hh,
                              const double *w_ih_r, const double *w_hh_r,
                              const double *b_ih_r, const double *b_hh_r,
                              const double *lin_w, const double *lin_b,
                              double *out, double *hcat_out) {
    if (B < 0 || T < 0 || I <= 0 || H <= 0 || O <= 0) return 1;
    size_t nh = (size_t)B * T * 2 * H;
    double *hcat = hcat_out ? hcat_out : (double *)malloc(sizeof(double) * (nh ? nh : 1));
    if (!hcat) return 2;
    lstm_dir(x, B, T, I, H, w_ih, w_hh, b_ih, b_hh, 0, hcat, 0);
    lstm_dir(x, B, T, I, H, w_ih_r, w_hh_r, b_ih_r, b_hh_r, 1, hcat, H);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long long bt = 0; bt < (long long)B * T; ++bt) {
        const double *hv = hcat + (size_t)bt * 2 * H;
        double *o = out + (size_t)bt * O;
        for (int r = 0; r < O; ++r) {
            double acc = lin_b[r];
            const double *w = lin_w + (size_t)r * 2 * H;
            for (int k = 0; k < 2 * H; ++k) acc += w[k] * hv[k];
            o[r] = acc;
        }
    }
    if (!hcat_out) free(hcat);
    return 0;
}
