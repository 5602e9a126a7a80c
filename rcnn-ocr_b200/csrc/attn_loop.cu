// The attention decoder's greedy step loop (model/model.py:89-108) as ONE host call: the three launches per step
// (attn.cu, gemm.cu) issued from C++, so that an eager (not graph-captured) decode is bound by the device and not by 81
// trips through the Python binding.  Nothing here touches the device itself.
#include "common.cuh"

extern "C" int rcnn_attn_greedy_decode(const void *projH, const float *v, const void *enc, int64_t enc_stride_b, int64_t enc_stride_t,
                                       const void *wcat_il, const float *bcat_il, const float *embT_il, const void *comb_w,
                                       const float *comb_b, int comb_rows, int B, int T, int H, int C, int V, int steps, int blank,
                                       int64_t *y, void *xcat0, void *xcat1, float *c, float *hg, float *probs, int chain,
                                       rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 1 && H >= 8 && C >= 8 && V >= 1 && steps >= 1 && comb_rows >= H + V,
                   "attn_greedy_decode: bad shape B=%d T=%d H=%d C=%d V=%d steps=%d comb_rows=%d", B, T, H, C, V, steps, comb_rows);
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(projH && v && enc && wcat_il && bcat_il && embT_il && comb_w && comb_b && y && xcat0 && xcat1 && c && hg && probs,
                   "attn_greedy_decode: null pointer");
    const int64_t K = (int64_t)C + H, Np = comb_rows;
    __nv_bfloat16 *xc[2] = {(__nv_bfloat16 *)xcat0, (__nv_bfloat16 *)xcat1};
    const float *logits = hg + H;                                 // columns [H, H + V) of the [h2h | generator] product
    rcnn_chain_launches(chain);
    int rc = RCNN_OK;
    for (int t = 0; t < steps && rc == RCNN_OK; ++t) {
        __nv_bfloat16 *cur = xc[t & 1], *nxt = xc[(t + 1) & 1];
        // score / softmax / context from proj_h = hg[:, :H]; the previous step's mask + copy + argmax rides along
        rc = rcnn_attn_step_bf16(projH, hg, Np, v, enc, enc_stride_b, enc_stride_t, B, T, H, C, nullptr, cur, K,
                                 t == 0 ? nullptr : logits, Np, V, blank, t == 0 ? nullptr : probs + (size_t)(t - 1) * V,
                                 (int64_t)steps * V, y, stream);
        if (rc != RCNN_OK) break;
        rc = rcnn_attn_gates_cell(cur, K, wcat_il, K, bcat_il, embT_il, y, B, H, (int)K, V, c, nxt + C, K, nullptr, 0, stream);
        if (rc != RCNN_OK) break;
        rc = rcnn_gemm_bf16(nxt + C, K, comb_w, H, hg, Np, RCNN_F32, comb_b, B, (int)Np, H, stream);
    }
    rcnn_chain_launches(0);
    if (rc != RCNN_OK) return rc;
    return rcnn_attn_argmax_ld(logits, Np, B, V, blank, probs + (size_t)(steps - 1) * V, (int64_t)steps * V, y, stream);
}

// The teacher-forced training pass (attention._TeacherForcedFn): forward and backward step loops, same kernels and order as the
// Python loops they replace.  Array shapes as in the header comment of the training entry points; S = steps.
extern "C" int rcnn_attn_train_forward(const void *projH, const float *v, const void *enc, int64_t enc_stride_b, int64_t enc_stride_t,
                                       const void *h2h_w, const float *h2h_b, const void *wcat_il, const float *bcat_il,
                                       const float *embT_il, const int64_t *tokens, const float *alpha_scale, int B, int T, int H,
                                       int C, int V, int S, void *xcat_all, float *c_all, float *gates_all, float *alpha_all,
                                       float *projh_all, float *out_hid, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 1 && H >= 8 && C >= 8 && V >= 1 && S >= 1, "attn_train_forward: bad shape");
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(projH && v && enc && h2h_w && h2h_b && wcat_il && bcat_il && embT_il && tokens && xcat_all && c_all && gates_all &&
                       alpha_all && projh_all && out_hid,
                   "attn_train_forward: null pointer");
    const int64_t K = (int64_t)C + H;
    __nv_bfloat16 *xall = (__nv_bfloat16 *)xcat_all;
    rcnn_chain_launches(1);
    int rc = RCNN_OK;
    for (int t = 0; t < S && rc == RCNN_OK; ++t) {
        __nv_bfloat16 *cur = xall + (size_t)t * B * K, *nxt = cur + (size_t)B * K;
        float *ph = projh_all + (size_t)t * B * H;
        rc = rcnn_gemm_bf16(cur + C, K, h2h_w, H, ph, H, RCNN_F32, h2h_b, B, H, H, stream);
        if (rc != RCNN_OK) break;
        rc = rcnn_attn_step_train(projH, ph, H, v, enc, enc_stride_b, enc_stride_t, B, T, H, C, alpha_all + (size_t)t * B * T,
                                  alpha_scale ? alpha_scale + (size_t)t * B * T : nullptr, cur, K, stream);
        if (rc != RCNN_OK) break;
        rc = rcnn_attn_gates_cell_train(cur, K, wcat_il, K, bcat_il, embT_il, tokens + (size_t)t * B, B, H, (int)K, V,
                                        c_all + (size_t)t * B * H, c_all + (size_t)(t + 1) * B * H, nxt + C, K,
                                        out_hid + (size_t)t * H, (int64_t)S * H, gates_all + (size_t)t * B * 4 * H, stream);
    }
    rcnn_chain_launches(0);
    return rc;
}

extern "C" int rcnn_attn_train_backward(const float *d_out, const float *gates_all, const float *c_all, const void *b1, const void *b2,
                                        const float *alpha_all, const float *alpha_scale, const void *enc, int64_t enc_stride_b,
                                        int64_t enc_stride_t, const void *projH, const float *projh_all, const float *v, int B, int T,
                                        int H, int C, int S, void *dg_all, float *dctx_all, float *de_all, float *dv_acc, float *dc,
                                        float *dh, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(B >= 0 && T >= 1 && H >= 8 && C >= 8 && S >= 1, "attn_train_backward: bad shape");
    if (B == 0) return RCNN_OK;
    RCNN_CHECK_ARG(d_out && gates_all && c_all && b1 && b2 && alpha_all && enc && projH && projh_all && v && dg_all && dctx_all &&
                       de_all && dv_acc && dc && dh,
                   "attn_train_backward: null pointer");
    const int64_t H5 = 5 * (int64_t)H;
    __nv_bfloat16 *dgall = (__nv_bfloat16 *)dg_all;
    rcnn_chain_launches(1);
    int rc = RCNN_OK;
    for (int t = S - 1; t >= 0 && rc == RCNN_OK; --t) {
        __nv_bfloat16 *dg = dgall + (size_t)t * B * H5;
        float *dctx = dctx_all + (size_t)t * B * C;
        rc = rcnn_attn_cell_bwd(gates_all + (size_t)t * B * 4 * H, c_all + (size_t)t * B * H, c_all + (size_t)(t + 1) * B * H,
                                d_out + (size_t)t * H, (int64_t)S * H, t == S - 1 ? nullptr : dh, H, dc, B, H, dg, H5, stream);
        if (rc != RCNN_OK) break;
        rc = rcnn_gemm_bf16(dg, H5, b1, 4 * (int64_t)H, dctx, C, RCNN_F32, nullptr, B, C, 4 * H, stream);   // dcontext
        if (rc != RCNN_OK) break;
        rc = rcnn_attn_step_bwd(dctx, C, alpha_all + (size_t)t * B * T, alpha_scale ? alpha_scale + (size_t)t * B * T : nullptr, enc,
                                enc_stride_b, enc_stride_t, projH, projh_all + (size_t)t * B * H, H, v, B, T, H, C,
                                de_all + (size_t)t * B * T, dg + 4 * (size_t)H, H5, dv_acc, stream);
        if (rc != RCNN_OK) break;
        if (t > 0) rc = rcnn_gemm_bf16(dg, H5, b2, H5, dh, H, RCNN_F32, nullptr, B, H, (int)H5, stream);     // dh_{t-1}
    }
    rcnn_chain_launches(0);
    return rc;
}
