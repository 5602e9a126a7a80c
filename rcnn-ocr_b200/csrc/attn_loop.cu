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
