"""Attention decoder of the reference (model/model.py:23-148) on the device, inference path.

``Attention`` keeps the reference's constructor, parameter names and shapes
(``attention_cell.i2h/h2h/score/rnn.*``, ``generator.*``), so ``attn.*`` of a reference checkpoint
loads with ``load_state_dict(strict=True)``.  ``forward(batch_H, text=None, is_train=True,
batch_max_length=25)`` returns what the reference returns: greedy-decoded ``probs [B, steps, V]``
for ``is_train=False`` (model/model.py:89-108), teacher-forced logits for ``is_train=True``
(:110-148).  The step loop runs on the tcgen05 GEMM (K1) plus three small kernels (K6, csrc/attn.cu);
``i2h(batch_H)`` is hoisted out of the loop.

Training (``is_train=True`` with autograd enabled): ``_train_forward`` records an autograd graph -- every product
(i2h, h2h, the LSTMCell gates over [context | h], the generator) runs on the tcgen05 GEMM through
``model._LinearFn`` (bf16 operands, fp32 accumulate, its backward on the K1 kernels as well), the one-hot half of
the LSTMCell input is a row gather of ``W_ih[:, C + y]``, the score / softmax / context and the cell update are fp32
elementwise torch ops.  Dropout (``dropout_p`` on alpha and on the generator input, model/model.py:41,136) and
scheduled sampling (``sampling_prob``: one ``torch.rand(1)`` per step on the CPU generator as at :139, the fed-back
token is the argmax of that step's logits) follow the reference; gradients are pinned to the reference module's own
(tests/golden/attn_train_*.npz).  With autograd disabled the fused inference kernels (K6) run.

Teacher forcing without scheduled sampling (the reference's default, ``sampling_prob = 0``) takes the fused training
path ``_TeacherForcedFn``: forward on the inference step kernels (alpha and the gate activations kept), backward on
``csrc/attn_bwd.cu`` + the tcgen05 GEMMs, alpha dropout as a multiplier drawn once for all steps.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib, ops


class AttentionCell(nn.Module):
    """Parameter holder with the reference's names (model/model.py:23-31)."""

    def __init__(self, input_size, hidden_size, num_embeddings, dropout_p=0.1):
        super().__init__()
        self.i2h = nn.Linear(input_size, hidden_size, bias=False)
        self.h2h = nn.Linear(hidden_size, hidden_size)
        self.score = nn.Linear(hidden_size, 1, bias=False)
        self.rnn = nn.LSTMCell(input_size + num_embeddings, hidden_size)
        self.hidden_size = hidden_size
        self.dropout_p = dropout_p



class _TeacherForcedFn(torch.autograd.Function):
    """out_hid [B, steps, H] of the teacher-forced pass (model/model.py:131-135 for every step), forward and backward on
    the fused kernels (csrc/attn.cu, attn_bwd.cu, the tcgen05 GEMMs): three launches per forward step, four per backward
    step, the weight gradients as products over all (step, sequence) rows after the loop.  bf16 operands, fp32 state and
    accumulation.  ``tokens`` [steps, B] are the step inputs (text[:, t]), ``alpha_scale`` [steps, B, T] the dropout
    multipliers of alpha (0 or 1 / (1 - p)) or None."""

    @staticmethod
    def forward(ctx, batch_H, tokens, alpha_scale, i2h_w, h2h_w, h2h_b, score_w, w_ih, w_hh, b_ih, b_hh):
        L = _lib.lib()
        B, T, C = batch_H.shape
        H = h2h_w.shape[0]
        V = w_ih.shape[1] - C
        S = tokens.shape[0]
        K = C + H
        dev = batch_H.device
        with torch.cuda.device(dev):
            enc = batch_H.detach().float()
            if enc.stride(2) != 1:
                enc = enc.contiguous()
            encb = ops.cast_bf16_3d(enc)
            i2h16 = ops.cast_bf16_2d(i2h_w.detach())
            h2h16 = ops.cast_bf16_2d(h2h_w.detach())
            projH = ops.gemm_bf16(encb.view(B * T, C), i2h16, None, torch.bfloat16)                       # [B*T, H]
            v = score_w.detach().float().reshape(-1).contiguous()
            wcat = torch.cat([w_ih.detach()[:, :C], w_hh.detach()], 1)                                     # [4H, C+H]
            wcat_il = ops.cast_bf16_2d(wcat.view(4, H, K).permute(1, 0, 2).reshape(4 * H, K).contiguous())
            bcat_il = (b_ih.detach() + b_hh.detach()).float().view(4, H).t().reshape(-1).contiguous()
            embT_il = w_ih.detach()[:, C:].t().float().reshape(V, 4, H).permute(0, 2, 1).reshape(V, 4 * H).contiguous()
            h2h_bias = h2h_b.detach().float().contiguous()
            xcat_all = torch.zeros((S + 1, B, K), dtype=torch.bfloat16, device=dev)     # row block t: [context_t | h_{t-1}]
            c_all = torch.zeros((S + 1, B, H), dtype=torch.float32, device=dev)         # c_all[t] = c_{t-1}
            gates_all = torch.empty((S, B, 4 * H), dtype=torch.float32, device=dev)
            alpha_all = torch.empty((S, B, T), dtype=torch.float32, device=dev)
            projh_all = torch.empty((S, B, H), dtype=torch.float32, device=dev)
            out_hid = torch.empty((B, S, H), dtype=torch.float32, device=dev)
            if alpha_scale is not None:
                alpha_scale = alpha_scale.to(device=dev, dtype=torch.float32).contiguous()
            tokens = tokens.to(device=dev, dtype=torch.int64).contiguous()
            # the step loop (h2h GEMM -> step kernel -> gate GEMM + cell, S times) as one host call
            _lib.check(L.rcnn_attn_train_forward(projH.data_ptr(), v.data_ptr(), encb.data_ptr(), encb.stride(0), encb.stride(1),
                                                 h2h16.data_ptr(), h2h_bias.data_ptr(), wcat_il.data_ptr(), bcat_il.data_ptr(),
                                                 embT_il.data_ptr(), tokens.data_ptr(),
                                                 alpha_scale.data_ptr() if alpha_scale is not None else None, B, T, H, C, V, S,
                                                 xcat_all.data_ptr(), c_all.data_ptr(), gates_all.data_ptr(), alpha_all.data_ptr(),
                                                 projh_all.data_ptr(), out_hid.data_ptr(), _lib.stream_ptr()),
                       "rcnn_attn_train_forward")
        ctx.save_for_backward(encb, projH, v, wcat_il, h2h16, i2h16, xcat_all, c_all, gates_all, alpha_all, projh_all, tokens,
                              alpha_scale if alpha_scale is not None else torch.empty(0, device=dev))
        ctx.dims = (B, T, C, H, V, S)
        ctx.in_dtype = batch_H.dtype
        return out_hid

    @staticmethod
    def backward(ctx, d_out):
        (encb, projH, v, wcat_il, h2h16, i2h16, xcat_all, c_all, gates_all, alpha_all, projh_all, tokens, alpha_scale) = ctx.saved_tensors
        B, T, C, H, V, S = ctx.dims
        K = C + H
        L = _lib.lib()
        dev = encb.device
        scale = alpha_scale if alpha_scale.numel() else None
        with torch.cuda.device(dev):
            d_out = d_out.float().contiguous()                                             # [B, S, H]
            b1 = wcat_il[:, :C].t().contiguous()                                           # [C, 4H]: dcontext = dgates @ W_ih[:, :C]
            b2 = torch.cat([wcat_il[:, C:].t(), h2h16.t()], 1).contiguous()                # [H, 5H]: dh = [dgates | dproj_h] @ [W_hh; W_h2h]
            dg_all = torch.zeros((S, B, 5 * H), dtype=torch.bfloat16, device=dev)          # [dgates (interleaved) | dproj_h]
            dctx_all = torch.empty((S, B, C), dtype=torch.float32, device=dev)
            de_all = torch.empty((S, B, T), dtype=torch.float32, device=dev)
            dv_acc = torch.zeros((B, H), dtype=torch.float32, device=dev)
            dc = torch.zeros((B, H), dtype=torch.float32, device=dev)
            dh = torch.empty((B, H), dtype=torch.float32, device=dev)
            s = _lib.stream_ptr()
            # the step loop (cell backward -> dcontext GEMM -> attention-step backward -> dh GEMM, S times) as one host call
            _lib.check(L.rcnn_attn_train_backward(d_out.data_ptr(), gates_all.data_ptr(), c_all.data_ptr(), b1.data_ptr(),
                                                  b2.data_ptr(), alpha_all.data_ptr(), scale.data_ptr() if scale is not None else None,
                                                  encb.data_ptr(), encb.stride(0), encb.stride(1), projH.data_ptr(),
                                                  projh_all.data_ptr(), v.data_ptr(), B, T, H, C, S, dg_all.data_ptr(),
                                                  dctx_all.data_ptr(), de_all.data_ptr(), dv_acc.data_ptr(), dc.data_ptr(),
                                                  dh.data_ptr(), s), "rcnn_attn_train_backward")
            dprojH = torch.empty((B * T, H), dtype=torch.bfloat16, device=dev)
            _lib.check(L.rcnn_attn_dprojH(de_all.data_ptr(), projh_all.data_ptr(), projH.data_ptr(), v.data_ptr(), S, B, T, H,
                                          dprojH.data_ptr(), s), "rcnn_attn_dprojH")
            rows_g = dg_all.view(S * B, 5 * H)
            dgates, dproj = rows_g[:, :4 * H], rows_g[:, 4 * H:]
            rows_x = xcat_all[:S].view(S * B, K)
            # parameter gradients
            d_i2h = ops.gemm_bf16_atb(dprojH, encb.view(B * T, C))                         # [H, C]
            d_h2h_w = ops.gemm_bf16_atb(dproj, rows_x[:, C:])                              # [H, H]
            d_score = dv_acc.sum(0).view(1, H)
            dwcat = ops.gemm_bf16_atb(dgates, rows_x).view(H, 4, K).permute(1, 0, 2).reshape(4 * H, K)   # back to torch's gate order
            # the one-hot half of W_ih and the three bias gradients in ONE product: [onehot(y) | 1]^T @ [dgates | dproj_h] --
            # row v < V = the sum of the rows whose input token was v, row V = the column sums
            Vp = (V + 1 + 7) // 8 * 8
            oh = torch.zeros((S * B, Vp), dtype=torch.bfloat16, device=dev)
            oh.scatter_(1, tokens.view(-1, 1).clamp(0, V - 1), 1.0)
            oh[:, V] = 1.0
            sums = ops.gemm_bf16_atb(oh, rows_g)                                            # [Vp, 5H]
            d_emb, d_b_il, d_h2h_b = sums[:V, :4 * H], sums[V, :4 * H], sums[V, 4 * H:].contiguous()
            d_w_ih = torch.cat([dwcat[:, :C], d_emb.reshape(V, H, 4).permute(2, 1, 0).reshape(4 * H, V)], 1)
            d_w_hh = dwcat[:, C:].contiguous()
            d_b = d_b_il.reshape(H, 4).t().reshape(-1)
            # encoder gradient: through proj_H and through the context
            d_enc = ops.gemm_bf16(dprojH, i2h16.t().contiguous(), None, torch.float32).view(B, T, C)
            alpha_d = alpha_all * scale if scale is not None else alpha_all
            d_enc = d_enc + torch.bmm(alpha_d.permute(1, 2, 0), dctx_all.permute(1, 0, 2))
        need = ctx.needs_input_grad
        return (d_enc.to(ctx.in_dtype) if need[0] else None, None, None, d_i2h if need[3] else None, d_h2h_w if need[4] else None,
                d_h2h_b if need[5] else None, d_score if need[6] else None, d_w_ih if need[7] else None,
                d_w_hh if need[8] else None, d_b if need[9] else None, d_b.clone() if need[10] else None)


class Attention(nn.Module):
    def __init__(self, input_size, hidden_size, num_classes, sos_id: int, eos_id: int, pad_id: int,
                 blank_id=None, dropout_p: float = 0.1, sampling_prob: float = 0.0):
        super().__init__()
        if input_size % 8 or hidden_size % 8:
            raise ValueError("input_size and hidden_size must be multiples of 8 (16-byte bf16 rows)")
        self.attention_cell = AttentionCell(input_size, hidden_size, num_classes, dropout_p=dropout_p)
        self.input_size, self.hidden_size, self.num_classes = input_size, hidden_size, num_classes
        self.sos_id, self.eos_id, self.pad_id, self.blank_id = sos_id, eos_id, pad_id, blank_id
        self.generator = nn.Linear(hidden_size, num_classes)
        self.dropout_p = dropout_p
        self.sampling_prob = sampling_prob
        self._prepared = None
        self._alpha_scale_override = None      # [steps, B, T] multipliers used instead of drawing the alpha dropout mask

    # ---- bf16 / transposed views of the parameters, rebuilt only when a parameter changed ------------
    def _weights(self):
        cell = self.attention_cell
        ps = [cell.i2h.weight, cell.h2h.weight, cell.h2h.bias, cell.score.weight, cell.rnn.weight_ih,
              cell.rnn.weight_hh, cell.rnn.bias_ih, cell.rnn.bias_hh, self.generator.weight, self.generator.bias]
        key = tuple((p.data_ptr(), p._version, p.device) for p in ps)
        if self._prepared is None or self._prepared[0] != key:
            C = self.input_size
            with torch.no_grad():
                w = {
                    "i2h": ops.cast_bf16_2d(cell.i2h.weight.detach()),                                    # [H, C]
                    "h2h": ops.cast_bf16_2d(cell.h2h.weight.detach()),                                    # [H, H]
                    "h2h_b": cell.h2h.bias.detach().float().contiguous(),
                    "v": cell.score.weight.detach().float().reshape(-1).contiguous(),                     # [H]
                    # gates = [context, h] @ [W_ih[:, :C] | W_hh]^T + (b_ih + b_hh) + W_ih[:, C + y]
                    "wcat": ops.cast_bf16_2d(torch.cat([cell.rnn.weight_ih.detach()[:, :C],
                                                        cell.rnn.weight_hh.detach()], 1).contiguous()),   # [4H, C+H]
                    "bcat": (cell.rnn.bias_ih.detach() + cell.rnn.bias_hh.detach()).float().contiguous(),
                    "embT": cell.rnn.weight_ih.detach()[:, C:].t().float().contiguous(),                  # [V, 4H]
                    "gen": ops.cast_bf16_2d(self.generator.weight.detach()),                              # [V, H]
                    "gen_b": self.generator.bias.detach().float().contiguous(),
                }
                if self.hidden_size % 8 == 0:
                    # gate-interleaved copies (row 4u + g = gate g of unit u) for the GEMM whose epilogue is the LSTMCell step
                    Hh, K = self.hidden_size, w["wcat"].shape[1]
                    w["wcat_il"] = w["wcat"].view(4, Hh, K).permute(1, 0, 2).reshape(4 * Hh, K).contiguous()
                    w["bcat_il"] = w["bcat"].view(4, Hh).t().reshape(-1).contiguous()
                    w["embT_il"] = w["embT"].view(-1, 4, Hh).permute(0, 2, 1).reshape(-1, 4 * Hh).contiguous()
                # greedy decode: h2h(h_t) (for step t+1) and generator(h_t) (for step t) in ONE product, N padded to 32 k
                H, V = self.hidden_size, self.num_classes
                Np = (H + V + 31) // 32 * 32
                comb = torch.zeros((Np, H), dtype=torch.float32, device=cell.h2h.weight.device)
                comb[:H] = cell.h2h.weight.detach()
                comb[H:H + V] = self.generator.weight.detach()
                comb_b = torch.zeros((Np,), dtype=torch.float32, device=comb.device)
                comb_b[:H] = cell.h2h.bias.detach()
                comb_b[H:H + V] = self.generator.bias.detach()
                w["comb"], w["comb_b"] = ops.cast_bf16_2d(comb), comb_b
            self._prepared = (key, w)
        return self._prepared[1]

    @torch.no_grad()
    def _decode(self, batch_H, steps, text):
        """Both paths of model/model.py:89-148.  text is None: greedy (returns probs [B,steps,V]);
        else teacher forcing with targets text[:, t] (returns generator(out_hid), blank-masked)."""
        _lib.require_cuda(batch_H, "batch_H")
        if batch_H.dim() != 3 or batch_H.shape[2] != self.input_size:
            raise RuntimeError(f"Attention expects [B, T, {self.input_size}], got {tuple(batch_H.shape)}")
        if self.training and self.dropout_p > 0:
            raise NotImplementedError("dropout belongs to the autograd path (_train_forward); the fused kernels run "
                                      "in eval() mode or with dropout_p == 0")
        B, T, C = batch_H.shape
        H, V = self.hidden_size, self.num_classes
        dev = batch_H.device
        w = self._weights()
        L = _lib.lib()
        blank = -1 if self.blank_id is None else int(self.blank_id)
        enc = batch_H.float()
        if enc.stride(2) != 1:
            enc = enc.contiguous()
        with torch.cuda.device(dev):
            encb = ops.cast_bf16_3d(enc)
            # hoisted i2h(batch_H).  The step kernel reads proj_H and batch_H as bf16 when the shapes allow 16-byte loads
            half = H % 8 == 0 and C % 8 == 0 and os.environ.get("RCNN_ATTN_F32", "0") != "1"
            projH = ops.gemm_bf16(encb.view(B * T, C), w["i2h"], None, torch.bfloat16 if half else torch.float32)

            def score_context(xcat, ph, ph_ld, prev=None):
                """K6a for this step; prev = (logits, ld, probs row block, y): K6c of the previous step rides along (bf16 kernel)"""
                if half:
                    pl, pld, pp, py = prev if prev is not None else (None, 0, None, None)
                    rc = L.rcnn_attn_step_bf16(projH.data_ptr(), ph.data_ptr(), ph_ld, w["v"].data_ptr(), encb.data_ptr(),
                                               encb.stride(0), encb.stride(1), B, T, H, C, None, xcat.data_ptr(), xcat.stride(0),
                                               pl, pld, V, blank, pp, probs.stride(0) if pp is not None else 0, py, s)
                else:
                    if prev is not None:
                        _lib.check(L.rcnn_attn_argmax_ld(prev[0], prev[1], B, V, blank, prev[2], probs.stride(0), prev[3], s),
                                   "rcnn_attn_argmax")
                    rc = L.rcnn_attn_score_context_ld(projH.data_ptr(), ph.data_ptr(), ph_ld, w["v"].data_ptr(), enc.data_ptr(),
                                                      enc.stride(0), enc.stride(1), B, T, H, C, None, xcat.data_ptr(),
                                                      xcat.stride(0), s)
                _lib.check(rc, "rcnn_attn_score_context")
            # [context | h] rows, h_0 = 0.  Two of them: the fused gate GEMM + cell writes h_t into the other one, because
            # CTAs that are still in their K loop read h_{t-1} from this one
            fused = half and os.environ.get("RCNN_ATTN_FUSED_CELL", "1") != "0"
            # the step loop is a chain of short dependent kernels: launched with programmatic stream serialization (the
            # bf16 step kernel and the tcgen05 GEMMs wait on the device for their predecessor; the fp32 kernels do not take part)
            chain = fused and os.environ.get("RCNN_ATTN_CHAIN", "1") != "0"
            xc = [torch.zeros((B, C + H), dtype=torch.bfloat16, device=dev) for _ in range(2 if fused else 1)]
            c = torch.zeros((B, H), dtype=torch.float32, device=dev)
            gates = None if fused else torch.empty((B, 4 * H), dtype=torch.float32, device=dev)
            greedy = text is None
            probs = torch.zeros((B, steps, V), dtype=torch.float32, device=dev) if greedy else None
            out_hid = None if greedy else torch.zeros((B, steps, H), dtype=torch.float32, device=dev)
            y = torch.full((B,), int(self.sos_id), dtype=torch.int64, device=dev) if greedy else None
            if not greedy:
                text = text.to(device=dev, dtype=torch.int64).contiguous()
            s = _lib.stream_ptr()

            def gates_cell(t, tokens, hid):
                """LSTMCell step t: reads [context | h_{t-1}] from xc[t % n], returns the buffer whose h half holds h_t"""
                cur = xc[t % len(xc)]
                hid_ptr, hid_ld = (hid.data_ptr(), out_hid.stride(0)) if hid is not None else (None, 0)
                if fused:
                    nxt = xc[(t + 1) % 2]
                    _lib.check(L.rcnn_attn_gates_cell(cur.data_ptr(), cur.stride(0), w["wcat_il"].data_ptr(), w["wcat_il"].stride(0),
                                                      w["bcat_il"].data_ptr(), w["embT_il"].data_ptr(), tokens.data_ptr(), B, H,
                                                      C + H, V, c.data_ptr(), nxt[:, C:].data_ptr(), nxt.stride(0), hid_ptr, hid_ld,
                                                      s), "rcnn_attn_gates_cell")
                    return nxt
                ops.gemm_bf16(cur, w["wcat"], w["bcat"], torch.float32, out=gates)
                _lib.check(L.rcnn_attn_cell(gates.data_ptr(), w["embT"].data_ptr(), tokens.data_ptr(), B, H, V, c.data_ptr(),
                                            cur.data_ptr(), cur.stride(0), C, hid_ptr, hid_ld, s), "rcnn_attn_cell")
                return cur

            if greedy:
                # step: [argmax of the previous step's logits +] score / context (proj_h of h_{t-1}) -> gates GEMM + cell (h_t)
                # -> ONE GEMM over [h2h | generator]: columns [0, H) = proj_h for the next step, [H, H + V) = this step's
                # logits.  3 launches per step (5 with fp32 operands), one argmax launch after the last step.
                Np = w["comb"].shape[0]
                hg = torch.empty((B, Np), dtype=torch.float32, device=dev)
                hg[:, :H] = w["h2h_b"]                                             # h2h(h_0 = 0) = its bias
                if fused and os.environ.get("RCNN_ATTN_PYLOOP", "0") != "1":
                    # the whole loop from C++ (one trip through the binding instead of 3 * steps + 1: an eager decode is then
                    # bound by the device)
                    _lib.check(L.rcnn_attn_greedy_decode(projH.data_ptr(), w["v"].data_ptr(), encb.data_ptr(), encb.stride(0),
                                                         encb.stride(1), w["wcat_il"].data_ptr(), w["bcat_il"].data_ptr(),
                                                         w["embT_il"].data_ptr(), w["comb"].data_ptr(), w["comb_b"].data_ptr(), Np,
                                                         B, T, H, C, V, steps, blank, y.data_ptr(), xc[0].data_ptr(),
                                                         xc[1].data_ptr(), c.data_ptr(), hg.data_ptr(), probs.data_ptr(),
                                                         int(chain), s), "rcnn_attn_greedy_decode")
                    return probs
                lg = hg[:, H:].data_ptr()
                L.rcnn_chain_launches(int(chain))
                try:
                    for t in range(steps):
                        score_context(xc[t % len(xc)], hg, hg.stride(0),
                                      None if t == 0 else (lg, hg.stride(0), probs[:, t - 1].data_ptr(), y.data_ptr()))
                        hbuf = gates_cell(t, y, None)
                        ops.gemm_bf16(hbuf[:, C:], w["comb"], w["comb_b"], torch.float32, out=hg)
                finally:
                    L.rcnn_chain_launches(0)
                _lib.check(L.rcnn_attn_argmax_ld(lg, hg.stride(0), B, V, blank, probs[:, steps - 1].data_ptr(), probs.stride(0),
                                                 y.data_ptr(), s), "rcnn_attn_argmax")
                return probs
            projh = torch.empty((B, H), dtype=torch.float32, device=dev)
            tokens = text[:, :steps].t().contiguous()                               # [steps, B]: row t = the step's input tokens
            L.rcnn_chain_launches(int(chain))
            try:
                for t in range(steps):
                    cur = xc[t % len(xc)]
                    ops.gemm_bf16(cur[:, C:], w["h2h"], w["h2h_b"], torch.float32, out=projh)
                    score_context(cur, projh, projh.stride(0))
                    gates_cell(t, tokens[t], out_hid[:, t])
            finally:
                L.rcnn_chain_launches(0)
            # teacher forcing: logits = generator(out_hid) in one GEMM, then the blank mask (model/model.py:146-148)
            out = ops.gemm_bf16(ops.cast_bf16_2d(out_hid.view(B * steps, H)), w["gen"], w["gen_b"], torch.float32)
            out = out.view(B, steps, V)
            if blank >= 0:
                out[:, :, blank] = -1e4
            return out

    def _train_forward(self, batch_H, text, steps):
        """model/model.py:110-148 with an autograd graph (see the module docstring)."""
        import torch.nn.functional as F
        from .model import _LinearFn
        _lib.require_cuda(batch_H, "batch_H")
        if batch_H.dim() != 3 or batch_H.shape[2] != self.input_size:
            raise RuntimeError(f"Attention expects [B, T, {self.input_size}], got {tuple(batch_H.shape)}")
        cell = self.attention_cell
        B, T, C = batch_H.shape
        H = self.hidden_size
        dev = batch_H.device
        text = text.to(device=dev, dtype=torch.int64)
        drop = self.dropout_p if self.training else 0.0
        sampling = self.sampling_prob if self.training else 0.0
        if sampling <= 0 and os.environ.get("RCNN_ATTN_FUSED_TRAIN", "1") != "0":
            # teacher forcing only (the reference's default, sampling_prob = 0): the fused forward / backward kernels
            scale = self._alpha_scale_override
            if scale is None and drop > 0:
                scale = (torch.rand((steps, B, T), device=dev) >= drop).to(torch.float32) / (1.0 - drop)
            out_hid = _TeacherForcedFn.apply(batch_H, text[:, :steps].t().contiguous(), scale, cell.i2h.weight, cell.h2h.weight,
                                             cell.h2h.bias, cell.score.weight, cell.rnn.weight_ih, cell.rnn.weight_hh,
                                             cell.rnn.bias_ih, cell.rnn.bias_hh)
            logits = _LinearFn.apply(out_hid, self.generator.weight, self.generator.bias, True)
            if self.blank_id is not None:
                mask = torch.arange(self.num_classes, device=dev) == int(self.blank_id)     # (device ops only: capturable)
                logits = logits.masked_fill(mask, -1e4)
            return logits
        enc = batch_H.float()
        projH = _LinearFn.apply(enc, cell.i2h.weight, None, True)                       # hoisted i2h(batch_H) [B,T,H]
        wcat = torch.cat([cell.rnn.weight_ih[:, :C], cell.rnn.weight_hh], 1)            # [4H, C+H]
        bcat = cell.rnn.bias_ih + cell.rnn.bias_hh
        embT = cell.rnn.weight_ih[:, C:].t()                                            # [V, 4H]: one-hot input = row gather
        v = cell.score.weight.reshape(1, 1, H)
        h = torch.zeros((B, H), dtype=torch.float32, device=dev)
        c = torch.zeros((B, H), dtype=torch.float32, device=dev)
        targets = text[:, 0]
        hids = []
        for t in range(steps):
            proj_h = _LinearFn.apply(h, cell.h2h.weight, cell.h2h.bias, True)
            e = (torch.tanh(projH + proj_h.unsqueeze(1)) * v).sum(-1)                   # [B,T]
            if self._alpha_scale_override is not None:                                  # (tests: the same mask on both paths)
                alpha = torch.softmax(e, dim=1) * self._alpha_scale_override[t]
            else:
                alpha = F.dropout(torch.softmax(e, dim=1), p=drop, training=drop > 0)
            context = (alpha.unsqueeze(2) * enc).sum(1)                                 # [B,C]
            gates = _LinearFn.apply(torch.cat([context, h], 1), wcat, bcat, True) + embT.index_select(0, targets)
            gi, gf, gg, go = gates.chunk(4, 1)
            c = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
            h = torch.sigmoid(go) * torch.tanh(c)
            hids.append(h)
            if t < steps - 1:
                if self.sampling_prob > 0 and torch.rand(1).item() < self.sampling_prob:
                    with torch.no_grad():                                               # (no gradient through the argmax)
                        out = F.dropout(h, p=drop, training=drop > 0)
                        targets = _LinearFn.apply(out, self.generator.weight, self.generator.bias, False).argmax(1)
                else:
                    targets = text[:, t + 1]
        out_hid = torch.stack(hids, 1)                                                  # [B,steps,H]
        logits = _LinearFn.apply(out_hid, self.generator.weight, self.generator.bias, True)
        if self.blank_id is not None:
            mask = torch.arange(self.num_classes, device=dev) == int(self.blank_id)
            logits = logits.masked_fill(mask, -1e4)
        return logits

    def forward(self, batch_H, text=None, is_train=True, batch_max_length=25):
        steps = batch_max_length + 1
        if not is_train:
            return self._decode(batch_H, steps, None)
        assert text is not None, "For training, `text` with <SOS> at text[:,0] is required"
        if text.shape[1] < steps:
            raise RuntimeError(f"text has {text.shape[1]} columns, {steps} steps need text[:, :{steps}]")
        needs_graph = torch.is_grad_enabled() and (batch_H.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_graph or (self.training and (self.dropout_p > 0 or self.sampling_prob > 0)):
            return self._train_forward(batch_H, text, steps)
        return self._decode(batch_H, steps, text)
