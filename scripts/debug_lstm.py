import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
from rcnn_ocr_b200 import ops
from test_lstm_gpu import _params, _hcat_oracle

def run(B, T, I, H):
    p = _params(I, H, H, seed=B + T + H)
    x = torch.randn(B, T, I, generator=torch.Generator().manual_seed(1))
    want = _hcat_oracle(x, p)
    pc = {k: v.cuda() for k, v in p.items()}
    packed = ops.lstm_pack(*[pc["rnn." + n + sfx] for sfx in ("", "_reverse")
                             for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")])
    xb = ops.cast_bf16_3d(x.cuda())
    xp = ops.gemm_bf16(xb.view(B * T, I), packed.wih_p, packed.bias_p, torch.float16)
    for rep in range(4):
        save = rep >= 2
        hcat = torch.full((B, T, 2 * H), float("nan"), dtype=torch.bfloat16, device="cuda")
        hcat, gates, cs = ops.lstm_forward(xp, packed, B, T, save, hcat=hcat)
        torch.cuda.synchronize()
        e = (hcat.float().cpu().double() - want).abs()
        e = torch.nan_to_num(e, nan=9.0)
        print(f"B={B} T={T} H={H} save={save} rep={rep} max err {e.max():.4f} nans {int(torch.isnan(hcat.float()).sum())}")
        if e.max() > 1e-2:
            bad = (e > 1e-2)
            for d in range(2):
                bd = bad[:, :, d * H:(d + 1) * H]
                # first processed step with an error
                ts = bd.any(dim=2).any(dim=0).nonzero().flatten().tolist()
                first = (min(ts) if d == 0 else max(ts)) if ts else None
                print(f"  dir {d}: first bad t (in processing order) = {first}")
                if first is not None:
                    fb = bd[:, first]
                    rows = fb.any(dim=1).nonzero().flatten()
                    print("    bad rows by warp quadrant (row%128)//32:", np.bincount(((rows % 128) // 32).numpy(), minlength=4))
                    print("    bad unit slices:", fb.any(dim=0).reshape(H // 32, 32).any(dim=1).nonzero().flatten().tolist())
                    print("    bad units within slice (count by j//8):", np.bincount((fb.any(dim=0).nonzero().flatten() % 32 // 8).numpy(), minlength=4))

for cfg in [(256, 8, 64, 512), (256, 64, 512, 512), (256, 64, 512, 256)]:
    run(*cfg)
