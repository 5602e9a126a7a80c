"""Two eager greedy decodes (B = 256, T_enc = 64, H = 512, V = 194) for `ncu -k regex:attn_score_context -s 30 -c 1`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rcnn_ocr_b200 as R
torch.manual_seed(1)
attn = R.Attention(512, 512, 194, 1, 2, 0, 3).cuda().eval()
enc = torch.randn(256, 64, 512, device="cuda")
for _ in range(2):
    attn(enc, is_train=False, batch_max_length=25)
torch.cuda.synchronize()
