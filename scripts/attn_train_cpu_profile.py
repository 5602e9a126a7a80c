"""Host-side cost of one eager attention training step (B = 256, 26 steps): cProfile of the launching thread, top entries."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import rcnn_ocr_b200 as R
B, T, H, V, S = 256, 64, 512, 194, 26
torch.manual_seed(0)
m = R.Attention(H, H, V, 1, 2, 0, 3, dropout_p=0.1).cuda().train()
enc = torch.randn(B, T, H, device="cuda")
text = torch.randint(4, V, (B, S + 1), device="cuda"); text[:, 0] = 1
def step():
    m.zero_grad(set_to_none=True)
    logits = m(enc.detach().requires_grad_(True), text[:, :S], is_train=True, batch_max_length=S - 1)
    loss = F.cross_entropy(logits.reshape(-1, V), text[:, 1:].reshape(-1))
    loss.backward()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20): step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(18)
