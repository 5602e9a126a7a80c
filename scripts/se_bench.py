"""SE gate kernel alone at the cfg-1 shapes: time per launch (50 launches per graph) for RCNN_SE_SLICES = 1, 2, 4, 8 CTAs per image."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rcnn_ocr_b200 as R
L = R.lib()
for (B, C, H, W, Cr) in ((32, 256, 8, 32, 16), (32, 512, 4, 16, 32)):
    y = torch.randn(B, C, H, W, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    w1 = torch.randn(Cr, C, device="cuda"); w2 = torch.randn(Cr, C, device="cuda"); yb = torch.randn(C, device="cuda")
    gate = torch.empty(B, C, device="cuda")
    ws = torch.zeros(int(L.rcnn_se_gate_workspace_bytes(B, C)), dtype=torch.uint8, device="cuda")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        call = lambda: L.rcnn_se_gate(y.data_ptr(), 1, B, H * W, C, w1.data_ptr(), w2.data_ptr(), Cr, yb.data_ptr(), gate.data_ptr(),
                                      ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        for _ in range(3): assert call() == 0
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(50): call()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3 / 50)
    print(f"slices={os.environ.get('RCNN_SE_SLICES', 'auto')} B={B} C={C} HW={H*W}: {np.median(ts):.2f} us per launch")
