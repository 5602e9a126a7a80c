"""Feasibility: capture the whole cfg-B train step (forward, fused CTC, backward, fused Adam) in a CUDA graph."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
step = bench.TrainStep(dev, 1, 0)
step.opt = torch.optim.Adam(step.params, lr=5.1e-4, weight_decay=1.95e-5, capturable=True, fused=True)
batches = [[t.to(dev) for t in bench.make_batch(256, 1234 + i)] for i in range(4)]
static = [t.clone() for t in batches[0]]
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(3):
        loss = step(*static)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    static_loss = step(*static)
torch.cuda.synchronize()
print("captured")
def run_graph(i):
    for d, src in zip(static, batches[i % 4]):
        d.copy_(src)
    g.replay()
    return static_loss
def run_eager(i):
    return step(*batches[i % 4])
for name, fn in (("eager", run_eager), ("graph", run_graph)):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(30):
        l = fn(i).item()
    torch.cuda.synchronize()
    print(name, "ms/step with loss.item() every step:", (time.perf_counter() - t0) / 30 * 1e3, "loss", l)
