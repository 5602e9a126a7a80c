"""Small driver for ncu: a few train steps + one inference step of the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CFG["B"]     # 512: the backward kernel works on two items per group
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    batches = [[t.to(dev) for t in bench.make_batch(B, 1234 + i)] for i in range(2)]
    step = bench.TrainStep(dev, 1, 0)
    infer = bench.InferStep(step)
    for i in range(steps):
        loss = step(*batches[i % 2])
    ids, lens = infer.device(batches[0][0])
    torch.cuda.synchronize()
    print("loss", float(loss), "decoded", int(lens.sum()))

if __name__ == "__main__":
    main()
