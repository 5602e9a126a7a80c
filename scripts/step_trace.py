"""Kernel-level trace of ONE CUDA-graph replay of the cfg-B train step (torch.profiler / CUPTI): per-kernel
busy time, launch count and the idle gaps between consecutive kernels.  Diagnostic only (numbers taken
under a profiler are never bench values); shows which small kernels and gaps make up the part of the
step that the K1/K2/K3 kernels do not."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench

# under torchrun (WORLD_SIZE > 1) the data-parallel step is traced on rank 0
world, rank, lrank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", lrank)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
step = bench.TrainStep(dev, world, rank)
batch = [t.to(dev) for t in bench.make_batch(256, 1234 + rank)]
step(*batch)
torch.cuda.synchronize()
gstep = step.R.GraphedStep(step, batch)
for _ in range(5):
    gstep(*batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    gstep(*batch)
    torch.cuda.synchronize()
if rank != 0:
    dist.barrier()
    os._exit(0)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, evs[-1].time_range.end
agg = collections.OrderedDict()
gap_after = collections.Counter()
prev_end = None
order = []
for e in evs:
    name = e.name.replace("(anonymous namespace)::", "").replace("rcnn::", "").replace("void ", "").split("(")[0][:60]
    dur = e.time_range.end - e.time_range.start
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += dur
    if prev_end is not None:
        a[2] += max(0.0, e.time_range.start - prev_end)
    prev_end = max(prev_end or 0, e.time_range.end)
    order.append((e.time_range.start - t0, dur, name))
busy = sum(a[1] for a in agg.values())
print(f"span {t1 - t0:.1f} us, kernel busy {busy:.1f} us, {len(evs)} device activities")
print(f"{'n':>4} {'busy us':>9} {'gap-before us':>13}  kernel")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]:4d} {a[1]:9.1f} {a[2]:13.1f}  {name}")
if os.environ.get("TRACE_ORDER"):
    for st, dur, name in order:
        print(f"{st:9.1f} {dur:8.1f}  {name}")

if world > 1:
    sys.stdout.flush()
    dist.barrier()
    os._exit(0)
