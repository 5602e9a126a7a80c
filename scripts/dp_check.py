"""2-GPU check of the data-parallel train step: after one step from identical weights on different
shards, all ranks hold identical parameters, equal to a 1-GPU step on the concatenated batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
B = 64
step = bench.TrainStep(dev, world, rank)
batch = [t.to(dev) for t in bench.make_batch(B, 100 + rank)]
loss = step(*batch)
flat = torch.cat([p.detach().flatten() for p in step.params])
ref = flat.clone(); dist.broadcast(ref, 0)
same = bool(torch.equal(flat, ref))
# single-GPU reference on the concatenated batch (rank 0 only)
ok = True
if rank == 0:
    single = bench.TrainStep(dev, 1, 0)
    parts = [bench.make_batch(B, 100 + r) for r in range(world)]
    cat = [torch.cat([p[i] for p in parts]).to(dev) for i in range(4)]
    single(*cat)
    f1 = torch.cat([p.detach().flatten() for p in single.params])
    err = (f1 - flat).abs().max().item()
    # Adam normalises the update: compare parameters after one step (lr 5e-4) loosely
    ok = err < 2e-4
    print(f"max |param_dp - param_single| after one Adam step = {err:.3e}")
res = torch.tensor([int(same and ok)], device=dev); dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DP_CHECK", "OK" if int(res) == 1 else "FAIL", "loss", float(loss))
dist.destroy_process_group()
