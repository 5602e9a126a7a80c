"""Host-side data-parallel logic on CPU: world_size-2 gloo run of the gradient all-reducer
(rcnn-ocr_b200/dist.py) and the batch sharding used by inference.  No GPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rcnn_ocr_b200.dist import GradAllReducer, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                  # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 8),
                                torch.nn.Tanh(), torch.nn.Linear(8, 4))
    layers = [list(m.parameters()) for m in model if isinstance(m, torch.nn.Linear)]
    # one bucket per producer (layer), split further by the byte cap: buckets never span two producers
    reducer = GradAllReducer(list(model.parameters()), bucket_bytes=600, producers=layers)
    assert len(reducer.buckets) >= 3
    for flat, layout in reducer.buckets:
        owners = {next(i for i, ps in enumerate(layers) if any(q is p for q in ps)) for p, _, _ in layout}
        assert len(owners) == 1
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(12, 16, generator=g)
    y_all = torch.randn(12, 4, generator=g)
    lo, hi = shard_range(12, rank, world)
    for it in range(2):                                   # two steps: hooks must re-arm
        for p in model.parameters():
            p.grad = None
        loss = ((model(x_all[lo:hi]) - y_all[lo:hi]) ** 2).mean()
        loss.backward()
        if it == 1:
            # staged finish: the last layers' buckets first (their optimizer step could run now), then the rest
            reducer.finish(layers[2] + layers[1])
            assert all(bi in {reducer._bucket_of[p] for p in layers[0]} for _, _, bi in reducer._works)
        reducer.finish()
        assert not reducer._works
    torch.save([p.grad.clone() for p in model.parameters()], os.path.join(out_dir, f"g{rank}.pt"))
    dist.destroy_process_group()


def test_gloo_gradient_allreduce_matches_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0 = torch.load(tmp_path / "g0.pt")
    g1 = torch.load(tmp_path / "g1.pt")
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 8),
                                torch.nn.Tanh(), torch.nn.Linear(8, 4))
    g = torch.Generator().manual_seed(100)
    x_all = torch.randn(12, 16, generator=g)
    y_all = torch.randn(12, 4, generator=g)
    ((model(x_all) - y_all) ** 2).mean().backward()       # equal shards: mean of means == global mean
    for a, b, p in zip(g0, g1, model.parameters()):
        assert torch.equal(a, b)
        torch.testing.assert_close(a, p.grad, rtol=1e-5, atol=1e-7)


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 256, 4097):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
