"""N>1 host logic on CPU: world_size-2 gloo processes exercise video sharding, the bucketed gradient allreduce(mean)
and the ragged integer gather of per-frame recall match sets."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nlvsgg_b200 import dist as D
    vids = D.shard_indices(7, rank, world)
    g = torch.Generator().manual_seed(rank)
    flat = torch.randn(1000, generator=g)
    mine = flat.clone()
    D.allreduce_mean_(flat, bucket_elems=300)           # 4 buckets
    ids = torch.tensor([10 * v + f for v in vids for f in range(2 + rank)], dtype=torch.int64)
    masks = (ids.view(-1, 1, 1, 1) * torch.ones(1, 3, 3, 8, dtype=torch.int64)).to(torch.int32)
    gi, gm = D.gather_frame_results(ids, masks)
    q.put((rank, vids, mine, flat, gi, gm))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_allreduce_and_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == [0, 2, 4, 6] and out[1][1] == [1, 3, 5]          # every video on exactly one rank
    mean = (out[0][2] + out[1][2]) / 2
    assert torch.allclose(out[0][3], mean) and torch.equal(out[0][3], out[1][3])
    ids0, ids1 = out[0][4], out[1][4]
    assert torch.equal(ids0, ids1) and torch.equal(ids0, torch.sort(ids0)[0]) and ids0.numel() == 4 * 2 + 3 * 3
    assert torch.equal(out[0][5][:, 0, 0, 0].long(), ids0)               # masks travelled with their frame ids
