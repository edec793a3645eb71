"""Data parallelism on real GPUs (needs 2; skipped otherwise): a 2-rank NCCL step — gradients of the temporal layers
all-reduced under the rest of the backward pass, the rest after it, the 1/world folded into AdamW — leaves the parameters
where ONE GPU stepping on the union of the two ranks' videos leaves them."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q, overlap):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from oracle import cref
    from nlvsgg_b200 import model as M, shapes, synth
    from nlvsgg_b200.trainer import Trainer
    sd = synth.make_state_dict(shapes.sttran_template(), 2)
    videos = [synth.synth_video(300 + i, 5 + i, 5, "sgdet", draw_fn=cref.draw_union_boxes)[0] for i in range(2 * world)]
    mine = videos[rank * 2:rank * 2 + 2]                      # two videos per rank: the mean over ranks of per-rank means is the global mean
    tr = Trainer({k: v.to(dev) for k, v in sd.items()}, "sgdet", "sttran", "fp32", lr=1e-3, device=dev)
    tr.overlap_allreduce = overlap
    losses = [float(tr.step(M.upload(M.collate(mine, "sgdet"), dev, rasterise=False))) for _ in range(2)]
    torch.cuda.synchronize()
    hooked = tr._tail_off is not None
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:                                             # the same two steps on one GPU over all four videos
        one = Trainer({k: v.to(dev) for k, v in sd.items()}, "sgdet", "sttran", "fp32", lr=1e-3, device=dev)
        ref_losses = [float(one.step(M.upload(M.collate(videos, "sgdet"), dev, rasterise=False))) for _ in range(2)]
        moved = 0.0
        err = (tr.flat_p - one.flat_p).abs().max().item()
        scale = one.flat_p.abs().max().item()
        q.put((err, scale, losses, ref_losses, hooked, moved))
    else:
        q.put(("rank1", losses))


@pytest.mark.parametrize("overlap", [True, False])
def test_two_gpu_step_equals_one_gpu_step_on_the_union_batch(overlap):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q, overlap)) for r in range(world)]
    for p in ps:
        p.start()
    out = [q.get(timeout=600) for _ in range(world)]
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0 = [o for o in out if o[0] != "rank1"][0]
    r1 = [o for o in out if o[0] == "rank1"][0]
    err, scale, losses, ref_losses, hooked, _ = r0
    assert hooked
    assert err <= 2e-5 * scale, (err, scale)                   # parameters after two DP steps == after two single-GPU steps
    # the single-GPU loss is the mean over the four videos = the mean of the two ranks' (two-video) losses
    for s in range(2):
        assert abs(0.5 * (losses[s] + r1[1][s]) - ref_losses[s]) <= 1e-4 * abs(ref_losses[s])
