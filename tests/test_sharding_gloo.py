"""World-size-2 (gloo, CPU) test of the sharding arithmetic of north_star (4): every rank draws the FULL RNG record,
computes the guidance gradient of ITS slice of the cutouts with the 1/N_total weighting, and one all-reduce(sum) of the
image gradient reproduces the single-process result.  The per-slice compute here is the CPU oracle (no GPU in this
container); the CUDA path uses the same shard_range / record.slice host logic (tests/test_cond_fn_gpu.py covers it on a GPU)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _slice_grad(x_in, rec, start, stop, clip, txt, n_total, scale):
    from oracle import cutouts as OC
    from oracle import losses as OL

    x = x_in.clone().requires_grad_()
    local = rec.slice(start, stop)
    emb = clip.encode_image(OC.clip_normalize(OC.make_cutouts(x, local)))
    d = OL.square_spherical_distance_loss(emb.unsqueeze(1), txt.unsqueeze(0))
    loss = d.sum() * (scale / n_total)  # mean over ALL cutouts, not over the local slice
    return torch.autograd.grad(loss, x)[0]


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from clip_diffusion_b200.rng_record import draw_cutout_record
    from clip_diffusion_b200.sample import shard_range
    from oracle import clip_vit

    torch.manual_seed(1234)  # set_seed semantics: identical on every rank
    clip = clip_vit.OracleCLIP("test-tiny/32")
    x_in = torch.tanh(torch.randn(1, 3, 96, 96))
    txt = torch.randn(1, 64)
    rec = draw_cutout_record(96, 96, 64, 2, 5, 5, 0.3, noise="cpu")  # full record on every rank
    sizes = torch.tensor(rec.size + rec.x0 + rec.y0)
    gathered = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(gathered, sizes)
    assert all(torch.equal(g, sizes) for g in gathered), "ranks drew different crops"
    start, stop = shard_range(rec.num_cuts, rank, world)
    g = _slice_grad(x_in, rec, start, stop, clip, txt, rec.num_cuts, 8000.0)
    dist.all_reduce(g)  # the one collective of the step
    if rank == 0:
        full = _slice_grad(x_in, rec, 0, rec.num_cuts, clip, txt, rec.num_cuts, 8000.0)
        torch.save({"rel": ((g - full).norm() / full.norm()).item()}, out_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gradient_sums_to_unsharded(tmp_path, world):
    out = str(tmp_path / "res.pt")
    port = 29600 + world + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert torch.load(out)["rel"] < 1e-5
