"""Generate tests/golden/*.pt from the REFERENCE's own modules executed in place (oracle.ref_stubs) -- run in the build
container only (needs /root/reference):   python -m oracle.gen_golden

Each fixture stores seeded inputs, the reference outputs, and (for cutouts) the RNG record our host logic draws for the
same seed, so that on the GPU box -- where /root/reference does not exist -- tests can still compare against what the
reference itself produced.  Files are kept small (float16 storage where exactness is not needed is NOT used: outputs are
compared tightly, so inputs are small images instead).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_stubs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CUTOUT_CASES = [
    # H, W, cs, n_over, n_inner, power, gray, seed
    (96, 96, 64, 4, 3, 5, 0.3, 0),
    (64, 128, 32, 2, 4, 5, 0.7, 1),
    (128, 64, 48, 0, 5, 2, 0.0, 2),
    (80, 80, 64, 6, 0, 5, 0.3, 3),
]


def main():
    cut, los, fun, cfg = ref_stubs.install()
    os.makedirs(OUT, exist_ok=True)
    # ---- cutouts: reference Cutouts.forward on CPU (noise drawn from the same CPU generator)
    items = []
    for (H, W, cs, no, ni, p, gp, seed) in CUTOUT_CASES:
        g = torch.Generator().manual_seed(1000 + seed)
        x = torch.tanh(torch.randn(1, 3, H, W, generator=g)) * 1.1
        torch.manual_seed(seed)
        out = cut.make_cutouts(x, cs, no, ni, p, gp)
        items.append({"args": (H, W, cs, no, ni, p, gp, seed), "x": x, "out": out})
    torch.save(items, os.path.join(OUT, "cutouts_reference.pt"))
    # ---- losses: values and autograd gradients from the reference module
    g = torch.Generator().manual_seed(7)
    x = (torch.tanh(torch.randn(2, 3, 40, 56, generator=g)) * 1.2).requires_grad_()
    tv = los.total_variational_loss(x)
    (gtv,) = torch.autograd.grad(tv.sum(), x)
    rg = los.rgb_range_loss(x)
    (grg,) = torch.autograd.grad(rg.sum(), x)
    e = torch.randn(6, 1, 48, generator=g).requires_grad_()
    t = torch.randn(1, 2, 48, generator=g)
    sp = los.square_spherical_distance_loss(e, t)
    (gsp,) = torch.autograd.grad(sp.sum(), e)
    torch.save({"x": x.detach(), "tv": tv.detach(), "gtv": gtv, "range": rg.detach(), "grange": grg, "emb": e.detach(), "txt": t, "sph": sp.detach(),
                "gsph": gsp, "clip_normalize": fun.CLIP_NORMALIZE(x.detach()[:, :, :8, :8])}, os.path.join(OUT, "losses_reference.pt"))
    # ---- Config schedules (config.py:30-38)
    C = cfg.Config
    torch.save({"over": C.num_overview_cuts_schedule, "inner": C.num_inner_cuts_schedule, "power": C.inner_cut_size_power_schedule,
                "gray": C.cut_gray_portion_schedule, "grad_threshold": C.grad_threshold, "clip_guidance_scale": C.clip_guidance_scale,
                "denoise_scale": C.denoise_scale, "num_cutout_batches": C.num_cutout_batches}, os.path.join(OUT, "config_reference.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
