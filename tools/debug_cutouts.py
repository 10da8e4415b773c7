"""Stage-by-stage error report of the cutout kernels against the oracle (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cutouts as OC
from clip_diffusion_b200.cutouts import cutouts_forward, make_cutouts_from_record
from clip_diffusion_b200.rng_record import draw_cutout_record
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_cutouts_gpu import CASES

for case in CASES:
    H, W, cs, no, ni, p, gp, seed = case
    g = torch.Generator().manual_seed(seed)
    x = torch.tanh(torch.randn(1, 3, H, W, generator=g)) * 1.1
    rec = draw_cutout_record(H, W, cs, no, ni, p, gp, generator=g, noise="cpu")
    base_ref = OC.base_cutouts(x.add(1).div(2), rec)
    base, _ = cutouts_forward(x.cuda(), rec, augment=False)
    eb = (base.cpu() - base_ref).abs()
    xr = x.clone().requires_grad_()
    ref = OC.make_cutouts(xr, rec)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(99))
    (gref,) = torch.autograd.grad((ref * w).sum(), xr)
    xc = x.cuda().requires_grad_()
    out = make_cutouts_from_record(xc, rec)
    (gout,) = torch.autograd.grad((out * w.cuda()).sum(), xc)
    e = (out.cpu() - ref).abs()
    per_cut = e.flatten(1).max(1).values
    rel = ((gout.cpu() - gref).norm() / gref.norm()).item()
    print(case, "base max %.2e | full max %.2e mean %.2e n>1e-5: %d | grad rel %.2e | perm %s flip %d gray %d" % (
        eb.max(), e.max(), e.mean(), int((e > 1e-5).sum()), rel, rec.perm, rec.flip, rec.gray))
    # gradient without jitter ops: isolate resample+affine backward
    import copy
    rec2 = copy.copy(rec)
    rec2.brightness = rec2.contrast = rec2.saturation = 1.0; rec2.hue = 0.0
