import sys, os, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import cutouts as OC
from clip_diffusion_b200.cutouts import make_cutouts_from_record
from clip_diffusion_b200.rng_record import draw_cutout_record
from test_cutouts_gpu import CASES

def run(x, rec, tag):
    xr = x.clone().requires_grad_()
    ref = OC.make_cutouts(xr, rec)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(99))
    (gref,) = torch.autograd.grad((ref * w).sum(), xr)
    xc = x.cuda().requires_grad_()
    out = make_cutouts_from_record(xc, rec)
    (gout,) = torch.autograd.grad((out * w.cuda()).sum(), xc)
    d = (gout.cpu() - gref).abs()
    rel = (d.norm() / gref.norm()).item()
    big = (d > 1e-3 * gref.abs().max()).sum().item()
    print(tag, "grad rel %.2e maxdiff %.2e max|g| %.2e  n_big %d  fwd max %.2e" % (rel, d.max(), gref.abs().max(), big, (out.detach().cpu()-ref.detach()).abs().max()))
    return d

H, W, cs, no, ni, p, gp, seed = CASES[4]
g = torch.Generator().manual_seed(seed)
x = torch.tanh(torch.randn(1, 3, H, W, generator=g)) * 1.1
rec = draw_cutout_record(H, W, cs, no, ni, p, gp, generator=g, noise="cpu")
d = run(x, rec, "full")
idx = torch.topk(d.flatten(), 8).indices
print("top diffs at", [(int(i) // (H * W), (int(i) % (H * W)) // W, int(i) % W, float(d.flatten()[i])) for i in idx])
for n in range(rec.num_cuts):
    r = rec.slice(n, n + 1)
    r.noise = [t[n:n+1] for t in rec.noise]
    run(x, r, "cut %d size %d flags %d" % (n, rec.size[n], rec.flags[n]))
    if n == 0:
        n = 15
