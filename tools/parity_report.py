"""Parity numbers at REAL tower sizes (random init): embedding cosine and input-gradient relative L2 of the B200 path
(bf16 operands, fp32 accumulate/residual) against the fp32 CPU oracle.  north_star: cosine >= 0.999, rel-L2 <= 1e-2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_diffusion_b200 import models
from clip_diffusion_b200.utils.functional import embed_image
from oracle.clip_vit import OracleCLIP
from oracle.cutouts import clip_normalize

torch.set_num_threads(os.cpu_count() or 1)
print("# tower, images, min cosine(emb, oracle), rel-L2(d loss/d image), oracle CPU seconds")
for name, n in [("ViT-B/32", 4), ("ViT-B/16", 4), ("ViT-L/14", 2), ("ViT-L/14@336px", 1)]:
    sd = models.random_clip_state_dict(name, seed=1)
    mine = models.CLIPModelB200(name, sd, "cuda")
    ref = OracleCLIP(name, state_dict=sd)
    res = mine.visual.input_resolution
    g = torch.Generator().manual_seed(0)
    img = torch.rand(n, 3, res, res, generator=g)
    txt = torch.randn(1, mine.visual.output_dim, generator=g)
    t0 = time.time()
    xr = img.clone().requires_grad_()
    er = ref.encode_image(clip_normalize(xr))
    # the guidance objective: spherical distance to a text embedding
    lr = (torch.nn.functional.normalize(er, dim=-1) - torch.nn.functional.normalize(txt, dim=-1)).norm(dim=-1).div(2).arcsin().pow(2).mul(2).sum()
    (gr,) = torch.autograd.grad(lr, xr)
    dt = time.time() - t0
    xc = img.cuda().requires_grad_()
    em = embed_image(mine, xc)
    tc = txt.cuda()
    lm = (torch.nn.functional.normalize(em, dim=-1) - torch.nn.functional.normalize(tc, dim=-1)).norm(dim=-1).div(2).arcsin().pow(2).mul(2).sum()
    (gm,) = torch.autograd.grad(lm, xc)
    cos = torch.nn.functional.cosine_similarity(em.detach().cpu(), er.detach(), dim=-1).min().item()
    rel = ((gm.cpu() - gr).norm() / gr.norm()).item()
    print("%-16s n=%d  cos_min=%.6f  grad_rel_l2=%.3e  oracle %.1fs" % (name, n, cos, rel, dt), flush=True)
    del mine, ref
    torch.cuda.empty_cache()
