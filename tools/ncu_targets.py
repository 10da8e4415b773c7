"""A short program that launches each HBM-bound kernel once at a saturating size (after a warm-up launch), for
`ncu --set full -k regex:... ` captures (profiles/).  Not a benchmark."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_diffusion_b200 import _lib
from clip_diffusion_b200.cutouts import cutouts_backward, cutouts_forward
from clip_diffusion_b200.rng_record import draw_cutout_record

P = _lib.ptr
B, H, W = 64, 1024, 1024
x = torch.tanh(torch.randn(B, 3, H, W, device="cuda")) * 1.1
g = torch.empty_like(x); loss = torch.empty(B, device="cuda")
for _ in range(2):
    _lib.call("cg_tv_loss_fwd_bwd", P(x), B, 3, H, W, 1.0, 0, P(loss), P(g))
    _lib.call("cg_range_loss_fwd_bwd", P(x), B, 3, H, W, 1.0, 0, P(loss), P(g))
N, E = 65536, 768
e = torch.randn(N, E, device="cuda"); t = torch.randn(1, E, device="cuda"); d = torch.empty_like(e)
for _ in range(2):
    _lib.call("cg_spherical_loss_fwd_bwd", P(e), P(t), None, N, 1, E, 1.0, None, P(d))
M, D = 131072, 1024
xx = torch.randn(M, D, device="cuda"); gam = torch.ones(D, device="cuda"); bet = torch.zeros(D, device="cuda")
y = torch.empty(M, D, device="cuda", dtype=torch.bfloat16); mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
dy = torch.randn(M, D, device="cuda"); dx = torch.zeros(M, D, device="cuda"); dxb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    _lib.call("cg_layernorm_fwd", P(xx), P(gam), P(bet), M, D, D, P(y), None, P(mean), P(rstd))
    _lib.call("cg_layernorm_bwd", P(dy), P(xx), P(gam), P(mean), P(rstd), M, D, D, 1, P(dx), P(dxb))
img = torch.tanh(torch.randn(1, 3, 512, 512, device="cuda"))
rec = draw_cutout_record(512, 512, 224, 32, 32, 5, 0.3, generator=torch.Generator().manual_seed(0), noise="device")
rec.noise_seed = 7
for _ in range(2):
    o, ctx = cutouts_forward(img, rec, fmt=_lib.CG_FMT_BF16_PATCH, patch=14, kpad=640, normalize=True)
    cutouts_backward(torch.randn_like(o), ctx, 1.0, torch.zeros(3, 512, 512, device="cuda"))
torch.cuda.synchronize()
print("done")
