"""Stand-alone roofline numbers for the HBM-bound kernels (losses, cutouts, LayerNorm) and the GEMM/attention kernels,
timed with CUDA events on the launching stream, L2 flushed between iterations (a 512 MB memset).
Usage (GPU box):  python tools/bench_kernels.py [losses] [cutouts] [layernorm] [gemm] [attention] > profiles/<name>.txt   (no argument = all)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from clip_diffusion_b200 import _lib, vit_ops
from clip_diffusion_b200.cutouts import cutouts_backward, cutouts_forward
from clip_diffusion_b200.rng_record import draw_cutout_record

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6650.0)
TF = peaks.get("bf16_tflops", 1590.0)  # burst figure: kernels timed alone
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def row(name, us, nbytes=None, flops=None):
    s = "%-58s %10.1f us" % (name, us)
    if nbytes:
        gbs = nbytes / us / 1e3
        s += "  %8.1f GB/s  %5.1f%% of %.0f (measured HBM copy)" % (gbs, 100 * gbs / HBM, HBM)
    if flops:
        tfs = flops / us / 1e6
        s += "  %8.1f TFLOP/s  %5.1f%% of %.0f (measured bf16 burst)" % (tfs, 100 * tfs / TF, TF)
    print(s, flush=True)


SECTIONS = set(sys.argv[1:])


def want(name):
    return not SECTIONS or name in SECTIONS


print("# median of 10, CUDA events, L2 flushed between iterations; algorithmic bytes/flops per DESIGN.md section 3")
P = _lib.ptr
# ---- losses
for (B, H, W) in [(1, 512, 512), (1, 768, 768), (16, 1024, 1024), (64, 1024, 1024)] if want("losses") else []:
    x = torch.tanh(torch.randn(B, 3, H, W, device="cuda")) * 1.1
    g = torch.empty_like(x); loss = torch.empty(B, device="cuda")
    nb = 2 * x.numel() * 4
    row("tv_loss value+grad [%d,3,%d,%d]" % (B, H, W), timeit(lambda: _lib.call("cg_tv_loss_fwd_bwd", P(x), B, 3, H, W, 1.0, 0, P(loss), P(g))), nb)
    row("range_loss value+grad [%d,3,%d,%d]" % (B, H, W), timeit(lambda: _lib.call("cg_range_loss_fwd_bwd", P(x), B, 3, H, W, 1.0, 0, P(loss), P(g))), nb)
for (B, H, W) in [(1, 512, 512), (1, 768, 768), (64, 1024, 1024)] if want("losses") else []:
    x = torch.tanh(torch.randn(B, 3, H, W, device="cuda")) * 1.1
    g = torch.zeros_like(x); loss2 = torch.empty(B, 2, device="cuda"); flag = torch.zeros(2, device="cuda")
    row("fused TV+range+NaN flag, grad += [%d,3,%d,%d]" % (B, H, W),
        timeit(lambda: _lib.call("cg_image_losses_fwd_bwd", P(x), B, 3, H, W, 1.0, 1.0, 1, P(loss2), P(g), P(flag))), 3 * x.numel() * 4)
for (N, E) in [(64, 768), (4096, 768), (65536, 768)] if want("losses") else []:
    e = torch.randn(N, E, device="cuda"); t = torch.randn(1, E, device="cuda"); d = torch.empty_like(e)
    row("spherical loss+grad N=%d E=%d" % (N, E), timeit(lambda: _lib.call("cg_spherical_loss_fwd_bwd", P(e), P(t), None, N, 1, E, 1.0, None, P(d))), (2 * N + 1) * E * 4)
# ---- cutouts
for (H, cs, no, ni, patch, kpad) in [(512, 224, 16, 16, 16, 768), (512, 224, 32, 32, 14, 640), (768, 336, 64, 64, 14, 640), (768, 336, 256, 256, 14, 640)] if want("cutouts") else []:
    x = torch.tanh(torch.randn(1, 3, H, H, device="cuda"))
    rec = draw_cutout_record(H, H, cs, no, ni, 5, 0.3, generator=torch.Generator().manual_seed(0), noise="device")
    rec.noise_seed = 7
    n = no + ni
    out_bytes = n * (cs // patch) ** 2 * kpad * 2
    state = {}

    def fwd():
        state["o"], state["ctx"] = cutouts_forward(x, rec, fmt=_lib.CG_FMT_BF16_PATCH, patch=patch, kpad=kpad, normalize=True)
    us = timeit(fwd)
    row("cutouts fwd  %d^2 -> %d x %d^2 bf16 patch-major (%d+%d)" % (H, n, cs, no, ni), us, 12 * H * H + out_bytes)
    dout = torch.randn_like(state["o"]); gx = torch.zeros(3, H, H, device="cuda")
    us = timeit(lambda: cutouts_backward(dout, state["ctx"], 1.0, gx))
    row("cutouts bwd  %d x %d^2 -> %d^2" % (n, cs, H), us, 12 * H * H + out_bytes)
# ---- LayerNorm
for (M, D) in [(6304, 768), (16448, 1024), (131072, 1024)] if want("layernorm") else []:
    x = torch.randn(M, D, device="cuda"); gam = torch.ones(D, device="cuda"); bet = torch.zeros(D, device="cuda")
    y = torch.empty(M, D, device="cuda", dtype=torch.bfloat16); mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    row("layernorm fwd [%d,%d] f32 -> bf16" % (M, D), timeit(lambda: _lib.call("cg_layernorm_fwd", P(x), P(gam), P(bet), M, D, D, P(y), None, P(mean), P(rstd))), M * D * 6)
    dy = torch.randn(M, D, device="cuda"); dx = torch.zeros(M, D, device="cuda"); dxb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    row("layernorm bwd [%d,%d] (+= dx, bf16 copy)" % (M, D), timeit(lambda: _lib.call("cg_layernorm_bwd", P(dy), P(x), P(gam), P(mean), P(rstd), M, D, D, 1, P(dx), P(dxb))), M * D * 18)
# ---- GEMM
for (M, N, K, epi, nm) in [(16448, 1024, 4096, _lib.EPI_F32, "f32 out"), (16448, 3072, 1024, _lib.EPI_BIAS_BF16, "bias->bf16 (qkv)"),
                           (16448, 4096, 1024, _lib.EPI_BIAS_QGELU_BF16, "bias+QuickGELU (c_fc)"), (6304, 2304, 768, _lib.EPI_BIAS_BF16, "bias->bf16 (B/16 qkv)"),
                           (36928, 1024, 4096, _lib.EPI_F32, "f32 out (L/14@336 x 64)")] if want("gemm") else []:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16(); bias = torch.zeros(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if epi == _lib.EPI_F32 else torch.bfloat16)
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    row("gemm %dx%dx%d %s" % (M, N, K, nm), timeit(lambda: vit_ops.gemm_bf16_tn(a, b, epi, bias=bias, out=out, aux=aux)), flops=2.0 * M * N * K)
# ---- attention
for (n, T, heads) in [(32, 197, 12), (64, 257, 16), (8, 257, 16), (64, 50, 12), (16, 577, 16)] if want("attention") else []:
    D = heads * 64
    qkv = torch.randn(n * T, 3 * D, device="cuda").bfloat16(); ctx = torch.empty(n * T, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(n, heads, T, device="cuda"); dctx = torch.randn(n * T, D, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
    row("attention fwd n=%d T=%d heads=%d" % (n, T, heads), timeit(lambda: _lib.call("cg_attention_fwd", P(qkv), n, T, heads, P(ctx), P(lse))), flops=4.0 * n * heads * T * T * 64)
    row("attention bwd n=%d T=%d heads=%d" % (n, T, heads), timeit(lambda: _lib.call("cg_attention_bwd", P(qkv), P(ctx), P(dctx), P(lse), n, T, heads, P(dqkv), P(delta))),
        flops=8.0 * n * heads * T * T * 64)
