import sys, torch
sys.path.insert(0, "/root/repo")
from clip_diffusion_b200 import _lib
P = _lib.ptr
for (n, T, heads) in [(1, 257, 16), (3, 257, 4), (1, 272, 1), (2, 256, 2)]:
    D = heads * 64
    g = torch.Generator().manual_seed(T + heads)
    qkv = (torch.randn(n * T, 3 * D, generator=g) * 0.8).bfloat16()
    dctx = (torch.randn(n * T, D, generator=g) * 0.5).bfloat16()
    q, k, v = [t.float().view(n, T, heads, 64).transpose(1, 2).requires_grad_() for t in qkv.split(D, dim=1)]
    s = (q @ k.transpose(-1, -2)) * 0.125
    ref = torch.softmax(s, -1) @ v
    do = dctx.float().view(n, T, heads, 64).transpose(1, 2)
    grads = torch.autograd.grad((ref * do).sum(), (q, k, v))
    qc, dc = qkv.cuda(), dctx.cuda()
    ctx = torch.full((n * T, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((n, heads, T), float("nan"), device="cuda")
    _lib.call("cg_attention_fwd", P(qc), n, T, heads, P(ctx), P(lse))
    for rep in range(3):
        dqkv = torch.full((n * T, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
        delta = torch.empty(n, heads, T, device="cuda")
        _lib.call("cg_attention_bwd", P(qc), P(ctx), P(dc), P(lse), n, T, heads, P(dqkv), P(delta))
        torch.cuda.synchronize()
        got = [t.float().cpu().view(n, T, heads, 64).transpose(1, 2) for t in dqkv.split(D, dim=1)]
        msg = []
        for name, a, b in zip("qkv", got, grads):
            per_tile = []
            for lo in range(0, T, 128):
                hi = min(T, lo + 128)
                per_tile.append("%.2e" % ((a[:, :, lo:hi] - b[:, :, lo:hi]).norm() / b[:, :, lo:hi].norm()).item())
            msg.append("d%s[%s]" % (name, " ".join(per_tile)))
        print((n, T, heads), rep, "  ".join(msg), flush=True)
