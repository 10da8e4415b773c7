"""A short program that launches the ViT tensor-core kernels once each at the target shapes (ViT-L/14 x 64 cutouts) after a warm-up
launch, for `ncu --set full -k regex:...` captures (profiles/).  Not a benchmark."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from clip_diffusion_b200 import _lib, vit_ops

P = _lib.ptr
M, D = 16448, 1024
for (N, K, epi) in [(3 * D, D, _lib.EPI_BIAS_BF16), (4 * D, D, _lib.EPI_BIAS_QGELU_BF16), (D, 4 * D, _lib.EPI_BIAS_RESID_F32), (D, 4 * D, _lib.EPI_F32)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); b = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16(); bias = torch.zeros(N, device="cuda")
    f32 = epi in (_lib.EPI_BIAS_RESID_F32, _lib.EPI_F32)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    aux = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi == _lib.EPI_BIAS_RESID_F32 else torch.bfloat16)
    for _ in range(2):
        vit_ops.gemm_bf16_tn(a, b, epi, bias=bias, out=out, aux=aux)
n, T, heads = 64, 257, 16
qkv = torch.randn(n * T, 3 * D, device="cuda").bfloat16(); ctx = torch.empty(n * T, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(n, heads, T, device="cuda"); dctx = torch.randn(n * T, D, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv); delta = torch.empty_like(lse)
for _ in range(2):
    _lib.call("cg_attention_fwd", P(qkv), n, T, heads, P(ctx), P(lse))
    _lib.call("cg_attention_bwd", P(qkv), P(ctx), P(dctx), P(lse), n, T, heads, P(dqkv), P(delta))
torch.cuda.synchronize()
print("done")
