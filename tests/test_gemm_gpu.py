"""GPU parity: the tcgen05/TMEM/TMA GEMM and its fused epilogues against a plain PyTorch fp32 reference on the
same bf16-rounded operands (floating-point kernel => torch fp32 reference, tolerance stated per case)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from clip_diffusion_b200 import _lib, vit_ops

    return _lib, vit_ops


def _mk(M, N, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) * (K ** -0.5)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) * 0.1
    ref = a.float() @ b.float().t()
    return a, b, bias, ref


SHAPES = [(128, 128, 64), (128, 256, 128), (200, 384, 192), (1576, 2304, 768), (6304, 768, 3072), (50, 128, 64), (16448, 1024, 640), (777, 512, 1024)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_plain_f32(M, N, K):
    _lib, ops = _ops()
    a, b, bias, ref = _mk(M, N, K)
    out = ops.gemm_bf16_tn(a, b, _lib.EPI_F32)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), err  # fp32 accumulation, different summation order


@pytest.mark.parametrize("M,N,K", SHAPES[:5])
def test_gemm_epilogues(M, N, K):
    _lib, ops = _ops()
    a, b, bias, ref = _mk(M, N, K, seed=1)
    # bias -> bf16
    out = ops.gemm_bf16_tn(a, b, _lib.EPI_BIAS_BF16, bias=bias)
    exp = (ref + bias).bfloat16().float()
    assert (out.float() - exp).abs().max().item() <= 2e-2 * max(1.0, exp.abs().max().item())
    # bias + residual (fp32, in place)
    resid = torch.randn(M, N, device="cuda")
    r0 = resid.clone()
    out2 = ops.gemm_bf16_tn(a, b, _lib.EPI_BIAS_RESID_F32, bias=bias, aux=resid)  # out of place
    assert torch.equal(resid, r0)
    assert (out2 - (r0 + ref + bias)).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
    ops.gemm_bf16_tn(a, b, _lib.EPI_BIAS_RESID_F32, bias=bias, out=resid, aux=resid)  # in place
    assert (resid - (r0 + ref + bias)).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
    # bias + QuickGELU (+ pre-activation)
    h, u = ops.gemm_bf16_tn(a, b, _lib.EPI_BIAS_QGELU_BF16, bias=bias)
    pre = ref + bias
    assert (u.float() - pre).abs().max().item() <= 2e-2 * max(1.0, pre.abs().max().item())
    assert (h.float() - pre * torch.sigmoid(1.702 * pre)).abs().max().item() <= 2e-2 * max(1.0, pre.abs().max().item())
    # dgrad through QuickGELU
    s = torch.sigmoid(1.702 * u.float())
    exp = ref * (s * (1 + 1.702 * u.float() * (1 - s)))
    dq = ops.gemm_bf16_tn(a, b, _lib.EPI_DQGELU_BF16, aux=u)
    assert (dq.float() - exp).abs().max().item() <= 2e-2 * max(1.0, exp.abs().max().item())
    # plain bf16
    out = ops.gemm_bf16_tn(a, b, _lib.EPI_BF16)
    assert (out.float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())


def test_gemm_patch_pos_epilogue():
    _lib, ops = _ops()
    n_img, g2, D, K = 5, 49, 256, 192
    a, b, _, ref = _mk(n_img * g2, D, K, seed=2)
    pos = torch.randn(g2 + 1, D, device="cuda")
    x = torch.full((n_img * (g2 + 1), D), 7.0, device="cuda")
    ops.gemm_bf16_tn(a, b, _lib.EPI_PATCH_POS_F32, out=x, pos=pos, g2=g2)
    x = x.view(n_img, g2 + 1, D)
    assert (x[:, 0] == 7.0).all()  # class-token rows untouched
    exp = ref.view(n_img, g2, D) + pos[1:]
    assert (x[:, 1:] - exp).abs().max().item() <= 1e-3 * max(1.0, exp.abs().max().item())


def test_gemm_strided_operands_and_repeat():
    """Leading dimensions larger than K (column slices of a wider buffer), called twice (tensor-map cache)."""
    _lib, ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    big = (torch.randn(300, 512, device="cuda", generator=g) * 0.5).bfloat16()
    a = big[:, 128:384]
    b = (torch.randn(128, 256, device="cuda", generator=g) / 16).bfloat16()
    ref = a.float() @ b.float().t()
    for _ in range(2):
        out = ops.gemm_bf16_tn(a, b, _lib.EPI_F32)
        assert (out - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())


def test_gemm_rejects_bad_shapes():
    _lib, ops = _ops()
    a = torch.zeros(16, 48, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(128, 48, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(_lib.ClipGuideError):
        ops.gemm_bf16_tn(a, b, _lib.EPI_F32)


@pytest.mark.parametrize("force", [{"CG_GEMM_PAIR": "1"}, {"CG_GEMM_PAIR": "0", "CG_GEMM_BN": "128"}, {"CG_GEMM_PAIR": "0", "CG_GEMM_BN": "256"}],
                         ids=["pair256x256", "bn128", "bn256"])
def test_gemm_forced_tile_variant_in_subprocess(force):
    """The tile variant is chosen per problem (choose_tile); CG_GEMM_PAIR / CG_GEMM_BN force one at library load.  Run the same parity
    cases on EVERY variant in a child process with that environment (and a timeout: a cluster deadlock must not hang the suite)."""
    import os
    import subprocess
    import sys

    code = r'''
import sys, torch
sys.path.insert(0, %r)
from clip_diffusion_b200 import _lib, vit_ops
torch.manual_seed(0)
worst = 0.0
for (M, N, K) in [(300, 256, 64), (1576, 2304, 768), (6304, 768, 3072), (16448, 1024, 640), (777, 512, 1024), (257, 256, 128)]:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda") * 0.1
    ref = a.float() @ b.float().t()
    out = vit_ops.gemm_bf16_tn(a, b, _lib.EPI_F32)
    worst = max(worst, ((out - ref).abs().max() / ref.abs().max().clamp_min(1)).item())
    h, u = vit_ops.gemm_bf16_tn(a, b, _lib.EPI_BIAS_QGELU_BF16, bias=bias)
    pre = ref + bias
    worst = max(worst, ((u.float() - pre).abs().max() / pre.abs().max().clamp_min(1)).item() / 20)
    r0 = torch.randn(M, N, device="cuda")
    o2 = vit_ops.gemm_bf16_tn(a, b, _lib.EPI_BIAS_RESID_F32, bias=bias, aux=r0)
    worst = max(worst, ((o2 - (r0 + pre)).abs().max() / pre.abs().max().clamp_min(1)).item())
torch.cuda.synchronize()
print("WORST", worst)
assert worst < 1e-3, worst
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, **force)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "WORST" in res.stdout
