"""GPU parity: loss kernels (value + analytic gradient) against the CPU oracle (oracle/losses.py =
restatement of clip_diffusion/losses.py:10-45, gradients from torch autograd).  Calls go through the
C ABI (ctypes) via the Python mirror of the reference interface."""
import pytest
import torch

from oracle import losses as O

pytestmark = pytest.mark.gpu

TOL_VALUE = 2e-6  # relative, fp32 reductions in a different order
TOL_GRAD = 1e-5   # relative L2


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("shape", [(1, 3, 256, 256), (1, 3, 512, 512), (2, 3, 64, 96), (1, 3, 67, 53), (1, 1, 1, 1), (1, 3, 768, 512)])
@pytest.mark.parametrize("which", ["tv", "range"])
def test_image_losses(shape, which):
    from clip_diffusion_b200 import losses as L

    g = torch.Generator().manual_seed(hash(shape) % 1000)
    x = (torch.tanh(torch.randn(shape, generator=g)) * 1.1).requires_grad_()
    ofn, mfn = (O.total_variational_loss, L.total_variational_loss) if which == "tv" else (O.rgb_range_loss, L.rgb_range_loss)
    ref = ofn(x)
    w = torch.arange(1, shape[0] + 1, dtype=torch.float32)
    (gref,) = torch.autograd.grad((ref * w).sum(), x)
    xc = x.detach().cuda().requires_grad_()
    out = mfn(xc)
    (gout,) = torch.autograd.grad((out * w.cuda()).sum(), xc)
    assert out.shape == ref.shape
    assert torch.allclose(out.cpu(), ref, rtol=TOL_VALUE * 10, atol=1e-9)
    if gref.norm() > 0:
        assert _rel(gout.cpu(), gref) < TOL_GRAD
    else:
        assert gout.abs().max().item() == 0


def test_north_star_aliases():
    from clip_diffusion_b200 import losses as L

    assert L.tv_loss is L.total_variational_loss and L.range_loss is L.rgb_range_loss
    assert L.spherical_dist_loss is L.square_spherical_distance_loss


@pytest.mark.parametrize("N,P,E", [(16, 1, 512), (64, 1, 768), (5, 3, 64), (1, 1, 8), (32, 2, 1024)])
def test_spherical(N, P, E):
    from clip_diffusion_b200 import losses as L

    g = torch.Generator().manual_seed(N * 7 + E)
    x = torch.randn(N, E, generator=g).requires_grad_()
    y = torch.randn(P, E, generator=g)
    ref = O.square_spherical_distance_loss(x.unsqueeze(1), y.unsqueeze(0))
    w = torch.randn(N, P, generator=g)
    (gref,) = torch.autograd.grad((ref * w).sum(), x)
    xc = x.detach().cuda().requires_grad_()
    out = L.square_spherical_distance_loss(xc.unsqueeze(1), y.cuda().unsqueeze(0))
    (gout,) = torch.autograd.grad((out * w.cuda()).sum(), xc)
    assert out.shape == ref.shape
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-6)
    assert _rel(gout.cpu(), gref) < 1e-5


def test_spherical_identical_and_opposite():
    from clip_diffusion_b200 import losses as L

    x = torch.randn(4, 32)
    same = L.square_spherical_distance_loss(x.cuda().unsqueeze(1), x[:1].cuda().unsqueeze(0)).cpu()
    ref = O.square_spherical_distance_loss(x.unsqueeze(1), x[:1].unsqueeze(0))
    assert torch.allclose(same, ref, atol=1e-6)
    assert abs(same[0, 0].item()) < 1e-6


def test_cpu_tensor_is_refused():
    from clip_diffusion_b200 import losses as L
    from clip_diffusion_b200._lib import ClipGuideError

    with pytest.raises(ClipGuideError):
        L.total_variational_loss(torch.zeros(1, 3, 8, 8))


@pytest.mark.parametrize("shape,q", [((1, 3, 512, 512), 0.995), ((2, 3, 64, 96), 0.995), ((1, 3, 37, 53), 0.5), ((3, 1000), 0.999), ((1, 3, 256, 256), 1.0),
                                     ((1, 7), 0.3)])
def test_dynamic_threshold_matches_torch_quantile(shape, q):
    """sample.py:116-132 (Imagen dynamic thresholding): radix-select kernel vs torch.quantile + clamp/div."""
    from clip_diffusion_b200 import _lib

    g = torch.Generator().manual_seed(len(shape) * 131 + shape[-1])
    x = (torch.randn(shape, generator=g) * 1.5).cuda()
    if shape == (1, 3, 37, 53):
        x[0, 0, :5] = 0.75  # ties around the selected rank
    b, n = shape[0], x.numel() // shape[0]
    thr_ref = torch.quantile(x.reshape(b, -1).abs(), q, dim=-1).clamp(min=1.0)
    ref = x.clamp(min=-thr_ref.view(-1, *([1] * (x.dim() - 1))), max=thr_ref.view(-1, *([1] * (x.dim() - 1)))) / thr_ref.view(-1, *([1] * (x.dim() - 1)))
    out = torch.empty_like(x)
    thr = torch.empty(b, device="cuda")
    ws = torch.empty(_lib.load().cg_dynamic_threshold_workspace_bytes(b), dtype=torch.uint8, device="cuda")
    _lib.call("cg_dynamic_threshold", _lib.ptr(x), b, n, q, 1.0, _lib.ptr(out), _lib.ptr(thr), _lib.ptr(ws))
    assert torch.allclose(thr, thr_ref, rtol=2e-7, atol=0), (thr, thr_ref)
    assert (out - ref).abs().max().item() <= 1e-6


def test_make_denoised_function_uses_the_kernel_and_matches_the_torch_path():
    from clip_diffusion_b200.sample import make_denoised_function

    f = make_denoised_function(0.995)
    x = torch.randn(1, 3, 128, 128, device="cuda") * 2
    a = f(x)
    b = f(x.double()).float()  # non-fp32 input takes the torch.quantile path
    assert (a - b).abs().max().item() <= 2e-6


@pytest.mark.parametrize("shape", [(1, 3, 256, 256), (1, 3, 512, 512), (2, 3, 96, 128), (1, 3, 768, 512)])
def test_fused_image_losses_match_the_oracle_and_are_deterministic(shape):
    """cg_image_losses_fwd_bwd: TV + range value and gradient + the NaN flag of the finished gradient in ONE launch
    (sample.py:217-228), against the oracle's autograd; values / gradients bit-identical run to run (no float atomics)."""
    from clip_diffusion_b200 import _lib

    B, C, H, W = shape
    g = torch.Generator().manual_seed(H + W)
    x = (torch.tanh(torch.randn(shape, generator=g)) * 1.2).requires_grad_()
    tv_scale, rg_scale = 10000.0, 150.0
    tv, rg = O.total_variational_loss(x), O.rgb_range_loss(x)
    (gref,) = torch.autograd.grad(tv.sum() * tv_scale + rg.sum() * rg_scale, x)
    base = torch.randn(shape, generator=g)
    xc = x.detach().cuda().contiguous()
    outs = []
    for rep in range(2):
        grad = base.cuda().clone()
        loss2 = torch.full((B, 2), float("nan"), device="cuda")
        flag = torch.full((2,), 7.0, device="cuda")
        _lib.call("cg_image_losses_fwd_bwd", _lib.ptr(xc), B, C, H, W, tv_scale, rg_scale, 1, _lib.ptr(loss2), _lib.ptr(grad), _lib.ptr(flag))
        outs.append((grad.cpu(), loss2.cpu(), flag.cpu()))
    grad, loss2, flag = outs[0]
    assert _rel(grad - base, gref) < TOL_GRAD
    assert torch.allclose(loss2[:, 0], tv.detach(), rtol=1e-5) and torch.allclose(loss2[:, 1], rg.detach(), rtol=1e-5, atol=1e-12)
    assert flag[0].item() == 0.0
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])  # deterministic
    # NaN anywhere in the finished gradient raises the flag (the reference's isnan(grad_tensor).any(), sample.py:228)
    bad = base.clone()
    bad[0, 1, 5, 7] = float("nan")
    grad = bad.cuda()
    flag = torch.zeros(2, device="cuda")
    _lib.call("cg_image_losses_fwd_bwd", _lib.ptr(xc), B, C, H, W, tv_scale, rg_scale, 1, None, _lib.ptr(grad), _lib.ptr(flag))
    assert flag[0].item() == 1.0
    # accumulate = 0 overwrites
    grad = torch.full(shape, float("nan"), device="cuda")
    _lib.call("cg_image_losses_fwd_bwd", _lib.ptr(xc), B, C, H, W, tv_scale, rg_scale, 0, None, _lib.ptr(grad), _lib.ptr(flag))
    assert _rel(grad.cpu(), gref) < TOL_GRAD and flag[0].item() == 0.0


def test_loss_values_and_clamp_factor_are_bit_reproducible():
    """loss values and the RMS of the guidance gradient were summed with float atomicAdd.  Now: per-block partials added in
    index order by the last block."""
    from clip_diffusion_b200 import _lib

    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 3, 512, 512, generator=g).cuda()
    vals = []
    for rep in range(3):
        loss = torch.empty(1, device="cuda")
        grad = torch.empty_like(x)
        _lib.call("cg_tv_loss_fwd_bwd", _lib.ptr(x), 1, 3, 512, 512, 1.0, 0, _lib.ptr(loss), _lib.ptr(grad))
        lr = torch.empty(1, device="cuda")
        _lib.call("cg_range_loss_fwd_bwd", _lib.ptr(x), 1, 3, 512, 512, 1.0, 0, _lib.ptr(lr), _lib.ptr(grad))
        out = torch.empty_like(x)
        scratch = torch.zeros(2, device="cuda")
        _lib.call("cg_grad_finalize", _lib.ptr(x), x.numel(), -1.0, 0.05, None, _lib.ptr(out), _lib.ptr(scratch))
        emb = torch.randn(64, 768, generator=torch.Generator().manual_seed(2)).cuda()
        txt = torch.randn(1, 768, generator=torch.Generator().manual_seed(3)).cuda()
        ls = torch.empty(1, device="cuda")
        demb = torch.empty_like(emb)
        _lib.call("cg_spherical_loss_fwd_bwd", _lib.ptr(emb), _lib.ptr(txt), None, 64, 1, 768, 2.0, _lib.ptr(ls), _lib.ptr(demb))
        vals.append((loss.item(), lr.item(), scratch[0].item(), ls.item(), out.cpu()))
    for v in vals[1:]:
        assert v[:4] == vals[0][:4] and torch.equal(v[4], vals[0][4])
    assert abs(vals[0][2] - float((x.double() ** 2).sum())) / vals[0][2] < 1e-5


@pytest.mark.parametrize("H,W", [(192, 192), (256, 320), (512, 512)])
def test_ms_ssim_dissimilarity_kernel_matches_the_float64_oracle(H, W):
    """N3 (sample.py:220-225, losses.py:48-54): 1 - MS-SSIM value and analytic gradient from csrc/msssim.cu against autograd of the
    float64 restatement of pytorch_msssim (oracle/ms_ssim.py; the package itself is un-vendored: parity unpinned)."""
    from clip_diffusion_b200 import losses as L
    from oracle.ms_ssim import structural_dissimilarity_loss as oracle_loss

    g = torch.Generator().manual_seed(H + W)
    base = torch.tanh(torch.randn(1, 3, H, W, generator=g))
    image = base.clone()
    x = (base + 0.15 * torch.randn(1, 3, H, W, generator=g)).clamp(-1, 1)
    xr = x.double().requires_grad_()
    ref = oracle_loss(xr, image.double())
    (gref,) = torch.autograd.grad(ref, xr)
    xc = x.cuda().requires_grad_()
    out = L.structural_dissimilarity_loss(xc, image.cuda())
    (gout,) = torch.autograd.grad(out * 3.0, xc)
    assert abs(out.item() - ref.item()) < 2e-5 * max(1.0, abs(ref.item())), (out.item(), ref.item())
    assert _rel(gout.cpu().double() / 3.0, gref) < 2e-4
    # identical images: dissimilarity 0
    same = L.structural_dissimilarity_loss(image.cuda(), image.cuda())
    assert abs(same.item()) < 1e-5
