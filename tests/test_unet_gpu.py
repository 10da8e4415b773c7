"""GPU parity of the fused NHWC GroupNorm32(+scale-shift)(+SiLU) op (csrc/unet_norm.cu, through the C ABI) against a plain
PyTorch fp32 reference of the same arithmetic, and of the channels_last UNet built on it against the fp32 model.
The blocks it serves are guided-diffusion's ResBlock / AttentionBlock (built at clip_diffusion/models.py:87-131)."""
import copy

import pytest
import torch
from torch.nn import functional as F

pytestmark = pytest.mark.gpu

# fp16 input/output (2^-11 rounding) + tanh.approx sigmoid (2^-11): stated tolerances
TOL_FWD = 2e-3   # relative L2 of y
TOL_BWD = 3e-3   # relative L2 of dx


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _reference(x_nchw_f32, gamma, beta, groups, eps, scale_shift, silu):
    y = F.group_norm(x_nchw_f32, groups, gamma, beta, eps)
    if scale_shift is not None:
        scale, shift = scale_shift.half().float()[:, :, None, None].chunk(2, dim=1)  # the reference rounds the projection to fp16 first
        y = y * (1 + scale) + shift
    return F.silu(y) if silu else y


CASES = [
    # N, C, H, W, groups
    (1, 128, 64, 64, 32),    # level-0 shape family (cg = 4)
    (1, 384, 24, 40, 32),    # concat width 256+128: C/8 = 48 does not divide 256 threads
    (2, 32, 16, 16, 32),     # cg = 1 (test-size UNet), batch 2
    (1, 64, 7, 9, 32),       # ragged H*W, cg = 2
    (1, 2048, 8, 8, 32),     # deepest concat: one row lane per CTA
    (1, 1536, 16, 16, 32),   # C/8 = 192
    (3, 256, 33, 31, 32),
    # > 2 MB per sample: the three-kernel path with the stand-alone finalize kernels (smaller ones fuse it into the apply kernels)
    (1, 64, 160, 168, 32),
    (2, 128, 96, 100, 32),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("silu", [False, True])
@pytest.mark.parametrize("with_ss", [False, True])
@pytest.mark.parametrize("out_f32", [False, True])
def test_group_norm_nhwc_forward_backward(case, silu, with_ss, out_f32):
    from clip_diffusion_b200.unet_ops import group_norm_nhwc

    N, C, H, W, G = case
    g = torch.Generator().manual_seed(N * 1000 + C + H)
    x = (torch.randn(N, C, H, W, generator=g) * 1.5 + 0.7 * torch.randn(1, C, 1, 1, generator=g)).half()
    gamma = 1 + 0.3 * torch.randn(C, generator=g)
    beta = 0.2 * torch.randn(C, generator=g)
    ss = 0.3 * torch.randn(N, 2 * C, generator=g) if with_ss else None
    dy = torch.randn(N, C, H, W, generator=g)
    if not out_f32:
        dy = dy.half()

    xr = x.float().cuda().requires_grad_()
    yr = _reference(xr, gamma.cuda(), beta.cuda(), G, 1e-5, None if ss is None else ss.cuda(), silu)
    (dxr,) = torch.autograd.grad(yr, xr, dy.float().cuda())

    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    y = group_norm_nhwc(xc, gamma.cuda(), beta.cuda(), G, 1e-5, scale_shift=None if ss is None else ss.cuda(), silu=silu,
                        out_dtype=torch.float32 if out_f32 else torch.float16)
    assert y.dtype == (torch.float32 if out_f32 else torch.float16) and y.shape == xc.shape
    assert y.is_contiguous(memory_format=torch.channels_last)
    (dx,) = torch.autograd.grad(y, xc, dy.cuda())  # dy arrives NCHW-contiguous: the op re-lays it out
    assert dx.dtype == torch.float16 and torch.isfinite(dx).all()
    assert _rel(y.float(), yr) <= TOL_FWD, _rel(y.float(), yr)
    assert _rel(dx.float(), dxr) <= TOL_BWD, _rel(dx.float(), dxr)


@pytest.mark.parametrize("case", [(1, 128, 32, 32, 32), (2, 384, 9, 13, 32), (1, 32, 16, 16, 32), (1, 1024, 8, 8, 32), (1, 128, 112, 96, 32)])
def test_group_norm_nhwc_deferred_conv_bias(case):
    """pre_bias: GN(x + b_c) with the producing convolution's bias folded into the statistics / affine (no memory pass)."""
    from clip_diffusion_b200.unet_ops import group_norm_nhwc

    N, C, H, W, G = case
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(N, C, H, W, generator=g).half()
    pb = torch.randn(C, generator=g) * 1.5  # as large as the signal: the cross terms of the folded variance matter
    gamma, beta = 1 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    ss = 0.3 * torch.randn(N, 2 * C, generator=g)
    dy = torch.randn(N, C, H, W, generator=g).half()
    # reference = the CPU oracle (oracle/unet_norm.py, float64 autograd of the block arithmetic; its folded algebra is pinned on the CPU
    # in tests/test_properties_cpu.py), with the embedding projection rounded to fp16 as the reference's `.type(h.dtype)` does
    from oracle import unet_norm as U

    xr = x.double().requires_grad_()
    yr = U.resblock_norm_reference(xr, gamma.double(), beta.double(), G, 1e-5, ss.half().double(), pb.double(), True)
    (dxr,) = torch.autograd.grad(yr, xr, dy.double())
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    y = group_norm_nhwc(xc, gamma.cuda(), beta.cuda(), G, 1e-5, scale_shift=ss.cuda(), silu=True, pre_bias=pb.cuda())
    (dx,) = torch.autograd.grad(y, xc, dy.cuda())
    assert _rel(y.float().cpu(), yr.detach()) <= TOL_FWD, _rel(y.float().cpu(), yr.detach())
    assert _rel(dx.float().cpu(), dxr) <= TOL_BWD, _rel(dx.float().cpu(), dxr)


@pytest.mark.parametrize("shape", [(1, 128, 64, 64), (2, 32, 6, 10), (1, 1024, 8, 8), (1, 384, 18, 22)])
def test_bias_residual_add_and_resample2x(shape):
    """The ResBlock tail add with deferred biases and the resblock_updown resamplers against the stock torch ops (fp16 rounding only)."""
    from clip_diffusion_b200 import unet_ops

    N, C, H, W = shape
    g = torch.Generator().manual_seed(C + H)
    a = torch.randn(shape, generator=g).half().cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    b = torch.randn(shape, generator=g).half().cuda().requires_grad_()  # NCHW on purpose: the op re-lays it out
    bias = torch.randn(C, generator=g).cuda()
    out = unet_ops.bias_residual_add(a, b, bias)
    want = a.float() + b.float() + bias.view(1, -1, 1, 1)
    assert out.is_contiguous(memory_format=torch.channels_last) and _rel(out.float(), want) <= 5e-4
    dy = torch.randn(shape, generator=g).half().cuda()
    ga, gb = torch.autograd.grad(out, (a, b), dy)
    assert torch.equal(ga, dy) and torch.equal(gb, dy)

    for mine, stock in ((unet_ops.avg_pool2x, lambda t: F.avg_pool2d(t, 2)), (unet_ops.upsample_nearest2x, lambda t: F.interpolate(t, scale_factor=2, mode="nearest"))):
        x = torch.randn(shape, generator=g).half().cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
        y = mine(x)
        xr = x.detach().float().requires_grad_()
        yr = stock(xr)
        assert y.shape == yr.shape and y.is_contiguous(memory_format=torch.channels_last)
        assert _rel(y.float(), yr) <= 5e-4
        dyy = torch.randn(yr.shape, generator=g).half().cuda()
        (gx,) = torch.autograd.grad(y, x, dyy)
        (gr,) = torch.autograd.grad(yr, xr, dyy.float())
        assert _rel(gx.float(), gr) <= 5e-4


@pytest.mark.parametrize("shape,cb", [((1, 128, 32, 32), 128), ((2, 256, 5, 7), 128), ((1, 1024, 8, 8), 512), ((1, 8, 3, 3), 24)])
def test_concat_channels_and_split_gradient(shape, cb):
    from clip_diffusion_b200 import unet_ops

    N, ca, H, W = shape
    g = torch.Generator().manual_seed(ca + cb)
    a = torch.randn(N, ca, H, W, generator=g).half().cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    b = torch.randn(N, cb, H, W, generator=g).half().cuda().requires_grad_()  # NCHW: re-laid out by the op
    out = unet_ops.concat_channels(a, b)
    assert out.is_contiguous(memory_format=torch.channels_last) and torch.equal(out, torch.cat([a, b], dim=1))
    dy = torch.randn(N, ca + cb, H, W, generator=g).half().cuda()
    ga, gb = torch.autograd.grad(out, (a, b), dy)
    assert torch.equal(ga, dy[:, :ca]) and torch.equal(gb, dy[:, ca:])
    assert ga.is_contiguous(memory_format=torch.channels_last) and gb.is_contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("shape", [(1, 128, 64, 64), (2, 64, 9, 14), (1, 1024, 8, 8), (1, 256, 96, 128)])
def test_producers_emit_the_next_groupnorm_statistics(shape):
    """bias_residual_add / concat_channels with stats=True: same result as without, and a GroupNorm fed their `._gn_partial`
    (no statistics pass of its own) equals the GroupNorm that reduces the tensor itself -- forward and backward."""
    from clip_diffusion_b200 import unet_ops

    N, C, H, W = shape
    g = torch.Generator().manual_seed(C + W)
    cl = torch.channels_last
    a = torch.randn(shape, generator=g).half().cuda().contiguous(memory_format=cl)
    b = (torch.randn(shape, generator=g) + 0.5).half().cuda().contiguous(memory_format=cl)
    bias = torch.randn(C, generator=g).cuda()
    gamma, beta = torch.rand(C, generator=g).cuda() + 0.5, torch.randn(C, generator=g).cuda()
    for make, width in ((lambda st: unet_ops.bias_residual_add(a, b, bias, stats=st), C), (lambda st: unet_ops.concat_channels(a, b, stats=st), 2 * C)):
        plain, withst = make(False), make(True)
        assert torch.equal(plain, withst) and not hasattr(plain, "_gn_partial") and withst._gn_partial.is_cuda
        gm, bt = (gamma, beta) if width == C else (torch.cat([gamma, gamma]), torch.cat([beta, beta]))
        x0 = plain.detach().requires_grad_()
        x1 = withst.detach().requires_grad_()
        y0 = unet_ops.group_norm_nhwc(x0, gm, bt, 32, 1e-5, silu=True)
        y1 = unet_ops.group_norm_nhwc(x1, gm, bt, 32, 1e-5, silu=True, input_partial=withst._gn_partial)
        assert _rel(y1.float(), y0.float()) <= 1e-6, _rel(y1.float(), y0.float())
        dy = torch.randn(y0.shape, generator=g).half().cuda()
        (g0,) = torch.autograd.grad(y0, x0, dy)
        (g1,) = torch.autograd.grad(y1, x1, dy)
        assert _rel(g1.float(), g0.float()) <= 1e-6
    with pytest.raises(ValueError):
        unet_ops.group_norm_nhwc(a, gamma, beta, 32, 1e-5, input_partial=torch.empty(16, dtype=torch.uint8, device="cuda"))


def test_resample2x_rejects_odd_sizes():
    from clip_diffusion_b200 import _lib, unet_ops

    with pytest.raises(_lib.ClipGuideError):
        unet_ops.avg_pool2x(torch.zeros(1, 16, 5, 4, device="cuda").half().contiguous(memory_format=torch.channels_last))


def test_group_norm_nhwc_passthrough_sums_both_gradients():
    """passthrough=True: (y, x') with x' an alias of x; d/dx of f(y) + g(x') comes out of ONE backward kernel."""
    from clip_diffusion_b200.unet_ops import group_norm_nhwc

    g = torch.Generator().manual_seed(11)
    shape = (2, 64, 24, 20)
    x = torch.randn(shape, generator=g).half().cuda().contiguous(memory_format=torch.channels_last).requires_grad_()
    gamma, beta = (1 + 0.2 * torch.randn(64, generator=g)).cuda(), (0.1 * torch.randn(64, generator=g)).cuda()
    w1, w2 = torch.randn(shape, generator=g).half().cuda(), torch.randn(shape, generator=g).half().cuda()
    y, xp = group_norm_nhwc(x, gamma, beta, 32, 1e-5, silu=True, passthrough=True)
    assert xp.data_ptr() == x.data_ptr()
    (gx,) = torch.autograd.grad((y * w1).sum() + (xp * w2).sum(), x)
    xr = x.detach().float().requires_grad_()
    yr = F.silu(F.group_norm(xr, 32, gamma, beta, 1e-5))
    (gr,) = torch.autograd.grad((yr * w1.float()).sum() + (xr * w2.float()).sum(), xr)
    assert _rel(gx.float(), gr) <= TOL_BWD, _rel(gx.float(), gr)
    # only one of the two outputs used
    y, xp = group_norm_nhwc(x, gamma, beta, 32, 1e-5, silu=True, passthrough=True)
    (g1,) = torch.autograd.grad((xp * w2).sum(), x)
    assert _rel(g1.float(), w2.float()) <= 1e-6
    y, xp = group_norm_nhwc(x, gamma, beta, 32, 1e-5, silu=True, passthrough=True)
    (g2,) = torch.autograd.grad((y * w1).sum(), x)
    (g2r,) = torch.autograd.grad((F.silu(F.group_norm(xr, 32, gamma, beta, 1e-5)) * w1.float()).sum(), xr)
    assert _rel(g2.float(), g2r) <= TOL_BWD


def test_group_norm_nhwc_tokens_and_determinism():
    """[N,T,C] token layout (AttentionBlock) and run-to-run bit-identical results (fixed-order reductions, no atomics)."""
    from clip_diffusion_b200.unet_ops import group_norm_nhwc

    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 1024, 512, generator=g).half().cuda()
    gamma, beta = torch.rand(512, generator=g).cuda(), torch.rand(512, generator=g).cuda()
    y0 = group_norm_nhwc(x, gamma, beta, 32, 1e-5)
    y1 = group_norm_nhwc(x, gamma, beta, 32, 1e-5)
    assert torch.equal(y0, y1)
    ref = F.group_norm(x.float().transpose(1, 2), 32, gamma, beta, 1e-5).transpose(1, 2)
    assert _rel(y0.float(), ref) <= TOL_FWD


def test_group_norm_nhwc_large_mean_and_big_plane():
    """Full-resolution level-0 plane (512x512x128, 67 MB) with a large common offset: fp64 merge of fp32 chunk sums keeps the
    variance accurate (E[x^2]-mean^2 cancellation)."""
    from clip_diffusion_b200.unet_ops import group_norm_nhwc

    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn(1, 128, 512, 512, device="cuda", generator=g) * 0.5 + 6.0).half().contiguous(memory_format=torch.channels_last)
    gamma, beta = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
    y = group_norm_nhwc(x, gamma, beta, 32, 1e-5)
    ref = F.group_norm(x.float(), 32, gamma, beta, 1e-5)
    assert _rel(y.float(), ref) <= TOL_FWD


def test_group_norm_nhwc_rejects_bad_input():
    from clip_diffusion_b200 import _lib
    from clip_diffusion_b200.unet_ops import group_norm_nhwc

    w = torch.ones(12, device="cuda")
    with pytest.raises(_lib.ClipGuideError):
        group_norm_nhwc(torch.zeros(1, 12, 4, 4, device="cuda").half().contiguous(memory_format=torch.channels_last), w, w, 4)  # C % 8
    with pytest.raises(_lib.ClipGuideError):
        group_norm_nhwc(torch.zeros(1, 16, 4, 4).half(), torch.ones(16), torch.ones(16), 4)  # CPU tensor: no fallback


@pytest.mark.parametrize("size", [64, 32])
def test_unet_nhwc_matches_fp32_model(size):
    """channels_last fp16 UNet on the fused ops vs the same weights in fp32 stock PyTorch: epsilon and d(sum eps*w)/dx."""
    from clip_diffusion_b200.unet import create_unet

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref = create_unet(size, seed=2, device="cpu", use_fp16=False)
    mine = create_unet(size, seed=2, device="cuda", use_fp16=True)       # default: channels_last + fused norm ops
    stock = create_unet(size, seed=2, device="cuda", use_fp16=True, channels_last=False)
    assert mine.channels_last and not stock.channels_last
    ref = copy.deepcopy(ref).cuda()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, 2 * size, 2 * size, generator=g).cuda()
    w = torch.randn(1, 6, 2 * size, 2 * size, generator=g).cuda()
    t = torch.tensor([412.0], device="cuda")
    outs = {}
    for name, m in (("ref", ref), ("mine", mine), ("stock", stock)):
        xi = x.clone().requires_grad_()
        eps = m(xi, t)
        (gx,) = torch.autograd.grad((eps.float() * w).sum(), xi)
        outs[name] = (eps.float(), gx.float())
    e_mine, e_stock = _rel(outs["mine"][0], outs["ref"][0]), _rel(outs["stock"][0], outs["ref"][0])
    g_mine, g_stock = _rel(outs["mine"][1], outs["ref"][1]), _rel(outs["stock"][1], outs["ref"][1])
    # fp16 trunk: both variants sit at the fp16 rounding level; the fused path (one rounding per norm) must not be worse
    assert e_mine <= 1e-2 and g_mine <= 2e-2, (e_mine, g_mine)
    assert e_mine <= 2.0 * e_stock + 1e-3 and g_mine <= 2.0 * g_stock + 1e-3, (e_mine, e_stock, g_mine, g_stock)


def test_unet_nhwc_graph_capture():
    """The fused ops are capturable (no syncs, current-stream launches): graphed callable == eager, forward and backward."""
    from clip_diffusion_b200.unet import create_unet, graph_unet

    m = create_unet(32, seed=2, device="cuda", use_fp16=True)
    x = torch.randn(1, 3, 64, 64, device="cuda")
    t = torch.tensor([100.0], device="cuda")
    xe = x.clone().requires_grad_()
    ee = m(xe, t)
    (ge,) = torch.autograd.grad(ee.sum(), xe)
    gm = graph_unet(m, 64, 64, "cuda")
    xg = x.clone().requires_grad_()
    eg = gm(xg, t)
    (gg,) = torch.autograd.grad(eg.sum(), xg)
    # same kernels of ours; cuDNN may pick other fp16 algorithms during capture => fp16-rounding-level differences only
    assert _rel(eg, ee) <= 5e-3 and _rel(gg, ge) <= 5e-3, (_rel(eg, ee), _rel(gg, ge))
