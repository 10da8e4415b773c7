"""Loss surface of clip_diffusion/losses.py, backed by single-pass sm_100a kernels.

Same names and argument meaning as the reference (losses.py:10-45) so sample.py's imports resolve
unchanged; north_star aliases (spherical_dist_loss, tv_loss, range_loss) are exported too.  Every op
is a torch.autograd.Function whose forward launches ONE kernel that also produces the analytic
gradient, so sample.py's own torch.autograd.grad calls (sample.py:201-226) work on it.
"""
import torch

from clip_diffusion_b200 import _lib


def _as_f32_contig(t):
    return t.contiguous().float() if (t.dtype != torch.float32 or not t.is_contiguous()) else t


class _ImageLossFn(torch.autograd.Function):
    """value [B] and unit-scale gradient [B,C,H,W] from one kernel launch."""

    @staticmethod
    def forward(ctx, input, entry):
        _lib.require_cuda(input)
        x = _as_f32_contig(input)
        if x.dim() != 4:
            raise ValueError("expected a [B,C,H,W] tensor")
        B, Cc, H, W = x.shape
        loss = torch.empty(B, device=x.device, dtype=torch.float32)
        grad = torch.empty_like(x) if input.requires_grad else None
        _lib.call(entry, _lib.ptr(x), B, Cc, H, W, 1.0, 0, _lib.ptr(loss), _lib.ptr(grad))
        ctx.save_for_backward(grad)
        ctx.in_dtype = input.dtype
        return loss

    @staticmethod
    def backward(ctx, gloss):
        (grad,) = ctx.saved_tensors
        if grad is None:
            return None, None
        return (grad * gloss.view(-1, 1, 1, 1)).to(ctx.in_dtype), None


def total_variational_loss(input):
    """L2 total variation, replicate pad, mean over (C,H,W) -> [B]   (losses.py:20-28)"""
    return _ImageLossFn.apply(input, "cg_tv_loss_fwd_bwd")


def rgb_range_loss(input):
    """mean((x - clamp(x,-1,1))^2) over (C,H,W) -> [B]   (losses.py:31-35)"""
    return _ImageLossFn.apply(input, "cg_range_loss_fwd_bwd")


def _pairwise_layout(xs, ys):
    """x [.., E] against y [.., E] must broadcast as an outer product with x's axes slowest --
    the reference's only call pattern is [N,1,E] x [1,P,E] (sample.py:179-182)."""
    out = torch.broadcast_shapes(xs[:-1], ys[:-1])
    nd = len(out)
    xl = (1,) * (nd - len(xs) + 1) + tuple(xs[:-1])
    yl = (1,) * (nd - len(ys) + 1) + tuple(ys[:-1])
    last_x = max([i for i in range(nd) if xl[i] > 1], default=-1)
    first_y = min([i for i in range(nd) if yl[i] > 1], default=nd)
    if last_x >= first_y:
        raise NotImplementedError(
            "square_spherical_distance_loss supports outer-product broadcasts ([N,1,E] x [1,P,E]); got %s x %s" % (tuple(xs), tuple(ys))
        )
    return out


class _SphericalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        _lib.require_cuda(x, y)
        E = x.shape[-1]
        if y.shape[-1] != E:
            raise ValueError("last dimensions differ: %s vs %s" % (tuple(x.shape), tuple(y.shape)))
        out_shape = _pairwise_layout(x.shape, y.shape)
        xe = _as_f32_contig(x).reshape(-1, E)
        ye = _as_f32_contig(y).reshape(-1, E)
        N, P = xe.shape[0], ye.shape[0]
        dist = torch.empty(N, P, device=x.device, dtype=torch.float32)
        _lib.call("cg_spherical_dist_fwd", _lib.ptr(xe), _lib.ptr(ye), N, P, E, _lib.ptr(dist))
        ctx.save_for_backward(xe, ye)
        ctx.x_shape, ctx.x_dtype = x.shape, x.dtype
        return dist.reshape(out_shape)

    @staticmethod
    def backward(ctx, g):
        xe, ye = ctx.saved_tensors
        N, P, E = xe.shape[0], ye.shape[0], xe.shape[1]
        g2 = _as_f32_contig(g.reshape(N, P))
        demb = torch.empty_like(xe)
        _lib.call("cg_spherical_dist_bwd", _lib.ptr(xe), _lib.ptr(ye), _lib.ptr(g2), N, P, E, _lib.ptr(demb))
        # the text side is a constant of the guidance step (preprocessing.py:11-24): no gradient
        return demb.reshape(ctx.x_shape).to(ctx.x_dtype), None


def square_spherical_distance_loss(x, y):
    """2*asin(||x^ - y^||/2)^2 over the last dim   (losses.py:10-16)"""
    return _SphericalFn.apply(x, y)


def aesthetic_loss(predictor, input):
    """predictor(normalize(emb)).mean()   (losses.py:43-45); the predictors are 1-5 tiny frozen Linear layers
    (models.py:188-217) on an [N,E] tensor -- plain torch."""
    return predictor(torch.nn.functional.normalize(input, dim=-1)).mean()


def LPIPS_loss(LPIPS_model, input, image):
    """pass-through to the caller's LPIPS network (losses.py:38-40); out of scope"""
    return LPIPS_model(input, image)


class _MsSsimFn(torch.autograd.Function):
    """1 - MS-SSIM(input, image), value and analytic gradient w.r.t. ``input`` from csrc/msssim.cu (the init image is a constant)."""

    @staticmethod
    def forward(ctx, input, image):
        _lib.require_cuda(input, image)
        if input.dim() != 4 or input.shape != image.shape:
            raise ValueError("expected two [B,C,H,W] images of the same shape, got %s and %s" % (tuple(input.shape), tuple(image.shape)))
        x, y = _as_f32_contig(input), _as_f32_contig(image)
        B, C, H, W = x.shape
        lib = _lib.load()
        ws_bytes = lib.cg_ms_ssim_workspace_bytes(C, H, W)
        if ws_bytes == 0:
            raise _lib.ClipGuideError(lib.cg_last_error().decode("utf-8", "replace") or "MS-SSIM: unsupported image size %dx%d" % (H, W))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        vals = torch.empty(B, device=x.device, dtype=torch.float32)
        grad = torch.empty_like(x)
        for b in range(B):  # the guidance path has batch 1 (sample.py:246-251)
            _lib.call("cg_ms_ssim_dissimilarity_fwd_bwd", _lib.ptr(x[b]), _lib.ptr(y[b]), C, H, W, 1.0, 0, _lib.ptr(vals[b:]), _lib.ptr(grad[b]), _lib.ptr(ws))
        ctx.save_for_backward(grad)
        ctx.batch, ctx.in_dtype = B, input.dtype
        return vals.mean()  # size_average=True: one value for the batch

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * (g / ctx.batch)).to(ctx.in_dtype), None


def structural_dissimilarity_loss(input, image):
    """1 - MS-SSIM of the two images mapped to [0,1] (losses.py:48-54; ``pytorch_msssim.MS_SSIM(win_size=11, win_sigma=1.5, data_range=1,
    size_average=True, channel=3)``, losses.py:7) as CUDA kernels with an analytic gradient.  CUDA tensors only (no CPU path); the
    float64 restatement used to check it lives in oracle/ms_ssim.py."""
    return _MsSsimFn.apply(input, image)


# north_star vocabulary
spherical_dist_loss = square_spherical_distance_loss
tv_loss = total_variational_loss
range_loss = rgb_range_loss
