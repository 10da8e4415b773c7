"""Fused NHWC GroupNorm32 (+ scale-shift) (+ SiLU) for the replicated guided-diffusion UNet, over csrc/unet_norm.cu.

SURVEY.md section 8(f) N1: once the CLIP side runs on tensor cores the guidance step is dominated by the element-wise
traffic of the stock-PyTorch UNet around cuDNN's NHWC convolutions.  `group_norm_nhwc` is one op for what the reference's
ResBlock spells `silu(GroupNorm32(x))` / `silu(GroupNorm32(h) * (1 + scale) + shift)` (un-vendored guided-diffusion,
App. A.3) on channels_last fp16 tensors; its autograd backward yields the input gradient only (UNet weights are frozen).
No fallback: a missing library raises (see _lib.load).
"""
import torch

from clip_diffusion_b200 import _lib


def _nhwc_dims(x):
    """(N, HW, C) of a tensor whose memory is [N, spatial..., C] row-major: 4-D channels_last or 3-D [N, T, C] contiguous."""
    if x.dim() == 4:
        if not x.is_contiguous(memory_format=torch.channels_last):
            raise _lib.ClipGuideError("group_norm_nhwc: 4-D input must be channels_last contiguous")
        return x.shape[0], x.shape[2] * x.shape[3], x.shape[1]
    if x.dim() == 3:
        if not x.is_contiguous():
            raise _lib.ClipGuideError("group_norm_nhwc: 3-D input must be [N, T, C] contiguous")
        return x.shape[0], x.shape[1], x.shape[2]
    raise _lib.ClipGuideError("group_norm_nhwc: expected [N,C,H,W] channels_last or [N,T,C]; got %s" % (tuple(x.shape),))


def _like_layout(x, t):
    """Bring a gradient to the memory layout of x (autograd may hand back NCHW-contiguous or expanded tensors)."""
    if x.dim() == 4:
        return t.contiguous(memory_format=torch.channels_last)
    return t.contiguous()


class _GroupNormNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, scale_shift, groups, eps, silu, out_f32, pre_bias, passthrough, input_partial=None):
        _lib.require_cuda(x, gamma, beta, scale_shift, pre_bias)
        if x.dtype != torch.float16:
            raise _lib.ClipGuideError("group_norm_nhwc: fp16 activations expected, got %s" % x.dtype)
        N, HW, C = _nhwc_dims(x)
        gamma = gamma.detach().float().contiguous()
        beta = beta.detach().float().contiguous()
        if scale_shift is not None:
            scale_shift = scale_shift.detach().float().contiguous()
            if tuple(scale_shift.shape) != (N, 2 * C):
                raise ValueError("scale_shift must be [N, 2C] = %s, got %s" % ((N, 2 * C), tuple(scale_shift.shape)))
        if pre_bias is not None:
            pre_bias = pre_bias.detach().float().contiguous()
            if pre_bias.numel() != C:
                raise ValueError("pre_bias must have C = %d elements" % C)
        y = torch.empty_like(x, dtype=torch.float32 if out_f32 else torch.float16)  # preserve_format: same NHWC strides
        stats = torch.empty(N, groups, 2, device=x.device, dtype=torch.float32)
        coef = torch.empty(2, N, C, device=x.device, dtype=torch.float32)
        ws = torch.empty(_lib.load().cg_groupnorm_nhwc_workspace_bytes(N, HW, C), device=x.device, dtype=torch.uint8)
        _lib.call("cg_groupnorm_nhwc_fwd", _lib.ptr(x), N, HW, C, int(groups), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(scale_shift),
                  _lib.ptr(pre_bias), _lib.ptr(input_partial), float(eps), int(bool(silu)), int(bool(out_f32)), _lib.ptr(y), _lib.ptr(stats), _lib.ptr(coef), _lib.ptr(ws))
        ctx.save_for_backward(x, stats, coef, pre_bias)
        ctx.cfg = (N, HW, C, int(groups), int(bool(silu)), bool(out_f32))
        ctx.set_materialize_grads(False)  # an unused output's gradient arrives as None, not as a zero tensor
        if passthrough:
            # second output = x itself, for the block's residual / skip path: both gradients then arrive in ONE backward call and
            # are summed inside the apply kernel (autograd would otherwise accumulate them in a separate read-read-write pass)
            return y, torch.ops.aten.alias(x)  # NOT view_as: that rewrites the stride of size-1 dims and cuDNN then stops seeing NHWC
        return y

    @staticmethod
    def backward(ctx, dy, dres=None):
        x, stats, coef, pre_bias = ctx.saved_tensors
        N, HW, C, groups, silu, out_f32 = ctx.cfg
        if dy is None:  # only the pass-through output was used
            return (dres,) + (None,) * 10
        dy = _like_layout(x, dy)
        if dy.dtype not in (torch.float16, torch.float32):
            dy = dy.float()
        if dres is not None:
            dres = _like_layout(x, dres.to(torch.float16))
        dx = torch.empty_like(x)
        ws = torch.empty(_lib.load().cg_groupnorm_nhwc_workspace_bytes(N, HW, C), device=x.device, dtype=torch.uint8)
        _lib.call("cg_groupnorm_nhwc_bwd", _lib.ptr(dy), int(dy.dtype == torch.float32), _lib.ptr(x), N, HW, C, groups, _lib.ptr(stats),
                  _lib.ptr(coef), _lib.ptr(pre_bias), silu, _lib.ptr(dres), _lib.ptr(dx), _lib.ptr(ws))
        return (dx,) + (None,) * 10


def group_norm_nhwc(x, gamma, beta, groups=32, eps=1e-5, scale_shift=None, silu=False, out_dtype=None, pre_bias=None, passthrough=False,
                    input_partial=None):
    """act(GroupNorm32(x + pre_bias) * (1 + scale) + shift) for x [N,C,H,W] channels_last (or [N,T,C]) fp16 on CUDA.

    scale_shift: [N, 2C] (the ResBlock's embedding projection, scale | shift) or None; silu: apply SiLU;
    out_dtype: torch.float16 (default) or torch.float32 (the UNet's fp32 output head);
    pre_bias: [C] bias of the convolution that produced x, deferred into this op (no extra memory pass);
    passthrough: return (y, x') with x' an alias of x to be used by the block's residual / skip path -- the two gradients of x are
    then summed inside the backward kernel;
    input_partial: the chunk partials of x written by the op that produced it (`bias_residual_add` / `concat_channels` attach them
    to their result as `._gn_partial`): the statistics pass over x is skipped."""
    out_f32 = out_dtype == torch.float32
    if input_partial is not None:
        N, HW, C = _nhwc_dims(x)
        if input_partial.numel() != _lib.load().cg_groupnorm_nhwc_workspace_bytes(N, HW, C):
            raise ValueError("input_partial does not belong to a tensor of this shape")
    return _GroupNormNHWC.apply(x, gamma, beta, scale_shift, groups, eps, silu, out_f32, pre_bias, passthrough, input_partial)


def _require_nhwc_half(x, what):
    if not (x.is_cuda and x.dtype == torch.float16 and x.dim() == 4):
        raise _lib.ClipGuideError("%s: [N,C,H,W] fp16 CUDA tensor expected (no CPU fallback); got %s %s on %s" % (what, tuple(x.shape), x.dtype, x.device))
    return x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)


class _BiasResidualAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, bias, partial):
        a, b = _require_nhwc_half(a, "bias_residual_add"), _require_nhwc_half(b, "bias_residual_add")
        if a.shape != b.shape:
            raise ValueError("shape mismatch: %s vs %s" % (tuple(a.shape), tuple(b.shape)))
        N, C, H, W = a.shape
        bias = bias.detach().float().contiguous()
        if bias.numel() != C:
            raise ValueError("bias must have C = %d elements" % C)
        out = torch.empty_like(a)
        if partial is None:
            _lib.call("cg_bias_residual_add_nhwc", _lib.ptr(a), _lib.ptr(b), _lib.ptr(bias), N * H * W, C, _lib.ptr(out))
        else:
            _lib.call("cg_bias_residual_add_stats_nhwc", _lib.ptr(a), _lib.ptr(b), _lib.ptr(bias), N, H * W, C, _lib.ptr(out), _lib.ptr(partial))
        return out

    @staticmethod
    def backward(ctx, dy):
        return dy, dy, None, None


def _partial_buffer(N, HW, C, device):
    return torch.empty(_lib.load().cg_groupnorm_nhwc_workspace_bytes(N, HW, C), device=device, dtype=torch.uint8)


def bias_residual_add(a, b, bias, stats=False):
    """a + b + bias[None, :, None, None] in one pass (ResBlock tail `skip(x) + out_conv(h)` with the conv biases deferred).

    stats=True: the kernel also writes the chunk partials of its result for the GroupNorm that consumes it next; they ride on the
    returned tensor as `._gn_partial` (see group_norm_nhwc(input_partial=...))."""
    partial = _partial_buffer(a.shape[0], a.shape[2] * a.shape[3], a.shape[1], a.device) if stats else None
    out = _BiasResidualAdd.apply(a, b, bias, partial)
    if stats:
        out._gn_partial = partial
    return out


def _resample2x(x, up, scale):
    x = _require_nhwc_half(x, "resample2x")
    N, C, H, W = x.shape
    y = torch.empty((N, C, 2 * H, 2 * W) if up else (N, C, H // 2, W // 2), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    _lib.call("cg_resample2x_nhwc", _lib.ptr(x), N, H, W, C, int(up), float(scale), _lib.ptr(y))
    return y


class _Resample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, up):
        ctx.up = up
        return _resample2x(x, up, 1.0 if up else 0.25)

    @staticmethod
    def backward(ctx, dy):
        # d(nearest upsample) = sum over each 2x2 window; d(avg_pool 2x2) = 0.25 * nearest upsample
        return (_resample2x(dy, False, 1.0) if ctx.up else _resample2x(dy, True, 0.25)), None


def upsample_nearest2x(x):
    """F.interpolate(x, scale_factor=2, mode="nearest") on channels_last fp16 (guided-diffusion Upsample without conv)."""
    return _Resample2x.apply(x, True)


def avg_pool2x(x):
    """F.avg_pool2d(x, 2) on channels_last fp16 (guided-diffusion Downsample without conv)."""
    return _Resample2x.apply(x, False)


def _split_channels(cat, ca, cb):
    cat = _require_nhwc_half(cat, "concat_channels backward")
    N, C, H, W = cat.shape
    a = torch.empty((N, ca, H, W), device=cat.device, dtype=cat.dtype, memory_format=torch.channels_last)
    b = torch.empty((N, cb, H, W), device=cat.device, dtype=cat.dtype, memory_format=torch.channels_last)
    _lib.call("cg_concat2_nhwc", _lib.ptr(a), ca, _lib.ptr(b), cb, N * H * W, _lib.ptr(cat), 1)
    return a, b


class _ConcatChannels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, partial):
        a, b = _require_nhwc_half(a, "concat_channels"), _require_nhwc_half(b, "concat_channels")
        if a.shape[0] != b.shape[0] or a.shape[2:] != b.shape[2:]:
            raise ValueError("shape mismatch: %s vs %s" % (tuple(a.shape), tuple(b.shape)))
        N, ca, H, W = a.shape
        cb = b.shape[1]
        out = torch.empty((N, ca + cb, H, W), device=a.device, dtype=a.dtype, memory_format=torch.channels_last)
        if partial is None:
            _lib.call("cg_concat2_nhwc", _lib.ptr(a), ca, _lib.ptr(b), cb, N * H * W, _lib.ptr(out), 0)
        else:
            _lib.call("cg_concat2_stats_nhwc", _lib.ptr(a), ca, _lib.ptr(b), cb, N, H * W, _lib.ptr(out), _lib.ptr(partial))
        ctx.widths = (ca, cb)
        return out

    @staticmethod
    def backward(ctx, dy):
        return _split_channels(dy, *ctx.widths) + (None,)  # two dense tensors in one pass (not strided views that every consumer re-copies)


def concat_channels(a, b, stats=False):
    """torch.cat([a, b], dim=1) on channels_last fp16 (the UNet's skip connections), with a one-pass split as its gradient.
    stats=True: also emits the next GroupNorm's chunk partials (`._gn_partial`, see bias_residual_add)."""
    partial = _partial_buffer(a.shape[0], a.shape[2] * a.shape[3], a.shape[1] + b.shape[1], a.device) if stats else None
    out = _ConcatChannels.apply(a, b, partial)
    if stats:
        out._gn_partial = partial
    return out
