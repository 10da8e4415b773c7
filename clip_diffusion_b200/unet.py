"""ADM UNet (the guided-diffusion epsilon model) -- the REPLICATED part of a guidance step.

north_star keeps "the guided-diffusion UNet epsilon forward" replicated on every rank; it is the workload around the
CLIP-guidance kernels (clip_diffusion/models.py:87-131 builds it from the un-vendored crowsonkb/guided-diffusion;
SURVEY.md App. A.3 restates the architecture) and, once the CLIP side ran on tensor cores, most of the step.

Two execution paths of the same modules and the same state dict:
  * stock PyTorch (CPU fp32 = the oracle / reference arm's model; CUDA fp16 NCHW kept for A/B): torch / cuDNN library calls only;
  * ``create_unet(channels_last=True)`` (default for CUDA fp16): NHWC trunk -- convolutions and the UNet's own attention stay
    cuDNN / cuBLAS calls, everything between them runs on the hand-written NHWC kernels of csrc/unet_norm.cu
    (clip_diffusion_b200.unet_ops: fused GroupNorm32+scale-shift+SiLU forward / input gradient with deferred conv biases and
    in-kernel skip-gradient sum, bias+residual add, 2x resample, skip concat / split).  No fallback: it raises without the library.

512 config (models.py:95-116): model_channels 256, 2 res blocks, head_channels 64, attention at 32/16/8,
channel_mult (0.5,1,1,2,2,4,4), resblock_updown, scale-shift norm, learn_sigma (6 output channels), fp16 trunk.
"""
import math

import torch
from torch import nn
from torch.nn import functional as F


class _SplitStatsGroupNorm(torch.autograd.Function):
    """GroupNorm for fp16 CUDA activations out of stock torch ops, arranged for parallelism and layout independence.

    ATen's group-norm statistics kernel launches one CTA per (sample, group) -- 32 CTAs for the batch-1 UNet, i.e. 22% of
    a B200's SMs -- and measured 38% of the whole guidance step (profiles/r01_c2_step_kernel_table_torchprofiler.txt).
    Here every group is split into S contiguous chunks whose mean/variance come from one `var_mean` over N*G*S rows and
    are merged exactly (parallel-variance formula, fp32); normalisation is one `addcmul` (fp32 math, one rounding), and the
    backward uses the closed form dx = a_c*dy + b_g*x + c_g with two per-channel reductions and two element-wise passes.
    A dimension-based fallback handles non-contiguous layouts.  This is the NCHW A/B variant (`bench.py --unet-layout nchw`,
    61.8 ms/step); the default CUDA path is the NHWC trunk on csrc/unet_norm.cu (23.0 ms/step)."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps):
        n, c = x.shape[0], x.shape[1]
        cg = c // groups
        dims = tuple(range(2, x.dim()))
        m = float(x.numel() // (n * groups))
        if x.is_contiguous():
            # NCHW: every group is one contiguous run -> split it into S chunks, ONE Welford pass over N*G*S rows, exact merge
            length, split = int(m), 1
            while split < 256 and length % (split * 2) == 0 and length // (split * 2) >= 2048:
                split *= 2
            var_s, mean_s = torch.var_mean(x.reshape(n * groups * split, length // split), dim=1, unbiased=False)
            mean_s, var_s = mean_s.float().view(n, groups, split), var_s.float().view(n, groups, split)
            mean = mean_s.mean(-1)
            var = (var_s + mean_s * mean_s).mean(-1) - mean * mean
        else:
            # any other layout (e.g. channels_last): per-channel sum / sum of squares, merged per group
            s1 = x.sum(dim=dims, dtype=torch.float32)
            s2 = torch.linalg.vector_norm(x, 2, dim=dims, dtype=torch.float32).square()
            mean = s1.view(n, groups, cg).sum(-1) / m
            var = s2.view(n, groups, cg).sum(-1) / m - mean * mean
        rstd = torch.rsqrt(var.clamp_min(0) + eps)
        w = weight.float().view(1, groups, cg)
        scale = (rstd.unsqueeze(-1) * w).view(n, c)
        shift = bias.float().view(1, c) - mean.repeat_interleave(cg, dim=1) * scale
        shp = (n, c) + (1,) * (x.dim() - 2)
        y = torch.addcmul(shift.to(x.dtype).view(shp), x, scale.to(x.dtype).view(shp))
        ctx.save_for_backward(x, mean, rstd, weight)
        ctx.groups = groups
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, weight = ctx.saved_tensors
        g = ctx.groups
        n, c = x.shape[0], x.shape[1]
        cg = c // g
        dims = tuple(range(2, x.dim()))
        m = float(x.numel() // (n * g))
        s_dy = dy.sum(dim=dims, dtype=torch.float32)                  # [n, c]
        s_dyx = (dy * x).sum(dim=dims, dtype=torch.float32)           # [n, c]
        w = weight.float().view(1, g, cg)
        s_dy_g = (s_dy.view(n, g, cg) * w).sum(-1)                    # sum over the group of gamma*dy
        s_dyx_g = (s_dyx.view(n, g, cg) * w).sum(-1)
        # x_hat = (x - mean) * rstd ;  dx = rstd * (gamma*dy - mean_g(gamma*dy) - x_hat * mean_g(gamma*dy*x_hat))
        c2 = (s_dyx_g - mean * s_dy_g) * rstd / m                     # mean_g(gamma*dy*x_hat)
        c1 = s_dy_g / m
        a = (rstd.unsqueeze(-1) * w).reshape(n, c)                    # coefficient of dy, per channel
        b = -(rstd * rstd * c2)                                       # coefficient of x, per group
        cc = -(rstd * c1) - b * mean                                  # constant, per group
        shp = (n, c) + (1,) * (x.dim() - 2)
        bx = b.repeat_interleave(cg, dim=1).to(x.dtype).view(shp)
        ccx = cc.repeat_interleave(cg, dim=1).to(x.dtype).view(shp)
        dx = torch.addcmul(torch.addcmul(ccx, x, bx), dy, a.to(x.dtype).view(shp))
        return dx, None, None, None, None


class GroupNorm32(nn.GroupNorm):
    """guided-diffusion's GroupNorm32 (fp32 statistics) with the block's follow-up element-wise work as arguments:
    ``forward(x, scale_shift=None, silu=False)`` = ``act(GN(x) * (1 + scale) + shift)``.

    ``nhwc`` (set by ``create_unet(channels_last=True)``): fp16 CUDA activations go through ONE fused sm_100a op
    (clip_diffusion_b200.unet_ops / csrc/unet_norm.cu); otherwise the same arithmetic is spelled in stock torch ops (the CPU
    fp32 model of the oracle / reference arm, and the NCHW fp16 variant kept for A/B measurements)."""

    nhwc = False

    def forward(self, x, scale_shift=None, silu=False, out_dtype=None, pre_bias=None, passthrough=False, input_partial=None):
        if self.nhwc and not (x.is_cuda and x.dtype == torch.float16):
            # a model built for the NHWC kernels never drops to the stock ops behind the caller's back
            raise RuntimeError("GroupNorm32(nhwc=True) needs fp16 CUDA activations (got %s on %s); build the model with "
                               "create_unet(channels_last=False) for the stock-PyTorch path" % (x.dtype, x.device))
        if self.nhwc:
            from clip_diffusion_b200.unet_ops import group_norm_nhwc

            if input_partial is None:  # statistics partials written by the op that produced x (bias_residual_add / concat_channels)
                input_partial = getattr(x, "_gn_partial", None)

            if x.dim() == 4 and not x.is_contiguous(memory_format=torch.channels_last):
                x = x.contiguous(memory_format=torch.channels_last)
            return group_norm_nhwc(x, self.weight, self.bias, self.num_groups, self.eps, scale_shift=scale_shift, silu=silu, out_dtype=out_dtype,
                                   pre_bias=pre_bias, passthrough=passthrough, input_partial=input_partial)
        assert not passthrough, "passthrough is a feature of the fused NHWC op"
        if pre_bias is not None:
            x = x + pre_bias.type(x.dtype).view(1, -1, *([1] * (x.dim() - 2)))
        if x.is_cuda and x.dtype == torch.float16:
            y = _SplitStatsGroupNorm.apply(x, self.weight, self.bias, self.num_groups, self.eps)
        else:
            y = super().forward(x.float()).type(x.dtype)
        if scale_shift is not None:
            scale, shift = scale_shift.type(y.dtype)[:, :, None, None].chunk(2, dim=1)
            y = y * (1 + scale) + shift
        if silu:
            y = F.silu(y)
        return y if out_dtype is None else y.type(out_dtype)


def timestep_embedding(timesteps, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class Resample(nn.Module):
    def __init__(self, up):
        super().__init__()
        self.up = up

    nhwc = False  # set by create_unet(channels_last=True): hand-written NHWC kernels instead of ATen's

    def forward(self, x):
        if self.nhwc and x.is_cuda and x.dtype == torch.float16:
            from clip_diffusion_b200 import unet_ops

            return unet_ops.upsample_nearest2x(x) if self.up else unet_ops.avg_pool2x(x)
        return F.interpolate(x, scale_factor=2, mode="nearest") if self.up else F.avg_pool2d(x, 2)


class ResBlock(nn.Module):
    def __init__(self, channels, emb_channels, out_channels, up=False, down=False):
        super().__init__()
        self.out_channels = out_channels
        self.in_norm = GroupNorm32(32, channels)
        self.in_conv = nn.Conv2d(channels, out_channels, 3, padding=1)
        self.resample = Resample(up) if (up or down) else None
        self.emb = nn.Linear(emb_channels, 2 * out_channels)  # scale-shift norm
        self.out_norm = GroupNorm32(32, out_channels)
        self.out_conv = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.skip = nn.Identity() if out_channels == channels else nn.Conv2d(channels, out_channels, 1)
        self._fused_bias = None
        self._scale_shift = None  # set per forward by UNetModel._project_embeddings (fused NHWC path)

    def forward(self, x, emb):
        if self.in_norm.nhwc and x.is_cuda and x.dtype == torch.float16:
            h, x = self.in_norm(x, silu=True, passthrough=True)  # x: alias for the skip path (gradients summed in the norm's backward kernel)
        else:
            h = self.in_norm(x, silu=True)
        if self.resample is not None:
            h = self.resample(h)
            x = self.resample(x)
        if self.in_norm.nhwc and x.is_cuda and x.dtype == torch.float16:
            return self._tail_nhwc(x, h, emb)
        h = self.in_conv(h)
        h = self.out_norm(h, scale_shift=self.emb(F.silu(emb)), silu=True)  # scale-shift norm: GN(h) * (1 + scale) + shift
        h = self.out_conv(h)
        return self.skip(x) + h


    def _tail_nhwc(self, x, h, emb):
        """Same arithmetic with the three convolution biases deferred: in_conv's into the following normalisation (folded into
        its statistics, free), out_conv's and the 1x1 skip's into the single residual-add pass (cuDNN through torch adds a
        bias in a separate full read+write of the activation: 2.9 of 32 ms per step)."""
        from clip_diffusion_b200.unet_ops import bias_residual_add

        if self._fused_bias is None:  # weights are frozen (create_unet): cache the fp32 bias vectors once
            tail = self.out_conv.bias.detach().float()
            if not isinstance(self.skip, nn.Identity):
                tail = tail + self.skip.bias.detach().float()
            self._fused_bias = (self.in_conv.bias.detach().float().contiguous(), tail.contiguous())
        b_in, b_tail = self._fused_bias
        h = F.conv2d(h, self.in_conv.weight, None, padding=1)
        # scale-shift of this block: a slice of the model-wide batched projection when UNetModel.forward prepared one
        ss = self._scale_shift if self._scale_shift is not None else self.emb(F.silu(emb))
        h = self.out_norm(h, scale_shift=ss, silu=True, pre_bias=b_in)
        h = F.conv2d(h, self.out_conv.weight, None, padding=1)
        skip = x if isinstance(self.skip, nn.Identity) else F.conv2d(x, self.skip.weight, None)
        return bias_residual_add(h, skip, b_tail, stats=True)  # whatever consumes a block's output normalises it first


class AttentionBlock(nn.Module):
    def __init__(self, channels, head_channels):
        super().__init__()
        self.heads = channels // head_channels
        self.norm = GroupNorm32(32, channels)
        self.qkv = nn.Conv1d(channels, channels * 3, 1)
        self.proj = nn.Conv1d(channels, channels, 1)

    def forward(self, x):
        b, c, hh, ww = x.shape
        if self.norm.nhwc and x.is_cuda and x.dtype == torch.float16:
            return self._forward_nhwc(x)
        xf = x.reshape(b, c, -1)
        qkv = self.qkv(self.norm(xf))  # [b, 3c, t], legacy order: heads x (q|k|v) x head_dim
        t = qkv.shape[-1]
        q, k, v = qkv.reshape(b * self.heads, 3, c // self.heads, t).unbind(1)
        a = F.scaled_dot_product_attention(q.transpose(1, 2).unsqueeze(0), k.transpose(1, 2).unsqueeze(0), v.transpose(1, 2).unsqueeze(0))
        a = a.squeeze(0).transpose(1, 2).reshape(b, c, t)
        return (xf + self.proj(a)).reshape(b, c, hh, ww)

    def _forward_nhwc(self, x):
        """Same block on channels_last activations: NHWC memory IS the token-major [t, c] matrix, so the 1x1 convolutions are
        plain linears on a view and nothing is transposed.  Legacy qkv channel order: heads x (q|k|v) x head_dim."""
        b, c, hh, ww = x.shape
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        t = hh * ww
        tok = x.permute(0, 2, 3, 1).reshape(b, t, c)  # a view
        normed, tok = self.norm(tok, passthrough=True, input_partial=getattr(x, "_gn_partial", None))
        qkv = F.linear(normed, self.qkv.weight.squeeze(-1), self.qkv.bias)
        q, k, v = qkv.view(b, t, self.heads, 3, c // self.heads).permute(3, 0, 2, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, t, c)
        out = tok + F.linear(a, self.proj.weight.squeeze(-1), self.proj.bias)
        return out.view(b, hh, ww, c).permute(0, 3, 1, 2)


class _Seq(nn.Module):
    """A list of layers where ResBlocks take the timestep embedding."""

    def __init__(self, *layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def forward(self, x, emb):
        for layer in self.layers:
            x = layer(x, emb) if isinstance(layer, ResBlock) else layer(x)
        return x


UNET_CONFIGS = {
    # image size: (model_channels, channel_mult, attention downsample rates)
    512: (256, (0.5, 1, 1, 2, 2, 4, 4), (16, 32, 64)),
    256: (256, (1, 1, 2, 2, 4, 4), (8, 16, 32)),
    # small variants for tests / CPU baselines
    64: (64, (1, 2, 2), (2, 4)),
    32: (32, (1, 2), (2,)),
}


class UNetModel(nn.Module):
    def __init__(self, image_size=512, num_res_blocks=2, head_channels=64, learn_sigma=True, use_fp16=True, config=None):
        super().__init__()
        mc, mult, attn_ds = config if config is not None else UNET_CONFIGS[image_size]
        self.model_channels = mc
        self.dtype = torch.float16 if use_fp16 else torch.float32
        self.channels_last = False  # set by create_unet(..., channels_last=True): NHWC activations/weights for cuDNN
        self._emb_cat = None
        emb_ch = mc * 4
        self.time_embed = nn.Sequential(nn.Linear(mc, emb_ch), nn.SiLU(), nn.Linear(emb_ch, emb_ch))
        ch = in_ch = int(mult[0] * mc)
        hc = min(head_channels, in_ch)
        self.input_blocks = nn.ModuleList([_Seq(nn.Conv2d(3, ch, 3, padding=1))])
        chans = [ch]
        ds = 1
        for level, m in enumerate(mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, emb_ch, int(m * mc))]
                ch = int(m * mc)
                if ds in attn_ds:
                    layers.append(AttentionBlock(ch, hc))
                self.input_blocks.append(_Seq(*layers))
                chans.append(ch)
            if level != len(mult) - 1:
                self.input_blocks.append(_Seq(ResBlock(ch, emb_ch, ch, down=True)))
                chans.append(ch)
                ds *= 2
        self.middle_block = _Seq(ResBlock(ch, emb_ch, ch), AttentionBlock(ch, hc), ResBlock(ch, emb_ch, ch))
        self.output_blocks = nn.ModuleList()
        for level, m in list(enumerate(mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + chans.pop(), emb_ch, int(m * mc))]
                ch = int(m * mc)
                if ds in attn_ds:
                    layers.append(AttentionBlock(ch, hc))
                if level and i == num_res_blocks:
                    layers.append(ResBlock(ch, emb_ch, ch, up=True))
                    ds //= 2
                self.output_blocks.append(_Seq(*layers))
        self.out_norm = GroupNorm32(32, ch)
        self.out_conv = nn.Conv2d(in_ch, 6 if learn_sigma else 3, 3, padding=1)
        if use_fp16:
            for part in (self.input_blocks, self.middle_block, self.output_blocks):
                for mod in part.modules():
                    if isinstance(mod, (nn.Conv1d, nn.Conv2d)):  # as guided-diffusion: convs only, Linear stays fp32
                        mod.half()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        invalidate_weight_caches(self)  # cached fused biases / embedding projection were built from the old weights
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if "_emb_cat" in self.__dict__:
            invalidate_weight_caches(self)
        return out

    def _project_embeddings(self, emb):
        """All ResBlocks' `emb_layers` (SiLU -> Linear(emb_ch, 2C)) as ONE fp32 GEMV over the concatenated weights instead of 49 SiLU +
        49 GEMV launches of ~7 us each (the timestep embedding is the same for every block; weights are frozen, so the
        concatenation is built once)."""
        if self._emb_cat is None:
            blocks = [m for m in self.modules() if isinstance(m, ResBlock)]
            w = torch.cat([b.emb.weight.detach() for b in blocks]).contiguous()
            bias = torch.cat([b.emb.bias.detach() for b in blocks]).contiguous()
            self._emb_cat = (blocks, w, bias, [b.emb.out_features for b in blocks])
        blocks, w, bias, widths = self._emb_cat
        all_ss = F.linear(F.silu(emb), w, bias)
        for blk, ss in zip(blocks, all_ss.split(widths, dim=1)):
            blk._scale_shift = ss

    def forward(self, x, timesteps, y=None):
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels))
        h = x.type(self.dtype)
        if self.channels_last:
            h = h.contiguous(memory_format=torch.channels_last)
            if h.is_cuda and h.dtype == torch.float16:
                self._project_embeddings(emb)
        hs = []
        for blk in self.input_blocks:
            h = blk(h, emb)
            hs.append(h)
        h = self.middle_block(h, emb)
        fused = self.channels_last and h.is_cuda and h.dtype == torch.float16
        if fused:
            from clip_diffusion_b200.unet_ops import concat_channels
        for blk in self.output_blocks:
            h = blk(concat_channels(h, hs.pop(), stats=True) if fused else torch.cat([h, hs.pop()], dim=1), emb)
        if self.channels_last and h.is_cuda and h.dtype == torch.float16:
            h = self.out_norm(h, silu=True, out_dtype=x.dtype)  # fp32 statistics and fp32 output from the fp16 trunk, one pass
        else:
            h = self.out_norm(h.type(x.dtype), silu=True)
        if self._emb_cat is not None:
            for blk in self._emb_cat[0]:
                blk._scale_shift = None  # per-forward slices: a ResBlock called on its own must not reuse them
        return self.out_conv(h)


def create_unet(image_size=512, seed=2, device="cuda", use_fp16=True, config=None, channels_last=None):
    """Random-init UNet (no checkpoints offline).  guided-diffusion zero-initialises the last conv of every ResBlock,
    the attention projections and the output conv; with random weights that would make epsilon identically 0, so
    they keep PyTorch's default init."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        model = UNetModel(image_size, use_fp16=use_fp16, config=config)
    model = model.to(device).eval().requires_grad_(False)
    if channels_last is None:
        channels_last = use_fp16 and torch.device(device).type == "cuda"
    if channels_last and not (use_fp16 and torch.device(device).type == "cuda"):
        raise ValueError("channels_last=True selects the sm_100a NHWC kernels: it needs device='cuda' and use_fp16=True (there is no CPU / fp32 variant)")
    if channels_last:
        # cuDNN's tensor-core convolutions are NHWC: keep the whole trunk NHWC instead of transposing around every conv, with the
        # normalisation / activation work between the convs in fused NHWC kernels (unet_ops.group_norm_nhwc)
        model = model.to(memory_format=torch.channels_last)
        model.channels_last = True
        for mod in model.modules():
            if isinstance(mod, (GroupNorm32, Resample)):
                mod.nhwc = True
    return model


_GD_RENAMES = (  # ours -> crowsonkb/guided-diffusion (the module the reference builds at models.py:90-117)
    (".in_norm.", ".in_layers.0."), (".in_conv.", ".in_layers.2."), (".emb.", ".emb_layers.1."), (".out_norm.", ".out_layers.0."),
    (".out_conv.", ".out_layers.3."), (".skip.", ".skip_connection."), (".proj.", ".proj_out."),
)


def guided_diffusion_key(ours):
    """Name of one of this UNet's parameters in a guided-diffusion ``UNetModel`` state dict (the checkpoint the reference loads
    with ``model.load_state_dict(torch.load(...))``, models.py:118-124)."""
    if ours.startswith("out_norm."):
        return "out.0." + ours[len("out_norm."):]
    if ours.startswith("out_conv."):
        return "out.2." + ours[len("out_conv."):]
    key = ours.replace(".layers.", ".")  # our _Seq wraps a ModuleList; guided-diffusion's TimestepEmbedSequential indexes directly
    for a, b in _GD_RENAMES:
        key = key.replace(a, b)
    return key


def invalidate_weight_caches(model):
    """The fused NHWC path caches tensors derived from the (frozen) weights: the concatenated embedding projection of all
    ResBlocks, each ResBlock's deferred biases and the per-forward scale-shift slices.  Call after changing weights in place;
    ``load_state_dict`` / ``_apply`` (``.to``, ``.half``...) do it automatically."""
    model._emb_cat = None
    for mod in model.modules():
        if isinstance(mod, ResBlock):
            mod._fused_bias = None
            mod._scale_shift = None


def load_guided_diffusion_state_dict(model, state_dict, strict=True):
    """Load a guided-diffusion checkpoint (512x512_diffusion_uncond_finetune_008100.pt and friends, models.py:118-124) into this
    UNet: same tensors, renamed (``guided_diffusion_key``).  Conv1d qkv/proj weights keep their [out, in, 1] shape.  Raises on
    missing / unexpected / mis-shaped entries when ``strict``."""
    own = model.state_dict()
    mapped, missing = {}, []
    for ours, ref in own.items():
        src = guided_diffusion_key(ours)
        if src not in state_dict:
            missing.append(src)
            continue
        t = state_dict[src]
        if tuple(t.shape) != tuple(ref.shape):
            raise ValueError("%s: checkpoint shape %s != model shape %s (%s)" % (src, tuple(t.shape), tuple(ref.shape), ours))
        mapped[ours] = t
    unexpected = sorted(set(state_dict) - {guided_diffusion_key(k) for k in own})
    if strict and (missing or unexpected):
        raise KeyError("guided-diffusion checkpoint does not match this UNet: %d missing (first %s), %d unexpected (first %s)"
                       % (len(missing), missing[:1], len(unexpected), unexpected[:1]))
    result = model.load_state_dict(mapped, strict=False)
    invalidate_weight_caches(model)
    return result


def graph_unet(model, height, width=None, device="cuda"):
    """Capture the replicated UNet's forward AND backward into CUDA graphs (``torch.cuda.make_graphed_callables``) for a fixed
    [1,3,H,W] input.  The UNet is ~7k small stock-PyTorch launches per guidance step and the host could not keep the GPU fed
    (profiles/r01_*): with graphs the step is bounded by device time instead of launch overhead.  Returns a callable with the
    module's signature ``(x, timesteps, y=None)``; ``x`` must require grad (use ``GuidanceStep.ddim_step``, which always
    evaluates the UNet once, with grad)."""
    from clip_diffusion_b200 import _lib

    width = width or height
    sx = torch.randn(1, 3, height, width, device=device, requires_grad=True)
    st = torch.full((1,), 500.0, device=device)
    # kernels of csrc/unet_norm.cu that one forward + backward launches (eager, counted by the ctypes binding): inside the graphs
    # they replay without passing through Python, so callers that report launch counts add this per replayed step
    k0 = _lib.kernel_launches
    torch.autograd.grad(model(sx, st).float().sum(), sx)
    own_kernels = _lib.kernel_launches - k0
    graphed = torch.cuda.make_graphed_callables(model, (sx, st))

    def call(x, timesteps, y=None):
        return graphed(x, timesteps.to(torch.float32))

    call.module = model
    call.own_kernels_per_replay = own_kernels
    return call
