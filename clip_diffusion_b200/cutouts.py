"""Drop-in for clip_diffusion/cutouts.py: ``Cutouts`` / ``make_cutouts`` on fused sm_100a kernels.

Same constructor/call signatures as the reference (cutouts.py:17-24, :47, :117-124).  The random
decisions are drawn on the host from torch's global CPU generator in the reference's order
(rng_record.py) -- crop sizes and offsets are therefore bit-exact with the reference on the same
seed -- and handed to ``cg_cutouts_fwd`` as plain integers/floats.  The result carries an autograd
graph back to ``input`` (sample.py differentiates through it at :201-213); its backward is
``cg_cutouts_bwd`` (gather-form scatter-add into d(loss)/d(x_in)).

``MakeCutouts`` is the north_star alias of ``Cutouts``.
"""
import ctypes as C

import numpy as np
import torch
from torch import nn

from clip_diffusion_b200 import _lib
from clip_diffusion_b200.rng_record import CutoutRecord, draw_cutout_record

__all__ = ["Cutouts", "MakeCutouts", "make_cutouts", "make_cutouts_from_record", "cutouts_forward", "cutouts_backward", "torch_noise_state"]

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def torch_noise_state(device, num_cuts, cut_size):
    """What the reference's three ``torch.randn_like(input)`` calls (cutouts.py:34,40,42) do to the CUDA generator of ``device`` for a
    ``[num_cuts, 3, cut_size, cut_size]`` float32 batch: returns ``(seed, [offset at call 0, 1, 2], launch threads)`` and advances the
    generator's Philox offset exactly like the three calls would.  The kernels regenerate those tensors element by element
    (``cg_aug_t.noise_mode = 1``), so on identical seeds the drop-in equals the reference running on CUDA -- not just statistically."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    numel = int(num_cuts) * 3 * int(cut_size) * int(cut_size)
    inc = C.c_uint64(0)
    threads = int(_lib.load().cg_randn_like_torch_geometry(numel, C.byref(inc)))
    off = int(gen.get_offset())
    offsets = [off, off + inc.value, off + 2 * inc.value]
    gen.set_offset(off + 3 * inc.value)
    return int(gen.initial_seed()), offsets, threads


def _device_noise_seed(device):
    """Legacy keying (noise_mode 0): a Philox key derived from the CUDA generator's (seed, offset); advances the offset by 4."""
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    seed = gen.initial_seed()
    off = gen.get_offset()
    gen.set_offset(off + 4)
    return (seed * 0x9E3779B97F4A7C15 + off) & 0xFFFFFFFFFFFFFFFF


def pack_record(rec: CutoutRecord, augment=True, normalize=False, input01=False):
    """CutoutRecord -> (CgCut array, CgAug) host structs of include/clipguide_b200.h."""
    n = rec.num_cuts
    cuts = (_lib.CgCut * n)()
    for i in range(n):
        cuts[i].y0, cuts[i].x0, cuts[i].size, cuts[i].flags = rec.y0[i], rec.x0[i], rec.size[i], rec.flags[i]
    aug = _lib.CgAug()
    aug.flip, aug.gray = int(rec.flip), int(rec.gray)
    for i in range(4):
        aug.perm[i] = rec.perm[i]
    m = rec.inverse_affine_matrix()
    # input -> output map = inverse of the (output -> input) matrix, in double
    a, b, tx, c, d, ty = m
    det = a * d - b * c
    ia, ib, ic, id_ = d / det, -b / det, -c / det, a / det
    fwd = [ia, ib, -(ia * tx + ib * ty), ic, id_, -(ic * tx + id_ * ty)]
    for i in range(6):
        aug.theta[i] = np.float32(m[i])
        aug.theta_fwd[i] = fwd[i]
    aug.brightness, aug.contrast, aug.saturation, aug.hue = rec.brightness, rec.contrast, rec.saturation, rec.hue
    aug.augment, aug.normalize = int(augment), int(normalize)
    for i in range(3):
        aug.mean[i], aug.stdv[i] = CLIP_MEAN[i], CLIP_STD[i]
    aug.noise_seed = rec.noise_seed & 0xFFFFFFFFFFFFFFFF
    aug.cut_index0 = rec.first_index()
    if rec.noise_torch is not None:
        seed, offsets, threads = rec.noise_torch
        aug.noise_mode, aug.noise_seed, aug.noise_threads = 1, seed & 0xFFFFFFFFFFFFFFFF, threads
        for i in range(3):
            aug.noise_offset[i] = offsets[i]
        aug.noise_total = rec.total_cuts()
    aug.noise_std = 0.01
    aug.input01 = int(input01)
    return cuts, aug


def cutouts_forward(x, rec, fmt=_lib.CG_FMT_F32_NCHW, patch=0, kpad=0, augment=True, normalize=False, input01=False):
    """Raw (no autograd) fused forward.  x: [1,3,H,W] or [3,H,W] fp32 CUDA.  Returns (out, ctx) where ctx is
    what ``cutouts_backward`` needs."""
    _lib.require_cuda(x)
    x3 = x.reshape(x.shape[-3:]).contiguous().float()
    if x3.shape[0] != 3:
        raise ValueError("expected a 3-channel image, got %s" % (tuple(x.shape),))
    H, W = int(x3.shape[1]), int(x3.shape[2])
    n, cs = rec.num_cuts, rec.cut_size
    if n == 0:
        raise ValueError("no cutouts requested (torch.cat of an empty list fails in the reference too, cutouts.py:111)")
    lib = _lib.load()
    ws_bytes = lib.cg_cutouts_workspace_bytes(n, cs, min(max(H, W), 2048))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    cuts, aug = pack_record(rec, augment, normalize, input01)
    noise = None
    if rec.noise is not None:
        noise = torch.stack([t.to(x.device, torch.float32) for t in rec.noise]).contiguous()
    if fmt == _lib.CG_FMT_F32_NCHW:
        out = torch.empty(n, 3, cs, cs, device=x.device, dtype=torch.float32)
    else:
        g = cs // patch
        out = torch.empty(n, g * g, kpad, device=x.device, dtype=torch.bfloat16)
    _lib.call("cg_cutouts_fwd", _lib.ptr(x3), H, W, cuts, n, cs, C.byref(aug), _lib.ptr(noise), _lib.ptr(out), fmt, patch, kpad, _lib.ptr(ws))
    return out, (ws, H, W, n, cs, fmt, patch, kpad, int(input01))


def cutouts_backward(dout, ctx, coef=1.0, dx_in=None):
    """dout -> d(x_in) [3,H,W]; accumulates into ``dx_in`` when given."""
    ws, H, W, n, cs, fmt, patch, kpad, input01 = ctx
    accumulate = dx_in is not None
    if dx_in is None:
        dx_in = torch.empty(3, H, W, device=dout.device, dtype=torch.float32)
    dout = dout.contiguous()
    if fmt == _lib.CG_FMT_BF16_PATCH and dout.dtype == torch.float32:
        fmt = _lib.CG_FMT_F32_PATCH  # the conv1 dgrad GEMM hands over fp32 patch-major gradients
    _lib.call("cg_cutouts_bwd", _lib.ptr(dout), H, W, n, cs, fmt, patch, kpad, float(coef), int(accumulate), input01, _lib.ptr(dx_in), _lib.ptr(ws))
    return dx_in


class _CutoutsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, rec, input01):
        if input.dim() != 4 or input.shape[0] != 1:
            raise ValueError("make_cutouts expects a [1,3,H,W] image (sample.py:246-251), got %s" % (tuple(input.shape),))
        out, saved = cutouts_forward(input, rec, augment=True, normalize=False, input01=input01)
        ctx.saved = saved
        ctx.in_shape, ctx.in_dtype = input.shape, input.dtype
        return out

    @staticmethod
    def backward(ctx, dout):
        dx = cutouts_backward(dout.float(), ctx.saved)
        return dx.reshape(ctx.in_shape).to(ctx.in_dtype), None, None


def _draw(input, cut_size, num_overview_cuts, num_inner_cuts, inner_cut_size_power, cut_gray_portion):
    height, width = int(input.shape[2]), int(input.shape[3])
    rec = draw_cutout_record(height, width, cut_size, num_overview_cuts, num_inner_cuts, inner_cut_size_power, cut_gray_portion, noise="device")
    rec.noise_torch = torch_noise_state(input.device, rec.num_cuts, cut_size)
    return rec


class Cutouts(nn.Module):
    """cutouts.py:10-114.  ``forward(input)`` takes the image already in [0,1] like the reference's."""

    def __init__(self, cut_size, num_overview_cuts, num_inner_cuts, inner_cut_size_power, cut_gray_portion):
        super().__init__()
        self.cut_size = cut_size
        self.num_overview_cuts = num_overview_cuts
        self.num_inner_cuts = num_inner_cuts
        self.inner_cut_size_power = inner_cut_size_power
        self.cut_gray_portion = cut_gray_portion
        self.last_record = None  # the RNG record of the most recent call (crop sizes / offsets / flags)

    def forward(self, input, _input01=True):
        _lib.require_cuda(input)
        rec = _draw(input, self.cut_size, self.num_overview_cuts, self.num_inner_cuts, self.inner_cut_size_power, self.cut_gray_portion)
        self.last_record = rec
        return _CutoutsFn.apply(input, rec, _input01)


MakeCutouts = Cutouts


def make_cutouts(input, cut_size, num_overview_cuts, num_inner_cuts, inner_cut_size_power, cut_gray_portion):
    """cutouts.py:117-134: ``input`` in [-1,1] -> [N,3,cs,cs] in [0,1] (denormalisation fused into the kernel)."""
    cutouts = Cutouts(cut_size, num_overview_cuts, num_inner_cuts, inner_cut_size_power, cut_gray_portion)
    return cutouts(input, _input01=False)


def make_cutouts_from_record(input, rec, input01=False):
    """Same op with an explicit RNG record (parity tests, sharded cond_fn)."""
    return _CutoutsFn.apply(input, rec, input01)
