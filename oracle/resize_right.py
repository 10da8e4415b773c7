"""Oracle restatement of ``resize_right.resize`` (assafshocher/ResizeRight), TEST-ONLY.

The reference calls ``resize(x, out_shape=[1, 3, cs, cs])`` with all defaults at
clip_diffusion/cutouts.py:64 and :105.  The package is not vendored in
/root/reference (requirements.txt:1, no version pin) so this is a restatement of
its published algorithm -- parity against the real package is
UNPINNED; tests cross-check it against torch's antialiased bicubic on interior
pixels and against exact identity when no scaling happens.

Defaults restated: cubic (Keys a=-0.5) kernel with support 4, antialiasing on,
by_convs off, pad_mode "constant" (zeros), float32 arithmetic, dimensions
processed in ascending scale order (stable => H before W for square crops).
"""
from math import ceil

import torch


def cubic(x: torch.Tensor) -> torch.Tensor:
    absx = x.abs()
    absx2 = absx ** 2
    absx3 = absx ** 3
    return (1.5 * absx3 - 2.5 * absx2 + 1.0) * (absx <= 1.0).to(x.dtype) + (
        -0.5 * absx3 + 2.5 * absx2 - 4.0 * absx + 2.0
    ) * ((1.0 < absx) & (absx <= 2.0)).to(x.dtype)


def dim_tables(in_sz: int, out_sz: int):
    """Per-dimension field of view and weights.

    Returns (left [out] int64 in UNPADDED input coordinates, weights [out, taps] fp32).
    A tap whose coordinate falls outside [0, in_sz) reads a zero (constant pad) but
    still takes part in the normalisation of the weights.
    """
    eps = torch.finfo(torch.float32).eps
    scale = out_sz / in_sz
    out_coordinates = torch.arange(out_sz)
    projected_grid = out_coordinates / float(scale) + (in_sz - 1) / 2 - (out_sz - 1) / (2 * float(scale))
    if scale < 1.0:
        support = 4.0 / scale
        kernel = lambda a: scale * cubic(scale * a)
    else:
        support = 4.0
        kernel = cubic
    left = (projected_grid - support / 2 - eps).ceil().long()
    taps = ceil(support - eps)
    fov = left[:, None] + torch.arange(taps)
    # the package pads then shifts grid and fov by the same amount; the difference is unchanged
    pad0 = -int(fov[0, 0])
    weights = kernel((projected_grid + pad0)[:, None] - (fov + pad0))
    s = weights.sum(1, keepdim=True)
    s[s == 0] = 1
    weights = weights / s
    return left, weights


def _resize_dim(x: torch.Tensor, dim: int, out_sz: int) -> torch.Tensor:
    in_sz = x.shape[dim]
    left, weights = dim_tables(in_sz, out_sz)
    taps = weights.shape[1]
    fov = left[:, None] + torch.arange(taps)
    pad = (max(0, -int(fov.min())), max(0, int(fov.max()) - in_sz + 1))
    t = x.transpose(dim, 0)
    padded = torch.zeros((t.shape[0] + pad[0] + pad[1],) + tuple(t.shape[1:]), dtype=x.dtype)
    padded[pad[0] : pad[0] + in_sz] = t
    neighbors = padded[fov + pad[0]]  # [out, taps, ...]
    w = weights.reshape(*weights.shape, *([1] * (x.ndim - 1))).to(x.dtype)
    out = (neighbors * w).sum(1)
    return out.transpose(0, dim)


def resize(input: torch.Tensor, scale_factors=None, out_shape=None, **_unused) -> torch.Tensor:
    """``resize_right.resize`` for the reference's usage: 4-D input, ``out_shape`` given."""
    assert out_shape is not None and scale_factors is None
    out_shape = list(out_shape)
    out_shape = list(input.shape[: input.ndim - len(out_shape)]) + out_shape
    scales = [o / i for o, i in zip(out_shape, input.shape)]
    order = [d for d in sorted(range(input.ndim), key=lambda d: scales[d]) if scales[d] != 1.0]
    out = input
    for d in order:
        out = _resize_dim(out, d, out_shape[d])
    return out
