"""Oracle for the replicated UNet's normalisation blocks -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two restatements, both float64 numpy/torch on the CPU:

* ``resblock_norm_reference``  -- the block arithmetic as the un-vendored guided-diffusion spells it (the
  model is built at clip_diffusion/models.py:87-131): ``GroupNorm32`` = ``F.group_norm(x.float(), 32, w, b, eps)`` (nn.py
  GroupNorm32), scale-shift norm ``out_norm(h) * (1 + scale) + shift`` and ``SiLU`` (unet.py ResBlock._forward, use_scale_shift_norm),
  with the producing convolution's bias written out explicitly (``x + conv_bias``).  **Parity unpinned** against the package
  itself (not on disk); it is the textbook definition and torch's own ``group_norm`` carries it.
* ``folded_forward`` / ``folded_backward`` -- the ALGEBRA csrc/unet_norm.cu uses: per-channel sums of the raw conv output,
  statistics of ``x + bias`` derived from them (``sum x' = s + HW*b``, ``sum x'^2 = q + 2*b*s + HW*b^2``), one per-channel affine
  ``y = act(a_c*x + b_c)``, and the closed-form input gradient ``dx = a_c*dv + B_g*x + C_g (+ dres)``.  tests/test_properties_cpu.py
  checks it against autograd of the reference above, so the kernels' formulas are pinned independently of any GPU run.
"""
import numpy as np
import torch
from torch.nn import functional as F


def resblock_norm_reference(x, gamma, beta, groups, eps=1e-5, scale_shift=None, conv_bias=None, silu=False):
    """x [N,C,H,W]; returns act(GN(x + conv_bias) * (1 + scale) + shift) in x's dtype (use float64 tensors for an oracle)."""
    if conv_bias is not None:
        x = x + conv_bias.view(1, -1, 1, 1)
    y = F.group_norm(x, groups, gamma, beta, eps)
    if scale_shift is not None:
        c = x.shape[1]
        y = y * (1 + scale_shift[:, :c, None, None]) + scale_shift[:, c:, None, None]
    return F.silu(y) if silu else y


def _silu(v):
    return v / (1.0 + np.exp(-v))


def _dsilu(v):
    s = 1.0 / (1.0 + np.exp(-v))
    return s * (1.0 + v * (1.0 - s))


def folded_coefficients(x, gamma, beta, groups, eps, scale_shift=None, conv_bias=None):
    """Per-channel sums -> group statistics of x + bias -> (a [N,C], b [N,C], mean [N,G], rstd [N,G]) as gn_finalize_fwd_kernel does."""
    n, c, h, w = x.shape
    hw, cg = h * w, c // groups
    s = x.reshape(n, c, hw).sum(-1)
    q = (x.reshape(n, c, hw) ** 2).sum(-1)
    pb = np.zeros(c) if conv_bias is None else conv_bias
    s1 = s + hw * pb
    q1 = q + 2.0 * pb * s + hw * pb * pb
    m = float(hw * cg)
    mean = s1.reshape(n, groups, cg).sum(-1) / m
    var = q1.reshape(n, groups, cg).sum(-1) / m - mean * mean
    rstd = 1.0 / np.sqrt(np.maximum(var, 0.0) + eps)
    a = gamma[None] * np.repeat(rstd, cg, axis=1)
    b = beta[None] - np.repeat(mean, cg, axis=1) * a
    if scale_shift is not None:
        sc = 1.0 + scale_shift[:, :c]
        a = a * sc
        b = b * sc + scale_shift[:, c:]
    b = b + a * pb[None]
    return a, b, mean, rstd


def folded_forward(x, gamma, beta, groups, eps=1e-5, scale_shift=None, conv_bias=None, silu=False):
    a, b, _, _ = folded_coefficients(x, gamma, beta, groups, eps, scale_shift, conv_bias)
    v = a[:, :, None, None] * x + b[:, :, None, None]
    return _silu(v) if silu else v


def folded_backward(dy, x, gamma, beta, groups, eps=1e-5, scale_shift=None, conv_bias=None, silu=False, dres=None):
    """dx = a_c*dv + B_g*x + C'_c (+ dres) with the per-channel sums S1 = sum dv, S2 = sum dv*x (gn_bwd_partial / gn_finalize_bwd)."""
    n, c, h, w = x.shape
    hw, cg = h * w, c // groups
    a, b, mean, rstd = folded_coefficients(x, gamma, beta, groups, eps, scale_shift, conv_bias)
    pb = np.zeros(c) if conv_bias is None else conv_bias
    v = a[:, :, None, None] * x + b[:, :, None, None]
    dv = dy * _dsilu(v) if silu else dy
    s1 = dv.reshape(n, c, hw).sum(-1)
    s2 = (dv * x).reshape(n, c, hw).sum(-1) + pb[None] * s1          # sum dv * (x + bias)
    wgt = a / np.repeat(rstd, cg, axis=1)                              # gamma_c * (1 + scale_c)
    t1 = (wgt * s1).reshape(n, groups, cg).sum(-1)
    t2 = (wgt * s2).reshape(n, groups, cg).sum(-1)
    m = float(hw * cg)
    c1 = t1 / m
    c2 = rstd * (t2 - mean * t1) / m
    bg = -rstd * rstd * c2
    cgp = -rstd * c1 - bg * mean
    bx = np.repeat(bg, cg, axis=1)
    cx = np.repeat(cgp, cg, axis=1) + bx * pb[None]
    dx = a[:, :, None, None] * dv + bx[:, :, None, None] * x + cx[:, :, None, None]
    return dx if dres is None else dx + dres
