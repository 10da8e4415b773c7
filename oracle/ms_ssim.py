"""Oracle (TEST-ONLY, CPU) for the init-image structural dissimilarity of conditon_function
(clip_diffusion/sample.py:220-225 -> clip_diffusion/losses.py:48-54 -> ``_ms_ssim_loss`` at losses.py:7).

The reference takes MS-SSIM from the third-party package ``pytorch_msssim`` (requirements.txt:19, no version pinned, NOT vendored
and not installable here: no network).  **Parity unpinned** against the real package: this is a restatement of its published algorithm
(``pytorch_msssim/ssim.py``: ``_fspecial_gauss_1d``, ``gaussian_filter``, ``_ssim``, ``ms_ssim``), anchored on the reference's own call
site -- ``MS_SSIM(win_size=11, win_sigma=1.5, data_range=1, size_average=True, channel=3)`` applied to both images mapped to [0,1]
(``denormalize_image_zero_to_one``, image_utils.py:40-42) and used as ``1 - ms_ssim``.  Written for float64 so that the CUDA kernel
(csrc/msssim.cu) and its analytic gradient are checked against autograd of an independent, higher-precision definition.

  * window: g[i] = exp(-(i - 5)^2 / (2 * 1.5^2)), i = 0..10, normalised to sum 1;
  * per level: mu = g*x (separable "valid" correlation), sigma11 = g*x^2 - mu1^2, ...; cs = (2 sigma12 + C2) / (sigma11 + sigma22 + C2),
    ssim = (2 mu1 mu2 + C1) / (mu1^2 + mu2^2 + C1) * cs, C1 = (0.01 L)^2, C2 = (0.03 L)^2; the spatial mean per channel;
  * 5 levels, 2x2 average pooling (padding = side % 2) between them, relu on every level's term, weights
    (0.0448, 0.2856, 0.3001, 0.2363, 0.1333): ms = prod_l term_l ^ w_l per channel, then the mean over batch and channels.
"""
import torch
from torch.nn import functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def gaussian_window(size=11, sigma=1.5, dtype=torch.float64):
    coords = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x, win):
    c = x.shape[1]
    k = win.to(x.dtype).view(1, 1, -1).repeat(c, 1, 1)
    if x.shape[2] >= win.numel():
        x = F.conv2d(x, k.unsqueeze(-1), groups=c)
    if x.shape[3] >= win.numel():
        x = F.conv2d(x, k.unsqueeze(-2), groups=c)
    return x


def ms_ssim(x, y, data_range=1.0, win_size=11, win_sigma=1.5, weights=WEIGHTS, k1=0.01, k2=0.03):
    if min(x.shape[-2:]) <= (win_size - 1) * 2 ** 4:
        raise ValueError("image side must exceed %d for 5-scale MS-SSIM" % ((win_size - 1) * 2 ** 4))
    win = gaussian_window(win_size, win_sigma, x.dtype)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    terms = []
    for level in range(len(weights)):
        mu1, mu2 = _filter(x, win), _filter(y, win)
        s11 = _filter(x * x, win) - mu1 * mu1
        s22 = _filter(y * y, win) - mu2 * mu2
        s12 = _filter(x * y, win) - mu1 * mu2
        cs_map = (2 * s12 + c2) / (s11 + s22 + c2)
        if level < len(weights) - 1:
            terms.append(torch.relu(cs_map.flatten(2).mean(-1)))
            pad = [s % 2 for s in x.shape[2:]]
            x = F.avg_pool2d(x, kernel_size=2, padding=pad)
            y = F.avg_pool2d(y, kernel_size=2, padding=pad)
        else:
            ssim_map = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs_map
            terms.append(torch.relu(ssim_map.flatten(2).mean(-1)))
    stack = torch.stack(terms, dim=0)  # [levels, B, C]
    w = torch.tensor(weights, dtype=x.dtype).view(-1, 1, 1)
    return torch.prod(stack ** w, dim=0).mean()  # size_average=True


def structural_dissimilarity_loss(input, image):
    """losses.py:48-54"""
    return 1.0 - ms_ssim((input + 1) / 2, (image + 1) / 2)
