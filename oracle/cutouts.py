"""Oracle restatement of clip_diffusion/cutouts.py:47-134 (TEST-ONLY, CPU torch).

Same tensor operations as the reference's ``Cutouts.forward`` -- including the real
torchvision functional ops the reference's transform classes dispatch to -- but every
random decision comes from an explicit ``CutoutRecord`` instead of the global
generator, so the CUDA kernels can be compared with it on identical parameters.
Pinned bit-exactly against the reference module executed in place
(tests/test_oracle_pins.py, build container) and by tests/golden/cutouts_*.pt.
"""
import torch
from torch.nn import functional as F
from torchvision.transforms import InterpolationMode
from torchvision.transforms import functional as TF

from oracle.resize_right import resize

CUT_GRAY_PRE, CUT_GRAY_POST, CUT_HFLIP, CUT_OVERVIEW = 1, 2, 4, 8


def _gray3(x):
    # torchvision T.Grayscale(3) on a float tensor (cutouts.py:49)
    return TF.rgb_to_grayscale(x, num_output_channels=3)


def base_cutouts(image01: torch.Tensor, rec) -> torch.Tensor:
    """cutouts.py:47-111: square-pad, overview resize + variants, inner crops, gray, resize, cat.
    ``image01`` is the [1,3,H,W] image already in [0,1] (cutouts.py:133)."""
    height, width = image01.shape[2:4]
    shorter_side = min(width, height)
    cs = rec.cut_size
    out_shape = [1, 3, cs, cs]
    pad_input = F.pad(
        image01,
        ((height - shorter_side) // 2, (height - shorter_side) // 2, (width - shorter_side) // 2, (width - shorter_side) // 2),
    )
    cut_size_input = resize(pad_input, out_shape=out_shape)
    cuts = []
    for n in range(rec.num_cuts):
        fl = rec.flags[n]
        is_overview = bool(fl & CUT_OVERVIEW)
        if is_overview:
            c = cut_size_input
            if fl & CUT_HFLIP:
                c = TF.hflip(c)
            if fl & CUT_GRAY_POST:
                c = _gray3(c)
        else:
            y, x, s = rec.y0[n], rec.x0[n], rec.size[n]
            c = image01[:, :, y : y + s, x : x + s]
            if fl & CUT_GRAY_PRE:
                c = _gray3(c)
            c = resize(c, out_shape=out_shape)
        cuts.append(c)
    return torch.cat(cuts)


def augment(cuts: torch.Tensor, rec) -> torch.Tensor:
    """cutouts.py:31-45 applied to the whole batch (cutouts.py:113) with recorded parameters."""
    n1, n2, n3 = rec.noise
    x = cuts
    if rec.flip:
        x = TF.hflip(x)
    x = x + n1 * 0.01
    x = TF.affine(
        x, angle=rec.angle, translate=[rec.tx, rec.ty], scale=1.0, shear=[0.0, 0.0],
        interpolation=InterpolationMode.BILINEAR, fill=[0.0, 0.0, 0.0],
    )
    x = x + n2 * 0.01
    if rec.gray:
        x = TF.rgb_to_grayscale(x, num_output_channels=3)
    x = x + n3 * 0.01
    for fn_id in rec.perm:
        if fn_id == 0:
            x = TF.adjust_brightness(x, rec.brightness)
        elif fn_id == 1:
            x = TF.adjust_contrast(x, rec.contrast)
        elif fn_id == 2:
            x = TF.adjust_saturation(x, rec.saturation)
        elif fn_id == 3:
            x = TF.adjust_hue(x, rec.hue)
    return x


def make_cutouts(input: torch.Tensor, rec) -> torch.Tensor:
    """cutouts.py:117-134: ``input`` in [-1,1] -> [N,3,cs,cs] in [0,1]."""
    return augment(base_cutouts(input.add(1).div(2), rec), rec)


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def clip_normalize(x: torch.Tensor) -> torch.Tensor:
    """utils/functional.py:16-18 CLIP_NORMALIZE."""
    return TF.normalize(x, CLIP_MEAN, CLIP_STD)
