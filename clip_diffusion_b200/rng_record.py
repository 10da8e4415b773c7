"""Host-side RNG record for one ``make_cutouts`` call.

The reference draws every random decision of ``Cutouts.forward`` from torch's global
CPU generator, interleaved with tensor work (clip_diffusion/cutouts.py:83-92 for the
inner-cut size/offsets, :31-45,113 for the torchvision augmentation chain).  To keep
crop sizes, offsets and augmentation parameters BIT-EXACT with the reference on the
same seed, this module consumes the generator in exactly the reference's order and
returns plain integers/floats that the fused CUDA kernels take as arguments:

  per inner cut i (cutouts.py:84-92):  torch.rand([]) -> size (float32 arithmetic,
      truncated), torch.randint(0, W-size+1, ()) -> x offset, torch.randint(0,
      H-size+1, ()) -> y offset
  RandomHorizontalFlip     torch.rand(1) < 0.5                 (whole batch)
  [noise 1]                randn_like: DEVICE generator on CUDA; on CPU the same CPU
                           generator (mode ``noise="cpu"`` reproduces that)
  RandomAffine.get_params  angle ~ U(-10,10); tx,ty = int(round(U(-.05cs,.05cs)))
  [noise 2]
  RandomGrayscale          torch.rand(1) < 0.1
  [noise 3]
  ColorJitter.get_params   torch.randperm(4); b,c,s ~ U(.9,1.1); h ~ U(-.1,.1)

All ranks of a multi-GPU job draw the FULL record (same seed => same record) and then
take their slice of the cutouts, so crops do not depend on the world size.
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch

# flag bits of one cutout descriptor (mirrors include/clipguide_b200.h)
CUT_GRAY_PRE = 1  # grayscale applied to the source crop before resampling (inner cuts, cutouts.py:102-103)
CUT_GRAY_POST = 2  # grayscale applied after resampling (overview variants, cutouts.py:72,76)
CUT_HFLIP = 4  # horizontal flip after resampling (overview variants, cutouts.py:74,76)
CUT_OVERVIEW = 8  # source is the whole image zero-padded to a square (cutouts.py:54-64)


@dataclass
class CutoutRecord:
    height: int
    width: int
    cut_size: int
    num_overview_cuts: int
    num_inner_cuts: int
    # per cutout, overview cuts first (cutouts.py:67-79 then :83-108)
    y0: List[int] = field(default_factory=list)
    x0: List[int] = field(default_factory=list)
    size: List[int] = field(default_factory=list)
    flags: List[int] = field(default_factory=list)
    # augmentation chain, one parameter set per call (the chain runs on the whole batch)
    flip: bool = False
    angle: float = 0.0
    tx: int = 0
    ty: int = 0
    gray: bool = False
    perm: List[int] = field(default_factory=lambda: [0, 1, 2, 3])
    brightness: float = 1.0
    contrast: float = 1.0
    saturation: float = 1.0
    hue: float = 0.0
    # explicit noise tensors [N,3,cs,cs] (parity mode) or None (generated in-kernel from noise_seed)
    noise: Optional[List[torch.Tensor]] = None
    noise_seed: int = 0
    # torch's CUDA randn stream for the three randn_like calls (cutouts.py:34,40,42): (generator seed, [offset at each call], threads of
    # torch's launch) as returned by cutouts.torch_noise_state(); None = the library's own keying of noise_seed
    noise_torch: Optional[tuple] = None

    @property
    def num_cuts(self) -> int:
        return self.num_overview_cuts + self.num_inner_cuts

    def inverse_affine_matrix(self) -> List[float]:
        """torchvision ``_get_inverse_affine_matrix(center=[0,0], angle, [tx,ty], scale=1, shear=[0,0])``
        (torchvision/transforms/functional.py:1006-1063), the matrix RandomAffine hands to the
        tensor ``affine`` path; python-float (double) arithmetic exactly as there."""
        rot = math.radians(self.angle)
        sx = sy = 0.0
        a = math.cos(rot - sy) / math.cos(sy)
        b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
        c = math.sin(rot - sy) / math.cos(sy)
        d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
        m = [d, -b, 0.0, -c, a, 0.0]
        m[2] += m[0] * (-float(self.tx)) + m[1] * (-float(self.ty))
        m[5] += m[3] * (-float(self.tx)) + m[4] * (-float(self.ty))
        return m

    def slice(self, start: int, stop: int) -> "CutoutRecord":
        """The record of cutouts [start, stop) -- one rank's shard.  Augmentation parameters
        are per call, hence shared by every shard."""
        r = CutoutRecord(self.height, self.width, self.cut_size, 0, stop - start)
        r.y0, r.x0 = self.y0[start:stop], self.x0[start:stop]
        r.size, r.flags = self.size[start:stop], self.flags[start:stop]
        for k in ("flip", "angle", "tx", "ty", "gray", "perm", "brightness", "contrast", "saturation", "hue", "noise_seed", "noise_torch"):
            setattr(r, k, getattr(self, k))
        if self.noise is not None:
            r.noise = [n[start:stop] for n in self.noise]
        r._slice_of = (start, self.num_cuts)
        return r

    def first_index(self) -> int:
        return getattr(self, "_slice_of", (0, self.num_cuts))[0]

    def total_cuts(self) -> int:
        """Cutouts of the whole batch this record (or shard) belongs to."""
        return getattr(self, "_slice_of", (0, self.num_cuts))[1]


def draw_cutout_record(
    height: int,
    width: int,
    cut_size: int,
    num_overview_cuts: int,
    num_inner_cuts: int,
    inner_cut_size_power,
    cut_gray_portion,
    generator: Optional[torch.Generator] = None,
    noise: str = "device",
) -> CutoutRecord:
    """Consume the (global, unless ``generator`` is given) CPU generator in the reference's order.

    noise="device": the three noise tensors are NOT drawn here (the reference draws them from
        the device generator when its tensors live on CUDA); the kernel generates them from
        ``noise_seed``, which the caller takes from the CUDA generator's Philox (seed, offset)
        so that -- as in the reference -- no extra CPU draw is consumed.
    noise="cpu":    draw them from the same CPU generator at the reference's positions, which
        reproduces the reference running on CPU bit for bit (parity tests).
    """
    assert noise in ("device", "cpu")
    g = generator
    rec = CutoutRecord(height, width, cut_size, num_overview_cuts, num_inner_cuts)
    shorter_side = min(width, height)
    longer_side = max(width, height)
    min_size = min(width, height, cut_size)
    assert (longer_side - shorter_side) % 2 == 0, "square padding needs an even side difference (cutouts.py:54-62)"

    # overview cuts: the whole image zero-padded to a square of the longer side (cutouts.py:54-79)
    pad_h = (longer_side - height) // 2
    pad_w = (longer_side - width) // 2
    if 0 < num_overview_cuts <= 4:
        variants = [0, CUT_GRAY_POST, CUT_HFLIP, CUT_GRAY_POST | CUT_HFLIP][:num_overview_cuts]
    else:
        variants = [0] * num_overview_cuts
    for v in variants:
        rec.y0.append(-pad_h)
        rec.x0.append(-pad_w)
        rec.size.append(longer_side)
        rec.flags.append(v | CUT_OVERVIEW)

    # inner cuts (cutouts.py:83-108)
    for i in range(num_inner_cuts):
        r = torch.rand([], generator=g)
        size = int(r ** inner_cut_size_power * (shorter_side - min_size) + min_size)
        x_off = int(torch.randint(0, width - size + 1, (), generator=g))
        y_off = int(torch.randint(0, height - size + 1, (), generator=g))
        rec.y0.append(y_off)
        rec.x0.append(x_off)
        rec.size.append(size)
        rec.flags.append(CUT_GRAY_PRE if i <= int(cut_gray_portion * num_inner_cuts) else 0)

    n = rec.num_cuts
    shape = (n, 3, cut_size, cut_size)
    noises = []

    def _noise():
        if noise == "cpu":
            noises.append(torch.randn(shape, generator=g))

    rec.flip = bool(torch.rand(1, generator=g) < 0.5)
    _noise()
    rec.angle = float(torch.empty(1).uniform_(-10.0, 10.0, generator=g).item())
    max_d = float(0.05 * cut_size)
    rec.tx = int(round(torch.empty(1).uniform_(-max_d, max_d, generator=g).item()))
    rec.ty = int(round(torch.empty(1).uniform_(-max_d, max_d, generator=g).item()))
    _noise()
    rec.gray = bool(torch.rand(1, generator=g) < 0.1)
    _noise()
    rec.perm = [int(v) for v in torch.randperm(4, generator=g)]
    rec.brightness = float(torch.empty(1).uniform_(0.9, 1.1, generator=g))
    rec.contrast = float(torch.empty(1).uniform_(0.9, 1.1, generator=g))
    rec.saturation = float(torch.empty(1).uniform_(0.9, 1.1, generator=g))
    rec.hue = float(torch.empty(1).uniform_(-0.1, 0.1, generator=g))
    if noise == "cpu":
        rec.noise = noises
    return rec
