"""Global settings read by the guidance path at CALL time (mirror of clip_diffusion/config.py:17-80).

The reference keeps its knobs as mutable class attributes that ``Config.update`` rewrites
(config.py:54-80) and that ``conditon_function`` reads on every call (sample.py:162-238), so the
drop-in keeps the same names and the same lazily-read class-attribute semantics.
"""
import torch


def create_schedule(values, steps):
    """(values[0],)*steps[0] + (values[1],)*steps[1] + ...   (config.py:4-14)"""
    if len(values) != len(steps):
        raise AssertionError("length of values and steps must be the same")
    out = []
    for v, n in zip(values, steps):
        out.extend([v] * n)
    return tuple(out)


_DEFAULTS = dict(
    width=768,
    height=512,
    num_cutout_batches=4,
    chosen_clip_models=("ViT-B/32", "ViT-B/16", "ViT-L/14", "RN101"),
    chosen_predictors=("ViT-B/32", "ViT-B/16", "ViT-L/14"),
    grad_threshold=0.05,
    clip_guidance_scale=8000,
    denoise_scale=10000,
    LPIPS_scale=1000,
    aesthetic_scale=0,
    MS_SSIM_scale=0,
)


class Config:
    device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")

    # 1000-entry schedules indexed by the original diffusion step (sample.py:157-171)
    num_overview_cuts_schedule = create_schedule((14, 12, 4, 0), (200, 200, 400, 200))
    num_inner_cuts_schedule = create_schedule((2, 4, 2, 12), (200, 200, 400, 200))
    inner_cut_size_power_schedule = create_schedule((5,), (1000,))
    cut_gray_portion_schedule = create_schedule((0.7, 0.6, 0.45, 0.3, 0), (100, 100, 100, 100, 600))

    @classmethod
    def update(cls, **kwargs):
        """Same keywords and defaults as config.py:54-80; width/height are floored to multiples of 64."""
        unknown = set(kwargs) - set(_DEFAULTS)
        if unknown:
            raise TypeError("Config.update() got unexpected keyword(s): %s" % sorted(unknown))
        merged = dict(_DEFAULTS)
        merged.update(kwargs)
        merged["width"] = (merged["width"] // 64) * 64
        merged["height"] = (merged["height"] // 64) * 64
        for k, v in merged.items():
            setattr(cls, k, v)


for _k, _v in _DEFAULTS.items():
    setattr(Config, _k, _v)
