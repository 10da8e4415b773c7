"""Oracle restatement of the OpenAI CLIP VisionTransformer (TEST-ONLY, CPU torch fp32).

The reference obtains its CLIP towers from the un-vendored ``clip`` package
(clip_diffusion/models.py:76-80, utils/functional.py:101); nothing of it is under
/root/reference, so this restates the published architecture:
conv1 patch embed (no bias) -> [cls] + positional -> ln_pre -> L x {x += MHA(ln_1 x);
x += c_proj(QuickGELU(c_fc(ln_2 x)))} -> ln_post(x[:,0]) @ proj.  Parity against
openai/CLIP itself is UNPINNED; tests cross-check against transformers'
CLIPVisionModelWithProjection(hidden_act="quick_gelu") with mapped weights.

State-dict keys follow OpenAI's names under ``visual.`` so real checkpoints map 1:1.
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

CONFIGS = {
    # name: (input_resolution, patch, width, layers, heads, embed_dim)   (models.py:33-37 for E)
    "ViT-B/32": (224, 32, 768, 12, 12, 512),
    "ViT-B/16": (224, 16, 768, 12, 12, 512),
    "ViT-L/14": (224, 14, 1024, 24, 16, 768),
    "ViT-L/14@336px": (336, 14, 1024, 24, 16, 768),
    # tiny towers for fast tests
    "test-tiny/32": (64, 32, 128, 2, 2, 64),
    "test-small/16": (64, 16, 256, 3, 4, 128),
    "test-k/14": (56, 14, 128, 2, 2, 64),
}


def random_state_dict(name: str, seed: int = 1):
    """Constructor-default random init of the OpenAI tower (App. A.1), as plain tensors."""
    res, patch, width, layers, heads, embed = CONFIGS[name]
    g = torch.Generator().manual_seed(seed)
    scale = width ** -0.5

    def randn(*s):
        return torch.randn(*s, generator=g)

    def kaiming_uniform(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound

    def bias_uniform(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        return (torch.rand(out_f, generator=g) * 2 - 1) * bound

    sd = {}
    fan_in = 3 * patch * patch
    sd["visual.conv1.weight"] = ((torch.rand(width, 3, patch, patch, generator=g) * 2 - 1) / math.sqrt(fan_in))
    sd["visual.class_embedding"] = scale * randn(width)
    sd["visual.positional_embedding"] = scale * randn((res // patch) ** 2 + 1, width)
    for n in ("ln_pre", "ln_post"):
        sd[f"visual.{n}.weight"] = 1.0 + 0.1 * randn(width)
        sd[f"visual.{n}.bias"] = 0.1 * randn(width)
    for i in range(layers):
        p = f"visual.transformer.resblocks.{i}."
        # nn.MultiheadAttention: xavier_uniform in_proj, zero biases; perturb biases/LN so tests see them
        bound = math.sqrt(6.0 / (width + 3 * width))
        sd[p + "attn.in_proj_weight"] = (torch.rand(3 * width, width, generator=g) * 2 - 1) * bound
        sd[p + "attn.in_proj_bias"] = 0.02 * randn(3 * width)
        sd[p + "attn.out_proj.weight"] = kaiming_uniform(width, width)
        sd[p + "attn.out_proj.bias"] = 0.02 * randn(width)
        sd[p + "ln_1.weight"] = 1.0 + 0.1 * randn(width)
        sd[p + "ln_1.bias"] = 0.1 * randn(width)
        sd[p + "mlp.c_fc.weight"] = kaiming_uniform(4 * width, width)
        sd[p + "mlp.c_fc.bias"] = bias_uniform(4 * width, width)
        sd[p + "mlp.c_proj.weight"] = kaiming_uniform(width, 4 * width)
        sd[p + "mlp.c_proj.bias"] = bias_uniform(width, 4 * width)
        sd[p + "ln_2.weight"] = 1.0 + 0.1 * randn(width)
        sd[p + "ln_2.bias"] = 0.1 * randn(width)
    sd["visual.proj"] = scale * randn(width, embed)
    return sd


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d, h):
        super().__init__()
        self.attn = nn.MultiheadAttention(d, h)
        self.ln_1 = nn.LayerNorm(d)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(d, 4 * d))
        self.mlp.add_module("gelu", QuickGELU())
        self.mlp.add_module("c_proj", nn.Linear(4 * d, d))
        self.ln_2 = nn.LayerNorm(d)

    def forward(self, x):  # x: [T, N, D]
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False)[0]
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, patch_size, patch_size, bias=False)
        self.class_embedding = nn.Parameter(torch.zeros(width))
        self.positional_embedding = nn.Parameter(torch.zeros((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(torch.zeros(width, output_dim))

    def forward(self, x):
        x = self.conv1(x)
        x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
        cls = self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype)
        x = torch.cat([cls, x], dim=1)
        x = x + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_post(x[:, 0, :])
        return x @ self.proj


class OracleCLIP(nn.Module):
    """Stands in for the object ``clip.load`` returns: ``.visual.input_resolution``, ``.encode_image``."""

    def __init__(self, name: str, state_dict=None, seed: int = 1):
        super().__init__()
        res, patch, width, layers, heads, embed = CONFIGS[name]
        self.visual = VisionTransformer(res, patch, width, layers, heads, embed)
        sd = state_dict if state_dict is not None else random_state_dict(name, seed)
        self.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
        self.eval().requires_grad_(False)  # models.py:67-71

    def encode_image(self, image):
        return self.visual(image)
