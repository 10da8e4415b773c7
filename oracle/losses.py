"""Oracle restatement of clip_diffusion/losses.py:10-45 (TEST-ONLY, CPU torch, autograd gives
the gradients the CUDA kernels' analytic gradients are checked against)."""
import torch
from torch.nn import functional as F


def L2_norm(input, dim=-1):
    """utils/functional.py:74-76"""
    return F.normalize(input, dim=dim)


def square_spherical_distance_loss(x, y):
    """losses.py:10-16"""
    x = L2_norm(x, dim=-1)
    y = L2_norm(y, dim=-1)
    return (x - y).norm(dim=-1).div(2).arcsin().pow(2).mul(2)


def total_variational_loss(input):
    """losses.py:20-28"""
    input = F.pad(input, (0, 1, 0, 1), "replicate")
    x_diff = input[..., :-1, 1:] - input[..., :-1, :-1]
    y_diff = input[..., 1:, :-1] - input[..., :-1, :-1]
    return (x_diff.pow(2) + y_diff.pow(2)).mean([1, 2, 3])


def rgb_range_loss(input):
    """losses.py:31-35"""
    return (input - input.clamp(min=-1, max=1)).pow(2).mean([1, 2, 3])


def aesthetic_loss(predictor, input):
    """losses.py:43-45"""
    return predictor(L2_norm(input, dim=-1)).mean()


class LinearAestheticPredictor(torch.nn.Module):
    """models.py:188-196"""

    def __init__(self, input_dim):
        super().__init__()
        self.linear = torch.nn.Linear(input_dim, 1)

    def forward(self, input):
        return self.linear(input)


class MLPAestheticPredictor(torch.nn.Module):
    """models.py:200-217 (dropouts are identity in eval; there are no activations)"""

    def __init__(self, input_dim):
        super().__init__()
        self.layers = torch.nn.Sequential(
            torch.nn.Linear(input_dim, 1024), torch.nn.Dropout(0.2), torch.nn.Linear(1024, 128), torch.nn.Dropout(0.2),
            torch.nn.Linear(128, 64), torch.nn.Dropout(0.1), torch.nn.Linear(64, 16), torch.nn.Linear(16, 1),
        )

    def forward(self, input):
        return self.layers(input)
