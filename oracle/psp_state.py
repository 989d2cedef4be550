"""Deterministic non-trivial state for a pSp GradualStyleEncoder (reference class or this repository's mirror).
TEST INFRASTRUCTURE ONLY.

The encoder has ~350 M parameters, far too many to store as a fixture: both sides are constructed under the same
`torch.manual_seed` (identical construction order => identical initial weights), then the tensors a default
initialisation leaves trivial — batch-norm statistics and affine terms, PReLU slopes, biases — are overwritten
here from a seeded generator, walking `named_modules()` (same order on both sides)."""
import torch
from torch import nn


def build(cls, seed=0, **kwargs):
    torch.manual_seed(seed)
    return cls(50, "ir_se", **kwargs).eval()


def randomize(module, seed=1):
    g = torch.Generator().manual_seed(seed)
    rnd = lambda t, lo, hi: t.copy_(torch.rand(t.shape, generator=g) * (hi - lo) + lo)  # noqa: E731
    with torch.no_grad():
        for _, m in module.named_modules():
            if isinstance(m, nn.BatchNorm2d):
                rnd(m.running_mean, -0.2, 0.2)
                rnd(m.running_var, 0.6, 1.6)
                rnd(m.weight, 0.7, 1.3)
                rnd(m.bias, -0.2, 0.2)
            elif isinstance(m, nn.PReLU):
                rnd(m.weight, 0.1, 0.4)
            elif isinstance(m, nn.Conv2d) and m.bias is not None:
                rnd(m.bias, -0.1, 0.1)
        for name, p in module.named_parameters():
            if name.endswith("linear.bias") or name == "adjust_style.bias":
                rnd(p, -0.5, 0.5)
    return module


def image(seed=2, batch=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, 256, 256, generator=g) * 2 - 1


def checksum(module):
    return float(sum(v.double().abs().sum() for v in module.state_dict().values()))
