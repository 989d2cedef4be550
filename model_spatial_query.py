"""Drop-in replacement for the reference's top-level `model_spatial_query` module.

Put this repository ahead of the TransEditor checkout on sys.path and
`from model_spatial_query import Generator, Discriminator` (train_spatial_query.py:27,
test_spatial_query.py, pSp/models/psp_new.py) resolves to the B200 implementation; class
names, constructor / forward signatures and state_dict keys are the reference's.
"""
from transeditor_b200.model import (  # noqa: F401
    Attention, AttentionBlock, Blur, ConstantInput, ConvLayer, Discriminator, Downsample,
    EqualConv2d, EqualLinear, Generator, ModulatedConv2d, NoiseInjection, PixelNorm, ResBlock,
    ScaledLeakyReLU, StyledConv, ToRGB, Upsample, make_kernel,
)
from transeditor_b200.op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d  # noqa: F401
