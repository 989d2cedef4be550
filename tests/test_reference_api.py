"""Drop-in fidelity against the reference tree itself: signatures, state_dict layout, same-seed
initialisation.  Runs only where /root/reference exists (the build container)."""
import inspect

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference_module()


def test_signatures_match(ref):
    import model_spatial_query as M
    for cls in ("Generator", "Discriminator", "EqualLinear", "EqualConv2d", "ModulatedConv2d",
                "StyledConv", "ToRGB", "ConvLayer", "ResBlock", "Attention", "AttentionBlock",
                "Blur", "Upsample", "PixelNorm", "NoiseInjection"):
        a, b = getattr(ref, cls), getattr(M, cls)
        assert str(inspect.signature(a.__init__)) == str(inspect.signature(b.__init__)), cls
    ra, rb = inspect.signature(ref.Generator.forward), inspect.signature(M.Generator.forward)
    assert list(ra.parameters) == list(rb.parameters)
    assert [p.default for p in ra.parameters.values()] == [p.default for p in rb.parameters.values()]
    assert list(inspect.signature(ref.Discriminator.forward).parameters) == \
        list(inspect.signature(M.Discriminator.forward).parameters)
    # ModulatedConv2d.forward keeps (input, style) as the leading positional arguments
    assert list(inspect.signature(M.ModulatedConv2d.forward).parameters)[:3] == ["self", "input", "style"]
    import utils.op as uop
    assert str(inspect.signature(uop.fused_leaky_relu)) == "(input, bias, negative_slope=0.2, scale=1.4142135623730951)"
    assert str(inspect.signature(uop.upfirdn2d)) == "(input, kernel, up=1, down=1, pad=(0, 0))"


@pytest.mark.parametrize("size,cm", [(64, 1), (256, 2), (1024, 2)])
def test_state_dict_layout_and_seeded_init(ref, size, cm):
    import model_spatial_query as M
    t = 2 * (size.bit_length() - 1) - 2
    torch.manual_seed(7)
    with ref_shim.cpu_mode():
        gr = ref.Generator(size, 512, 512, t, channel_multiplier=cm, n_trans=8, pixel_norm_op_dim=1)
        dr = ref.Discriminator(size, channel_multiplier=cm)
    torch.manual_seed(7)
    g = M.Generator(size, 512, 512, t, channel_multiplier=cm, n_trans=8, pixel_norm_op_dim=1)
    d = M.Discriminator(size, channel_multiplier=cm)
    for a, b in ((gr, g), (dr, d)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        for k in sa:
            assert sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]), k
        assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]
    b.load_state_dict(a.state_dict(), strict=True)
