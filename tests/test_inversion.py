"""pSp encoder mirror + fused inference form (transeditor_b200/inversion.py) against the reference's own
GradualStyleEncoder run on CPU (tests/golden/psp_encoder.npz, written by oracle/make_golden_psp.py)."""
import numpy as np
import pytest
import torch

from oracle import psp_state
from tests.conftest import load_golden


@pytest.fixture(scope="module")
def encoder():
    from transeditor_b200.inversion import GradualStyleEncoder
    return psp_state.randomize(psp_state.build(GradualStyleEncoder))


def test_same_seed_state_matches_the_reference(encoder):
    gold = load_golden("psp_encoder")
    assert sorted(encoder.state_dict().keys()) == list(gold["keys"])
    assert abs(psp_state.checksum(encoder) - float(gold["checksum"])) < 1e-6 * float(gold["checksum"])


def test_module_forward_matches_reference_on_cpu(encoder):
    gold = load_golden("psp_encoder")
    with torch.no_grad():
        z, p = encoder(psp_state.image())
    assert z.shape == (1, 512, 16) and p.shape == (1, 512, 16)
    assert np.abs(z.numpy() - gold["z"]).max() < 1e-4 * max(1.0, np.abs(gold["z"]).max())
    assert np.abs(p.numpy() - gold["p"]).max() < 1e-4 * max(1.0, np.abs(gold["p"]).max())


def test_fused_form_fp32_matches_reference_on_cpu(encoder):
    """Folded batch norms, strided-sampling shortcut, batched heads (grouped convolutions) in f32: same numbers."""
    from transeditor_b200.inversion import FusedEncoder
    gold = load_golden("psp_encoder")
    z, p = FusedEncoder(encoder, dtype=torch.float32)(psp_state.image())
    assert np.abs(z.numpy() - gold["z"]).max() < 2e-3 * max(1.0, np.abs(gold["z"]).max())
    assert np.abs(p.numpy() - gold["p"]).max() < 2e-3 * max(1.0, np.abs(gold["p"]).max())


def test_get_blocks_layout():
    from transeditor_b200.inversion import get_blocks
    b = get_blocks(50)
    assert [len(s) for s in b] == [3, 4, 14, 3]
    assert b[1][0] == (64, 128, 2) and b[1][1] == (128, 128, 1) and b[3][0] == (256, 512, 2)
    with pytest.raises(ValueError):
        get_blocks(18)


@pytest.mark.gpu
def test_fused_bf16_on_gpu_close_to_reference(encoder):
    from transeditor_b200.inversion import FusedEncoder
    gold = load_golden("psp_encoder")
    enc = encoder.to("cuda")
    try:
        z, p = FusedEncoder(enc, dtype=torch.bfloat16)(psp_state.image().cuda())
        z32, p32 = FusedEncoder(enc, dtype=torch.float32)(psp_state.image().cuda())
    finally:
        encoder.to("cpu")
    for got, name in ((z32, "z"), (p32, "p")):
        assert np.abs(got.cpu().numpy() - gold[name]).max() < 5e-3 * max(1.0, np.abs(gold[name]).max())
    for got, name in ((z, "z"), (p, "p")):      # bf16 through 50 layers: compare on the scale of the codes
        err = np.abs(got.cpu().numpy() - gold[name])
        assert err.mean() < 0.03 * np.abs(gold[name]).std() + 1e-3 and err.max() < 0.25 * np.abs(gold[name]).std() + 1e-2


@pytest.mark.gpu
def test_inversion_pipeline_graph_equals_eager(encoder):
    import model_spatial_query as M
    from transeditor_b200 import model as te_model
    from transeditor_b200.inversion import InversionPipeline
    te_model.set_precision("bf16")
    try:
        torch.manual_seed(0)
        g = M.Generator(64, 512, 512, 10, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).to("cuda").eval()
        enc = encoder.to("cuda")
        z_avg = torch.randn(1, 512, 16, device="cuda") * 0.1
        x = psp_state.image(batch=2).cuda()
        eager = InversionPipeline(enc, g, z_avg=z_avg, p_avg=None, resize=False, graph=False)
        img0, z0, p0 = [t.clone() for t in eager(x)]
        graphed = InversionPipeline(enc, g, z_avg=z_avg, p_avg=None, resize=False, graph=True)
        graphed(x)
        img1, z1, p1 = graphed(x)
        assert img1.shape == (2, 3, 64, 64) and z1.shape == (2, 512, 16)
        assert (z1 - z0).abs().max().item() < 1e-5 and (p1 - p0).abs().max().item() < 1e-5
        assert (img1.float() - img0.float()).abs().max().item() < 1e-3
        pooled = InversionPipeline(enc, g, resize=True, graph=False)(x)[0]
        assert pooled.shape == (2, 3, 256, 256)
    finally:
        te_model.set_precision("fp32")
        encoder.to("cpu")
