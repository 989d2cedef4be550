"""Editing / interpolation latency path: linear_interpolate against the numpy restatement
of the reference helper; CUDA-graph replay of Generator.forward against the eager call."""
import numpy as np
import pytest
import torch

from oracle import ops_cpu


@pytest.mark.parametrize("shape", [(1, 512), (1, 16, 512), (1, 14, 512)])
def test_linear_interpolate_matches_reference_semantics(shape):
    from transeditor_b200.inference import linear_interpolate
    rng = np.random.default_rng(3)
    code = rng.standard_normal(shape).astype(np.float32)
    normal = rng.standard_normal((1, 512)).astype(np.float32)
    normal /= np.linalg.norm(normal)
    ref = ops_cpu.linear_interpolate_np(code, normal, -3.0, 3.0, 7)
    out_np = linear_interpolate(code, normal, start_distance=-3.0, end_distance=3.0, steps=7)
    assert isinstance(out_np, np.ndarray) and out_np.shape == ref.shape
    assert np.abs(out_np - ref).max() < 1e-5
    out_t = linear_interpolate(torch.from_numpy(code), torch.from_numpy(normal), -3.0, 3.0, 7)
    assert isinstance(out_t, torch.Tensor) and np.abs(out_t.numpy() - ref).max() < 1e-5
    if len(shape) == 2:  # absolute distances to the hyperplane
        assert np.allclose(out_np @ normal.T, np.linspace(-3, 3, 7).reshape(-1, 1), atol=1e-4)
    with pytest.raises(AssertionError):
        linear_interpolate(np.zeros((2, 512), np.float32), normal)


def _generator(size=64):
    import model_spatial_query as M
    torch.manual_seed(0)
    t = 2 * int(np.log2(size)) - 2
    return M.Generator(size, 512, 512, t, channel_multiplier=2, n_trans=8, pixel_norm_op_dim=1).cuda().eval()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_generator_replays_the_eager_forward(precision):
    from transeditor_b200 import model as te_model
    from transeditor_b200.inference import GraphedGenerator
    te_model.set_precision(precision)
    try:
        g = _generator()
        gg = GraphedGenerator(g, batch=2)
        for seed in (1, 2):
            gen = torch.Generator(device="cuda").manual_seed(seed)
            z = torch.randn(2, 512, 16, device="cuda", generator=gen)
            p = torch.randn(2, 512, 16, device="cuda", generator=gen)
            with torch.no_grad():
                ref = g(z, p)[0]
            out = gg(z, p)[0]
            assert torch.equal(out, ref)
        with pytest.raises(RuntimeError):
            gg(z[:1], p[:1])
    finally:
        te_model.set_precision("fp32")


@pytest.mark.gpu
def test_edit_frames_on_device():
    from transeditor_b200 import model as te_model
    from transeditor_b200.inference import GraphedGenerator, edit_frames, linear_interpolate, to_uint8
    te_model.set_precision("bf16")
    try:
        g = _generator()
        gen = torch.Generator(device="cuda").manual_seed(5)
        z = torch.randn(1, 512, 16, device="cuda", generator=gen)
        p = torch.randn(1, 512, 16, device="cuda", generator=gen)
        with torch.no_grad():
            z_plus, p_plus = g(z, p, return_mapped_codes=True)        # [1, 512, 16]
        z_tok, p_tok = z_plus.transpose(1, 2).contiguous(), p_plus.transpose(1, 2).contiguous()
        bz = torch.nn.functional.normalize(torch.randn(1, 512, device="cuda", generator=gen), dim=1)
        gg = GraphedGenerator(g, batch=1, use_style_mapping=False, use_spatial_mapping=False)
        frames = edit_frames(gg, z_tok, p_tok, z_boundary=bz, z_distance=2.0, steps=5)
        assert frames.shape == (5, 64, 64, 3) and frames.dtype == torch.uint8 and frames.is_cuda
        zs = linear_interpolate(z_tok, bz, -2.0, 2.0, 5)
        with torch.no_grad():
            ref = g(zs[3:4].transpose(1, 2), p_plus, use_style_mapping=False, use_spatial_mapping=False)[0]
        assert torch.equal(frames[3:4], to_uint8(ref))
        assert not torch.equal(frames[0], frames[4])
    finally:
        te_model.set_precision("fp32")
