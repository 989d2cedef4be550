"""Data formats either side of the path: te_image_prep / te_image_quantize against the golden
vectors written from torchvision's own transforms / save_image (oracle/make_golden_data.py) and the numpy oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import data_cpu
from tests.conftest import ROOT

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "data.npz"))


def test_oracle_reproduces_torchvision_goldens():
    for t in "abc":
        got = data_cpu.image_prep(GOLD["prep_%s_u8" % t], GOLD["prep_%s_flip" % t])
        assert np.array_equal(got, GOLD["prep_%s_out" % t])
    for t in "ab":
        assert np.array_equal(data_cpu.image_quantize(GOLD["quant_%s_x" % t]), GOLD["quant_%s_out" % t])


def test_host_logic_on_cpu_emulation(cpu_emulation):
    from transeditor_b200 import data
    u8 = torch.from_numpy(GOLD["prep_b_u8"])
    fl = torch.from_numpy(GOLD["prep_b_flip"])
    pad = torch.empty(u8.shape[:3] + (8,), dtype=torch.bfloat16)
    out = data.image_prep(u8, fl, nhwc8=pad)
    assert np.array_equal(out.numpy(), GOLD["prep_b_out"])
    assert torch.equal(pad[..., :3].float(), out.permute(0, 2, 3, 1).to(torch.bfloat16).float())
    assert pad[..., 3:].abs().sum() == 0
    q = data.quantize(torch.from_numpy(GOLD["quant_b_x"]))
    assert np.array_equal(q.numpy(), GOLD["quant_b_out"])
    with pytest.raises(TypeError):
        data.image_prep(u8.float())
    with pytest.raises(TypeError):
        data.image_prep(u8, fl[:1])
    with pytest.raises(TypeError):
        data.quantize(torch.zeros(1, 4, 2, 2))
    torch.manual_seed(0)
    want = [bool(torch.rand(1) < 0.5) for _ in range(6)]
    torch.manual_seed(0)
    assert data.draw_flips(6).bool().tolist() == want


def test_decode_uint8_round_trips_a_png():
    import io
    from PIL import Image
    from transeditor_b200 import data
    img = GOLD["prep_a_u8"][0]
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="png")
    assert np.array_equal(data.decode_uint8(buf.getvalue(), 16).numpy(), img)
    with pytest.raises(ValueError):
        data.decode_uint8(buf.getvalue(), 32)


def test_no_cpu_fallback():
    from transeditor_b200 import data
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        data.image_prep(torch.zeros(1, 4, 4, 3, dtype=torch.uint8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        data.quantize(torch.zeros(1, 3, 4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_gpu_image_prep_bit_exact(tag):
    from transeditor_b200 import data
    u8 = torch.from_numpy(GOLD["prep_%s_u8" % tag]).cuda()
    fl = torch.from_numpy(GOLD["prep_%s_flip" % tag]).cuda()
    for dt in (torch.bfloat16, torch.float32):
        pad = torch.full(tuple(u8.shape[:3]) + (8,), 7.0, dtype=dt, device="cuda")
        out = data.image_prep(u8, fl, nhwc8=pad)
        assert np.array_equal(out.cpu().numpy(), GOLD["prep_%s_out" % tag])
        want = torch.from_numpy(data_cpu.image_prep_nhwc8(GOLD["prep_%s_u8" % tag], GOLD["prep_%s_flip" % tag]))
        assert torch.equal(pad.cpu().float(), want.to(dt).float())
    noflip = data.image_prep(u8)
    assert np.array_equal(noflip.cpu().numpy(), data_cpu.image_prep(GOLD["prep_%s_u8" % tag]))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_gpu_image_quantize_bit_exact(tag):
    from transeditor_b200 import data
    x = torch.from_numpy(GOLD["quant_%s_x" % tag]).cuda()
    assert np.array_equal(data.quantize(x).cpu().numpy(), GOLD["quant_%s_out" % tag])
    cl = x.contiguous(memory_format=torch.channels_last)
    assert np.array_equal(data.quantize(cl).cpu().numpy(), GOLD["quant_%s_out" % tag])
    xb = x.to(torch.bfloat16)
    assert np.array_equal(data.quantize(xb).cpu().numpy(), data_cpu.image_quantize(xb.float().cpu().numpy()))


@pytest.mark.gpu
def test_gpu_full_size_round_trip_and_edges():
    """256^2 batch 16 (BASELINE configs[1]): every 8-bit value survives prep -> quantize; a flip applied twice is the
    identity; empty batch is a no-op."""
    from transeditor_b200 import data
    g = torch.Generator().manual_seed(1)
    u8 = torch.randint(0, 256, (16, 256, 256, 3), generator=g, dtype=torch.uint8).cuda()
    fl = torch.randint(0, 2, (16,), generator=g, dtype=torch.uint8).cuda()
    x = data.image_prep(u8, fl)
    assert x.min() >= -1 and x.max() <= 1
    back = data.quantize(x)
    want = torch.where(fl.bool()[:, None, None, None], u8.flip(2), u8)
    assert torch.equal(back, want)
    assert np.array_equal(x.cpu().numpy(), data_cpu.image_prep(u8.cpu().numpy(), fl.cpu().numpy()))
    assert data.image_prep(u8[:0]).shape == (0, 3, 256, 256)
    pipe = data.DeviceImagePipeline(16, 256, "cuda")
    got = pipe.load(u8.cpu().pin_memory(), fl.cpu().pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(got, x)


def test_uint8_dataset_reads_the_reference_lmdb_layout(monkeypatch):
    """Uint8Dataset mirrors utils/dataset.py:9-45: keys `{resolution}-{index:05d}`, `length`, decode with PIL, and a
    corrupt record falls through to another index — checked against an in-memory stand-in for the lmdb module."""
    import io
    import sys
    import types
    from PIL import Image
    imgs = [GOLD["prep_a_u8"][i] for i in range(3)]
    store = {b"length": b"3"}
    for i, im in enumerate(imgs):
        buf = io.BytesIO()
        Image.fromarray(im).save(buf, format="png")
        store[("16-%05d" % i).encode()] = buf.getvalue()
    store[b"16-00001"] = b"not an image"

    class Txn:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def get(self, key):
            return store.get(key)

    class Env:
        def begin(self, write=False):
            return Txn()

    fake = types.ModuleType("lmdb")
    fake.open = lambda path, **kw: Env()
    monkeypatch.setitem(sys.modules, "lmdb", fake)
    from transeditor_b200 import data
    ds = data.Uint8Dataset("ignored", resolution=16)
    assert len(ds) == 3
    assert np.array_equal(ds[0].numpy(), imgs[0]) and np.array_equal(ds[2].numpy(), imgs[2])
    import random
    random.seed(0)
    got = ds[1].numpy()          # corrupt record: retried at a random index, like the reference
    assert any(np.array_equal(got, im) for im in (imgs[0], imgs[2]))
