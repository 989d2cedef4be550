"""Data formats either side of the path, on the device.

Input side — the reference's loader (`MultiResolutionDataset.__getitem__`, utils/dataset.py:32-45) decodes an
image with PIL and runs the transform of train_spatial_query.py:511-517 (RandomHorizontalFlip, ToTensor,
Normalize(0.5, 0.5)) on the HOST, then copies float32 NCHW to the GPU (`real_img.to(device)`, :180).  Here the
host keeps only what is inherently serial (entropy decode with PIL, the coin flips); the batch crosses PCIe as
uint8 HWC pixels — a quarter of the bytes — and ONE kernel (`te_image_prep`) mirrors, converts, normalises and
lays the batch out, bit-identical to the torch ops.

Output side — `utils.save_image(sample, normalize=True, range=(-1, 1))` (test_spatial_query.py:82-88,
train_spatial_query.py:345-351) quantises on the host after a float32 D2H copy; `quantize` does the same
arithmetic on the device (`te_image_quantize`) so uint8 pixels cross instead.
"""
import io

import torch

from . import lib


def decode_uint8(img_bytes, size=None):
    """Encoded image bytes (what the LMDB record holds, utils/dataset.py:36-39) -> uint8 [H, W, 3] tensor."""
    import numpy as np
    from PIL import Image
    img = Image.open(io.BytesIO(img_bytes)).convert("RGB")
    if size is not None and img.size != (size, size):
        raise ValueError("decode_uint8: record is %s, expected %dx%d" % (img.size, size, size))
    return torch.from_numpy(np.array(img, dtype=np.uint8))


def draw_flips(n, p=0.5):
    """The coins of RandomHorizontalFlip (torchvision: `torch.rand(1) < p`, one draw per image, in order)."""
    return torch.tensor([bool(torch.rand(1) < p) for _ in range(n)], dtype=torch.uint8)


def image_prep(u8_hwc, flip=None, out=None, nhwc8=None):
    """uint8 [B, H, W, 3] CUDA tensor (+ uint8 [B] flip flags) -> float32 [B, 3, H, W] in [-1, 1].
    `nhwc8`: optional [B, H, W, 8] float32 / bfloat16 buffer that receives the same values channels-last with the
    channels zero-padded to 8 (the tensor-core from-RGB operand).  Returns `out`."""
    lib.require_cuda(u8_hwc, flip, out, nhwc8)
    if u8_hwc.dtype != torch.uint8 or u8_hwc.dim() != 4 or u8_hwc.shape[3] != 3 or not u8_hwc.is_contiguous():
        raise TypeError("image_prep: expected a contiguous uint8 [B, H, W, 3] tensor, got %s %s"
                        % (u8_hwc.dtype, tuple(u8_hwc.shape)))
    b, h, w, _ = u8_hwc.shape
    if flip is not None and (flip.dtype != torch.uint8 or flip.shape != (b,) or not flip.is_contiguous()):
        raise TypeError("image_prep: flip must be a contiguous uint8 [B] tensor")
    if out is None:
        out = torch.empty((b, 3, h, w), dtype=torch.float32, device=u8_hwc.device)
    elif out.shape != (b, 3, h, w) or out.dtype != torch.float32 or not out.is_contiguous():
        raise TypeError("image_prep: out must be a contiguous float32 [B, 3, H, W] tensor")
    if nhwc8 is not None and (nhwc8.shape != (b, h, w, 8) or not nhwc8.is_contiguous()
                              or nhwc8.dtype not in (torch.float32, torch.bfloat16)):
        raise TypeError("image_prep: nhwc8 must be a contiguous float32 / bfloat16 [B, H, W, 8] tensor")
    lib.image_prep(out, nhwc8, u8_hwc, flip, b, h, w)
    return out


def quantize(x, low=-1.0, high=1.0, out=None):
    """[B, 3, H, W] float32 / bfloat16 CUDA tensor (any strides) -> uint8 [B, H, W, 3] with save_image's arithmetic."""
    lib.require_cuda(x, out)
    if x.dim() != 4 or x.shape[1] != 3 or x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("quantize: expected a float32 / bfloat16 [B, 3, H, W] tensor, got %s %s"
                        % (x.dtype, tuple(x.shape)))
    b, _, h, w = x.shape
    if out is None:
        out = torch.empty((b, h, w, 3), dtype=torch.uint8, device=x.device)
    elif out.shape != (b, h, w, 3) or out.dtype != torch.uint8 or not out.is_contiguous():
        raise TypeError("quantize: out must be a contiguous uint8 [B, H, W, 3] tensor")
    lib.image_quantize(out, x, float(low), float(high))
    return out


class DeviceImagePipeline:
    """Static staging for one training process: pinned host uint8 batch -> device uint8 -> te_image_prep.
    `load(u8, flip)` returns the float32 NCHW batch in a persistent buffer (CUDA-graph friendly: the address never
    changes); the copies are asynchronous on the current stream."""

    def __init__(self, batch, size, device):
        self.device = torch.device(device)
        self.u8 = torch.empty((batch, size, size, 3), dtype=torch.uint8, device=self.device)
        self.flip = torch.zeros((batch,), dtype=torch.uint8, device=self.device)
        self.out = torch.empty((batch, 3, size, size), dtype=torch.float32, device=self.device)
        self.h2d_bytes = self.u8.numel() + self.flip.numel()

    def load(self, u8_host, flip_host=None, out=None):
        self.u8.copy_(u8_host, non_blocking=True)
        if flip_host is not None:
            self.flip.copy_(flip_host, non_blocking=True)
        return image_prep(self.u8, self.flip if flip_host is not None else None, self.out if out is None else out)


class Uint8Dataset(torch.utils.data.Dataset):
    """`MultiResolutionDataset` (utils/dataset.py:9-45) without the host transform: same LMDB layout and keys
    (`{resolution}-{index:05d}`, `length`), same retry-on-corrupt-record behaviour, but __getitem__ returns the
    decoded uint8 [H, W, 3] pixels; flip / ToTensor / Normalize run on the device (image_prep)."""

    def __init__(self, path, resolution=256):
        try:
            import lmdb
        except ImportError as e:  # the data path is a caller of the hot path; fail loudly, never fake data
            raise RuntimeError("Uint8Dataset needs the `lmdb` package the reference's dataset uses") from e
        self.env = lmdb.open(path, max_readers=32, readonly=True, lock=False, readahead=False, meminit=False)
        if not self.env:
            raise IOError("Cannot open lmdb dataset", path)
        with self.env.begin(write=False) as txn:
            self.length = int(txn.get("length".encode("utf-8")).decode("utf-8"))
        self.resolution = resolution

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        import random
        with self.env.begin(write=False) as txn:
            img_bytes = txn.get(f"{self.resolution}-{str(index).zfill(5)}".encode("utf-8"))
        try:
            return decode_uint8(img_bytes, self.resolution)
        except Exception as e:  # utils/dataset.py:43-45
            print(e)
            return self.__getitem__(random.randint(0, self.length - 1))
