"""CPU restatement of the data formats either side of the path.  TEST INFRASTRUCTURE ONLY (same rule as
oracle/ops_cpu.py: only tests/, smoke() and bench.py's CPU legs import it).

* image_prep      the loader transform of train_spatial_query.py:511-517 applied at utils/dataset.py:38-41:
                  RandomHorizontalFlip (the coin is an argument here), ToTensor (uint8 HWC -> f32 CHW / 255),
                  Normalize(0.5, 0.5): (t - 0.5) / 0.5, every step rounded to float32 like the torch ops.
* image_quantize  torchvision.utils.save_image(normalize=True, range=(low, high)) as the reference calls it
                  (test_spatial_query.py:82-88, train_spatial_query.py:345-351): clamp, (x - low) / max(high - low,
                  1e-5), * 255, + 0.5, clamp(0, 255), truncate to uint8, CHW -> HWC.

Pinned by tests/golden/data.npz, written by oracle/make_golden_data.py from torchvision's own transforms / save_image.
"""
import numpy as np


def image_prep(u8_hwc, flip=None):
    """u8_hwc uint8 [B, H, W, 3], flip bool/uint8 [B] or None -> float32 [B, 3, H, W]."""
    x = np.asarray(u8_hwc)
    if flip is not None:
        x = np.stack([img[:, ::-1] if f else img for img, f in zip(x, flip)])
    t = x.astype(np.float32) / np.float32(255.0)
    t = (t - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(t.transpose(0, 3, 1, 2))


def image_prep_nhwc8(u8_hwc, flip=None):
    """The same values channels-last with the channels zero-padded to 8: float32 [B, H, W, 8]."""
    t = image_prep(u8_hwc, flip).transpose(0, 2, 3, 1)
    out = np.zeros(t.shape[:3] + (8,), np.float32)
    out[..., :3] = t
    return out


def image_quantize(x_nchw, low=-1.0, high=1.0):
    """float32 [B, 3, H, W] -> uint8 [B, H, W, 3]."""
    x = np.asarray(x_nchw, np.float32)
    low, high = np.float32(low), np.float32(high)
    t = np.clip(x, low, high)
    t = (t - low) / np.maximum(high - low, np.float32(1e-5))
    t = t * np.float32(255.0) + np.float32(0.5)   # two rounded float32 steps (numpy does not fuse)
    t = np.clip(t, np.float32(0.0), np.float32(255.0))
    return np.ascontiguousarray(t.astype(np.uint8).transpose(0, 2, 3, 1))
