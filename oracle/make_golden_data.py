"""Writes tests/golden/data.npz from torchvision's OWN transforms / save_image (what the reference calls):

    python oracle/make_golden_data.py

* prep_*:  PIL images -> transforms.Compose([RandomHorizontalFlip(p), ToTensor(), Normalize(0.5, 0.5, inplace=True)])
           with p forced to 0 or 1 per image (the coin itself is host logic), train_spatial_query.py:511-517.
* quant_*: utils.save_image(x, buf, nrow=1, padding=0, normalize=True, value_range=(-1, 1)) decoded back from PNG
           (`range=` in the reference's torchvision 0.8, test_spatial_query.py:82-88).
Runs in the build container only (torchvision + PIL); the .npz travels.
"""
import io
import os

import numpy as np
import torch
from PIL import Image
from torchvision import transforms, utils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rng = np.random.default_rng(7)
    out = {}
    for tag, (b, h, w) in {"a": (3, 16, 16), "b": (2, 9, 13), "c": (2, 32, 20)}.items():
        u8 = rng.integers(0, 256, size=(b, h, w, 3), dtype=np.uint8)
        u8[0, 0, :6, 0] = [0, 1, 127, 128, 254, 255]
        flip = rng.integers(0, 2, size=b).astype(np.uint8)
        flip[0] = 1
        got = []
        for img, f in zip(u8, flip):
            tf = transforms.Compose([transforms.RandomHorizontalFlip(p=float(f)), transforms.ToTensor(),
                                     transforms.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5), inplace=True)])
            got.append(tf(Image.fromarray(img)))
        out["prep_%s_u8" % tag] = u8
        out["prep_%s_flip" % tag] = flip
        out["prep_%s_out" % tag] = torch.stack(got).numpy()
    g = torch.Generator().manual_seed(3)
    for tag, (b, h, w) in {"a": (2, 16, 16), "b": (3, 7, 10)}.items():
        x = torch.randn(b, 3, h, w, generator=g) * 0.8
        x[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, -3.0, 3.0])
        # values sitting on rounding boundaries of the 8-bit grid
        k = torch.arange(0, min(w, 8), dtype=torch.float32)
        x[0, 1, 1, :k.numel()] = (k * 16 + 0.5) / 255.0 * 2 - 1
        buf = io.BytesIO()
        utils.save_image(x.clone(), buf, nrow=1, padding=0, normalize=True, value_range=(-1, 1), format="png")
        img = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
        out["quant_%s_x" % tag] = x.numpy()
        out["quant_%s_out" % tag] = img.reshape(b, h, w, 3)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "data.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
