#!/usr/bin/env python
"""Writes tests/golden/jpeg.npz: a few small JPEG files (encoded with Pillow / libjpeg-turbo) together with the pixels
libjpeg decodes them to.  tests/test_jpeg.py checks pathtracer_rs_b200/host/jpeg_decode.cpp against them without needing
Pillow at test time.     usage: python tests/golden/make_jpeg_golden.py"""
import io
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))


def picture(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(x / 3.0) * np.cos(y / 2.5), 128 + 90 * np.sin((x + y) / 4.0), (x * 13 + y * 7) % 256], -1)
    return np.clip(base + rng.normal(0, 20, (h, w, 3)), 0, 255).astype(np.uint8)


CASES = {
    "baseline_420": (picture(19, 13, 1), dict(quality=85, subsampling=2)),
    "baseline_422_restart": (picture(21, 9, 2), dict(quality=70, subsampling=1, restart_marker_blocks=1)),
    "progressive_444": (picture(24, 10, 3), dict(quality=92, subsampling=0, progressive=True)),
    "progressive_420": (picture(33, 17, 4), dict(quality=60, subsampling=2, progressive=True)),
    "grey": (picture(9, 11, 5)[..., 0], dict(quality=80)),
    "narrow_420": (picture(3, 7, 6), dict(quality=90, subsampling=2)),
}

out = {}
for name, (px, kw) in CASES.items():
    b = io.BytesIO()
    Image.fromarray(px).save(b, "JPEG", **kw)
    data = b.getvalue()
    ref = np.asarray(Image.open(io.BytesIO(data)))
    out[name + "_file"] = np.frombuffer(data, dtype=np.uint8)
    out[name + "_pixels"] = ref if ref.ndim == 3 else ref[..., None]
np.savez_compressed(os.path.join(HERE, "jpeg.npz"), **out)
print({k: v.shape for k, v in out.items()})
