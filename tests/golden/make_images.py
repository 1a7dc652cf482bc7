#!/usr/bin/env python3
"""Writes the small image files the golden scene image_maps.pov maps onto its objects (deterministic; committed with their outputs):
img_ramp.ppm (24 x 16 RGB) and img_disc.png (16 x 16 RGBA, an opaque disc with a half-transparent rim on a transparent ground)."""
import os
import struct
import zlib

import numpy as np

d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scenes")
rng = np.random.RandomState(7)
w, h = 24, 16
y, x = np.mgrid[0:h, 0:w]
img = np.stack([(x * 255 // (w - 1)), (y * 255 // (h - 1)), ((x // 4 + y // 4) % 2) * 200 + 30], axis=-1).astype(np.uint8)
img = (img.astype(int) + rng.randint(-20, 20, size=img.shape)).clip(0, 255).astype(np.uint8)
open(os.path.join(d, "img_ramp.ppm"), "wb").write(b"P6\n%d %d\n255\n" % (w, h) + img.tobytes())


def png(path, a):
    hh, ww, _ = a.shape
    raw = b"".join(b"\x00" + a[r].tobytes() for r in range(hh))

    def chunk(t, dat):
        return struct.pack(">I", len(dat)) + t + dat + struct.pack(">I", zlib.crc32(t + dat) & 0xffffffff)
    open(path, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", ww, hh, 8, 6, 0, 0, 0)) +
                           chunk(b"IDAT", zlib.compress(raw, 9)) + chunk(b"IEND", b""))


w = h = 16
y, x = np.mgrid[0:h, 0:w]
r = np.hypot(x - 7.5, y - 7.5)
a = np.zeros((h, w, 4), np.uint8)
a[..., 0] = (x * 16).clip(0, 255)
a[..., 1] = 200 - (y * 10)
a[..., 2] = ((x + y) % 3) * 100 + 40
a[..., 3] = np.where(r < 5, 255, np.where(r < 8, 140, 0))
png(os.path.join(d, "img_disc.png"), a)
