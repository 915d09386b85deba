#!/usr/bin/env python
"""Derives tests/golden/render_bmp_8x8.npz from the reference's committed render, PathTracerAP/Render.bmp.

Render.bmp (1000x800, 24 bpp) is the only golden artefact the reference ships: the image its author's GPU produced
for the scene coded in Scene.cpp:3-224.  The iteration count it was rendered with is not recorded, so it pins the path statistically:
the fixture holds the 8x8 box-filtered image (100 x 125 x 3 means of the stored bytes, in FILE channel order, i.e. the film's R,G,B as
Renderer.cpp:41-52 stores them) and the per-channel means.  Run in the build container, where /root/reference exists:
    python tools/make_render_bmp_fixture.py
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/PathTracerAP/Render.bmp"


def read_bmp24(path):
    """Rows as stored (first row = y 0 of the film, Renderer.cpp:41), channels as stored."""
    b = open(path, "rb").read()
    off, = struct.unpack("<I", b[10:14])
    W, H, planes, bpp = struct.unpack("<iiHH", b[18:30])
    assert b[:2] == b"BM" and bpp == 24 and (3 * W) % 4 == 0, "Renderer.cpp:15-63 writes unpadded 24-bpp rows"
    return np.frombuffer(b, np.uint8, W * H * 3, off).reshape(H, W, 3)


def box8(img):
    H, W, _ = img.shape
    return img.astype(np.float64).reshape(H // 8, 8, W // 8, 8, 3).mean(axis=(1, 3))


def main():
    img = read_bmp24(SRC)
    out = os.path.join(ROOT, "tests", "golden", "render_bmp_8x8.npz")
    np.savez_compressed(out, box8=box8(img).astype(np.float32), channel_means=img.reshape(-1, 3).mean(0), shape=np.array(img.shape[:2], np.int32))
    print(out, img.shape, img.reshape(-1, 3).mean(0))


if __name__ == "__main__":
    sys.exit(main())
