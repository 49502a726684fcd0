"""Generate the synthetic 2048x2048 texture atlases the bundled scenes use.

The reference ships three 2048^2 RGB checker atlases (8x8 labelled cells) under resources/;
they are input data of the reference repository and are not copied here.  These stand-ins
have the same size, format and character (flat cells, sharp cell borders, small dark
high-frequency marks inside every cell), which is what matters for the bilinear lookup
path (src/rendering/texture.rs:62-90).  Deterministic: same bytes on every run.
"""
import os
import sys

import numpy as np
from PIL import Image

PALETTES = {
    "celestial": [(233, 150, 170), (150, 180, 215), (240, 225, 140), (135, 195, 160)],
    "disk": [(255, 110, 60), (245, 150, 40), (190, 40, 50), (255, 160, 30)],
    "sphere": [(120, 200, 230), (90, 120, 200), (200, 230, 240), (60, 170, 160)],
}


def atlas(name, size=2048, cells=8):
    pal = np.array(PALETTES[name], dtype=np.uint8)
    cell = size // cells
    yy, xx = np.mgrid[0:size, 0:size]
    cy, cx = yy // cell, xx // cell
    idx = (cx + 3 * cy + (cy // 2)) % 4
    img = pal[idx]
    # marks: a ring and a bar per cell whose size depends on the cell index (stand-in for the labels)
    ly, lx = (yy % cell) - cell // 2, (xx % cell) - cell // 2
    rad = np.sqrt(lx * lx + ly * ly)
    ring_r = 30 + 6 * ((cx + cy) % 5)
    ring = np.abs(rad - ring_r) < 5
    bar = (np.abs(ly + 60) < 4) & (np.abs(lx) < 20 + 8 * (cx % 4))
    tick = (np.abs(lx - 70) < 3) & (np.abs(ly) < 10 + 6 * (cy % 4))
    img = img.copy()
    img[ring | bar | tick] = (25, 25, 25)
    return img


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name in PALETTES:
        path = os.path.join(out_dir, name + ".png")
        Image.fromarray(atlas(name), "RGB").save(path, optimize=False, compress_level=6)
    # 1x1 swatches (the reference has a few of these as well)
    Image.fromarray(np.array([[[255, 160, 40, 128]]], dtype=np.uint8), "RGBA").save(
        os.path.join(out_dir, "orange_transparent.png"))
    Image.fromarray(np.array([[[255, 255, 255]]], dtype=np.uint8), "RGB").save(os.path.join(out_dir, "white.png"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "resources"))
