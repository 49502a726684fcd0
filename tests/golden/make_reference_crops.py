"""Cuts the crops of the REFERENCE's own rendered output that tests/test_reference_image.py compares against.

    python tests/golden/make_reference_crops.py      (needs /root/reference; writes tests/golden/reference_hero_crops.npz)

Source: /root/reference/kerr_black_hole_with_stars.png (1280x720, 8-bit sRGB), the README image of
mdreem/gr_raytracer, rendered by the reference binary with the frozen recipe of
/root/reference/images/create-main-image.sh:40-79 (KerrBL a = 0.499, VolumetricDisc 0.8-16, T = 12000 K, Perlin seed 1,
camera -17,0,1.5, theta = -3.14159, default 4x4 adaptive supersampling, Reinhard tone mapping).  Only rendered pixel
values are kept (two crops, 60 x 220 and 40 x 200 pixels), no reference source.  The star-field background of that image
is not in the reference repository, so the tests compare against a black sky and use the crops one-sidedly where stars
shine through (see the test)."""
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
CROPS = {"disc_left": (380, 440, 200, 420), "disc_far_side": (320, 360, 330, 530)}   # row0, row1, col0, col1

if __name__ == "__main__":
    ref = np.asarray(Image.open("/root/reference/kerr_black_hole_with_stars.png").convert("RGB"))
    assert ref.shape == (720, 1280, 3)
    out = {}
    for name, (r0, r1, c0, c1) in CROPS.items():
        out[name] = ref[r0:r1, c0:c1].copy()
        out[name + "_box"] = np.array([r0, r1, c0, c1], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "reference_hero_crops.npz"), **out)
    print({k: v.shape for k, v in out.items()})
