"""Golden vector for the output stage (b200rt_quantise_rgba8): a float RGBA image with the edge values of the clamp,
written by the reference's own write_image_png (through oracle/_ref) and decoded back from the PNG.

    python tests/golden/make_golden_quantise.py        (in the container that has /root/reference)
"""
import ctypes as C
import os
import sys
import tempfile

import numpy as np
from PIL import Image as PILImage

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.oracle import RefOracle  # noqa: E402


def main():
    o = RefOracle()
    L = o.lib
    L.refo_write_png.restype = C.c_int
    L.refo_write_png.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_char_p, C.c_int]
    rng = np.random.default_rng(11)
    w, h = 53, 29
    img = rng.random((h, w, 4), dtype=np.float32) * 1.3 - 0.15              # below 0 and above 1 included
    edge = np.array([0.0, 1.0, 0.999999, 1.0 / 255.0, 0.5, 254.999 / 255.0, 255.0 / 255.0, 2.5, -1.0, 1e-9, 0.00390625, 0.99609375], np.float32)
    img[0, :len(edge), 0] = edge; img[1, :len(edge), 1] = edge; img[2, :len(edge), 2] = edge; img[3, :len(edge), 3] = edge
    out = {"image": img}
    for flip in (1, 0):
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "q.png")
            rc = L.refo_write_png(img.ctypes.data_as(C.POINTER(C.c_float)), w, h, path.encode(), flip)
            assert rc == 0
            out[f"rgba8_flip{flip}"] = np.asarray(PILImage.open(path).convert("RGBA"), np.uint8).copy()
    np.savez_compressed(os.path.join(HERE, "quantise.npz"), **out)
    print("quantise.npz written", out["rgba8_flip1"].shape)


if __name__ == "__main__":
    main()
