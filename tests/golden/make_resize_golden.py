"""Generates tests/golden/resize.npz with PILLOW ITSELF -- the third-party library in which the arithmetic of the reference's
`_image.resize((width, height))` (detect.py:65) lives -- for both the filter this container's Pillow applies to that call
(BICUBIC, the default since Pillow 7) and the one the Pillow of the reference's time applied (NEAREST).
Run once, here:   python tests/golden/make_resize_golden.py"""
import os

import numpy as np
import PIL
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rs = np.random.RandomState(53)
    out = {"pillow_version": np.array(PIL.__version__)}
    # (in_h, in_w, out_h, out_w): shrink both, enlarge both, mixed, one axis unchanged, tiny input, strong shrink (wide antialias window)
    cases = {"shrink": (120, 160, 96, 96), "enlarge": (40, 56, 96, 128), "mixed": (150, 50, 64, 96), "same_w": (75, 96, 64, 96),
             "tiny": (3, 2, 32, 32), "strong": (160, 240, 16, 32)}
    for name, (h, w, oh, ow) in cases.items():
        img = rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        if name == "strong":
            img[::7] = 255                                      # structure the antialiasing window must average correctly
        out[name + "_in"] = img
        out[name + "_size"] = np.array([oh, ow])
        default = np.asarray(Image.fromarray(img).resize((ow, oh)))
        out[name + "_bicubic"] = np.asarray(Image.fromarray(img).resize((ow, oh), Image.Resampling.BICUBIC))
        out[name + "_nearest"] = np.asarray(Image.fromarray(img).resize((ow, oh), Image.Resampling.NEAREST))
        assert np.array_equal(default, out[name + "_bicubic"])   # what the reference's call gets from THIS Pillow
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **out)
    print("wrote resize.npz %.0f KiB (Pillow %s)" % (os.path.getsize(os.path.join(HERE, "resize.npz")) / 1024, PIL.__version__))


if __name__ == "__main__":
    main()
