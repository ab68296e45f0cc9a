#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the compiled reference
(oracle/_ref/libchafa_ref.so, built by oracle/Makefile from /root/reference).

    python tests/golden/gen_golden.py

Each .npz holds: the source image, the canvas configuration, the compiled symbol tables in
evaluation order, the prepared pixmap, the cell records and the printed bytes, all produced
by the reference's own code with one thread.  They let the oracle restatement and the CUDA
path be checked where /root/reference (and even oracle/_ref) is absent."""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity  # noqa: E402
from chafa_b200 import capi  # noqa: E402

CASES = {
    # name: (image kind, src w, src h, cells w, cells h, config)
    "truecolor_block": ("shapes", 192, 80, 24, 10, dict(symbols="block+border+space")),
    "truecolor_all_wide": ("shapes", 192, 80, 24, 10, dict(symbols="all+wide")),
    "truecolor_noise_all": ("noise", 160, 64, 20, 8, dict(symbols="all")),
    "truecolor_alpha": ("alpha", 192, 80, 24, 10, dict(symbols="block+border+space")),
    "truecolor_exhaustive": ("shapes", 96, 48, 12, 6, dict(symbols="all+wide", work=1.0)),
    "idx256_all_wide": ("shapes", 192, 80, 24, 10, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_256)),
    "idx240_noise": ("noise", 160, 64, 20, 8, dict(symbols="block+border+space", mode=capi.CANVAS_MODE_INDEXED_240)),
    "idx16_alpha": ("alpha", 192, 80, 24, 10, dict(symbols="ascii", mode=capi.CANVAS_MODE_INDEXED_16,
                                                    preprocessing=False)),
    "truecolor_downscale": ("shapes", 500, 300, 24, 10, dict(symbols="block+border+space")),
    "idx256_ordered": ("shapes", 192, 80, 24, 10, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_256,
                                                       dither=capi.DITHER_ORDERED)),
}
SIXEL_CASES = {
    "sixel_fs": ("shapes", 160, 96, 16, 6, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
    "sixel_fs_alpha": ("alpha", 160, 96, 16, 6, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
}


def compiled(lib, sel, wide):
    sm = lib.chafa_symbol_map_new()
    assert lib.chafa_symbol_map_apply_selectors(sm, sel.encode(), None)
    c = np.zeros(4096, np.uint32)
    b0 = np.zeros(4096, np.uint64)
    b1 = np.zeros(4096, np.uint64)
    n = lib.ref_probe_symbol_map_compiled(sm, wide, 4096, c.ctypes.data, b0.ctypes.data, b1.ctypes.data)
    lib.chafa_symbol_map_unref(sm)
    return c[:n].copy(), b0[:n].copy(), b1[:n].copy()


def main():
    ref = capi.load_reference()
    ref.chafa_set_n_threads(1)
    for name, (kind, w, h, cw, ch, cfg) in CASES.items():
        img = parity.make_image(kind, w, h, seed=42)
        c = capi.Canvas(ref, cw, ch, **cfg)
        c.draw(img, w, h)
        pixmap = c.prepared_pixels(img, w, h)
        chars, bm, _ = compiled(ref, cfg["symbols"], 0)
        chars2, bm2a, bm2b = compiled(ref, cfg["symbols"], 1)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), src=img, config=json.dumps(dict(cells=[cw, ch], cfg=cfg)),
            pixmap=pixmap, chars=chars, bitmaps=bm, chars2=chars2, bitmaps2=np.stack([bm2a, bm2b], axis=1).ravel(),
            info=np.array(c.info(), np.int32), cells=c.cells(), printed=np.frombuffer(c.print(), np.uint8))
        c.close()
        print("wrote", name)
    for name, (kind, w, h, cw, ch, cfg) in SIXEL_CASES.items():
        img = parity.make_image(kind, w, h, seed=42)
        c = capi.Canvas(ref, cw, ch, pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, **cfg)
        c.draw(img, w, h)
        indexed, pal, n = parity.sixel_image(c)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"), src=img, config=json.dumps(dict(cells=[cw, ch], cfg=cfg)),
            indexed=indexed, palette=pal, n_colors=np.array([n], np.int32),
            printed=np.frombuffer(c.print(), np.uint8))
        c.close()
        print("wrote", name)


if __name__ == "__main__":
    main()
