"""Tensor-core matcher vs the warp-per-cell kernels on large canvases (8K sources, rows that do not
divide into tiles): cell records and printed bytes must be equal."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from chafa_b200 import capi, synth
prod = capi.load_product()
prod.lib.chafa_b200_canvas_set_tensor_match.argtypes = [C.c_void_p, C.c_int]
for (w, h, cw, ch, cfg) in [(7680, 4320, 960, 540, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_256)),
                            (7672, 4312, 959, 539, dict(symbols="all+wide")),
                            (1920, 1080, 240, 67, dict(symbols="block+border", work=0.5))]:
    img = synth.KINDS["shapes"](w, h, 7)
    out = []
    for flag in (1, 0):
        c = capi.Canvas(prod, cw, ch, **cfg)
        prod.lib.chafa_b200_canvas_set_tensor_match(c.handle, flag)
        c.draw(img, w, h)
        out.append((c.cells().copy(), c.print()))
        c.close()
    print(w, h, cw, ch, "cells equal", bool((out[0][0] == out[1][0]).all()), "print equal", out[0][1] == out[1][1], len(out[0][1]))
