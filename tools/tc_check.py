"""A/B check of the tensor-core matcher (match_tc.cu) against the warp-per-cell kernels and the
compiled reference on a few configurations; prints mismatch counts and kernel times."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from chafa_b200 import capi, synth  # noqa: E402

CASES = [
    ("noise", 256, 128, 32, 16, dict(symbols="all+wide")),
    ("shapes", 256, 128, 32, 16, dict(symbols="all+wide")),
    ("shapes", 320, 160, 40, 20, dict()),
    ("alpha", 1024, 512, 128, 64, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_256)),
    ("noise", 1024, 512, 128, 64, dict(symbols="all", mode=capi.CANVAS_MODE_INDEXED_240)),
    ("shapes", 1016, 504, 127, 63, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_16)),
    ("shapes", 256, 128, 32, 16, dict(symbols="all+wide", work=1.0)),
    ("noise", 1024, 512, 128, 64, dict(symbols="all+wide", work=1.0, mode=capi.CANVAS_MODE_INDEXED_256)),
    ("alpha", 1016, 504, 127, 63, dict(symbols="all", work=1.0, mode=capi.CANVAS_MODE_INDEXED_240,
                                       color_space=capi.COLOR_SPACE_DIN99D)),
    ("shapes", 320, 160, 40, 20, dict(work=0.9, mode=capi.CANVAS_MODE_FGBG)),
    ("alpha", 640, 320, 80, 40, dict(symbols="block+border+wide", work=1.0, mode=capi.CANVAS_MODE_INDEXED_16)),
    ("noise", 3840, 2160, 480, 270, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_256,
                                         dither=capi.DITHER_ORDERED)),
]


def main():
    prod = capi.load_product()
    prod.lib.chafa_b200_canvas_set_tensor_match.argtypes = [C.c_void_p, C.c_int]
    ref = capi.load_reference() if os.path.exists(capi.REFERENCE_LIB) else None
    bad = 0
    for kind, w, h, cw, ch, cfg in CASES:
        img = synth.KINDS[kind](w, h, 1)
        out = {}
        for name, flag in (("tensor", 1), ("warp", 0)):
            c = capi.Canvas(prod, cw, ch, **cfg)
            prod.lib.chafa_b200_canvas_set_tensor_match(c.handle, flag)
            c.draw(img, w, h)
            cells = c.cells().copy()
            t0 = time.perf_counter()
            for _ in range(5):
                c.draw(img, w, h)
            c.cells()
            dt = (time.perf_counter() - t0) / 5
            out[name] = (cells, c.print(), dt)
            c.close()
        a, b = out["tensor"][0], out["warp"][0]
        diff = (a != b).any(axis=2)
        msg = f"{kind} {cw}x{ch} {cfg}: tensor vs warp cells differ {int(diff.sum())}/{diff.size}, " \
              f"print equal {out['tensor'][1] == out['warp'][1]}, draw {out['tensor'][2]*1e3:.2f} / {out['warp'][2]*1e3:.2f} ms"
        if ref is not None and cw * ch <= 130 * 70:
            ref.chafa_set_n_threads(4)
            cr = capi.Canvas(ref, cw, ch, **cfg)
            cr.draw(img, w, h)
            r = cr.cells()
            msg += f"; vs reference {int((a != r).any(axis=2).sum())}"
            if out["tensor"][1] != cr.print():
                msg += " PRINT DIFFERS"
                bad += 1
            cr.close()
        if diff.any():
            bad += 1
            ys, xs = np.nonzero(diff)
            for y, x in list(zip(ys, xs))[:6]:
                msg += f"\n   ({x},{y}) tensor {[hex(v) for v in a[y, x]]} warp {[hex(v) for v in b[y, x]]}"
        print(msg, flush=True)
    print("FAILED" if bad else "ALL OK")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
