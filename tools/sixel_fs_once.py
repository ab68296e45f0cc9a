"""One 1080p sixel frame with Floyd-Steinberg (BASELINE configs[2]) for profiling runs:
    python tools/sixel_fs_once.py [shapes|noise|alpha] [n_frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import parity  # noqa: E402
from chafa_b200 import capi  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "shapes"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = capi.load_product()
img = parity.make_image(kind, 1920, 1080, seed=100)
c = capi.Canvas(lib, 240, 135, pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1))
for i in range(n):
    t0 = time.perf_counter()
    c.draw(img, 1920, 1080)
    t1 = time.perf_counter()
    out = c.print()
    t2 = time.perf_counter()
    print(f"{kind} frame {i}: draw call {1e3 * (t1 - t0):.2f} ms, print {1e3 * (t2 - t1):.2f} ms, {len(out)} bytes")
