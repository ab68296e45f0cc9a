import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import parity
from chafa_b200 import capi
lib = capi.load_product()
img = parity.make_image("shapes", 1920, 1080)
for dith, name in ((capi.DITHER_DIFFUSION, "fs"), (capi.DITHER_NOISE, "noise")):
    c = capi.Canvas(lib, 240, 135, pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=dith, grain=(1, 1))
    for i in range(3):
        t0 = time.perf_counter(); c.draw(img, 1920, 1080); t1 = time.perf_counter(); s = c.print(); t2 = time.perf_counter()
        print(name, "draw %.1f ms print %.1f ms len %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, len(s)))
ref = capi.load_reference(); ref.chafa_set_n_threads(-1)
for dith, name in ((capi.DITHER_DIFFUSION, "fs"), (capi.DITHER_NOISE, "noise")):
    c = capi.Canvas(ref, 240, 135, pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=dith, grain=(1, 1))
    for i in range(2):
        t0 = time.perf_counter(); c.draw(img, 1920, 1080); t1 = time.perf_counter(); s = c.print(); t2 = time.perf_counter()
        print("ref", name, "draw %.1f ms print %.1f ms len %d" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, len(s)))
