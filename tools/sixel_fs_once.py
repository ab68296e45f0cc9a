import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import parity
from chafa_b200 import capi
lib = capi.load_product()
img = parity.make_image("shapes", 1920, 1080)
c = capi.Canvas(lib, 240, 135, pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1))
c.draw(img, 1920, 1080)
print(len(c.print()))
