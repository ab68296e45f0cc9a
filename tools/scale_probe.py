"""Scaler A/B probe: one geometry per process (a faulting kernel kills the context).
usage: scale_probe.py SRC_W SRC_H CELLS_W CELLS_H [kind]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import parity
from chafa_b200 import capi

sw, sh, cw, ch = map(int, sys.argv[1:5])
kind = sys.argv[5] if len(sys.argv) > 5 else "shapes"
img = parity.make_image(kind, sw, sh, seed=7)
res = parity.run_pair(img, cw, ch)
print(sw, sh, cw, ch, kind, "gather" if os.environ.get("CHAFA_B200_SCALE_GATHER") else "tma", res,
      capi.load_product().chafa_b200_last_error())
