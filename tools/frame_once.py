"""A few device-resident frames of a bench workload, for profiling runs:
    python tools/frame_once.py WORKLOAD [image kind] [n_frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from chafa_b200 import capi, synth  # noqa: E402

name = sys.argv[1]
kind = sys.argv[2] if len(sys.argv) > 2 else "noise"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
wl = bench.WORKLOADS[name]
(w, h), (cw, ch) = wl["src"], wl["cells"]
lib = capi.load_product()
src = torch.from_numpy(synth.KINDS[kind](w, h, seed=100)).cuda()
cv = capi.Canvas(lib, cw, ch, **wl["cfg"])
for i in range(n):
    cv.draw_device(src.data_ptr(), w, h)
    if bench.is_sixel(wl):
        cv.print_len()
    else:
        cv.print_device_enqueue()
torch.cuda.synchronize()
print("done", name, kind, n)
