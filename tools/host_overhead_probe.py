"""How much host time does one frame's enqueue take?  (device-resident source, one canvas, one
stream, nothing waited for inside the loop)"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch
import bench
from chafa_b200 import capi, synth

lib = capi.load_product()
lib.chafa_b200_set_device(0)
name = sys.argv[1] if len(sys.argv) > 1 else "c1_1080p_truecolor_block"
wl = bench.WORKLOADS[name]
(w, h), (cw, ch) = wl["src"], wl["cells"]
srcs = [torch.from_numpy(synth.shapes(w, h, seed=s)).cuda() for s in range(4)]
st = torch.cuda.Stream()
cv = capi.Canvas(lib, cw, ch, **wl["cfg"])
lib.chafa_b200_canvas_set_stream(cv.handle, st.cuda_stream)
N = int(os.environ.get("N", "150"))
for mode in ("draw+print", "draw", "print"):
    for i in range(20):
        cv.draw_device(srcs[i % 4].data_ptr(), w, h); cv.print_device_enqueue()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        if mode != "print":
            cv.draw_device(srcs[i % 4].data_ptr(), w, h)
        if mode != "draw":
            cv.print_device_enqueue()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name} {mode:10s}: enqueue {1e6 * (t1 - t0) / N:7.2f} us/frame, with drain {1e6 * (t2 - t0) / N:7.2f} us/frame")
