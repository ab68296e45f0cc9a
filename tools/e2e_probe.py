"""Where does the end-to-end time go?  Per-stage CUDA-event times through the host API."""
import ctypes as C, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chafa_b200 import capi, synth
import bench
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2_4k_256c_ordered_allwide"]
lib = capi.load_product()
(w, h), (cw, ch) = wl["src"], wl["cells"]
src = torch.from_numpy(synth.noise(w, h, 5)).pin_memory()
c = capi.Canvas(lib, cw, ch, **wl["cfg"])
lib.lib.chafa_b200_canvas_set_kernel_timing.argtypes = [C.c_void_p, C.c_int]
lib.lib.chafa_b200_canvas_get_kernel_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
for i in range(3):
    c.draw(src.data_ptr(), w, h); c.print()
lib.lib.chafa_b200_canvas_set_kernel_timing(c.handle, 1)
N = 10
t_draw = t_print = 0.0
for i in range(N):
    t0 = time.perf_counter(); c.draw(src.data_ptr(), w, h); t1 = time.perf_counter(); s = c.print(); t2 = time.perf_counter()
    t_draw += t1 - t0; t_print += t2 - t1
ms = (C.c_double * 16)(); cnt = (C.c_uint64 * 16)()
lib.lib.chafa_b200_canvas_get_kernel_times(c.handle, ms, cnt, 16)
print({bench.STAGES[i]: round(ms[i] / max(cnt[i], 1), 4) for i in range(9)})
print("host wall: draw %.3f ms, print %.3f ms, out %d bytes" % (t_draw / N * 1e3, t_print / N * 1e3, len(s)))
