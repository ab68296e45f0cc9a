"""Does rendering a stream of frames on TWO canvases (two CUDA streams) beat one canvas?  Device-resident
sources, text left in HBM, CUDA events across both streams."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from chafa_b200 import capi, synth  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2_4k_256c_ordered_allwide"]
kind = sys.argv[2] if len(sys.argv) > 2 else "noise"
lib = capi.load_product()
(w, h), (cw, ch) = wl["src"], wl["cells"]
srcs = [torch.from_numpy(synth.KINDS[kind](w, h, seed=100 + i)).cuda() for i in range(8)]
for n_canvas in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(n_canvas)]
    canvases = [capi.Canvas(lib, cw, ch, **wl["cfg"]) for _ in range(n_canvas)]
    for c, s in zip(canvases, streams):
        lib.chafa_b200_canvas_set_stream(c.handle, s.cuda_stream)

    def frame(i):
        c = canvases[i % n_canvas]
        c.draw_device(srcs[i % 8].data_ptr(), w, h)
        c.print_device_enqueue()

    for i in range(6):
        frame(i)
    torch.cuda.synchronize()
    K = 60
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]:
        s.wait_event(e0)
    for i in range(K):
        frame(i)
    for s in streams[1:]:
        ev = torch.cuda.Event()
        ev.record(s)
        streams[0].wait_event(ev)
    e1.record(streams[0])
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"{n_canvas} canvas(es): {ms:.4f} ms/frame, {cw * ch / ms / 1e3:.1f} Mcell/s")
    for c in canvases:
        c.close()
