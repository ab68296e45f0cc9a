"""Sixel diffusion throughput against the number of canvases in flight:
    python tools/replica_scaling.py [shapes|noise] [counts...]"""
import os
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from chafa_b200 import capi, synth  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "shapes"
counts = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8, 16, 32]
lib = capi.load_product()
src = torch.from_numpy(synth.KINDS[kind](1920, 1080, seed=100)).cuda()
cfg = dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1))
for n in counts:
    streams = [torch.cuda.Stream() for _ in range(n)]
    canvases = [capi.Canvas(lib, 240, 135, **cfg) for _ in range(n)]
    for cv, st in zip(canvases, streams):
        lib.chafa_b200_canvas_set_stream(cv.handle, st.cuda_stream)
    lat = [0.0] * n
    start = threading.Barrier(n + 1)

    def worker(t):
        canvases[t].draw_device(src.data_ptr(), 1920, 1080)
        canvases[t].print_len()
        start.wait()
        for _ in range(2):
            t0 = time.perf_counter()
            canvases[t].draw_device(src.data_ptr(), 1920, 1080)
            canvases[t].print_len()
            lat[t] = time.perf_counter() - t0

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n)]
    for th in threads:
        th.start()
    start.wait()
    t0 = time.perf_counter()
    for th in threads:
        th.join()
    dt = time.perf_counter() - t0
    print(f"{kind}: {n:3d} canvases in flight: {2 * n / dt:7.1f} frames/s, frame latency {1e3 * sum(lat) / n:7.1f} ms", flush=True)
    for cv in canvases:
        cv.close()
