"""Sixel diffusion throughput with the chains batched into one launch per group of canvases:
    python tools/batch_scaling.py [shapes|noise] G:M [G:M ...]     (G groups of M canvases, one host thread per group)
BATCH_GROUP_STREAM=1: all canvases of a group share ONE CUDA stream (a process has 32 hardware queues; with a stream
per canvas, streams of one group share queues with the stream another group's long chain launch sits on)."""
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
configs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(1, 64), (2, 64), (1, 128)]
lib = capi.load_product()
srcs = [torch.from_numpy(synth.KINDS[kind](1920, 1080, seed=100 + i)).cuda() for i in range(4)]
cfg = dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1))
ROUNDS = 2
for G, M in configs:
    groups = []
    for g in range(G):
        cvs = [capi.Canvas(lib, 240, 135, **cfg) for _ in range(M)]
        shared = torch.cuda.Stream() if os.environ.get("BATCH_GROUP_STREAM") else None
        for cv in cvs:
            st = shared or torch.cuda.Stream()
            lib.chafa_b200_canvas_set_stream(cv.handle, st.cuda_stream)
            cv._stream = st
        groups.append(cvs)
    start = threading.Barrier(G + 1)

    def worker(g):
        lib.chafa_b200_set_device(0)
        cvs = groups[g]
        ptrs = [srcs[(g + i) % len(srcs)].data_ptr() for i in range(M)]
        capi.draw_batch_device(lib, cvs, ptrs, 1920, 1080)      # warm-up: allocations, tables
        for cv in cvs:
            cv.print_len()
        start.wait()
        for r in range(ROUNDS):
            capi.draw_batch_device(lib, cvs, ptrs, 1920, 1080)
            for cv in cvs:
                cv.print_len()

    threads = [threading.Thread(target=worker, args=(g,)) for g in range(G)]
    for th in threads:
        th.start()
    start.wait()
    t0 = time.perf_counter()
    for th in threads:
        th.join()
    dt = time.perf_counter() - t0
    print(f"{kind}: {G} group(s) of {M}: {ROUNDS * G * M / dt:7.1f} frames/s ({dt / ROUNDS * 1e3:7.1f} ms per round)", flush=True)
    for cvs in groups:
        for cv in cvs:
            cv.close()
    torch.cuda.empty_cache()
