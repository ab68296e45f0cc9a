"""2+ GPUs: a still in a canvas mode that normalises against a whole-image histogram
(16 colours), rendered as row bands with the histogram all-reduced over NCCL, must equal the
single-GPU render.  Run:  torchrun --nproc-per-node 2 tools/band_allreduce_check.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from chafa_b200 import capi, shard, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
lib = capi.load_product()
lib.chafa_b200_set_device(int(os.environ["LOCAL_RANK"]))

ok = True
for mode, name in ((capi.CANVAS_MODE_INDEXED_16, "16"), (capi.CANVAS_MODE_INDEXED_8, "8"), (capi.CANVAS_MODE_FGBG, "fgbg")):
    w, h, cw, ch = 1920, 1080, 240, 135
    img = synth.shapes(w, h, seed=77)
    d_img = torch.from_numpy(img).cuda()
    cfg = dict(mode=mode, symbols="all+wide", dither=capi.DITHER_ORDERED)

    full = capi.Canvas(lib, cw, ch, **cfg)
    full.draw_device(d_img.data_ptr(), w, h)
    expected = full.print()

    stream = torch.cuda.current_stream()
    band = capi.Canvas(lib, cw, ch, **cfg)
    lib.chafa_b200_canvas_set_stream(band.handle, stream.cuda_stream)
    hist = torch.zeros(2051, dtype=torch.int32, device="cuda")
    lib.chafa_b200_canvas_set_histogram_buffer(band.handle, hist.data_ptr())
    first, n = shard.split_rows(ch, world)[rank]
    lib.chafa_b200_canvas_set_band(band.handle, first, n)
    ptr, cnt = C.c_void_p(), C.c_size_t()
    need = lib.chafa_b200_canvas_draw_begin_device(band.handle, capi.PIXEL_RGBA8_UNASSOCIATED, d_img.data_ptr(), w, h,
                                                   w * 4, C.byref(ptr), C.byref(cnt))
    if need:
        dist.all_reduce(hist[:2049])            # exact: integer sums
    lib.chafa_b200_canvas_draw_finish(band.handle)
    text = band.print()
    parts = [None] * world
    dist.all_gather_object(parts, text)
    got = shard.join_bands(parts)
    same = got == expected
    ok = ok and same
    if rank == 0:
        print(f"mode {name}: reduction needed {bool(need)}, bands == full: {same} ({len(got)} bytes)")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
