"""BASELINE.json configurations at FULL size against the compiled reference (oracle/_ref),
through the C-ABI: the workloads bench.py times are the workloads checked here.

* configs[1]  3840x2160 -> 480x270 cells, all+wide, 256 colours, ordered dither, on the three
  synthetic distributions: 1021 tiles of 127 cells on 592 tensor-core groups, i.e. the matcher's
  persistent loop runs a second round (the small sweeps never get there).
* configs[2]  1920x1080 sixels, generated 256-colour palette, Floyd-Steinberg.
* configs[3]  frames of the 1080p animation (shapes translated by one pixel per frame).
* configs[4]  7680x4320 -> 960x540 cells, DIN99d, exhaustive search: tolerance path, the
  differing fraction is printed and bounded.
* the device's DIN99d table, all 2^24 entries, against chafa_color_rgb_to_din99d ().
"""

import ctypes as C

import numpy as np
import pytest

import parity
from chafa_b200 import capi, synth

pytestmark = pytest.mark.gpu

C2_CFG = dict(mode=capi.CANVAS_MODE_INDEXED_256, symbols="all+wide", dither=capi.DITHER_ORDERED, work=0.5)


@pytest.mark.parametrize("kind", ["noise", "shapes", "alpha"])
def test_baseline_config2_full_frame(gpu, kind):
    """BASELINE configs[1] at 3840x2160: prepared pixmap, all 129 600 cell records and the
    printed bytes identical to the reference (60 ms of CPU on 16 threads)."""
    img = parity.make_image(kind, 3840, 2160, seed=100)
    res = parity.run_pair(img, 480, 270, n_threads=-1, **C2_CFG)
    assert res.n_cells == 129600
    assert res.pixels_equal, res
    assert res.cells_equal, res
    assert res.print_equal, res


def test_baseline_config2_full_frame_queued_canvases(gpu, reference):
    """The bench's device leg: three canvases on three CUDA streams rendering distinct 4K frames
    queued back to back, text left in HBM; every frame's text equals the reference's."""
    import torch
    frames = [parity.make_image("noise", 3840, 2160, seed=100 + i) for i in range(3)]
    d_frames = [torch.from_numpy(f).cuda() for f in frames]
    streams = [torch.cuda.Stream() for _ in range(3)]
    canvases = [capi.Canvas(gpu, 480, 270, **C2_CFG) for _ in range(3)]
    for cv, st in zip(canvases, streams):
        gpu.chafa_b200_canvas_set_stream(cv.handle, st.cuda_stream)
    torch.cuda.synchronize()
    outs = []
    for rnd in range(2):        # second round: same canvases again while the first round may still be running
        outs = []
        for cv, d in zip(canvases, d_frames):
            cv.draw_device(d.data_ptr(), 3840, 2160)
            outs.append(cv.print_device_enqueue())
    torch.cuda.synchronize()
    reference.chafa_set_n_threads(-1)
    cr = capi.Canvas(reference, 480, 270, **C2_CFG)
    for f, (d_text, d_len) in zip(frames, outs):
        n = C.c_uint64(0)
        assert gpu.chafa_b200_copy_from_device(C.addressof(n), d_len, 8)
        buf = (C.c_char * n.value)()
        assert gpu.chafa_b200_copy_from_device(C.addressof(buf), d_text, n.value)
        cr.draw(f, 3840, 2160)
        assert bytes(buf) == cr.print()
    cr.close()
    for cv in canvases:
        cv.close()


@pytest.mark.parametrize("kind", ["noise", "shapes", "alpha"])
def test_baseline_config3_sixel_full_frame(gpu, kind):
    """BASELINE configs[2] at 1920x1080: generated palette, 2 073 600 pens and the sixel text
    identical to the reference."""
    img = parity.make_image(kind, 1920, 1080, seed=100)
    res = parity.run_sixel_pair(img, 240, 135, n_threads=-1, dither=capi.DITHER_DIFFUSION, grain=(1, 1))
    assert res.n_pixels == 1920 * 1080
    assert res.palette_equal, res
    assert res.image_equal, res
    assert res.print_equal, res


def test_baseline_config4_animation_frames(gpu, reference):
    """BASELINE configs[3]: frames of the 1080p animation (shapes translated by one pixel per
    frame) on ONE reused canvas, as bench.py renders them: every frame's text
    identical to the reference's."""
    cfg = dict(mode=capi.CANVAS_MODE_TRUECOLOR, symbols="block+border", work=0.5)
    reference.chafa_set_n_threads(-1)
    cr = capi.Canvas(reference, 240, 67, **cfg)
    cp = capi.Canvas(gpu, 240, 67, **cfg)
    frame = synth.animation(1920, 1080, 300, seed=100)
    for f in (0, 1, 2, 150, 299):
        img = frame(f)
        cr.draw(img, 1920, 1080)
        cp.draw(img, 1920, 1080)
        assert (cr.cells() == cp.cells()).all(), f
        assert cr.print() == cp.print(), f
    cr.close()
    cp.close()


# Tolerance of the DIN99d path (double-precision transcendental code: glibc libm on the CPU,
# CUDA's math library on the device): the prepared pixmap may differ by at most 1 per channel in
# at most 1e-4 of the pixels, and at most 0.5 % of the cells may carry a different symbol or pen.
DIN99D_PIXEL_FRACTION = 1e-4
DIN99D_CELL_FRACTION = 5e-3


@pytest.mark.parametrize("kind", ["noise", "shapes"])
def test_baseline_config5_full_frame_din99d(gpu, kind):
    """BASELINE configs[4] at 7680x4320 -> 960x540 cells, DIN99d, INDEXED_256, symbols all,
    work 1.0 (every symbol scored for every cell: 6.5e8 evaluations)."""
    img = parity.make_image(kind, 7680, 4320, seed=100)
    res = parity.run_pair(img, 960, 540, n_threads=-1, mode=capi.CANVAS_MODE_INDEXED_256, symbols="all", work=1.0,
                          color_space=capi.COLOR_SPACE_DIN99D)
    n_px = 7680 * 4320
    print(f"\nconfigs[4] {kind}: symbol choices differing {res.symbol_mismatches}/{res.n_cells} "
          f"({res.symbol_mismatches / res.n_cells:.2e}), cells differing {res.cell_mismatches}/{res.n_cells}, "
          f"prepared pixels differing {res.pixel_mismatches}/{n_px} ({res.pixel_mismatches / n_px:.2e}), "
          f"printed bytes equal: {res.print_equal}")
    assert res.n_cells == 518400
    assert res.pixel_mismatches <= n_px * DIN99D_PIXEL_FRACTION, res
    assert res.cell_mismatches <= res.n_cells * DIN99D_CELL_FRACTION, res
    if res.pixel_mismatches == 0:
        # same pixmap in: everything after it is integer arithmetic and must be identical
        assert res.cells_equal and res.print_equal, res


def test_device_din99d_table_exhaustive(gpu, reference):
    """All 2^24 entries of the device's DIN99d table against chafa_color_rgb_to_din99d ()
    (chafa-color.c:130-168) evaluated by the compiled reference: exact mismatch count, and no
    entry further than 1 in any channel."""
    reference.lib.ref_probe_din99d_table.argtypes = [C.c_void_p]
    reference.lib.ref_probe_din99d_table.restype = None
    gpu.lib.chafa_b200_get_din99d_table.argtypes = [C.c_void_p]
    gpu.lib.chafa_b200_get_din99d_table.restype = C.c_int
    ref = np.zeros(1 << 24, np.uint32)
    dev = np.zeros(1 << 24, np.uint32)
    reference.lib.ref_probe_din99d_table(ref.ctypes.data)
    assert gpu.lib.chafa_b200_get_din99d_table(dev.ctypes.data)
    bad = np.nonzero(ref != dev)[0]
    worst = 0
    if bad.size:
        a = ref[bad].view(np.uint8).reshape(-1, 4).astype(np.int16)
        b = dev[bad].view(np.uint8).reshape(-1, 4).astype(np.int16)
        worst = int(np.abs(a - b).max())
    print(f"\nDIN99d table: {bad.size} of {1 << 24} entries differ from the reference (largest channel "
          f"difference {worst})" + (f"; first: rgb {int(bad[0]):06x} ref {int(ref[bad[0]]):06x} dev {int(dev[bad[0]]):06x}"
                                    if bad.size else ""))
    assert worst <= 1
    assert bad.size <= (1 << 24) * DIN99D_PIXEL_FRACTION
