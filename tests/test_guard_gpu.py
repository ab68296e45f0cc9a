"""Out-of-bounds writes, checked by the library itself: compute-sanitizer is closed on the GPU pool
these tests run on (profiles/sanitizer/), so a cross-section of the parity cases -- the tensor-core
matchers with full and ragged tiles, the printer's look-back and staging routes, resolve scans, the
scaler's reducing and enlarging routes, preprocessing with statistics, both diffusion kernels, the
palette builder at every cube resolution, the sixel band encoder, image text -- runs with every device
allocation exact-sized and framed by 4 KB guard bands (chafa_b200_set_guard_mode), and no band
may have been written to."""

import pytest

import parity
from chafa_b200 import capi

pytestmark = pytest.mark.gpu

SYMBOL_CASES = ["truecolor_all_wide", "idx256_ordered_allwide", "truecolor_work9_exhaustive", "idx16_fs", "fgbg_ordered",
                "truecolor_down3x", "truecolor_down20x_box", "truecolor_upscale", "idx16_8_wide", "truecolor_odd_cells",
                "truecolor_median_noise_wide", "idx256_fill_alpha", "truecolor_placement_big_canvas", "idx8_fgonly"]
SIXEL_CASES = ["sixel_fs_default", "sixel_fs_alpha", "sixel_fs_intensity", "sixel_fs_16", "sixel_fs_cell10x20",
               "sixel_none", "sixel_ordered", "sixel_noise_din99d"]


@pytest.fixture()
def guarded(gpu):
    gpu.lib.chafa_b200_check_guards.restype = int
    before = gpu.lib.chafa_b200_check_guards()
    gpu.lib.chafa_b200_set_guard_mode(1)
    yield lambda: gpu.lib.chafa_b200_check_guards() - before
    gpu.lib.chafa_b200_set_guard_mode(0)


@pytest.mark.parametrize("name", SYMBOL_CASES)
def test_symbols_no_guard_band_touched(guarded, name):
    res = parity.run_named(name)
    assert res.ok, res
    assert guarded() == 0


@pytest.mark.parametrize("name", SIXEL_CASES)
def test_sixel_no_guard_band_touched(guarded, name):
    res = parity.run_sixel_named(name, n_threads=3)
    assert res.ok, res
    assert guarded() == 0


def test_tiles_rows_and_tables_no_guard_band_touched(guarded, gpu, reference):
    """Ragged last tile (255 x 64 cells = 128.5 tiles of 127), rows longer than the printer's staging
    buffer, a 900-row look-back, every cube resolution of the palette builder, Kitty text."""
    cases = [
        ("alpha", 2040, 512, 255, 64, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_256)),
        ("noise", 1016, 504, 127, 63, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_16, work=0.7)),
        ("alpha", 1016, 504, 127, 63, dict(symbols="all", work=1.0, mode=capi.CANVAS_MODE_INDEXED_240,
                                          color_space=capi.COLOR_SPACE_DIN99D)),
        ("alpha", 3000, 24, 3000, 3, dict(symbols="all")),
        ("alpha", 56, 900, 7, 900, dict(symbols="all")),
    ]
    for kind, w, h, cw, ch, cfg in cases:
        res = parity.run_pair(parity.make_image(kind, w, h, seed=5), cw, ch, check_pixels=False, n_threads=4, **cfg)
        assert res.cells_equal and res.print_equal, (cfg, res)
        assert guarded() == 0, cfg
    for work in (0.0, 0.5, 1.0):
        img = parity.make_image("noise", 512, 384, seed=43)
        res = parity.run_sixel_pair(img, 64, 48, n_threads=4, dither=capi.DITHER_DIFFUSION, grain=(1, 1), work=work)
        assert res.ok, res
        assert guarded() == 0, work
    img = parity.make_image("alpha", 333, 211)
    for mode in (capi.PIXEL_MODE_KITTY, capi.PIXEL_MODE_ITERM2):
        outs = []
        for lib in (reference, gpu):
            c = capi.Canvas(lib, 25, 11, pixel_mode=mode, symbols=None)
            c.draw(img, 333, 211)
            outs.append(c.print())
            c.close()
        assert outs[0] == outs[1]
        assert guarded() == 0, mode


def test_guard_mode_detects_a_stray_write(gpu):
    """The checker itself: a write one byte past an allocation made in guard mode is reported."""
    import ctypes as C
    gpu.lib.chafa_b200_check_guards.restype = int
    before = gpu.lib.chafa_b200_check_guards()
    gpu.lib.chafa_b200_set_guard_mode(1)
    try:
        c = capi.Canvas(gpu, 8, 4)
        c.draw(parity.make_image("shapes", 64, 32), 64, 32)
        gpu.lib.chafa_b200_debug_cells_address.restype = C.c_void_p
        gpu.lib.chafa_b200_debug_cells_address.argtypes = [C.c_void_p]
        cells = gpu.lib.chafa_b200_debug_cells_address(c.handle)
        assert cells
        # 8 x 4 cells of 12 bytes: the first byte after them
        gpu.lib.chafa_b200_debug_poke.argtypes = [C.c_void_p, C.c_int]
        gpu.lib.chafa_b200_debug_poke(cells + 8 * 4 * 12, 7)
        assert gpu.lib.chafa_b200_check_guards() - before == 1
        c.close()
    finally:
        gpu.lib.chafa_b200_set_guard_mode(0)
