"""Parity of the CUDA path against the oracle (oracle/_ref: the compiled reference), through
the C-ABI, on seeded synthetic inputs.  Integer colour spaces: byte-identical prepared pixmap,
cell records and chafa_canvas_print output."""

import os

import numpy as np
import pytest

import parity
from chafa_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(parity.SWEEP))
def test_sweep_byte_identical(gpu, name):
    res = parity.run_named(name)
    assert res.pixels_equal is not False, f"prepared pixmap differs: {res}"
    assert res.cells_equal, f"cells differ: {res}"
    assert res.print_equal, f"printed bytes differ: {res}"


@pytest.mark.parametrize("name", sorted(parity.SWEEP_DIN99D))
def test_din99d_within_tolerance(gpu, name):
    """DIN99d runs through double-precision transcendental code on both sides (libm vs CUDA
    math): tolerance = at most 0.5 % of cells may pick a different symbol, and prepared pixels
    may differ by at most 1 per channel."""
    kind, w, h, cw, ch, cfg = parity.SWEEP_DIN99D[name]
    res = parity.run_pair(parity.make_image(kind, w, h), cw, ch, check_pixels=True, **cfg)
    assert res.symbol_mismatches <= max(1, res.n_cells // 200), res
    print(f"{name}: symbol choices differing: {res.symbol_mismatches}/{res.n_cells}, "
          f"pixels differing: {res.pixel_mismatches}")


@pytest.mark.parametrize("kind", ["noise", "shapes", "alpha"])
def test_baseline_config1_geometry(gpu, kind):
    """BASELINE configs[0]: 1920x1080 -> 240x67 cells, truecolor, block+border, work 5."""
    img = parity.make_image(kind, 1920, 1080, seed=7)
    res = parity.run_pair(img, 240, 67, n_threads=-1, symbols="block+border", work=0.5)
    assert res.ok, res


def test_baseline_config2_band(gpu):
    """BASELINE configs[1] arithmetic on a band of the 4K frame: 3840x432 -> 480x54 cells,
    all+wide, 256 colours, ordered dither (the full frame is in test_fullsize_gpu.py)."""
    img = parity.make_image("shapes", 3840, 432, seed=9)
    res = parity.run_pair(img, 480, 54, n_threads=-1, mode=capi.CANVAS_MODE_INDEXED_256, symbols="all+wide",
                          dither=capi.DITHER_ORDERED)
    assert res.ok, res


@pytest.mark.parametrize("ptype", [capi.PIXEL_RGBA8_PREMULTIPLIED, capi.PIXEL_BGRA8_PREMULTIPLIED,
                                   capi.PIXEL_ARGB8_PREMULTIPLIED, capi.PIXEL_ABGR8_PREMULTIPLIED,
                                   capi.PIXEL_RGBA8_UNASSOCIATED, capi.PIXEL_BGRA8_UNASSOCIATED,
                                   capi.PIXEL_ARGB8_UNASSOCIATED, capi.PIXEL_ABGR8_UNASSOCIATED])
def test_pixel_formats_32bit(gpu, ptype):
    img = parity.make_image("alpha", 200, 120, seed=11)
    if ptype < 4:   # premultiplied formats need colour <= alpha
        a = img[..., 3:4].astype(np.uint16)
        img = np.concatenate([(img[..., :3].astype(np.uint16) * a // 255).astype(np.uint8), img[..., 3:4]], axis=2)
        img = np.ascontiguousarray(img)
    res = parity.run_pair(img, 30, 14, pixel_type=ptype)
    assert res.ok, res


@pytest.mark.parametrize("ptype", [capi.PIXEL_RGB8, capi.PIXEL_BGR8])
def test_pixel_formats_24bit(gpu, ptype):
    img = np.ascontiguousarray(parity.make_image("shapes", 200, 120, seed=12)[..., :3])
    res = parity.run_pair(img, 30, 14, pixel_type=ptype)
    assert res.ok, res


@pytest.mark.parametrize("geom", [(1, 1, 1, 1), (1, 1, 10, 5), (3, 2, 8, 4), (17, 31, 5, 9), (640, 3, 20, 2),
                                  (5, 700, 3, 20), (64, 64, 1, 1)])
def test_ragged_geometries(gpu, geom):
    w, h, cw, ch = geom
    img = parity.make_image("noise", w, h, seed=13)
    res = parity.run_pair(img, cw, ch)
    assert res.ok, res


def test_canvas_test_known_answers(gpu, reference):
    """The reference's own known-answer test (tests/canvas-test.c:72-206): a 1x1 black or white
    source in FGBG modes must give an all-blank or all-solid canvas."""
    import ctypes as C
    cases = [
        (capi.CANVAS_MODE_FGBG, "[ a]", 0.5, 1, 1), (capi.CANVAS_MODE_FGBG_BGFG, "[ a]", 0.5, 3, 3),
        (capi.CANVAS_MODE_FGBG, "block+border+space", 1.0, 10, 10),
        (capi.CANVAS_MODE_FGBG_BGFG, "block+border+space", 0.05, 100, 100),
    ]
    for mode, sel, work, cw, ch in cases:
        for value in (0, 255):
            px = np.array([[[value, value, value, 255]]], dtype=np.uint8)
            res = parity.run_pair(px, cw, ch, mode=mode, symbols=sel, work=work)
            assert res.ok, (mode, sel, work, cw, ch, value, res)
            chars = set(res.cells_prod[..., 0].ravel().tolist())
            assert len(chars) == 1, chars


def test_print_rows_and_accessors(gpu, reference):
    img = parity.make_image("alpha", 160, 96, seed=5)
    for lib_name in ("ref", "prod"):
        pass
    cr = capi.Canvas(reference, 20, 12, symbols="all+wide")
    cp = capi.Canvas(gpu, 20, 12, symbols="all+wide")
    cr.draw(img, 160, 96)
    cp.draw(img, 160, 96)
    assert cr.print_rows() == cp.print_rows()
    import ctypes as C
    for (x, y) in [(0, 0), (5, 3), (19, 11)]:
        assert reference.chafa_canvas_get_char_at(cr.handle, x, y) == gpu.chafa_canvas_get_char_at(cp.handle, x, y)
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        reference.chafa_canvas_get_colors_at(cr.handle, x, y, C.byref(a), C.byref(b))
        gpu.chafa_canvas_get_colors_at(cp.handle, x, y, C.byref(c), C.byref(d))
        assert (a.value, b.value) == (c.value, d.value)
    for lib, cv in ((reference, cr), (gpu, cp)):
        lib.chafa_canvas_set_char_at(cv.handle, 2, 2, ord("Z"))
        lib.chafa_canvas_set_colors_at(cv.handle, 2, 2, 0x102030, -1)
        lib.chafa_canvas_set_char_at(cv.handle, 4, 2, 0x3042)
        lib.chafa_canvas_set_raw_colors_at(cv.handle, 7, 5, 0xFF00FF, 0x00FF00)
    assert cr.print() == cp.print()
    assert (cr.cells() == cp.cells()).all()
    # drawing again replaces everything
    img2 = parity.make_image("shapes", 160, 96, seed=6)
    cr.draw(img2, 160, 96)
    cp.draw(img2, 160, 96)
    assert cr.print() == cp.print()


def test_empty_canvas_print(gpu, reference):
    """print before any draw: every cell is a space with colour 0 (maybe_clear)."""
    for mode in (capi.CANVAS_MODE_TRUECOLOR, capi.CANVAS_MODE_INDEXED_256, capi.CANVAS_MODE_FGBG):
        cr = capi.Canvas(reference, 7, 3, mode=mode)
        cp = capi.Canvas(gpu, 7, 3, mode=mode)
        assert cr.print() == cp.print()


def test_custom_term_info(gpu):
    """Caller-supplied sequences (no REPEAT_CHAR): same bytes as the reference."""
    S = capi.SEQ
    seqs = {S["RESET_ATTRIBUTES"]: "<R>", S["INVERT_COLORS"]: "<I>", S["SET_COLOR_FG_DIRECT"]: "<f%1,%2,%3>",
            S["SET_COLOR_BG_DIRECT"]: "<b%1,%2,%3>", S["SET_COLOR_FGBG_DIRECT"]: "<fb%1,%2,%3/%4,%5,%6>"}
    img = parity.make_image("alpha", 160, 96, seed=21)
    res = parity.run_pair(img, 20, 12, term_info_seqs=seqs)
    assert res.ok, res


def test_full_size_properties(gpu):
    """BASELINE configs[1] at full size: size-independent properties of the CUDA path --
    determinism, device-resident and host paths agree, rows concatenate to the full print,
    band rendering equals the matching rows of the full render."""
    import torch
    img = parity.make_image("shapes", 3840, 2160, seed=3)
    cfg = dict(mode=capi.CANVAS_MODE_INDEXED_256, symbols="all+wide", dither=capi.DITHER_ORDERED)
    c = capi.Canvas(gpu, 480, 270, **cfg)
    c.draw(img, 3840, 2160)
    full = c.print()
    c.draw(img, 3840, 2160)
    assert c.print() == full
    rows = c.print_rows()
    assert b"\n".join(rows) == full
    d = torch.from_numpy(img).cuda()
    c.draw_device(d.data_ptr(), 3840, 2160)
    ptr, n = c.print_device()
    assert n == len(full)
    import ctypes as C
    buf = (C.c_char * n)()
    assert gpu.chafa_b200_copy_from_device(C.addressof(buf), ptr, n)
    assert bytes(buf) == full
    # band of rows 100..149
    gpu.chafa_b200_canvas_set_band(c.handle, 100, 50)
    c.draw(img, 3840, 2160)
    band = c.print()
    assert band == b"\n".join(rows[100:150]) + b"\n"
    cells = c.cells()
    assert (cells[..., 0] != 0).any()


# ---- tensor-core matchers vs the warp-per-cell / table-driven kernels ------------------------

TENSOR_CASES = {
    "fast_allwide_truecolor": ("noise", 256, 128, 32, 16, dict(symbols="all+wide")),
    "fast_block_truecolor": ("shapes", 320, 160, 40, 20, dict()),
    "fast_allwide_256_tiles": ("alpha", 2040, 512, 255, 64, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_256)),
    "fast_all_240_din99d": ("shapes", 1024, 512, 128, 64, dict(symbols="all", mode=capi.CANVAS_MODE_INDEXED_240,
                                                             color_space=capi.COLOR_SPACE_DIN99D)),
    "fast_wide_16_work7": ("noise", 1016, 504, 127, 63, dict(symbols="all+wide", mode=capi.CANVAS_MODE_INDEXED_16, work=0.7)),
    "fast_one_symbol": ("noise", 128, 64, 16, 8, dict(symbols="vhalf")),
    "fast_flat_image": ("flat", 512, 256, 64, 32, dict(symbols="all+wide")),
    "exhaustive_allwide_truecolor": ("shapes", 256, 128, 32, 16, dict(symbols="all+wide", work=1.0)),
    "exhaustive_allwide_256": ("noise", 1024, 512, 128, 64, dict(symbols="all+wide", work=1.0,
                                                               mode=capi.CANVAS_MODE_INDEXED_256)),
    "exhaustive_all_240_din99d": ("alpha", 1016, 504, 127, 63, dict(symbols="all", work=1.0,
                                                                  mode=capi.CANVAS_MODE_INDEXED_240,
                                                                  color_space=capi.COLOR_SPACE_DIN99D)),
    "exhaustive_fgbg": ("shapes", 320, 160, 40, 20, dict(work=0.9, mode=capi.CANVAS_MODE_FGBG)),
    "exhaustive_wide_16": ("alpha", 640, 320, 80, 40, dict(symbols="block+border+wide", work=1.0,
                                                         mode=capi.CANVAS_MODE_INDEXED_16)),
}


@pytest.mark.parametrize("name", sorted(TENSOR_CASES))
def test_tensor_matchers_equal_scalar_kernels_and_reference(gpu, reference, name):
    """match_tc.cu (tcgen05 distances / masked sums, thread per cell) against the kernels in match.cu
    (popcount scan, subset tables) on the same canvas, and both against the compiled reference:
    cell records and printed bytes identical.  Tiles of 127 cells, rows that end inside a tile,
    a single-symbol table (degenerate bound), flat cells (massive ties)."""
    import ctypes as C
    import numpy as np
    kind, w, h, cw, ch, cfg = TENSOR_CASES[name]
    img = np.full((h, w, 4), (90, 120, 200, 255), dtype=np.uint8) if kind == "flat" else parity.make_image(kind, w, h, seed=5)
    gpu.lib.chafa_b200_canvas_set_tensor_match.argtypes = [C.c_void_p, C.c_int]
    out = []
    for flag in (1, 0):
        c = capi.Canvas(gpu, cw, ch, **cfg)
        gpu.lib.chafa_b200_canvas_set_tensor_match(c.handle, flag)
        c.draw(img, w, h)
        out.append((c.cells().copy(), c.print()))
        c.close()
    reference.chafa_set_n_threads(4)
    cr = capi.Canvas(reference, cw, ch, **cfg)
    cr.draw(img, w, h)
    ref_cells, ref_print = cr.cells().copy(), cr.print()
    cr.close()
    assert (out[0][0] == out[1][0]).all()
    assert out[0][1] == out[1][1]
    assert (out[0][0] == ref_cells).all()
    assert out[0][1] == ref_print


def test_printer_very_wide_rows(gpu, reference):
    """Rows too long for the printer's shared-memory staging (straight-to-output route) and rows
    of very different lengths (look-back over many rows): byte-identical to the reference."""
    for cw, ch, w, h in ((3000, 3, 3000, 24), (7, 900, 56, 900)):
        img = parity.make_image("alpha", w, h, seed=11)
        res = parity.run_pair(img, cw, ch, symbols="all", check_pixels=False)
        assert res.ok, res


# ---- sixel mode -------------------------------------------------------------------------

@pytest.mark.parametrize("name", sorted(parity.SIXEL_SWEEP))
def test_sixel_byte_identical(gpu, name):
    """Generated palette, pen image and printed sixel text identical to the oracle (diffusion
    dither and fixed palettes: no lookup cache involved)."""
    res = parity.run_sixel_named(name)
    assert res.palette_equal, res
    assert res.image_equal, res
    assert res.print_equal, res


@pytest.mark.parametrize("passthrough", [capi.PASSTHROUGH_SCREEN, capi.PASSTHROUGH_TMUX])
@pytest.mark.parametrize("name", ["sixel_fs_default", "sixel_fs_256"])
def test_sixel_passthrough_framing(gpu, name, passthrough):
    """screen / tmux passthrough (chafa-passthrough-encoder.c): 200-byte and 1 MB packets, ESC
    doubling for tmux, one packet per byte of the end-of-sixels sequence under screen."""
    res = parity.run_sixel_named(name, passthrough=passthrough)
    assert res.image_equal, res
    assert res.print_equal, res
    plain = parity.run_sixel_named(name)
    assert len(res.prod_bytes) > len(plain.prod_bytes)


def test_sixel_tmux_passthrough_splits_packets_at_one_megabyte(gpu):
    """A body longer than one tmux packet (1 000 000 escaped bytes)."""
    img = parity.make_image("noise", 800, 600)
    res = parity.run_sixel_pair(img, 100, 50, dither=capi.DITHER_DIFFUSION, grain=(1, 1),
                                passthrough=capi.PASSTHROUGH_TMUX)
    assert res.prod_len > 1000000, res.prod_len
    assert res.print_equal, res


@pytest.mark.parametrize("n_threads", [1, 3, 16])
@pytest.mark.parametrize("name", sorted(parity.SIXEL_SWEEP_CACHED))
def test_sixel_cached_modes_byte_identical(gpu, name, n_threads):
    """none / ordered / noise dither: the reference answers repeated near-identical colours from
    a lossy per-batch cache, so its output depends on the thread count.  The
    CUDA path replays the cache for the same chafa_set_n_threads () setting: identical palette,
    pens and text."""
    res = parity.run_sixel_named(name, n_threads=n_threads)
    assert res.palette_equal, res
    assert res.image_equal, res
    assert res.print_equal, res


def test_sixel_encoder_exact_on_reference_image(gpu, reference):
    """The band encoder alone: whatever pens differ in the cached modes, encoding is a pure
    function of the pen image -- checked by the byte-identical diffusion cases above and here on
    a wide image that spans several 64-column filter banks and long runs."""
    img = np.zeros((60, 700, 4), np.uint8)
    img[..., 3] = 255
    img[10:40, 100:650, 0] = 255
    img[20:25, :, 1] = 200
    res = parity.run_sixel_pair(img, 70, 6, dither=capi.DITHER_DIFFUSION, grain=(1, 1), cell=(10, 10))
    assert res.ok, res


# ---- committed golden vectors (no reference library needed) ------------------------------

import oracle_api  # noqa: E402


@pytest.mark.parametrize("name", [n for n in oracle_api.golden_names() if not n.startswith("sixel")])
def test_golden_symbols(gpu, name):
    g = oracle_api.load_golden(name)
    cw, ch = g["config"]["cells"]
    src = np.ascontiguousarray(g["src"])
    c = capi.Canvas(gpu, cw, ch, **g["config"]["cfg"])
    c.draw(src, src.shape[1], src.shape[0])
    assert (c.prepared_pixels() == g["pixmap"]).all()
    assert (c.cells() == g["cells"]).all()
    assert c.print() == g["printed"].tobytes()
    # and the plain-C restatement agrees with the CUDA path on the same pixmap
    cells, printed = oracle_api.run_oracle(g)
    assert (cells == c.cells()).all()


@pytest.mark.parametrize("name", oracle_api.golden_names("sixel"))
def test_golden_sixel(gpu, name):
    g = oracle_api.load_golden(name)
    cw, ch = g["config"]["cells"]
    cfg = dict(g["config"]["cfg"])
    cfg["grain"] = tuple(cfg.get("grain", (4, 4)))
    src = np.ascontiguousarray(g["src"])
    c = capi.Canvas(gpu, cw, ch, pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, **cfg)
    c.draw(src, src.shape[1], src.shape[0])
    indexed, pal, n = parity.sixel_image(c)
    assert n == int(g["n_colors"][0])
    assert (pal[:n] == g["palette"][:n]).all()
    assert (indexed == g["indexed"]).all()
    assert c.print() == g["printed"].tobytes()


# ---- user glyphs (chafa_symbol_map_add_glyph) ----------------------------------------------

def _glyph_image(rng, w, h, fmt):
    bpp = 3 if fmt in (capi.PIXEL_RGB8, capi.PIXEL_BGR8) else 4
    yy, xx = np.mgrid[0:h, 0:w]
    shape = ((xx - w / 2) ** 2 / (w / 2.2) ** 2 + (yy - h / 2) ** 2 / (h / 2.5) ** 2 < 1) ^ (xx > w * 0.6)
    img = np.zeros((h, w, bpp), np.uint8)
    noise = rng.integers(0, 40, size=(h, w), dtype=np.uint8)
    if bpp == 3:
        img[...] = (shape * 215)[..., None] + noise[..., None]
    else:
        img[..., :3] = 255
        img[..., 3] = shape * 215 + noise
        if fmt < 4:   # premultiplied: colour <= alpha
            img[..., :3] = img[..., 3:4]
        if fmt in (capi.PIXEL_ARGB8_PREMULTIPLIED, capi.PIXEL_ARGB8_UNASSOCIATED,
                   capi.PIXEL_ABGR8_PREMULTIPLIED, capi.PIXEL_ABGR8_UNASSOCIATED):
            img = np.ascontiguousarray(img[..., [3, 0, 1, 2]])
    return np.ascontiguousarray(img)


def test_glyph_import_matches_reference(gpu, reference):
    import ctypes as C
    rng = np.random.default_rng(5)
    cases = [(ord("A"), 20, 30, capi.PIXEL_RGBA8_UNASSOCIATED), (ord("B"), 8, 8, capi.PIXEL_RGBA8_PREMULTIPLIED),
             (ord("C"), 64, 64, capi.PIXEL_RGB8), (0x3042, 40, 30, capi.PIXEL_BGRA8_UNASSOCIATED),
             (ord("D"), 5, 7, capi.PIXEL_ARGB8_UNASSOCIATED), (0x30A2, 16, 8, capi.PIXEL_RGBA8_PREMULTIPLIED),
             (ord("E"), 200, 300, capi.PIXEL_ABGR8_PREMULTIPLIED), (ord("F"), 13, 9, capi.PIXEL_BGR8)]
    maps = {}
    for lib in (reference, gpu):
        sm = lib.chafa_symbol_map_new()
        lib.chafa_symbol_map_set_allow_builtin_glyphs(sm, 0)
        assert lib.chafa_symbol_map_apply_selectors(sm, b"all+wide", None)
        maps[lib] = sm
    images = {}
    for cp, w, h, fmt in cases:
        img = _glyph_image(rng, w, h, fmt)
        images[cp] = img
        for lib in (reference, gpu):
            lib.chafa_symbol_map_add_glyph(maps[lib], cp, fmt, img.ctypes.data, w, h, img.strides[0])

    def compiled(lib, wide):
        c = np.zeros(64, np.uint32)
        b0 = np.zeros(64, np.uint64)
        b1 = np.zeros(64, np.uint64)
        f = lib.ref_probe_symbol_map_compiled if lib.is_reference else lib.chafa_b200_symbol_map_get_compiled
        n = f(maps[lib], wide, 64, c.ctypes.data, b0.ctypes.data, b1.ctypes.data)
        return n, c[:n].tolist(), b0[:n].tolist(), b1[:n].tolist()

    for wide in (0, 1):
        assert compiled(gpu, wide) == compiled(reference, wide)
    assert compiled(gpu, 0)[0] == 6 and compiled(gpu, 1)[0] == 2

    # read a glyph back in a couple of formats
    for cp, fmt in ((ord("A"), capi.PIXEL_ARGB8_PREMULTIPLIED), (0x3042, capi.PIXEL_RGBA8_UNASSOCIATED)):
        got = {}
        for lib in (reference, gpu):
            px, w, h, rs = C.c_void_p(), C.c_int(), C.c_int(), C.c_int()
            assert lib.chafa_symbol_map_get_glyph(maps[lib], cp, fmt, C.byref(px), C.byref(w), C.byref(h), C.byref(rs))
            got[lib] = (w.value, h.value, rs.value, C.string_at(px.value, rs.value * h.value))
            lib.free_array(px)
        assert got[gpu] == got[reference]
    for lib in (reference, gpu):
        lib.chafa_symbol_map_unref(maps[lib])


# ---- row-staged scaler (TMA tiles): geometries, pixel orders, placements, bands ---------------------

STAGED = {
    # name: (src w, src h, cells w, h, kind, pixel type, config): all have 16-byte aligned source rows
    "c1_copy_x_1h": (1920, 1080, 240, 67, "shapes", capi.PIXEL_RGBA8_UNASSOCIATED, {}),
    "c1_alpha_mixed_tiles": (1920, 1080, 240, 67, "alpha", capi.PIXEL_BGRA8_UNASSOCIATED, {}),
    "c1_float_walk_dither": (1920, 1080, 240, 67, "shapes", capi.PIXEL_ARGB8_UNASSOCIATED,
                             dict(dither=capi.DITHER_ORDERED, mode=capi.CANVAS_MODE_INDEXED_256)),
    "v3_float_walk": (512, 4096, 64, 40, "noise", capi.PIXEL_RGBA8_UNASSOCIATED, {}),
    "alpha_2h_x_2h": (1920, 1080, 80, 24, "alpha", capi.PIXEL_RGBA8_UNASSOCIATED, {}),
    "upscale_0h": (100, 60, 40, 20, "alpha", capi.PIXEL_RGBA8_UNASSOCIATED, {}),
    "wide_upscale": (128, 64, 100, 9, "noise", capi.PIXEL_BGRA8_UNASSOCIATED, {}),
    "h2_v3": (3840, 2160, 100, 30, "shapes", capi.PIXEL_ARGB8_UNASSOCIATED, {}),
    "h1_v2_ragged_tiles": (1000, 700, 33, 11, "alpha", capi.PIXEL_ABGR8_UNASSOCIATED, {}),
    "copy_x_copy_offset": (640, 320, 80, 40, "alpha", capi.PIXEL_BGRA8_UNASSOCIATED, dict(dither=capi.DITHER_ORDERED,
                                                                                          mode=capi.CANVAS_MODE_INDEXED_256)),
    "tall_1h_x_copy": (1280, 96, 80, 12, "noise", capi.PIXEL_RGBA8_UNASSOCIATED, {}),
    "v_only": (256, 2048, 32, 30, "shapes", capi.PIXEL_RGBA8_UNASSOCIATED, {}),
    "noise_dither_256": (1280, 720, 120, 40, "shapes", capi.PIXEL_RGBA8_UNASSOCIATED,
                         dict(dither=capi.DITHER_NOISE, mode=capi.CANVAS_MODE_INDEXED_256)),
    "one_cell": (64, 64, 1, 1, "alpha", capi.PIXEL_RGBA8_UNASSOCIATED, {}),
    "one_row_of_cells": (4096, 64, 200, 1, "alpha", capi.PIXEL_RGBA8_UNASSOCIATED, {}),
}


def _reorder(img, pixel_type):
    order = {capi.PIXEL_RGBA8_UNASSOCIATED: [0, 1, 2, 3], capi.PIXEL_BGRA8_UNASSOCIATED: [2, 1, 0, 3],
             capi.PIXEL_ARGB8_UNASSOCIATED: [3, 0, 1, 2], capi.PIXEL_ABGR8_UNASSOCIATED: [3, 2, 1, 0]}[pixel_type]
    return np.ascontiguousarray(img[..., order])


@pytest.mark.parametrize("name", sorted(STAGED))
def test_scaler_row_staged(gpu, name):
    """The TMA-staged scaler against the reference: prepared pixmap, cells and text identical, and
    the launch really went to that kernel."""
    sw, sh, cw, ch, kind, ptype, cfg = STAGED[name]
    img = _reorder(parity.make_image(kind, sw, sh, seed=11), ptype)
    before = gpu.chafa_b200_staged_scale_count()
    res = parity.run_pair(img, cw, ch, pixel_type=ptype, **cfg)
    assert res.ok, res
    assert res.pixels_equal is True
    if name != "copy_x_copy_offset":      # 1:1 placements have a kernel of their own
        assert gpu.chafa_b200_staged_scale_count() > before


def test_scaler_row_staged_placement_with_borders(gpu, reference):
    """tuck = FIT leaves borders: tiles that straddle the placement, rows above and below it."""
    img = parity.make_image("alpha", 1600, 400, seed=5)
    out = {}
    before = gpu.chafa_b200_staged_scale_count()
    for lib in (reference, gpu):
        frame = lib.chafa_frame_new(img.ctypes.data, capi.PIXEL_RGBA8_UNASSOCIATED, 1600, 400, 6400)
        image = lib.chafa_image_new()
        lib.chafa_image_set_frame(image, frame)
        for halign, valign in ((capi.ALIGN_CENTER, capi.ALIGN_CENTER), (capi.ALIGN_END, capi.ALIGN_START)):
            pl = lib.chafa_placement_new(image, 1)
            lib.chafa_placement_set_tuck(pl, capi.TUCK_FIT)
            lib.chafa_placement_set_halign(pl, halign)
            lib.chafa_placement_set_valign(pl, valign)
            c = capi.Canvas(lib, 60, 30)
            lib.chafa_canvas_set_placement(c.handle, pl)
            out.setdefault(lib.is_reference, []).append(c.print())
            c.close()
            lib.chafa_placement_unref(pl)
        lib.chafa_image_unref(image)
        lib.chafa_frame_unref(frame)
    assert out[False] == out[True]
    assert gpu.chafa_b200_staged_scale_count() > before


def test_scaler_row_staged_bands(gpu):
    """Row bands through the staged scaler: every band equals its rows of the full render."""
    img = parity.make_image("shapes", 1920, 1080, seed=3)
    full = capi.Canvas(gpu, 240, 67)
    full.draw(img, 1920, 1080)
    want = full.print()
    full.close()
    parts = []
    for first, n in ((0, 1), (1, 30), (31, 29), (60, 7)):
        c = capi.Canvas(gpu, 240, 67)
        gpu.chafa_b200_canvas_set_band(c.handle, first, n)
        c.draw(img, 1920, 1080)
        parts.append(c.print())
        c.close()
    assert b"".join(parts) == want


def _font_sized_map(n_narrow, n_wide, seed):
    """A symbol-map builder with thousands of imported glyphs (a whole font's worth): random
    8x8 / 16x8 coverage, narrow ones in supplementary private use plane A, wide ones on CJK
    ideographs.  The same glyph images go to whichever library asks."""
    rng = np.random.default_rng(seed)
    narrow = rng.integers(0, 2, size=(n_narrow, 8, 8), dtype=np.uint8)
    wide = rng.integers(0, 2, size=(n_wide, 8, 16), dtype=np.uint8)
    # smooth a little so that the glyphs are not all at Hamming distance ~32 from everything
    narrow[:, :, :4] = narrow[:, :, :1]
    wide[:, 4:, :] = wide[:, 3:4, :]

    def build(lib):
        sm = lib.chafa_symbol_map_new()
        lib.chafa_symbol_map_set_allow_builtin_glyphs(sm, 0)
        sel = f"0xf0000..0x{0xf0000 + n_narrow - 1:x}"
        if n_wide:
            sel += f"+0x4e00..0x{0x4e00 + n_wide - 1:x}"
        assert lib.chafa_symbol_map_apply_selectors(sm, sel.encode(), None)
        for cp0, cov in ((0xf0000, narrow), (0x4e00, wide)):
            for i in range(cov.shape[0]):
                img = np.zeros(cov.shape[1:] + (4,), np.uint8)
                img[..., :3] = 255
                img[..., 3] = cov[i] * 255
                lib.chafa_symbol_map_add_glyph(sm, cp0 + i, capi.PIXEL_RGBA8_UNASSOCIATED, img.ctypes.data,
                                               img.shape[1], img.shape[0], img.strides[0])
        return sm
    return build


@pytest.mark.parametrize("cfg", [
    dict(),                                             # K3 table variant (too many symbols for K3t)
    dict(fg_only=True),
    dict(mode=capi.CANVAS_MODE_INDEXED_16),
    dict(mode=capi.CANVAS_MODE_INDEXED_8),
    dict(extractor=capi.EXTRACTOR_MEDIAN),
    dict(work=1.0),                                     # K3x with the tables read in place
], ids=["default", "fg_only", "idx16", "idx8", "median", "work9"])
def test_font_sized_symbol_table(gpu, reference, cfg):
    """the scalar matchers' shared memory grew with the symbol count and the launch
    failed beyond a few thousand symbols.  6000 narrow + 2500 wide imported glyphs: tables are
    read in place and every variant still matches the reference cell for cell."""
    img = parity.make_image("shapes", 192, 96, seed=77)
    res = parity.run_pair(img, 24, 12, symbols=_font_sized_map(6000, 2500, 9), **cfg)
    assert res.ok, res
    assert len(set(res.cells_prod[..., 0].ravel().tolist())) > 8


@pytest.mark.parametrize("mode", [capi.CANVAS_MODE_TRUECOLOR, capi.CANVAS_MODE_INDEXED_256, capi.CANVAS_MODE_INDEXED_16,
                                  capi.CANVAS_MODE_FGBG, capi.CANVAS_MODE_FGBG_BGFG])
@pytest.mark.parametrize("kind", ["shapes", "alpha"])
def test_repeat_char_compression(gpu, mode, kind):
    """A caller-supplied term-info with REPEAT_CHAR: runs of equal characters are compressed
    exactly as the reference does (chafa-canvas-printer.c:56-111)."""
    S = capi.SEQ
    seqs = {S["RESET_ATTRIBUTES"]: "\x1b[0m", S["INVERT_COLORS"]: "\x1b[7m", S["REPEAT_CHAR"]: "\x1b[%1b",
            S["SET_COLOR_FG_DIRECT"]: "\x1b[38;2;%1;%2;%3m", S["SET_COLOR_BG_DIRECT"]: "\x1b[48;2;%1;%2;%3m",
            S["SET_COLOR_FGBG_DIRECT"]: "\x1b[38;2;%1;%2;%3;48;2;%4;%5;%6m",
            S["SET_COLOR_FG_256"]: "\x1b[38;5;%1m", S["SET_COLOR_BG_256"]: "\x1b[48;5;%1m",
            S["SET_COLOR_FGBG_256"]: "\x1b[38;5;%1;48;5;%2m", S["SET_COLOR_FG_16"]: "\x1b[%1m",
            S["SET_COLOR_BG_16"]: "\x1b[%1m", S["SET_COLOR_FGBG_16"]: "\x1b[%1;%2m"}
    img = parity.make_image(kind, 512, 96, seed=31)
    img[20:60, 40:400] = (10, 200, 30, 255) if kind == "shapes" else (0, 0, 0, 0)     # long flat runs
    res = parity.run_pair(np.ascontiguousarray(img), 64, 12, mode=mode, term_info_seqs=seqs, symbols="all+wide")
    assert res.ok, res
    assert b"b" in res.prod_bytes


def test_sixel_bins_beyond_float_exact_range(gpu):
    """Large flat areas put > 65793 samples of one colour into one bin: the reference's float
    sums leave the exact range (2^24) and round in sample order.  The CUDA path re-accumulates
    such bins in float, in order: palette and text still identical."""
    img = np.zeros((1080, 1920, 4), np.uint8)
    img[...] = (201, 77, 33, 255)
    img[300:700, 200:1500] = (255, 253, 251, 255)
    img[::7, ::5, 1] = 99
    res = parity.run_sixel_pair(img, 240, 135, dither=capi.DITHER_NOISE, grain=(1, 1), work=1.0)
    assert res.ok, res


def test_sixel_palette_from_sparse_samples(gpu):
    """Fewer than 256 opaque samples at the regular sampling step: the palette is rebuilt from
    every pixel (chafa-palette.c:561-569) -- decided on the device."""
    img = parity.make_image("noise", 640, 400, seed=41)
    img[..., 3] = 0
    img[100:112, 200:230, 3] = 255          # 360 opaque pixels
    res = parity.run_sixel_pair(np.ascontiguousarray(img), 80, 50, dither=capi.DITHER_DIFFUSION, grain=(1, 1))
    assert res.ok, res
    img[..., 3] = 0                         # nothing opaque at all: no generated colours
    res = parity.run_sixel_pair(np.ascontiguousarray(img), 80, 50, dither=capi.DITHER_DIFFUSION, grain=(1, 1))
    assert res.ok, res


@pytest.mark.parametrize("work,kind", [(1.0, "noise"), (0.65, "noise"), (0.0, "noise"), (1.0, "shapes"), (0.25, "alpha")])
def test_sixel_palette_cube_resolutions(gpu, work, kind):
    """3, 4 and 5 bits per channel (512 / 4096 / 32768 colour-cube bins).  On the noise image at 5
    bits every bin is occupied: the merge runs from global memory instead of shared memory, with
    flat channel weights and unquantised counts (more than 33 bins per colour), 32 513 merges."""
    img = parity.make_image(kind, 1024, 768, seed=43)
    res = parity.run_sixel_pair(img, 128, 96, n_threads=-1, dither=capi.DITHER_DIFFUSION, grain=(1, 1), work=work)
    assert res.palette_equal, res
    assert res.image_equal, res
    assert res.print_equal, res


# ---- randomised configurations ---------------------------------------------------------------

def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    pick = lambda xs: xs[int(rng.integers(0, len(xs)))]      # noqa: E731
    mode = pick([capi.CANVAS_MODE_TRUECOLOR, capi.CANVAS_MODE_INDEXED_256, capi.CANVAS_MODE_INDEXED_240,
                 capi.CANVAS_MODE_INDEXED_16, capi.CANVAS_MODE_INDEXED_8, capi.CANVAS_MODE_INDEXED_16_8,
                 capi.CANVAS_MODE_FGBG, capi.CANVAS_MODE_FGBG_BGFG])
    cfg = dict(
        mode=mode,
        symbols=pick(["block+border+space", "all", "all+wide", "ascii", "braille+space", "half+quad+space",
                      "sextant+wedge", "wide+space", "[ a]", "alpha+digit+space-inverted", "octant+solid"]),
        fill=pick([None, None, "all", "block+space", "ascii"]),
        work=float(pick([0.0, 0.1, 0.25, 0.4, 0.5, 0.6, 0.75, 0.8, 1.0])),
        dither=pick([capi.DITHER_NONE, capi.DITHER_ORDERED, capi.DITHER_DIFFUSION, capi.DITHER_NOISE]),
        grain=(pick([1, 2, 4, 8]), pick([1, 2, 4, 8])),
        dither_intensity=float(pick([0.3, 1.0, 1.0, 2.5])),
        cell=(int(rng.integers(4, 14)), int(rng.integers(6, 24))),
        fg=int(rng.integers(0, 1 << 24)), bg=int(rng.integers(0, 1 << 24)),
        threshold=float(pick([0.0, 0.3, 0.5, 0.5, 0.9, 1.0])),
        preprocessing=bool(rng.integers(0, 2)),
        fg_only=bool(rng.integers(0, 4) == 0),
        optimizations=pick([None, None, 0, capi.OPTIMIZATION_REUSE_ATTRIBUTES]),
        extractor=pick([capi.EXTRACTOR_AVERAGE, capi.EXTRACTOR_AVERAGE, capi.EXTRACTOR_MEDIAN]),
    )
    cw, ch = int(rng.integers(1, 40)), int(rng.integers(1, 16))
    w, h = int(rng.integers(1, 400)), int(rng.integers(1, 300))
    kind = pick(["noise", "shapes", "alpha"])
    placement = dict(tuck=pick([capi.TUCK_STRETCH, capi.TUCK_FIT, capi.TUCK_SHRINK_TO_FIT]),
                     halign=pick([capi.ALIGN_START, capi.ALIGN_CENTER, capi.ALIGN_END]),
                     valign=pick([capi.ALIGN_START, capi.ALIGN_CENTER, capi.ALIGN_END])) if rng.integers(0, 2) else None
    return kind, w, h, cw, ch, cfg, placement


def _draw_with_placement(lib, canvas, img, placement):
    h, w = img.shape[:2]
    if placement is None:
        canvas.draw(img, w, h)
        return
    frame = lib.chafa_frame_new(img.ctypes.data, capi.PIXEL_RGBA8_UNASSOCIATED, w, h, w * 4)
    image = lib.chafa_image_new()
    lib.chafa_image_set_frame(image, frame)
    pl = lib.chafa_placement_new(image, 1)
    lib.chafa_placement_set_tuck(pl, placement["tuck"])
    lib.chafa_placement_set_halign(pl, placement["halign"])
    lib.chafa_placement_set_valign(pl, placement["valign"])
    lib.chafa_canvas_set_placement(canvas.handle, pl)
    lib.chafa_placement_unref(pl)
    lib.chafa_image_unref(image)
    lib.chafa_frame_unref(frame)


@pytest.mark.parametrize("seed", range(80))
def test_random_configuration(gpu, reference, seed):
    """Seeded random canvas configurations, geometries and placements: cells and printed bytes
    identical to the reference (DIN99d is exercised by its own tolerance tests)."""
    kind, w, h, cw, ch, cfg, placement = _random_case(seed)
    img = parity.make_image(kind, w, h, seed=seed)
    reference.chafa_set_n_threads(1)
    cr = capi.Canvas(reference, cw, ch, **cfg)
    cp = capi.Canvas(gpu, cw, ch, **cfg)
    _draw_with_placement(reference, cr, img, placement)
    _draw_with_placement(gpu, cp, img, placement)
    a, b = cr.cells(), cp.cells()
    assert (a == b).all(), (cfg, placement, (w, h, cw, ch), int((a != b).any(axis=2).sum()))
    assert cr.print() == cp.print(), (cfg, placement)
    assert cr.print_rows() == cp.print_rows()


def test_band_sharded_histogram_allreduce_two_gpus(gpu):
    """Row bands of a 16 / 8-colour / FGBG still on two GPUs with the intensity histogram
    all-reduced over NCCL equal the single-GPU render (needs >= 2 GPUs)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = parity.ROOT
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531",
                        os.path.join(root, "tools", "band_allreduce_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_argument_guards(gpu, reference):
    """Bad arguments are silent no-ops as in the reference (g_return_if_fail); an image mode
    prints nothing before the first draw; nothing crashes."""
    import ctypes as C
    img = parity.make_image("shapes", 64, 48, seed=2)
    c = capi.Canvas(gpu, 8, 4)
    before = c.print()
    gpu.chafa_canvas_draw_all_pixels(c.handle, capi.PIXEL_RGBA8_UNASSOCIATED, None, 64, 48, 256)     # NULL source
    gpu.chafa_canvas_draw_all_pixels(c.handle, 99, img.ctypes.data, 64, 48, 256)                     # bad pixel type
    gpu.chafa_canvas_draw_all_pixels(c.handle, capi.PIXEL_RGBA8_UNASSOCIATED, img.ctypes.data, 0, 0, 0)  # empty
    gpu.chafa_canvas_draw_all_pixels(c.handle, capi.PIXEL_RGBA8_UNASSOCIATED, img.ctypes.data, -1, 48, 256)
    assert c.print() == before
    assert gpu.chafa_canvas_get_char_at(c.handle, 100, 0) == 0
    assert gpu.chafa_canvas_set_char_at(c.handle, 7, 0, 0x3042) == 0          # wide char does not fit in the last column

    cfg = gpu.chafa_canvas_config_new()
    gpu.chafa_canvas_config_set_geometry(cfg, 0, 10)       # rejected: geometry keeps its default
    w, h = C.c_int(), C.c_int()
    gpu.lib.chafa_canvas_config_get_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    gpu.lib.chafa_canvas_config_get_geometry(cfg, C.byref(w), C.byref(h))
    rcfg = reference.chafa_canvas_config_new()
    reference.chafa_canvas_config_set_geometry(rcfg, 0, 10)
    rw, rh = C.c_int(), C.c_int()
    reference.lib.chafa_canvas_config_get_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    reference.lib.chafa_canvas_config_get_geometry(rcfg, C.byref(rw), C.byref(rh))
    assert (w.value, h.value) == (rw.value, rh.value)
    gpu.chafa_canvas_config_unref(cfg)
    reference.chafa_canvas_config_unref(rcfg)

    kitty = capi.Canvas(gpu, 8, 4, pixel_mode=capi.PIXEL_MODE_KITTY)
    rkitty = capi.Canvas(reference, 8, 4, pixel_mode=capi.PIXEL_MODE_KITTY)
    assert kitty.print() == rkitty.print() == b""

    big = capi.Canvas(gpu, 1500, 700)
    one = np.array([[[9, 200, 30, 255]]], np.uint8)
    big.draw(one, 1, 1)
    rbig = capi.Canvas(reference, 1500, 700)
    rbig.draw(one, 1, 1)
    assert big.print() == rbig.print()


# ---- Kitty and iTerm2 pixel modes -----------------------------------------------------------

def _image_text_pair(reference, gpu, img, cw, ch, *, pixel_type=capi.PIXEL_RGBA8_UNASSOCIATED, placement=None,
                     placement_id=1, **cfg):
    """Printed bytes of the same draw on both libraries."""
    cfg.setdefault("symbols", None)
    out = []
    for lib in (reference, gpu):
        c = capi.Canvas(lib, cw, ch, **cfg)
        try:
            if placement is None:
                h, w = img.shape[:2]
                c.draw(img, w, h, pixel_type=pixel_type)
            else:
                h, w = img.shape[:2]
                frame = lib.chafa_frame_new(img.ctypes.data, pixel_type, w, h, img.strides[0])
                image = lib.chafa_image_new()
                lib.chafa_image_set_frame(image, frame)
                pl = lib.chafa_placement_new(image, placement_id)
                lib.chafa_placement_set_tuck(pl, placement["tuck"])
                lib.chafa_placement_set_halign(pl, placement["halign"])
                lib.chafa_placement_set_valign(pl, placement["valign"])
                lib.chafa_canvas_set_placement(c.handle, pl)
                lib.chafa_placement_unref(pl)
                lib.chafa_image_unref(image)
                lib.chafa_frame_unref(frame)
            out.append(c.print())
        finally:
            c.close()
    return out


def _first_difference(a, b):
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] != b[i]:
            return i, a[max(0, i - 24):i + 24], b[max(0, i - 24):i + 24]
    return n, a[n - 24:n + 24], b[n - 24:n + 24]


IMAGE_MODES = [capi.PIXEL_MODE_KITTY, capi.PIXEL_MODE_ITERM2]

# name -> (image kind, src w, src h, canvas w, canvas h, config)
IMAGE_TEXT_SWEEP = {
    "down_alpha": ("alpha", 320, 200, 30, 12, dict()),
    "down_shapes": ("shapes", 333, 211, 25, 11, dict()),
    "up_noise": ("noise", 40, 30, 20, 9, dict()),
    "one_to_one": ("alpha", 160, 96, 20, 12, dict()),                     # same format, same size: bytes pass through
    "flatten": ("alpha", 320, 200, 30, 12, dict(threshold=1.0, bg=0x204060)),
    "flatten_one_to_one": ("alpha", 160, 96, 20, 12, dict(threshold=1.0, bg=0xC08040)),
    "cell_10x20": ("alpha", 300, 300, 17, 7, dict(cell=(10, 20))),
    "odd_size": ("shapes", 127, 61, 7, 3, dict(cell=(9, 17))),            # pixmap bytes not a multiple of 3
    "huge_ratio": ("noise", 2000, 9, 1, 1, dict(cell=(3, 2))),
}


@pytest.mark.parametrize("pixel_mode", IMAGE_MODES)
@pytest.mark.parametrize("name", sorted(IMAGE_TEXT_SWEEP))
def test_image_text_byte_identical(gpu, reference, name, pixel_mode):
    """Kitty immediate images and iTerm2 inline TIFFs: scaled pixmap, base64 and framing
    identical to the reference (chafa-kitty-renderer.c, chafa-iterm2-renderer.c)."""
    kind, w, h, cw, ch, cfg = IMAGE_TEXT_SWEEP[name]
    img = parity.make_image(kind, w, h)
    a, b = _image_text_pair(reference, gpu, img, cw, ch, pixel_mode=pixel_mode, **cfg)
    assert len(a) > 60
    assert a == b, _first_difference(a, b)


@pytest.mark.parametrize("pixel_mode", IMAGE_MODES)
@pytest.mark.parametrize("pixel_type", [capi.PIXEL_RGBA8_PREMULTIPLIED, capi.PIXEL_BGRA8_UNASSOCIATED,
                                        capi.PIXEL_ARGB8_PREMULTIPLIED, capi.PIXEL_RGB8])
@pytest.mark.parametrize("flatten", [False, True])
def test_image_text_pixel_formats(gpu, reference, pixel_mode, pixel_type, flatten):
    img = parity.make_image("alpha", 200, 120)
    if pixel_type == capi.PIXEL_RGB8:
        img = np.ascontiguousarray(img[..., :3])
    elif pixel_type < 4:   # premultiplied formats need colour <= alpha
        a = img[..., 3:4].astype(np.uint16)
        img = np.concatenate([(img[..., :3].astype(np.uint16) * a // 255).astype(np.uint8), img[..., 3:4]], axis=2)
        if pixel_type == capi.PIXEL_ARGB8_PREMULTIPLIED:
            img = img[..., [3, 0, 1, 2]]
        img = np.ascontiguousarray(img)
    cfg = dict(threshold=1.0, bg=0x336699) if flatten else dict()
    a, b = _image_text_pair(reference, gpu, img, 20, 9, pixel_mode=pixel_mode, pixel_type=pixel_type, **cfg)
    assert a == b, _first_difference(a, b)


@pytest.mark.parametrize("passthrough", [capi.PASSTHROUGH_SCREEN, capi.PASSTHROUGH_TMUX])
@pytest.mark.parametrize("placement_id", [1, 77, 300])
def test_kitty_virtual_image_under_passthrough(gpu, reference, passthrough, placement_id):
    """Under screen / tmux the image is transmitted as a virtual one in passthrough packets
    and shown through unicode placeholders carrying the row and column as combining marks."""
    img = parity.make_image("alpha", 200, 120)
    pl = dict(tuck=capi.TUCK_FIT, halign=capi.ALIGN_CENTER, valign=capi.ALIGN_END)
    a, b = _image_text_pair(reference, gpu, img, 40, 37, pixel_mode=capi.PIXEL_MODE_KITTY, cell=(4, 4),
                            passthrough=passthrough, placement=pl, placement_id=placement_id)
    assert "\U0010eeee".encode() in b
    assert a == b, _first_difference(a, b)


@pytest.mark.parametrize("pixel_mode", IMAGE_MODES)
def test_image_text_with_placement(gpu, reference, pixel_mode):
    img = parity.make_image("alpha", 300, 100)
    for tuck in (capi.TUCK_FIT, capi.TUCK_SHRINK_TO_FIT, capi.TUCK_STRETCH):
        pl = dict(tuck=tuck, halign=capi.ALIGN_CENTER, valign=capi.ALIGN_CENTER)
        a, b = _image_text_pair(reference, gpu, img, 24, 12, pixel_mode=pixel_mode, placement=pl,
                                threshold=1.0, bg=0x102030)
        assert a == b, (tuck, _first_difference(a, b))


def test_kitty_1080p_frame(gpu, reference):
    """A full 1080p pixmap (8.3 MB -> 11 MB of text) from a 4K source."""
    img = parity.make_image("shapes", 3840, 2160)
    a, b = _image_text_pair(reference, gpu, img, 240, 135, pixel_mode=capi.PIXEL_MODE_KITTY)
    assert len(b) > 11000000
    assert a == b, _first_difference(a, b)
