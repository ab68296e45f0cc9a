"""Shared parity machinery: run one configuration through the product library (CUDA) and
through the checker (oracle/_ref, the compiled reference) and compare the prepared pixmap,
the cell records and the printed bytes."""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from chafa_b200 import capi, synth  # noqa: E402

M = capi


def make_image(kind, w, h, seed=1):
    return synth.KINDS[kind](w, h, seed)


class Result:
    def __init__(self):
        self.pixels_equal = None
        self.pixel_mismatches = 0
        self.cells_equal = None
        self.cell_mismatches = 0
        self.symbol_mismatches = 0
        self.n_cells = 0
        self.print_equal = None
        self.ref_len = 0
        self.prod_len = 0
        self.first_diff = None

    @property
    def ok(self):
        return bool(self.cells_equal and self.print_equal and self.pixels_equal is not False)

    def __repr__(self):
        return (f"pixels={self.pixels_equal}({self.pixel_mismatches}) cells={self.cells_equal}"
                f"({self.cell_mismatches}/{self.n_cells}, sym {self.symbol_mismatches}) "
                f"print={self.print_equal} (ref {self.ref_len} B, prod {self.prod_len} B, first diff {self.first_diff})")


def run_pair(img, cw, ch, *, n_threads=1, check_pixels=True, pixel_type=capi.PIXEL_RGBA8_UNASSOCIATED,
             term_info_seqs=None, **cfg):
    """Render @img on a cw x ch canvas with both libraries."""
    ref = capi.load_reference()
    prod = capi.load_product()
    h, w = img.shape[0], img.shape[1]
    ref.chafa_set_n_threads(n_threads)
    res = Result()

    cr = capi.Canvas(ref, cw, ch, **cfg)
    cp = capi.Canvas(prod, cw, ch, **cfg)
    try:
        cr.draw(img, w, h, pixel_type=pixel_type)
        cp.draw(img, w, h, pixel_type=pixel_type)

        if check_pixels and cfg.get("pixel_mode", capi.PIXEL_MODE_SYMBOLS) == capi.PIXEL_MODE_SYMBOLS:
            pr = cr.prepared_pixels(img, w, h, pixel_type)
            pp = cp.prepared_pixels()
            res.pixel_mismatches = int((pr != pp).sum())
            res.pixels_equal = res.pixel_mismatches == 0

        if cfg.get("pixel_mode", capi.PIXEL_MODE_SYMBOLS) == capi.PIXEL_MODE_SYMBOLS:
            a, b = cr.cells(), cp.cells()
            res.n_cells = a.shape[0] * a.shape[1]
            diff = (a != b).any(axis=2)
            res.cell_mismatches = int(diff.sum())
            res.symbol_mismatches = int((a[..., 0] != b[..., 0]).sum())
            res.cells_equal = res.cell_mismatches == 0
            res.cells_ref, res.cells_prod = a, b
        else:
            res.cells_equal = True

        tr = tp = None
        if term_info_seqs:
            tr, tp = ref.chafa_term_info_new(), prod.chafa_term_info_new()
            for seq, s in term_info_seqs.items():
                ref.chafa_term_info_set_seq(tr, seq, s.encode(), None)
                prod.chafa_term_info_set_seq(tp, seq, s.encode(), None)
        sr, sp = cr.print(tr), cp.print(tp)
        if tr:
            ref.chafa_term_info_unref(tr)
            prod.chafa_term_info_unref(tp)
        res.ref_len, res.prod_len = len(sr), len(sp)
        res.print_equal = sr == sp
        res.ref_bytes, res.prod_bytes = sr, sp
        if not res.print_equal:
            n = min(len(sr), len(sp))
            d = next((i for i in range(n) if sr[i] != sp[i]), n)
            res.first_diff = d
    finally:
        cr.close()
        cp.close()
    return res


# name -> (image kind, src w, src h, canvas w, canvas h, config)
SWEEP = {
    "truecolor_block_1to1": ("shapes", 320, 160, 40, 20, dict()),
    "truecolor_noise": ("noise", 256, 128, 32, 16, dict()),
    "truecolor_alpha": ("alpha", 256, 128, 32, 16, dict()),
    "truecolor_c1_geometry": ("shapes", 480, 270, 60, 17, dict(work=0.5)),
    "truecolor_all": ("shapes", 256, 128, 32, 16, dict(symbols="all")),
    "truecolor_all_wide": ("shapes", 256, 128, 32, 16, dict(symbols="all+wide")),
    "truecolor_wide_noise": ("noise", 256, 128, 32, 16, dict(symbols="all+wide")),
    "truecolor_upscale": ("shapes", 100, 60, 40, 20, dict()),
    "truecolor_down3x": ("noise", 960, 480, 40, 20, dict()),
    "truecolor_down20x_box": ("noise", 2000, 1200, 12, 7, dict()),
    "truecolor_work1_nearest": ("shapes", 300, 200, 30, 12, dict(work=0.1)),
    "truecolor_work9_exhaustive": ("shapes", 128, 64, 16, 8, dict(work=1.0, symbols="all")),
    "truecolor_fill": ("shapes", 256, 128, 32, 16, dict(fill="block+space")),
    "truecolor_fgonly": ("alpha", 256, 128, 32, 16, dict(fg_only=True)),
    "truecolor_noreuse": ("alpha", 256, 128, 32, 16, dict(optimizations=0)),
    "idx256": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_256)),
    "idx256_ordered_allwide": ("shapes", 256, 128, 32, 16,
                               dict(mode=capi.CANVAS_MODE_INDEXED_256, symbols="all+wide",
                                    dither=capi.DITHER_ORDERED)),
    "idx256_noise_dither": ("shapes", 256, 128, 32, 16,
                            dict(mode=capi.CANVAS_MODE_INDEXED_256, dither=capi.DITHER_NOISE)),
    "idx256_fs": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_256, dither=capi.DITHER_DIFFUSION)),
    "idx240_alpha": ("alpha", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_240)),
    "idx16": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_16)),
    "idx16_fs": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_16, dither=capi.DITHER_DIFFUSION)),
    "idx8": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_8)),
    "idx16_8": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_16_8)),
    "idx16_8_alpha": ("alpha", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_16_8)),
    "fgbg": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_FGBG)),
    "fgbg_bgfg": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_FGBG_BGFG)),
    "fgbg_bgfg_ascii": ("noise", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_FGBG_BGFG, symbols="ascii")),
    "fgbg_ordered": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_FGBG, dither=capi.DITHER_ORDERED)),
    "idx256_fill_alpha": ("alpha", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_256, fill="all")),
    "idx16_fill": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_16, fill="block+space+ascii")),
    "idx256_nopreproc": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_16, preprocessing=False)),
    "truecolor_braille": ("noise", 256, 128, 32, 16, dict(symbols="braille")),
    "truecolor_bg_white": ("alpha", 256, 128, 32, 16, dict(bg=0xFFFFFF, fg=0x000000)),
    "truecolor_odd_cells": ("shapes", 333, 177, 37, 11, dict(cell=(10, 20))),
    "truecolor_median": ("shapes", 256, 128, 32, 16, dict(extractor=capi.EXTRACTOR_MEDIAN)),
    "truecolor_median_noise_wide": ("noise", 256, 128, 32, 16, dict(extractor=capi.EXTRACTOR_MEDIAN,
                                                                     symbols="all+wide")),
    "idx256_median_alpha": ("alpha", 256, 128, 32, 16, dict(extractor=capi.EXTRACTOR_MEDIAN,
                                                            mode=capi.CANVAS_MODE_INDEXED_256)),
    "truecolor_median_exhaustive": ("shapes", 128, 64, 16, 8, dict(extractor=capi.EXTRACTOR_MEDIAN, work=1.0,
                                                                   symbols="block+border+space")),
    "idx16_8_wide": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_16_8, symbols="all+wide")),
    "idx8_fgonly": ("alpha", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_8, fg_only=True)),
    "idx256_fgonly_fill": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_256, fg_only=True,
                                                           fill="all")),
    "truecolor_threshold": ("alpha", 256, 128, 32, 16, dict(threshold=0.8)),
    "truecolor_placement_big_canvas": ("shapes", 100, 100, 40, 30, dict(cell=(8, 16))),
}

# DIN99d goes through float transcendental code: tolerance cases, not byte-identity
SWEEP_DIN99D = {
    "din99d_256": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_256,
                                                     color_space=capi.COLOR_SPACE_DIN99D)),
    "din99d_240_exhaustive": ("noise", 128, 64, 16, 8, dict(mode=capi.CANVAS_MODE_INDEXED_240, work=1.0,
                                                            symbols="all", color_space=capi.COLOR_SPACE_DIN99D)),
    "din99d_16": ("shapes", 256, 128, 32, 16, dict(mode=capi.CANVAS_MODE_INDEXED_16,
                                                    color_space=capi.COLOR_SPACE_DIN99D)),
}


class SixelResult:
    def __init__(self):
        self.palette_equal = self.image_equal = self.print_equal = None
        self.n_colors = (0, 0)
        self.pixel_mismatches = 0
        self.n_pixels = 0
        self.ref_len = self.prod_len = 0

    @property
    def ok(self):
        return bool(self.palette_equal and self.image_equal and self.print_equal)

    def __repr__(self):
        return (f"palette={self.palette_equal} n_colors={self.n_colors} image={self.image_equal}"
                f"({self.pixel_mismatches}/{self.n_pixels}) print={self.print_equal} "
                f"(ref {self.ref_len} B, prod {self.prod_len} B)")


def sixel_image(canvas):
    """(indexed image as (H, W) uint8, palette as 256 uint32, n_colors) of the last sixel draw."""
    import ctypes as C
    lib = canvas.lib
    pal = np.zeros(256, np.uint32)
    w, h, n = C.c_int(), C.c_int(), C.c_int()
    if lib.is_reference:
        ptr = C.c_void_p()
        ok = lib.ref_probe_sixel_image(canvas.handle, C.byref(w), C.byref(h), C.byref(ptr), C.byref(n),
                                       pal.ctypes.data)
        assert ok
        img = np.frombuffer(C.string_at(ptr.value, w.value * h.value), np.uint8).reshape(h.value, w.value).copy()
    else:
        ok = lib.chafa_b200_canvas_get_sixel_image(canvas.handle, None, C.byref(w), C.byref(h), C.byref(n),
                                                   pal.ctypes.data)
        assert ok
        img = np.zeros((h.value, w.value), np.uint8)
        assert lib.chafa_b200_canvas_get_sixel_image(canvas.handle, img.ctypes.data, None, None, None, None)
    return img, pal, n.value


def run_sixel_pair(img, cw, ch, *, n_threads=1, **cfg):
    ref = capi.load_reference()
    prod = capi.load_product()
    h, w = img.shape[0], img.shape[1]
    ref.chafa_set_n_threads(n_threads)
    prod.chafa_set_n_threads(n_threads)     # only selects the batch split the lookup cache is replayed for
    cfg = dict(cfg)
    cfg.setdefault("pixel_mode", capi.PIXEL_MODE_SIXELS)
    cfg.setdefault("symbols", None)
    res = SixelResult()
    cr = capi.Canvas(ref, cw, ch, **cfg)
    cp = capi.Canvas(prod, cw, ch, **cfg)
    try:
        cr.draw(img, w, h)
        cp.draw(img, w, h)
        ir, pr, nr = sixel_image(cr)
        ip, pp, npd = sixel_image(cp)
        res.n_colors = (nr, npd)
        res.palette_equal = nr == npd and bool((pr[:nr] == pp[:npd]).all())
        res.n_pixels = ir.size
        res.pixel_mismatches = int((ir != ip).sum()) if ir.shape == ip.shape else ir.size
        res.image_equal = res.pixel_mismatches == 0
        sr, sp = cr.print(), cp.print()
        res.ref_len, res.prod_len = len(sr), len(sp)
        res.print_equal = sr == sp
        res.ref_bytes, res.prod_bytes = sr, sp
        res.images = (ir, ip)
        res.palettes = (pr, pp)
    finally:
        cr.close()
        cp.close()
    return res


# name -> (image kind, src w, src h, canvas w, canvas h, config); sixel canvases are cw*cell x ch*cell pixels
SIXEL_SWEEP = {
    "sixel_fs_default": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
    "sixel_fs_noise_img": ("noise", 200, 120, 20, 9, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
    "sixel_fs_alpha": ("alpha", 320, 200, 30, 12, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
    "sixel_fs_upscale": ("shapes", 64, 40, 30, 12, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
    "sixel_fs_cell10x20": ("shapes", 300, 300, 17, 7, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1), cell=(10, 20))),
    "sixel_fs_work9": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1), work=1.0)),
    "sixel_fs_work1": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1), work=0.0)),
    "sixel_fs_intensity": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1),
                                                           dither_intensity=0.5)),
    "sixel_fs_256": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1),
                                                     mode=capi.CANVAS_MODE_INDEXED_256)),
    "sixel_fs_16": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1),
                                                    mode=capi.CANVAS_MODE_INDEXED_16)),
    "sixel_flat": ("flat", 64, 64, 8, 4, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
    "sixel_fs_din99d": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1),
                                                        color_space=capi.COLOR_SPACE_DIN99D)),
    "sixel_fs_din99d_240": ("alpha", 200, 120, 20, 9, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1),
                                                          color_space=capi.COLOR_SPACE_DIN99D,
                                                          mode=capi.CANVAS_MODE_INDEXED_240)),
}

# Without diffusion the reference caches pen lookups on the colour with its low bits masked
# (chafa-indexed-image.c:65-82), so a pixel can inherit the pen of a near-identical colour seen
# earlier in its thread batch.  The CUDA path replays that cache for the same thread setting.
SIXEL_SWEEP_CACHED = {
    "sixel_none": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_NONE)),
    "sixel_ordered": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_ORDERED, grain=(1, 1))),
    "sixel_noise": ("alpha", 320, 200, 30, 12, dict(dither=capi.DITHER_NOISE, grain=(1, 1))),
    "sixel_none_noise_img": ("noise", 400, 300, 40, 20, dict(dither=capi.DITHER_NONE)),
    "sixel_ordered_256": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_ORDERED, grain=(2, 2),
                                                          mode=capi.CANVAS_MODE_INDEXED_256)),
    "sixel_noise_din99d": ("shapes", 320, 200, 30, 12, dict(dither=capi.DITHER_NOISE, grain=(1, 1),
                                                           color_space=capi.COLOR_SPACE_DIN99D)),
}


def run_sixel_named(name, n_threads=1, **extra):
    table = SIXEL_SWEEP if name in SIXEL_SWEEP else SIXEL_SWEEP_CACHED
    kind, w, h, cw, ch, cfg = table[name]
    cfg = dict(cfg, **extra)
    if kind == "flat":
        img = np.zeros((h, w, 4), np.uint8)
        img[...] = (200, 100, 50, 255)
    else:
        img = make_image(kind, w, h)
    return run_sixel_pair(img, cw, ch, n_threads=n_threads, **cfg)


def run_named(name, table=None):
    kind, w, h, cw, ch, cfg = (table or SWEEP)[name]
    return run_pair(make_image(kind, w, h), cw, ch, **cfg)


if __name__ == "__main__":
    names = sys.argv[1:] or list(SWEEP) + list(SWEEP_DIN99D) + list(SIXEL_SWEEP) + list(SIXEL_SWEEP_CACHED)
    bad = 0
    for name in names:
        table = SWEEP if name in SWEEP else SWEEP_DIN99D
        try:
            r = run_sixel_named(name) if name.startswith("sixel") else run_named(name, table)
            print(("OK   " if r.ok else "FAIL ") + name, r, flush=True)
            bad += 0 if r.ok else 1
        except Exception as e:  # noqa: BLE001
            print("ERR  " + name, repr(e), flush=True)
            bad += 1
    print(f"{bad} failing of {len(names)}")
