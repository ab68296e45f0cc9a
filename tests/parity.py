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


def run_named(name, table=None):
    kind, w, h, cw, ch, cfg = (table or SWEEP)[name]
    return run_pair(make_image(kind, w, h), cw, ch, **cfg)


if __name__ == "__main__":
    names = sys.argv[1:] or list(SWEEP) + list(SWEEP_DIN99D)
    bad = 0
    for name in names:
        table = SWEEP if name in SWEEP else SWEEP_DIN99D
        try:
            r = run_named(name, table)
            print(("OK   " if r.ok else "FAIL ") + name, r, flush=True)
            bad += 0 if r.ok else 1
        except Exception as e:  # noqa: BLE001
            print("ERR  " + name, repr(e), flush=True)
            bad += 1
    print(f"{bad} failing of {len(names)}")
