"""CPU-only: the plain-C oracle restatement against the golden vectors the compiled reference
produced (tests/golden/, generator tests/golden/gen_golden.py), and live against the compiled
reference when oracle/_ref/libchafa_ref.so is present."""

import os

import numpy as np
import pytest

import oracle_api
from chafa_b200 import capi

SYMBOL_CASES = [n for n in oracle_api.golden_names() if not n.startswith("sixel") and n != "idx256_ordered"]


@pytest.mark.parametrize("name", SYMBOL_CASES)
def test_oracle_matches_golden(name):
    g = oracle_api.load_golden(name)
    cells, printed = oracle_api.run_oracle(g)
    assert (cells == g["cells"]).all(), f"{int((cells != g['cells']).any(axis=2).sum())} cells differ"
    assert printed == g["printed"].tobytes()


def test_oracle_ordered_dither_case_from_prepared_pixmap():
    """The restatement does not dither; from the reference's prepared (dithered) pixmap it must
    still reproduce the cells and the text."""
    g = oracle_api.load_golden("idx256_ordered")
    cells, printed = oracle_api.run_oracle(g)
    assert (cells == g["cells"]).all()
    assert printed == g["printed"].tobytes()


@pytest.mark.skipif(not os.path.exists(capi.REFERENCE_LIB), reason="compiled reference not built")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_matches_compiled_reference_live(seed):
    """Fresh random inputs, not stored anywhere: restatement vs the reference's own code."""
    import ctypes as C
    import parity
    ref = capi.load_reference()
    ref.chafa_set_n_threads(1)
    rng = np.random.default_rng(seed)
    cw, ch = int(rng.integers(3, 20)), int(rng.integers(2, 9))
    mode = [capi.CANVAS_MODE_TRUECOLOR, capi.CANVAS_MODE_INDEXED_256, capi.CANVAS_MODE_INDEXED_240][seed % 3]
    sel = ["all+wide", "block+border+space", "ascii+braille"][seed % 3]
    img = parity.make_image(["noise", "shapes", "alpha"][seed % 3], cw * 8, ch * 8, seed=seed + 50)
    c = capi.Canvas(ref, cw, ch, mode=mode, symbols=sel, work=[0.5, 0.1, 0.9][seed % 3])
    c.draw(img, cw * 8, ch * 8)

    def compiled(wide):
        sm = ref.chafa_symbol_map_new()
        ref.chafa_symbol_map_apply_selectors(sm, sel.encode(), None)
        cc = np.zeros(4096, np.uint32)
        b0 = np.zeros(4096, np.uint64)
        b1 = np.zeros(4096, np.uint64)
        n = ref.ref_probe_symbol_map_compiled(sm, wide, 4096, cc.ctypes.data, b0.ctypes.data, b1.ctypes.data)
        ref.chafa_symbol_map_unref(sm)
        return cc[:n].copy(), b0[:n].copy(), b1[:n].copy()

    chars, bm, _ = compiled(0)
    chars2, b2a, b2b = compiled(1)
    g = dict(config=dict(cells=[cw, ch], cfg=dict(mode=mode)), info=np.array(c.info(), np.int32),
             bitmaps=bm, chars=chars, bitmaps2=np.stack([b2a, b2b], axis=1).ravel() if len(chars2) else np.zeros(0, np.uint64),
             chars2=chars2, pixmap=c.prepared_pixels(img, cw * 8, ch * 8))
    cells, printed = oracle_api.run_oracle(g)
    assert (cells == c.cells()).all()
    assert printed == c.print()
