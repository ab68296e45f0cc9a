"""CPU-only checks: the C-ABI library loads and exports every declared symbol, host-side
table building matches the oracle, the product refuses to run without a GPU."""

import ctypes as C
import os
import re

import numpy as np
import pytest

from chafa_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for header in ("chafa.h", "chafa_b200.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"CHAFA_API\s+[^;(]*?\b(chafa_\w+)\s*\(", text):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol(product):
    names = declared_symbols()
    assert len(names) > 100
    missing = [n for n in names if not hasattr(product.lib, n)]
    assert not missing, missing


SELECTORS = ["block+border+space-wide", "all", "all+wide", "ascii", "[ a]", "braille", "half", "quad+sextant",
             "octant", "wide", "alpha-latin", "0..fff", "legacy+wedge-inverted", "all-braille-0x2580..0x259f",
             "none", "extra+diagonal+dot+stipple", "vhalf+hhalf+solid", "narrow-ascii+digit"]


def compiled(lib, sel, wide):
    sm = lib.chafa_symbol_map_new()
    assert lib.chafa_symbol_map_apply_selectors(sm, sel.encode(), None)
    c = np.zeros(4096, np.uint32)
    b0 = np.zeros(4096, np.uint64)
    b1 = np.zeros(4096, np.uint64)
    f = lib.ref_probe_symbol_map_compiled if lib.is_reference else lib.chafa_b200_symbol_map_get_compiled
    n = f(sm, wide, 4096, c.ctypes.data, b0.ctypes.data, b1.ctypes.data)
    lib.chafa_symbol_map_unref(sm)
    return n, c[:n].tolist(), b0[:n].tolist(), b1[:n].tolist()


@pytest.mark.parametrize("sel", SELECTORS)
def test_symbol_table_order_matches_oracle(product, reference, sel):
    """Evaluation order decides every tie: the compiled table must be
    identical, element for element."""
    for wide in (0, 1):
        assert compiled(product, sel, wide) == compiled(reference, sel, wide)


def test_bad_selectors_rejected(product, reference):
    for lib in (product, reference):
        sm = lib.chafa_symbol_map_new()
        assert not lib.chafa_symbol_map_apply_selectors(sm, b"nosuchtag", None)
        assert not lib.chafa_symbol_map_apply_selectors(sm, b"block++", None)
        lib.chafa_symbol_map_unref(sm)


def test_no_cpu_fallback(product):
    """Without a CUDA device canvas creation must fail loudly, not fall back."""
    if product.chafa_b200_available():
        pytest.skip("a GPU is present")
    cfg = product.chafa_canvas_config_new()
    assert not product.chafa_canvas_new(cfg)
    assert b"no CUDA device" in product.chafa_b200_last_error()
    product.chafa_canvas_config_unref(cfg)


def test_term_info_known_answers(product, reference):
    """chafa_term_info_emit_seq formatting incl. aixterm offsets (reference tests/term-info-test.c:8-62),
    compared call by call against the oracle."""
    for lib in (product, reference):
        lib.lib.chafa_term_info_emit_seq.restype = C.c_void_p
    def emit(lib, ti, seq, *args):
        lib.lib.chafa_term_info_emit_seq.argtypes = [C.c_void_p, C.c_int] + [C.c_int] * (len(args) + 1)
        p = lib.lib.chafa_term_info_emit_seq(ti, seq, *args, -1)
        if not p:
            return None
        s = C.string_at(p)
        lib.free_array(p)
        return s
    tp = product.chafa_term_db_get_fallback_info(product.chafa_term_db_get_default())
    tr = reference.chafa_term_db_get_fallback_info(reference.chafa_term_db_get_default())
    S = capi.SEQ
    cases = [(S["RESET_ATTRIBUTES"],), (S["INVERT_COLORS"],), (S["SET_COLOR_FG_DIRECT"], 1, 2, 3),
             (S["SET_COLOR_FGBG_DIRECT"], 255, 0, 128, 7, 77, 200), (S["SET_COLOR_FG_256"], 200),
             (S["SET_COLOR_FGBG_256"], 1, 254), (S["SET_COLOR_FG_16"], 3), (S["SET_COLOR_FG_16"], 12),
             (S["SET_COLOR_BG_16"], 9), (S["SET_COLOR_FGBG_16"], 7, 8), (S["SET_COLOR_FG_8"], 5),
             (S["SET_COLOR_FGBG_8"], 1, 2), (S["CURSOR_UP"], 10), (S["CURSOR_TO_POS"], 3, 4),
             (S["REPEAT_CHAR"], 5), (S["ENABLE_BOLD"],), (S["RESET_COLOR_FG"],)]
    for case in cases:
        assert emit(product, tp, *case) == emit(reference, tr, *case), case
    assert emit(product, tp, S["SET_COLOR_FGBG_DIRECT"], 255, 0, 128, 7, 77, 200) \
        == b"\x1b[38;2;255;0;128;48;2;7;77;200m"
    # the generic emitter passes pens through; the aixterm offsets belong to the typed emitters
    # the printer uses (checked byte for byte by the GPU print parity tests)
    assert emit(product, tp, S["SET_COLOR_FG_16"], 12) == b"\x1b[12m"


def test_canvas_geometry_helper(product, reference):
    for lib in (product, reference):
        lib.lib.chafa_calc_canvas_geometry.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                                       C.c_float, C.c_int, C.c_int]
    for args in [(1920, 1080, 80, 24, 0.5, 0, 0), (100, 400, 80, -1, 0.5, 1, 0), (640, 480, -1, 30, 0.45, 0, 1),
                 (10, 10, 200, 200, 1.0, 0, 0), (3000, 10, 40, 40, 0.5, 0, 0), (0, 0, 10, 10, 0.5, 0, 0)]:
        out = []
        for lib in (product, reference):
            w, h = C.c_int(args[2]), C.c_int(args[3])
            lib.lib.chafa_calc_canvas_geometry(args[0], args[1], C.byref(w), C.byref(h), args[4], args[5], args[6])
            out.append((w.value, h.value))
        assert out[0] == out[1], (args, out)


def test_scaler_luts_round_trip_opaque_pixels():
    """The 1:1 fast path of the scale kernel copies opaque pixels through unchanged; that is
    exact because linearise -> premultiply -> composite (zero background weight) ->
    un-premultiply -> sRGB is the identity for every byte value at alpha 255 (SURVEY.md
    finding 5).  Checked here for both internal formats from the tables the kernel uses."""
    text = open(os.path.join(ROOT, "chafa_b200", "csrc", "gen_tables.inc")).read()

    def table(name):
        m = re.search(name + r"\[[^\]]*\] = \{(.*?)\};", text, re.S)
        return [int(x, 0) for x in re.findall(r"0x[0-9a-fA-F]+|\d+", m.group(1))]

    fs, ts = table("gen_from_srgb_lut"), table("gen_to_srgb_lut")
    p8, p8l, p16l = table("gen_inv_div_p8_lut"), table("gen_inv_div_p8l_lut"), table("gen_inv_div_p16l_lut")
    a = 255
    for c in range(256):
        t = ((fs[c] * (a + 1)) >> 8 & 0xffff) * (a + 1)
        assert ts[((t * p16l[a]) >> 19) & 0x7ff] == c
        u = ((c * p8[a]) >> 13) & 0xff
        t = ((((fs[u] * (a + 1)) >> 8) & 0x7ff) & 0xfff) * (a + 1) >> 8 & 0xfff
        assert ts[((t * p8l[a]) >> 11) & 0x7ff] == c


def test_all_selector_symbol_count(product, reference):
    """`all` = 1251 narrow symbols (1182 distinct bitmaps) + 181 wide, and why SURVEY.md 7.3-b's
    1252 / 1183 is one too many: the reference's glyph table (chafa-symbols.c:572-666) holds 1261
    narrow definitions; 9 carry a static AMBIGUOUS tag and are excluded by `all`; of the 1252 left,
    U+00B7 appears twice and the symbol map keeps one entry per code point
    (g_hash_table_replace, chafa-symbol-map.c:586): the later definition wins."""
    f = reference.lib.ref_probe_builtin_symbol
    f.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                  C.POINTER(C.c_uint64)]
    defs = []
    while True:
        c, t, b0, b1 = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        if not f(0, len(defs), C.byref(c), C.byref(t), C.byref(b0), C.byref(b1)):
            break
        defs.append((c.value, t.value, b0.value))
    TAG_BAD = (1 << 19) | (1 << 20)
    assert len(defs) == 1261
    good = [d for d in defs if not d[1] & TAG_BAD]
    assert sorted(d[0] for d in defs if d[1] & TAG_BAD) == [0x25AA, 0x25B6, 0x25C0, 0x25C6, 0x25E2, 0x25E3, 0x25E4,
                                                           0x25E5, 0x25FC]
    assert len(good) == 1252 and len({d[2] for d in good}) == 1183
    twice = [d for d in good if sum(1 for e in good if e[0] == d[0]) > 1]
    assert [d[0] for d in twice] == [0xB7, 0xB7]
    for lib in (reference, product):
        n, cps, b0, _ = compiled(lib, "all", 0)
        assert n == 1251 and len(set(b0)) == 1182
        assert b0[cps.index(0xB7)] == twice[1][2]        # the later definition
        assert compiled(lib, "all+wide", 1)[0] == 181


def test_sixel_diffusion_integer_form_is_exact():
    """The sixel Floyd-Steinberg chain (sixel.cu, intensity 1) computes the compensated channel as
    trunc ((160 px + 9 e) / 160) in integers.  The reference computes (gint16) (px + e * 0.9f / 16.0f)
    in float (chafa-indexed-image.c:84-93 with chafa-palette.c:1048-1050): identical for every 16-bit
    error and every channel value."""
    e = np.arange(-32768, 32768, dtype=np.int32)
    t = (e.astype(np.float32) * np.float32(0.9)) * np.float32(0.0625)
    for px in range(256):
        ref = np.trunc((np.float32(px) + t).astype(np.float32)).astype(np.int32).astype(np.int16).astype(np.int32)
        s = 160 * px + 9 * e
        q = np.where(s >= 0, s // 160, -((-s) // 160))
        assert (q == ref).all(), px


def _emitter_prototypes():
    """(name, [argument C types] or 'varargs') for every typed emitter declared in the header."""
    text = open(os.path.join(ROOT, "include", "chafa-term-seq-emit.inc")).read()
    out = []
    for m in re.finditer(r"gchar \*chafa_term_info_emit_(\w+) \(const ChafaTermInfo \*term_info, gchar \*dest(.*?)\);", text):
        tail = m.group(2)
        if "guint *args" in tail:
            out.append((m.group(1), "varargs"))
        else:
            out.append((m.group(1), re.findall(r"(guint8|guint16|guint) arg\d", tail)))
    return out


def test_typed_emitters_match_reference(product, reference):
    """Every chafa_term_info_emit_<seq> () of the public ABI (chafa-term-seq-def.h; argument
    transforms chafa-term-info.c:1699-1746), on the fallback terminal description and on one with
    caller-defined sequences: same bytes as the reference for a spread of argument values."""
    protos = _emitter_prototypes()
    assert len(protos) == len(capi.SEQ) == 157
    ctype = {"guint": C.c_uint, "guint8": C.c_uint8, "guint16": C.c_uint16}
    values = {"guint": [0, 1, 7, 42, 999, 9999], "guint8": [0, 1, 7, 8, 15, 100, 255], "guint16": [0, 0x12, 0xabcd, 0xffff]}

    def infos(lib):
        db = lib.chafa_term_db_get_default()
        fb = lib.chafa_term_db_get_fallback_info(db)
        custom = lib.chafa_term_info_new()
        for name, args in protos:
            n = 7 if args == "varargs" else len(args)
            text = f"<{name}" + "".join(f":%{i + 1}" for i in range(n if args != "varargs" else 0)) + ">"
            if args == "varargs":
                text = f"<{name}:%v>"
            assert lib.chafa_term_info_set_seq(custom, capi.SEQ[name.upper()], text.encode(), None), name
        return [fb, custom]

    ti = {lib: infos(lib) for lib in (product, reference)}
    checked = 0
    for name, args in protos:
        for which in (0, 1):
            outs = {}
            for lib in (product, reference):
                fn = getattr(lib.lib, "chafa_term_info_emit_" + name)
                fn.restype = C.c_void_p
                res = []
                if args == "varargs":
                    fn.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint), C.c_int]
                    for vals in ([1], [62, 4, 9], [0, 1, 2, 3, 4, 5]):
                        buf = C.create_string_buffer(256)
                        arr = (C.c_uint * len(vals))(*vals)
                        end = fn(ti[lib][which], buf, arr, len(vals))
                        res.append(buf.raw[:end - C.addressof(buf)])
                else:
                    fn.argtypes = [C.c_void_p, C.c_void_p] + [ctype[a] for a in args]
                    n_cases = max([len(values[a]) for a in args], default=1)
                    for k in range(n_cases):
                        buf = C.create_string_buffer(256)
                        call = [values[a][(k + i) % len(values[a])] for i, a in enumerate(args)]
                        end = fn(ti[lib][which], buf, *call)
                        res.append(buf.raw[:end - C.addressof(buf)])
                outs[lib] = res
            assert outs[product] == outs[reference], (name, which, outs[product], outs[reference])
            checked += len(outs[product])
    assert checked > 600


def test_n_actual_threads_follows_reference(product, reference):
    """chafa_get_n_actual_threads (chafa-features.c:266-282): the setting, or min(processors, 24)
    when automatic, never below 1."""
    for n in (-1, 0, 1, 3, 16, 40):
        product.chafa_set_n_threads(n)
        reference.chafa_set_n_threads(n)
        assert product.chafa_get_n_actual_threads() == reference.chafa_get_n_actual_threads(), n
    product.chafa_set_n_threads(-1)
    reference.chafa_set_n_threads(-1)


def test_split_sixel_rows_lie_on_sixel_bands():
    """Row bands of a sixel canvas start on multiples of six pixel rows whatever the cell height,
    cover the canvas exactly once and leave surplus ranks empty."""
    from chafa_b200 import shard
    for cell_h in (8, 10, 12, 16, 20, 6, 3):
        for n_rows in (1, 2, 3, 7, 45, 135, 270):
            for world in (1, 2, 3, 4, 8, 16):
                bands = shard.split_sixel_rows(n_rows, cell_h, world)
                assert len(bands) == world
                at = 0
                for first, n in bands:
                    assert first == at or n == 0
                    if n:
                        assert (first * cell_h) % 6 == 0
                        assert ((first + n) * cell_h) % 6 == 0 or first + n == n_rows
                    at += n
                assert at == n_rows
                sizes = [n for _, n in bands if n]
                if len(sizes) > 1:
                    assert max(sizes[:-1]) - min(sizes[:-1]) <= 6      # near-equal apart from the last
