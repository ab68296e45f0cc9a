"""ctypes binding of the Chafa canvas C API.

The same binding class is used for two shared libraries that export the same symbols:

* ``chafa_b200/libchafa_b200.so`` -- the product: the B200-native implementation of
  ``chafa_canvas_draw_all_pixels`` / ``chafa_canvas_print`` (include/chafa.h,
  include/chafa_b200.h).
* ``oracle/_ref/libchafa_ref.so`` -- TEST INFRASTRUCTURE: the reference's own C files
  compiled by oracle/Makefile.  Only tests/, ``__graft_entry__.smoke()`` and the CPU-baseline
  legs of bench.py may load it.

Python is not the host side of the product (that is C++ above the C-ABI, see
INTEGRATION.md); this module exists so that tests and the benchmark can drive both
libraries through identical calls.
"""

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PRODUCT_LIB = os.path.join(HERE, "libchafa_b200.so")
REFERENCE_LIB = os.path.join(ROOT, "oracle", "_ref", "libchafa_ref.so")

# enums (chafa/chafa-canvas-config.h, chafa/chafa-common.h, chafa/chafa-symbol-map.h)
CANVAS_MODE_TRUECOLOR, CANVAS_MODE_INDEXED_256, CANVAS_MODE_INDEXED_240, CANVAS_MODE_INDEXED_16, \
    CANVAS_MODE_FGBG_BGFG, CANVAS_MODE_FGBG, CANVAS_MODE_INDEXED_8, CANVAS_MODE_INDEXED_16_8 = range(8)
COLOR_SPACE_RGB, COLOR_SPACE_DIN99D = 0, 1
DITHER_NONE, DITHER_ORDERED, DITHER_DIFFUSION, DITHER_NOISE = range(4)
PIXEL_MODE_SYMBOLS, PIXEL_MODE_SIXELS, PIXEL_MODE_KITTY, PIXEL_MODE_ITERM2 = range(4)
EXTRACTOR_AVERAGE, EXTRACTOR_MEDIAN = 0, 1
PASSTHROUGH_NONE, PASSTHROUGH_SCREEN, PASSTHROUGH_TMUX = 0, 1, 2
PIXEL_RGBA8_PREMULTIPLIED, PIXEL_BGRA8_PREMULTIPLIED, PIXEL_ARGB8_PREMULTIPLIED, \
    PIXEL_ABGR8_PREMULTIPLIED, PIXEL_RGBA8_UNASSOCIATED, PIXEL_BGRA8_UNASSOCIATED, \
    PIXEL_ARGB8_UNASSOCIATED, PIXEL_ABGR8_UNASSOCIATED, PIXEL_RGB8, PIXEL_BGR8 = range(10)
OPTIMIZATION_REUSE_ATTRIBUTES, OPTIMIZATION_SKIP_CELLS, OPTIMIZATION_REPEAT_CELLS = 1, 2, 4
OPTIMIZATION_NONE, OPTIMIZATION_ALL = 0, 0x7FFFFFFF
ALIGN_START, ALIGN_END, ALIGN_CENTER = range(3)
TUCK_STRETCH, TUCK_FIT, TUCK_SHRINK_TO_FIT = range(3)


def _term_seq_names():
    import re
    path = os.path.join(ROOT, "include", "chafa-term-seq-enum.inc")
    return [m.group(1) for m in re.finditer(r"^\s*CHAFA_TERM_SEQ_(\w+)\s*,", open(path).read(), re.M)]


# ChafaTermSeq enumerators in ABI order: SEQ["RESET_ATTRIBUTES"] == 2, ...
SEQ = {name: i for i, name in enumerate(_term_seq_names())}


class GString(C.Structure):
    _fields_ = [("str", C.c_void_p), ("len", C.c_size_t), ("allocated_len", C.c_size_t)]


class Cell(C.Structure):
    """Layout of ChafaCanvasCell (chafa/internal/chafa-canvas-internal.h:30-37)."""
    _fields_ = [("c", C.c_uint32), ("fg", C.c_uint32), ("bg", C.c_uint32)]


_P = C.c_void_p


class ChafaLib:
    """One loaded library (product or reference)."""

    def __init__(self, path, is_reference=False):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        self.path = path
        self.is_reference = is_reference
        self.lib = lib = C.CDLL(path)
        d = self._declare
        d("chafa_canvas_config_new", _P)
        d("chafa_canvas_config_unref", None, _P)
        d("chafa_canvas_config_set_geometry", None, _P, C.c_int, C.c_int)
        d("chafa_canvas_config_set_cell_geometry", None, _P, C.c_int, C.c_int)
        d("chafa_canvas_config_set_canvas_mode", None, _P, C.c_int)
        d("chafa_canvas_config_set_color_extractor", None, _P, C.c_int)
        d("chafa_canvas_config_set_color_space", None, _P, C.c_int)
        d("chafa_canvas_config_set_symbol_map", None, _P, _P)
        d("chafa_canvas_config_set_fill_symbol_map", None, _P, _P)
        d("chafa_canvas_config_set_transparency_threshold", None, _P, C.c_float)
        d("chafa_canvas_config_set_fg_color", None, _P, C.c_uint32)
        d("chafa_canvas_config_set_bg_color", None, _P, C.c_uint32)
        d("chafa_canvas_config_set_work_factor", None, _P, C.c_float)
        d("chafa_canvas_config_set_preprocessing_enabled", None, _P, C.c_int)
        d("chafa_canvas_config_set_dither_mode", None, _P, C.c_int)
        d("chafa_canvas_config_set_dither_grain_size", None, _P, C.c_int, C.c_int)
        d("chafa_canvas_config_set_dither_intensity", None, _P, C.c_float)
        d("chafa_canvas_config_set_pixel_mode", None, _P, C.c_int)
        d("chafa_canvas_config_set_optimizations", None, _P, C.c_int)
        d("chafa_canvas_config_set_fg_only_enabled", None, _P, C.c_int)
        d("chafa_canvas_config_set_passthrough", None, _P, C.c_int)
        d("chafa_symbol_map_new", _P)
        d("chafa_symbol_map_unref", None, _P)
        d("chafa_symbol_map_add_by_tags", None, _P, C.c_int)
        d("chafa_symbol_map_remove_by_tags", None, _P, C.c_int)
        d("chafa_symbol_map_add_by_range", None, _P, C.c_uint32, C.c_uint32)
        d("chafa_symbol_map_apply_selectors", C.c_int, _P, C.c_char_p, _P)
        d("chafa_symbol_map_add_glyph", None, _P, C.c_uint32, C.c_int, _P, C.c_int, C.c_int, C.c_int)
        d("chafa_symbol_map_get_glyph", C.c_int, _P, C.c_uint32, C.c_int, C.POINTER(_P), C.POINTER(C.c_int),
          C.POINTER(C.c_int), C.POINTER(C.c_int))
        d("chafa_symbol_map_set_allow_builtin_glyphs", None, _P, C.c_int)
        d("chafa_canvas_new", _P, _P)
        d("chafa_canvas_new_similar", _P, _P)
        d("chafa_canvas_unref", None, _P)
        d("chafa_canvas_draw_all_pixels", None, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int)
        d("chafa_canvas_print", C.POINTER(GString), _P, _P)
        d("chafa_canvas_print_rows", None, _P, _P, C.POINTER(C.POINTER(C.POINTER(GString))),
          C.POINTER(C.c_int))
        d("chafa_canvas_get_char_at", C.c_uint32, _P, C.c_int, C.c_int)
        d("chafa_canvas_set_char_at", C.c_int, _P, C.c_int, C.c_int, C.c_uint32)
        d("chafa_canvas_get_colors_at", None, _P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int))
        d("chafa_canvas_set_colors_at", None, _P, C.c_int, C.c_int, C.c_int, C.c_int)
        d("chafa_canvas_get_raw_colors_at", None, _P, C.c_int, C.c_int, C.POINTER(C.c_int),
          C.POINTER(C.c_int))
        d("chafa_canvas_set_raw_colors_at", None, _P, C.c_int, C.c_int, C.c_int, C.c_int)
        d("chafa_canvas_set_placement", None, _P, _P)
        d("chafa_frame_new", _P, _P, C.c_int, C.c_int, C.c_int, C.c_int)
        d("chafa_frame_unref", None, _P)
        d("chafa_image_new", _P)
        d("chafa_image_unref", None, _P)
        d("chafa_image_set_frame", None, _P, _P)
        d("chafa_placement_new", _P, _P, C.c_int)
        d("chafa_placement_unref", None, _P)
        d("chafa_placement_set_tuck", None, _P, C.c_int)
        d("chafa_placement_set_halign", None, _P, C.c_int)
        d("chafa_placement_set_valign", None, _P, C.c_int)
        d("chafa_term_info_new", _P)
        d("chafa_term_info_unref", None, _P)
        d("chafa_term_info_set_seq", C.c_int, _P, C.c_int, C.c_char_p, _P)
        d("chafa_term_db_get_default", _P)
        d("chafa_term_db_get_fallback_info", _P, _P)
        d("chafa_set_n_threads", None, C.c_int)
        d("chafa_get_n_actual_threads", C.c_int)
        if is_reference:
            d("g_string_free", _P, C.POINTER(GString), C.c_int)
            d("g_free", None, _P)
            d("ref_probe_canvas_cells", _P, _P)
            d("ref_probe_canvas_info", None, _P, C.POINTER(C.c_int))
            d("ref_probe_prepare_pixels", _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int,
              C.c_int, C.c_int)
            d("ref_probe_free", None, _P)
            d("ref_probe_scale", C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int,
              C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)
            d("ref_probe_symbol_map_compiled", C.c_int, _P, C.c_int, C.c_int, _P, _P, _P)
            d("ref_probe_rgb_to_din99d", C.c_uint32, C.c_uint32)
            d("ref_probe_sixel_image", C.c_int, _P, C.POINTER(C.c_int), C.POINTER(C.c_int),
              C.POINTER(_P), C.POINTER(C.c_int), _P)
        else:
            d("chafa_b200_available", C.c_int)
            d("chafa_b200_set_device", C.c_int, C.c_int)
            d("chafa_b200_last_error", C.c_char_p)
            d("chafa_b200_launch_count", C.c_uint64)
            d("chafa_b200_staged_scale_count", C.c_uint64)
            d("chafa_b200_copy_from_device", C.c_int, _P, _P, C.c_size_t)
            d("chafa_b200_string_free", None, C.POINTER(GString))
            d("chafa_b200_free", None, _P)
            d("chafa_b200_canvas_set_stream", None, _P, _P)
            d("chafa_b200_canvas_draw_all_pixels_device", None, _P, C.c_int, _P, C.c_int, C.c_int,
              C.c_int)
            d("chafa_b200_canvas_print_device", C.c_size_t, _P, _P, C.POINTER(_P))
            d("chafa_b200_canvas_get_cells", C.c_int, _P, _P)
            d("chafa_b200_canvas_get_prepared_pixels", C.c_int, _P, _P, C.POINTER(C.c_int),
              C.POINTER(C.c_int))
            d("chafa_b200_symbol_map_get_compiled", C.c_int, _P, C.c_int, C.c_int, _P, _P, _P)
            d("chafa_b200_canvas_get_info", None, _P, C.POINTER(C.c_int))
            d("chafa_b200_canvas_get_sixel_image", C.c_int, _P, _P, C.POINTER(C.c_int),
              C.POINTER(C.c_int), C.POINTER(C.c_int), _P)
            d("chafa_b200_canvas_set_band", None, _P, C.c_int, C.c_int)
            d("chafa_b200_canvas_print_into", C.c_size_t, _P, _P, _P, C.c_size_t)
            d("chafa_b200_canvas_draw_begin_device", C.c_int, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int,
              C.POINTER(_P), C.POINTER(C.c_size_t))
            d("chafa_b200_canvas_draw_finish", None, _P)
            d("chafa_b200_canvas_set_histogram_buffer", None, _P, _P)
            d("chafa_b200_canvas_draw_float_sums_device", C.c_int, _P, C.POINTER(_P), C.POINTER(C.c_size_t))
            d("chafa_b200_canvas_exchange_buffer_bytes", C.c_size_t, _P)
            d("chafa_b200_copy_to_device", C.c_int, _P, _P, C.c_size_t)
            d("chafa_b200_device_alloc", _P, C.c_size_t)
            d("chafa_b200_canvas_synchronize", C.c_int, _P)
            d("chafa_b200_canvases_draw_batch_device", C.c_int, C.POINTER(_P), C.c_int, C.c_int, C.POINTER(_P),
              C.c_int, C.c_int, C.c_int)
            d("chafa_b200_canvases_draw_batch", C.c_int, C.POINTER(_P), C.c_int, C.c_int, C.POINTER(_P),
              C.c_int, C.c_int, C.c_int)
            d("chafa_b200_device_free", None, _P)

    def _declare(self, name, restype, *argtypes):
        fn = getattr(self.lib, name)
        fn.restype = restype
        fn.argtypes = list(argtypes)

    def __getattr__(self, name):
        return getattr(self.lib, name)

    # -- helpers ---------------------------------------------------------------

    def take_string(self, gs):
        """GString* -> bytes, freeing it with the library's own deallocator."""
        if not gs:
            return b""
        data = C.string_at(gs.contents.str, gs.contents.len) if gs.contents.len else b""
        if self.is_reference:
            self.lib.g_string_free(gs, 1)
        else:
            self.lib.chafa_b200_string_free(gs)
        return data

    def take_string_len(self, gs):
        """GString* -> its length; the text stays where the library put it (host memory the caller
        owns) and is freed without a Python-side copy."""
        if not gs:
            return 0
        n = int(gs.contents.len)
        if self.is_reference:
            self.lib.g_string_free(gs, 1)
        else:
            self.lib.chafa_b200_string_free(gs)
        return n

    def free_array(self, p):
        if self.is_reference:
            self.lib.g_free(p)
        else:
            self.lib.chafa_b200_free(p)


def draw_batch_device(lib, canvases, dev_ptrs, width, height, rowstride=None, pixel_type=PIXEL_RGBA8_UNASSOCIATED,
                      on_device=True):
    """chafa_b200_canvases_draw_batch_device (or, with on_device=False, chafa_b200_canvases_draw_batch for
    host images) over Canvas objects and integer addresses."""
    n = len(canvases)
    handles = (_P * n)(*[c.handle for c in canvases])
    srcs = (_P * n)(*dev_ptrs)
    bpp = 3 if pixel_type in (PIXEL_RGB8, PIXEL_BGR8) else 4
    fn = lib.chafa_b200_canvases_draw_batch_device if on_device else lib.chafa_b200_canvases_draw_batch
    return fn(handles, n, pixel_type, srcs, width, height, rowstride or width * bpp)


_libs = {}


def load_product():
    """The B200 library.  Raises if it has not been built: there is no fallback."""
    if "product" not in _libs:
        _libs["product"] = ChafaLib(PRODUCT_LIB, False)
    return _libs["product"]


def load_reference():
    """TEST INFRASTRUCTURE: the compiled reference (oracle/_ref)."""
    if "reference" not in _libs:
        _libs["reference"] = ChafaLib(REFERENCE_LIB, True)
    return _libs["reference"]


class Canvas:
    """A configured canvas on one library; mirrors the call sequence of the reference's
    examples/example.c:30-62."""

    def __init__(self, lib, width, height, *, mode=CANVAS_MODE_TRUECOLOR, symbols="block+border+space-wide",
                 fill=None, work=0.5, color_space=COLOR_SPACE_RGB, dither=DITHER_NONE, grain=(4, 4),
                 dither_intensity=1.0, pixel_mode=PIXEL_MODE_SYMBOLS, cell=(8, 8), fg=0xFFFFFF, bg=0x000000,
                 threshold=None, preprocessing=True, fg_only=False, optimizations=None,
                 extractor=EXTRACTOR_AVERAGE, passthrough=0):
        self.lib = lib
        self.width, self.height = width, height
        cfg = lib.chafa_canvas_config_new()
        lib.chafa_canvas_config_set_geometry(cfg, width, height)
        lib.chafa_canvas_config_set_cell_geometry(cfg, cell[0], cell[1])
        lib.chafa_canvas_config_set_canvas_mode(cfg, mode)
        lib.chafa_canvas_config_set_color_space(cfg, color_space)
        lib.chafa_canvas_config_set_color_extractor(cfg, extractor)
        lib.chafa_canvas_config_set_dither_mode(cfg, dither)
        lib.chafa_canvas_config_set_dither_grain_size(cfg, grain[0], grain[1])
        lib.chafa_canvas_config_set_dither_intensity(cfg, dither_intensity)
        lib.chafa_canvas_config_set_pixel_mode(cfg, pixel_mode)
        lib.chafa_canvas_config_set_work_factor(cfg, work)
        lib.chafa_canvas_config_set_fg_color(cfg, fg)
        lib.chafa_canvas_config_set_bg_color(cfg, bg)
        lib.chafa_canvas_config_set_preprocessing_enabled(cfg, int(preprocessing))
        lib.chafa_canvas_config_set_fg_only_enabled(cfg, int(fg_only))
        lib.chafa_canvas_config_set_passthrough(cfg, passthrough)
        if threshold is not None:
            lib.chafa_canvas_config_set_transparency_threshold(cfg, threshold)
        if optimizations is not None:
            lib.chafa_canvas_config_set_optimizations(cfg, optimizations)
        if callable(symbols):
            # a builder: (lib) -> new ChafaSymbolMap (user glyphs, explicit ranges ...)
            sm = symbols(lib)
            lib.chafa_canvas_config_set_symbol_map(cfg, sm)
            lib.chafa_symbol_map_unref(sm)
        elif symbols is not None:
            sm = lib.chafa_symbol_map_new()
            if not lib.chafa_symbol_map_apply_selectors(sm, symbols.encode(), None):
                raise ValueError(f"bad symbol selectors {symbols!r}")
            lib.chafa_canvas_config_set_symbol_map(cfg, sm)
            lib.chafa_symbol_map_unref(sm)
        if fill is not None:
            fm = lib.chafa_symbol_map_new()
            if not lib.chafa_symbol_map_apply_selectors(fm, fill.encode(), None):
                raise ValueError(f"bad fill selectors {fill!r}")
            lib.chafa_canvas_config_set_fill_symbol_map(cfg, fm)
            lib.chafa_symbol_map_unref(fm)
        self.handle = lib.chafa_canvas_new(cfg)
        lib.chafa_canvas_config_unref(cfg)
        if not self.handle:
            err = b"" if lib.is_reference else lib.chafa_b200_last_error()
            raise RuntimeError(f"chafa_canvas_new failed: {err.decode(errors='replace')}")

    def close(self):
        if self.handle:
            self.lib.chafa_canvas_unref(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def draw(self, pixels, width, height, rowstride=None, pixel_type=PIXEL_RGBA8_UNASSOCIATED):
        """pixels: bytes / bytearray / numpy array / integer host address."""
        bpp = 3 if pixel_type in (PIXEL_RGB8, PIXEL_BGR8) else 4
        if rowstride is None:
            rowstride = width * bpp
        if isinstance(pixels, int):
            addr = pixels
        elif hasattr(pixels, "ctypes"):
            addr = pixels.ctypes.data
        else:
            self._keep = (C.c_char * len(pixels)).from_buffer_copy(pixels)
            addr = C.addressof(self._keep)
        self.lib.chafa_canvas_draw_all_pixels(self.handle, pixel_type, addr, width, height, rowstride)

    def draw_device(self, dev_ptr, width, height, rowstride=None, pixel_type=PIXEL_RGBA8_UNASSOCIATED):
        bpp = 3 if pixel_type in (PIXEL_RGB8, PIXEL_BGR8) else 4
        self.lib.chafa_b200_canvas_draw_all_pixels_device(self.handle, pixel_type, dev_ptr, width, height,
                                                          rowstride or width * bpp)

    def print(self, term_info=None):
        return self.lib.take_string(self.lib.chafa_canvas_print(self.handle, term_info))

    def print_len(self, term_info=None):
        """chafa_canvas_print () as a C caller uses it: the GString is produced in host memory and
        freed; only its length crosses into Python."""
        return self.lib.take_string_len(self.lib.chafa_canvas_print(self.handle, term_info))

    def print_device(self, term_info=None):
        out = _P()
        n = self.lib.chafa_b200_canvas_print_device(self.handle, term_info, C.byref(out))
        return out.value, n

    def print_device_enqueue(self, term_info=None):
        """Queue the printer and return (device address of the text, device address of its 64-bit
        length) without waiting for anything."""
        out, dlen = _P(), _P()
        f = self.lib.lib.chafa_b200_canvas_print_device_enqueue
        f.restype = C.c_int
        f.argtypes = [_P, _P, C.POINTER(_P), C.POINTER(_P)]
        if not f(self.handle, term_info, C.byref(out), C.byref(dlen)):
            raise RuntimeError("chafa_b200_canvas_print_device_enqueue failed")
        return out.value, dlen.value

    def print_rows(self, term_info=None):
        arr = C.POINTER(C.POINTER(GString))()
        n = C.c_int(0)
        self.lib.chafa_canvas_print_rows(self.handle, term_info, C.byref(arr), C.byref(n))
        rows = [self.lib.take_string(arr[i]) for i in range(n.value)]
        self.lib.free_array(C.cast(arr, _P))
        return rows

    def cells(self):
        """numpy array (height, width, 3) of uint32: c, fg, bg."""
        import numpy as np
        n = self.width * self.height
        out = np.zeros((n, 3), dtype=np.uint32)
        if self.lib.is_reference:
            p = self.lib.ref_probe_canvas_cells(self.handle)
            C.memmove(out.ctypes.data, p, n * 12)
        else:
            self.lib.chafa_b200_canvas_get_cells(self.handle, out.ctypes.data)
        return out.reshape(self.height, self.width, 3)

    def info(self):
        out = (C.c_int * 12)()
        if self.lib.is_reference:
            self.lib.ref_probe_canvas_info(self.handle, out)
        else:
            self.lib.chafa_b200_canvas_get_info(self.handle, out)
        return list(out)

    def prepared_pixels(self, src=None, src_w=0, src_h=0, pixel_type=PIXEL_RGBA8_UNASSOCIATED):
        """Canvas pixmap after scaling and preprocessing, (H, W) uint32 little-endian RGBA.
        The product returns the pixmap of the last draw; the reference probe recomputes it
        from @src."""
        import numpy as np
        info = self.info()
        w, h = info[0], info[1]
        out = np.zeros((h, w), dtype=np.uint32)
        if self.lib.is_reference:
            bpp = 3 if pixel_type in (PIXEL_RGB8, PIXEL_BGR8) else 4
            p = self.lib.ref_probe_prepare_pixels(self.handle, pixel_type, src.ctypes.data, src_w, src_h,
                                                  src_w * bpp, 0, 0, 0)
            C.memmove(out.ctypes.data, p, w * h * 4)
            self.lib.ref_probe_free(p)
        else:
            ww, hh = C.c_int(), C.c_int()
            ok = self.lib.chafa_b200_canvas_get_prepared_pixels(self.handle, out.ctypes.data, C.byref(ww),
                                                                C.byref(hh))
            if not ok:
                raise RuntimeError("no prepared pixels: draw first")
        return out
