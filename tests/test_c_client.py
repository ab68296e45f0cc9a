"""The drop-in as a C caller sees it: tests/c/client.c (and, where the reference tree is present,
the reference's own examples/example.c, compiled where it lies) built with gcc against
include/chafa.h and linked against libchafa_b200.so -- link only on CPU; on the GPU box the client
runs and its output is compared with the same calls made on the compiled reference."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from chafa_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "c", "build")
REF_EXAMPLE = "/root/reference/examples/example.c"


def _cc(src, out, strict=True):
    os.makedirs(BUILD, exist_ok=True)
    libdir = os.path.join(ROOT, "chafa_b200")
    cmd = ["gcc", "-std=c99", "-Wall"] + (["-Wextra", "-Werror"] if strict else []) + ["-I", os.path.join(ROOT, "include"), src, "-o", out,
           "-L", libdir, "-lchafa_b200", f"-Wl,-rpath,{libdir}", "-Wl,--no-undefined"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return out


def test_c_client_links():
    _cc(os.path.join(ROOT, "tests", "c", "client.c"), os.path.join(BUILD, "client"))


@pytest.mark.skipif(not os.path.exists(REF_EXAMPLE), reason="reference tree not present (GPU box)")
def test_reference_example_links_unmodified():
    """examples/example.c:30-62, as it is, against this repo's header and library."""
    _cc(REF_EXAMPLE, os.path.join(BUILD, "ref_example"), strict=False)


def _sections(blob):
    out, pos = {}, 0
    while pos < len(blob):
        eol = blob.index(b"\n", pos)
        head = blob[pos:eol].decode().split()
        assert head[0] == "=="
        if head[1] == "n_rows":
            out["n_rows"], out["total"] = int(head[2]), int(head[4])
            pos = eol + 1
            continue
        n = int(head[2])
        out[head[1]] = blob[eol + 1:eol + 1 + n]
        pos = eol + 1 + n + 1
    return out


def _client_pixels():
    W, H = 64, 48
    y, x = np.mgrid[0:H, 0:W]
    px = np.zeros((H, W, 4), np.uint8)
    px[..., 0] = (x * 4) & 255
    px[..., 1] = (y * 5) & 255
    px[..., 2] = ((x ^ y) * 3) & 255
    px[..., 3] = np.where((x + y) % 11 == 0, 0, 255)
    return np.ascontiguousarray(px)


@pytest.mark.gpu
def test_c_client_output_matches_reference(gpu, reference):
    exe = _cc(os.path.join(ROOT, "tests", "c", "client.c"), os.path.join(BUILD, "client"))
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode == 0, r.stderr.decode(errors="replace")
    got = _sections(r.stdout)

    ref = reference
    px = _client_pixels()
    c = capi.Canvas(ref, 16, 6, mode=capi.CANVAS_MODE_INDEXED_256, symbols="all")
    c.draw(px, 64, 48)
    assert got["print"] == c.print()

    ref.lib.chafa_canvas_print_rows_strv.restype = C.POINTER(C.c_char_p)
    ref.lib.chafa_canvas_print_rows_strv.argtypes = [C.c_void_p, C.c_void_p]
    rows = ref.lib.chafa_canvas_print_rows_strv(c.handle, None)
    n = 0
    while rows[n] is not None:
        assert got[f"row{n}"] == rows[n]
        n += 1
    assert got["n_rows"] == n == 6
    c.close()

    ti = ref.chafa_term_info_new()
    fb = ref.chafa_term_db_get_fallback_info(ref.chafa_term_db_get_default())
    buf = C.create_string_buffer(256)
    for name, args in (("cursor_to_pos", (12, 34)), ("set_color_fgbg_direct", (1, 2, 3, 250, 251, 252)),
                       ("reset_attributes", ()), ("begin_sixels", (0, 1, 0))):
        fn = getattr(ref.lib, "chafa_term_info_emit_" + name)
        fn.restype = C.c_void_p
        fn.argtypes = [C.c_void_p, C.c_void_p] + [C.c_uint] * len(args)
        end = fn(fb, C.addressof(buf), *args)
        assert got[name] == buf.raw[:end - C.addressof(buf)], name
    ref.chafa_term_info_unref(fb)
    ref.chafa_term_info_unref(ti)


@pytest.mark.gpu
def test_print_rows_strv(gpu, reference):
    """chafa_canvas_print_rows_strv (chafa-canvas.c:1000-1032): NULL-terminated array of NUL-terminated
    rows, equal to the reference's and to print_rows."""
    import parity
    img = parity.make_image("shapes", 256, 128, seed=3)
    for lib in (gpu, reference):
        lib.lib.chafa_canvas_print_rows_strv.restype = C.POINTER(C.c_void_p)
        lib.lib.chafa_canvas_print_rows_strv.argtypes = [C.c_void_p, C.c_void_p]
    out = {}
    for lib in (gpu, reference):
        c = capi.Canvas(lib, 32, 16, mode=capi.CANVAS_MODE_INDEXED_256, symbols="all+wide")
        c.draw(img, 256, 128)
        rows = lib.lib.chafa_canvas_print_rows_strv(c.handle, None)
        got = []
        i = 0
        while rows[i]:
            got.append(C.string_at(rows[i]))
            lib.free_array(rows[i])
            i += 1
        lib.free_array(C.cast(rows, C.c_void_p))
        out[lib.is_reference] = got
        c.close()
    assert len(out[False]) == 16
    assert out[False] == out[True]
