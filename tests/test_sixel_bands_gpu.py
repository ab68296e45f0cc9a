"""Row-band sharding of sixel stills (north_star: "NCCL ... to all-reduce the global palette
histogram").  The band owners are emulated on ONE GPU so that the check runs on the driver's
single-GPU box: one canvas per band, the colour-cube bins summed on the host with numpy in place of
ncclAllReduce (exact integer sums either way), the float sums of the bins beyond float's exact range
handed from owner to owner in band order.  The concatenated band prints must equal the reference's
print of the whole image, byte for byte.  bench.py --gpus N runs the same protocol over NCCL."""

import ctypes as C

import numpy as np
import pytest

import parity
from chafa_b200 import capi, shard

pytestmark = pytest.mark.gpu


def _dev_source(lib, img):
    """The source image in device memory (the two-phase draw takes a device pointer)."""
    d = lib.chafa_b200_device_alloc(img.nbytes)
    assert d
    assert lib.chafa_b200_copy_to_device(d, img.ctypes.data, img.nbytes)
    return d


def render_banded(lib, img, cw, ch, world, cell=(8, 8), **cfg):
    """-> (concatenated text, list of per-band facts)"""
    h, w = img.shape[:2]
    cfg = dict(cfg)
    cfg.setdefault("pixel_mode", capi.PIXEL_MODE_SIXELS)
    cfg.setdefault("symbols", None)
    d_src = _dev_source(lib, img)
    bands = [b for b in shard.split_sixel_rows(ch, cell[1], world) if b[1] > 0]
    canvases, heads = [], []
    try:
        for first, n in bands:
            c = capi.Canvas(lib, cw, ch, cell=cell, **cfg)
            lib.chafa_b200_canvas_set_band(c.handle, first, n)
            canvases.append(c)

        # phase 1 on every owner, then the "all-reduce"
        need = []
        for c in canvases:
            buf, n_words = C.c_void_p(), C.c_size_t()
            r = lib.chafa_b200_canvas_draw_begin_device(c.handle, capi.PIXEL_RGBA8_UNASSOCIATED, d_src, w, h, w * 4,
                                                        C.byref(buf), C.byref(n_words))
            need.append(r)
            heads.append((buf.value, n_words.value))
            assert lib.chafa_b200_canvas_synchronize(c.handle)
        assert len(set(need)) == 1
        reduced = bool(need[0])
        if reduced:
            total = None
            for buf, n_words in heads:
                a = np.zeros(n_words, np.uint32)
                assert lib.chafa_b200_copy_from_device(a.ctypes.data, buf, a.nbytes)
                total = a if total is None else total + a          # wraps like the device sums
            for buf, n_words in heads:
                assert lib.chafa_b200_copy_to_device(buf, total.ctypes.data, total.nbytes)

            # float sums, owner after owner from the top; the last owner's go to everybody
            state = None
            ptrs = []
            for c in canvases:
                p, nb = C.c_void_p(), C.c_size_t()
                if state is not None:
                    # what the owner above handed on
                    buf, n_words = heads[len(ptrs)]
                    assert lib.chafa_b200_copy_to_device(buf + n_words * 4, state.ctypes.data, state.nbytes)
                assert lib.chafa_b200_canvas_draw_float_sums_device(c.handle, C.byref(p), C.byref(nb))
                assert lib.chafa_b200_canvas_synchronize(c.handle)
                state = np.zeros(nb.value // 4, np.float32)
                assert lib.chafa_b200_copy_from_device(state.ctypes.data, p.value, state.nbytes)
                ptrs.append(p.value)
            for p in ptrs:
                assert lib.chafa_b200_copy_to_device(p, state.ctypes.data, state.nbytes)

        texts = []
        for c in canvases:
            lib.chafa_b200_canvas_draw_finish(c.handle)
            texts.append(c.print())
        return b"".join(texts), dict(reduced=reduced, n_bands=len(bands), lens=[len(t) for t in texts],
                                     risky=int((state != 0).any()) if reduced else 0)
    finally:
        for c in canvases:
            c.close()
        lib.chafa_b200_device_free(d_src)


def reference_print(img, cw, ch, n_threads, cell=(8, 8), **cfg):
    ref = capi.load_reference()
    ref.chafa_set_n_threads(n_threads)
    cfg = dict(cfg)
    cfg.setdefault("pixel_mode", capi.PIXEL_MODE_SIXELS)
    cfg.setdefault("symbols", None)
    c = capi.Canvas(ref, cw, ch, cell=cell, **cfg)
    try:
        c.draw(img, img.shape[1], img.shape[0])
        return c.print()
    finally:
        c.close()


CASES = {
    # name: (image kind, src w, src h, canvas cells w, h, cell, threads, config)
    "noise_dither": ("shapes", 640, 400, 80, 48, (8, 8), 16, dict(dither=capi.DITHER_NOISE, grain=(1, 1))),
    "no_dither_alpha": ("alpha", 640, 400, 80, 48, (8, 8), 16, dict(dither=capi.DITHER_NONE)),
    "ordered_noise_img": ("noise", 512, 512, 60, 45, (8, 8), 5, dict(dither=capi.DITHER_ORDERED, grain=(2, 2))),
    "cell_10x20": ("shapes", 500, 300, 40, 21, (10, 20), 3, dict(dither=capi.DITHER_NOISE, grain=(1, 1))),
    "work9_dense_sampling": ("shapes", 400, 240, 50, 30, (8, 8), 16, dict(dither=capi.DITHER_NOISE, grain=(1, 1), work=1.0)),
    "one_thread_one_batch": ("shapes", 320, 240, 40, 30, (8, 8), 1, dict(dither=capi.DITHER_NOISE, grain=(1, 1))),
    "fixed_palette_16": ("shapes", 320, 240, 40, 30, (8, 8), 16, dict(dither=capi.DITHER_NOISE, grain=(1, 1),
                                                                       mode=capi.CANVAS_MODE_INDEXED_16)),
    "diffusion_replicas": ("shapes", 320, 200, 30, 12, (8, 8), 16, dict(dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
}


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", sorted(CASES))
def test_sixel_bands_concatenate_to_reference(gpu, reference, name, world):
    kind, sw, sh, cw, ch, cell, threads, cfg = CASES[name]
    img = parity.make_image(kind, sw, sh, seed=23)
    gpu.chafa_set_n_threads(threads)
    try:
        want = reference_print(img, cw, ch, threads, cell=cell, **cfg)
        got, facts = render_banded(gpu, img, cw, ch, world, cell=cell, **cfg)
    finally:
        gpu.chafa_set_n_threads(-1)
        reference.chafa_set_n_threads(-1)
    assert facts["n_bands"] >= 2
    dynamic = cfg.get("mode", capi.CANVAS_MODE_TRUECOLOR) not in (capi.CANVAS_MODE_INDEXED_16,)
    assert facts["reduced"] == (dynamic and cfg["dither"] != capi.DITHER_DIFFUSION)
    assert got == want, (facts, len(got), len(want))


def test_sixel_bands_float_sums_travel(gpu, reference):
    """A large flat area puts more than 65 793 samples into one bin: its float sums leave the exact
    range and depend on the order of the samples, which spans the bands."""
    img = np.zeros((1080, 1920, 4), np.uint8)
    img[...] = (201, 77, 33, 255)
    img[300:700, 200:1500] = (255, 253, 251, 255)
    img[::7, ::5, 1] = 99
    cfg = dict(dither=capi.DITHER_NOISE, grain=(1, 1), work=1.0)
    want = reference_print(img, 240, 135, 16, **cfg)
    gpu.chafa_set_n_threads(16)
    try:
        for world in (2, 4):
            got, facts = render_banded(gpu, img, 240, 135, world, **cfg)
            assert facts["reduced"] and facts["risky"]
            assert got == want, facts
    finally:
        gpu.chafa_set_n_threads(-1)
        reference.chafa_set_n_threads(-1)


def test_sixel_band_must_lie_on_sixel_rows(gpu):
    c = capi.Canvas(gpu, 20, 10, pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None)
    gpu.chafa_b200_canvas_set_band(c.handle, 1, 3)          # 8 px: not a multiple of 6
    img = parity.make_image("shapes", 160, 80, seed=2)
    c.draw(img, 160, 80)
    assert b"sixel bands" in gpu.chafa_b200_last_error()
    assert c.print() == b""
    c.close()


# ---- many stills, one launch for their diffusion chains ---------------------------------------

def test_batched_diffusion_chains_match_reference(gpu, reference):
    """chafa_b200_canvases_draw_batch_device: the Floyd-Steinberg chains of several sixel canvases
    run as blocks of one launch (one launch per kind of chain); every canvas prints what the
    reference prints for its image.  Canvases of other kinds in the same batch are drawn as usual."""
    w, h = 320, 200
    specs = [
        ("shapes", dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
        ("noise", dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
        ("alpha", dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1))),
        ("shapes", dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1),
                        dither_intensity=0.5)),
        ("shapes", dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1),
                        mode=capi.CANVAS_MODE_INDEXED_16)),
        ("alpha", dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_NOISE, grain=(1, 1))),
        ("shapes", dict(mode=capi.CANVAS_MODE_INDEXED_256, symbols="all")),
        ("noise", dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1),
                       color_space=capi.COLOR_SPACE_DIN99D)),
    ]
    images = [parity.make_image(kind, w, h, seed=50 + i) for i, (kind, _) in enumerate(specs)]
    reference.chafa_set_n_threads(1)
    gpu.chafa_set_n_threads(1)
    want = []
    for img, (_, cfg) in zip(images, specs):
        c = capi.Canvas(reference, 30, 12, **cfg)
        c.draw(img, w, h)
        want.append(c.print())
        c.close()

    d_srcs, canvases = [], []
    try:
        for img, (_, cfg) in zip(images, specs):
            d = gpu.chafa_b200_device_alloc(img.nbytes)
            assert d and gpu.chafa_b200_copy_to_device(d, img.ctypes.data, img.nbytes)
            d_srcs.append(d)
            canvases.append(capi.Canvas(gpu, 30, 12, **cfg))
        for rep in range(2):        # the second round reuses every buffer
            before = gpu.chafa_b200_launch_count()
            assert capi.draw_batch_device(gpu, canvases, d_srcs, w, h)
            got = [c.print() for c in canvases]
            assert got == want, [i for i in range(len(want)) if got[i] != want[i]]
        # the same from host images
        assert capi.draw_batch_device(gpu, canvases, [im.ctypes.data for im in images], w, h, on_device=False)
        assert [c.print() for c in canvases] == want
        # an individual draw after a batched one
        canvases[0].draw(images[1], w, h)
        assert canvases[0].print() == want[1]
    finally:
        gpu.chafa_set_n_threads(-1)
        reference.chafa_set_n_threads(-1)
        for c in canvases:
            c.close()
        for d in d_srcs:
            gpu.chafa_b200_device_free(d)
