"""Deterministic synthetic RGBA8 images: noise, gradient+shapes, alpha."""

import numpy as np


def noise(w, h, seed=1):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    img[..., 3] = 255
    return np.ascontiguousarray(img)


def shapes(w, h, seed=2, shift=0):
    """Smooth gradients with hard-edged rectangles and discs; @shift translates by pixels."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    xx = xx + shift
    img = np.zeros((h, w, 4), dtype=np.uint8)
    img[..., 0] = (xx * 255 // max(w - 1, 1)) & 255
    img[..., 1] = (yy * 255 // max(h - 1, 1)) & 255
    img[..., 2] = ((xx + yy) * 255 // max(w + h - 2, 1)) & 255
    for _ in range(24):
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        sw, sh = int(rng.integers(w // 32 + 1, w // 4 + 2)), int(rng.integers(h // 32 + 1, h // 4 + 2))
        col = rng.integers(0, 256, size=3, dtype=np.uint8)
        if rng.integers(0, 2):
            m = (xx >= x0) & (xx < x0 + sw) & (yy >= y0) & (yy < y0 + sh)
        else:
            r = min(sw, sh) // 2 + 1
            m = (xx - x0) ** 2 + (yy - y0) ** 2 <= r * r
        img[m, :3] = col
    img[..., 3] = 255
    return np.ascontiguousarray(img)


def alpha(w, h, seed=3):
    """shapes() with a soft-edged alpha mask including fully transparent regions."""
    img = shapes(w, h, seed)
    yy, xx = np.mgrid[0:h, 0:w]
    d = np.sqrt(((xx - w / 2) / (w / 2 + 1e-9)) ** 2 + ((yy - h / 2) / (h / 2 + 1e-9)) ** 2)
    a = np.clip((1.1 - d) * 400.0, 0, 255).astype(np.uint8)
    a[: max(h // 8, 1), : max(w // 8, 1)] = 0
    img[..., 3] = a
    return np.ascontiguousarray(img)


def animation(w, h, n_frames, seed=2):
    """BASELINE configs[3]: @n_frames frames of shapes () translated by one pixel per frame
    (SURVEY.md section 8d): frame f is the window [f, f + w) of one wide image.  Returns a function
    f -> contiguous (h, w, 4) array."""
    wide = shapes(w + n_frames, h, seed)

    def frame(f):
        return np.ascontiguousarray(wide[:, f:f + w])

    return frame


KINDS = {"noise": noise, "shapes": shapes, "alpha": alpha}
