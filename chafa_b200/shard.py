"""Host-side sharding rules for multi-GPU runs.

Animations shard by frame, stills by cell-row band; neither needs a collective on the data
path.  Each rank renders its share on its own GPU (chafa_b200_canvas_set_band for bands) and
rank 0 gathers the text.  These helpers hold the arithmetic so that it can be tested on CPU
with the gloo backend."""

RESET = b"\x1b[0m"


def split_rows(n_rows, world):
    """Contiguous, near-equal bands of cell rows: list of (first_row, n_rows), one per rank."""
    base, extra = divmod(n_rows, world)
    out, first = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((first, n))
        first += n
    return out


def split_sixel_rows(n_cell_rows, cell_height, world):
    """Bands of a sixel canvas: a band starts on a multiple of six pixel rows (sixel bands,
    chafa/internal/chafa-sixel-canvas.c) and of the cell height (the unit of
    chafa_b200_canvas_set_band), i.e. on a multiple of lcm(6, cell_height) / cell_height cell
    rows; the last band takes what is left.  Ranks beyond the available units get (n, 0)."""
    from math import gcd
    unit = 6 // gcd(6, cell_height)             # cell rows per boundary unit
    n_units = (n_cell_rows + unit - 1) // unit
    base, extra = divmod(n_units, world)
    out, first = [], 0
    for r in range(world):
        n = (base + (1 if r < extra else 0)) * unit
        n = max(0, min(n, n_cell_rows - first))
        out.append((first, n))
        first += n
    return out


def frames_of_rank(n_frames, world, rank):
    """Frames rendered by @rank: round-robin, so every rank gets work from the first frame on."""
    return range(rank, n_frames, world)


def band_text_from_standalone(text, is_first, is_last, leading=RESET):
    """Turn the print of a standalone canvas holding just one band into what the band
    contributes to the print of the whole canvas: only the first canvas row carries the leading
    reset, and every row but the last of the whole canvas ends in a newline
    (chafa/internal/chafa-canvas-printer.c:681-688, 755-756)."""
    if not is_first and leading and text.startswith(leading):
        text = text[len(leading):]
    if not is_last:
        text += b"\n"
    return text


def join_bands(texts):
    """Bands printed by chafa_b200 band canvases (or converted with band_text_from_standalone)
    concatenate to the full print."""
    return b"".join(texts)


def max_over_ranks(value, dist=None, device=None):
    """Multi-GPU timings are reported as the slowest rank's."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
