"""The N > 1 host logic on CPU: two gloo ranks render two row bands of a still (with the
compiled reference standing in for the GPU renderer) and frame shares of an animation;
rank 0 gathers and the result must equal the single-process render."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity
from chafa_b200 import capi, shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = capi.load_reference()
    ref.chafa_set_n_threads(1)
    cw, ch = 24, 10
    img = parity.make_image("shapes", cw * 8, ch * 8, seed=11)

    # still: band of cell rows per rank (1:1 scale, so a band is a sub-image)
    first, n = shard.split_rows(ch, world)[rank]
    band = np.ascontiguousarray(img[first * 8:(first + n) * 8])
    c = capi.Canvas(ref, cw, n, symbols="all+wide")
    c.draw(band, cw * 8, n * 8)
    text = shard.band_text_from_standalone(c.print(), first == 0, first + n == ch)
    gathered = [None] * world
    dist.all_gather_object(gathered, text)

    # animation: frames round-robin
    n_frames = 5
    mine = {}
    for f in shard.frames_of_rank(n_frames, world, rank):
        frame = parity.make_image("shapes", cw * 8, ch * 8, seed=100 + f)
        cf = capi.Canvas(ref, cw, ch)
        cf.draw(frame, cw * 8, ch * 8)
        mine[f] = cf.print()
    frames = [None] * world
    dist.all_gather_object(frames, mine)

    slowest = shard.max_over_ranks(float(rank + 1), dist)
    if rank == 0:
        merged = {}
        for d in frames:
            merged.update(d)
        np.save(os.path.join(out_dir, "still.npy"), np.frombuffer(shard.join_bands(gathered), np.uint8))
        np.save(os.path.join(out_dir, "frames.npy"),
                np.frombuffer(b"\0".join(merged[f] for f in range(n_frames)), np.uint8))
        np.save(os.path.join(out_dir, "slowest.npy"), np.array([slowest]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.exists(capi.REFERENCE_LIB), reason="compiled reference not built")
def test_two_rank_band_and_frame_sharding(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)

    ref = capi.load_reference()
    ref.chafa_set_n_threads(1)
    cw, ch = 24, 10
    img = parity.make_image("shapes", cw * 8, ch * 8, seed=11)
    c = capi.Canvas(ref, cw, ch, symbols="all+wide")
    c.draw(img, cw * 8, ch * 8)
    assert np.load(tmp_path / "still.npy").tobytes() == c.print()

    expected = []
    for f in range(5):
        frame = parity.make_image("shapes", cw * 8, ch * 8, seed=100 + f)
        cf = capi.Canvas(ref, cw, ch)
        cf.draw(frame, cw * 8, ch * 8)
        expected.append(cf.print())
    assert np.load(tmp_path / "frames.npy").tobytes() == b"\0".join(expected)
    assert float(np.load(tmp_path / "slowest.npy")[0]) == 2.0


def test_split_rows_covers_everything():
    for n in (1, 2, 7, 67, 270, 540):
        for world in (1, 2, 3, 4, 8):
            parts = shard.split_rows(n, world)
            assert len(parts) == world
            assert parts[0][0] == 0 and sum(p[1] for p in parts) == n
            for a, b in zip(parts, parts[1:]):
                assert a[0] + a[1] == b[0]
            assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1
    assert sorted(sum((list(shard.frames_of_rank(10, 4, r)) for r in range(4)), [])) == list(range(10))
