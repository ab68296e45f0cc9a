#!/usr/bin/env python
"""Benchmark of the canvas rendering hot path (chafa_canvas_draw_all_pixels + chafa_canvas_print).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME] [--image KIND]

One "step" is one full frame through the hot path: scale/composite (+ elementwise
preprocessing) -> cell matching (narrow + wide symbols, one tensor-core kernel) -> row resolve ->
print (one kernel).  Default workload is BASELINE.json configs[1]:
3840x2160 RGBA -> 480x270 cells, symbols all+wide, 256 colours, ordered dither.

Numbers on the JSON line:
  value      cells/s, whole job, source images already resident in HBM and the printed text (and its
             length) left in HBM (device-side API of include/chafa_b200.h): the K frames are queued
             back to back, CUDA events around them; median of --repeats such windows
  e2e        cells/s through the reference-facing C API with HOST buffers: H2D of the frame from
             pinned memory and D2H of the printed text inside the timed region; e2e.pageable is the
             same with the frames in ordinary (malloc) memory, as a drop-in caller has them
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event duration vs measured HBM peak
  cpu_baseline  the compiled reference (oracle/_ref) on this box's host cores, bounded sample
  parity_check  one frame of the timed workload against the compiled reference
  c4_frames  BASELINE configs[3]: the 300-frame 1080p animation, frames sharded over the ranks
  strong_c5_bands  BASELINE configs[4]: one 8K DIN99d still as cell-row bands over the ranks, the
             text assembled on rank 0 by peer stores over NVLink (no host round trip)

With N > 1 (torchrun) every rank renders its own frames (frame sharding, no collective on the
data path); value = all ranks' cells / max-over-ranks time.
"""

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
from concurrent.futures import ThreadPoolExecutor
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# Canvases in flight on their own CUDA streams only overlap if the streams do not share a hardware
# queue: the driver's default is 8 queues, the most it offers is 32 (read when the context is made).
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from chafa_b200 import capi, shard, synth  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2_4k_256c_ordered_allwide": dict(
        src=(3840, 2160), cells=(480, 270),
        cfg=dict(mode=capi.CANVAS_MODE_INDEXED_256, symbols="all+wide", dither=capi.DITHER_ORDERED, work=0.5),
        desc="3840x2160 RGBA8 -> 480x270 cells, symbols all+wide (1251 narrow + 181 wide), INDEXED_256, "
             "ordered dither, work 0.5"),
    # BASELINE.json configs[0] / the per-frame unit of configs[3]
    "c1_1080p_truecolor_block": dict(
        src=(1920, 1080), cells=(240, 67),
        cfg=dict(mode=capi.CANVAS_MODE_TRUECOLOR, symbols="block+border", work=0.5),
        desc="1920x1080 RGBA8 -> 240x67 cells, truecolor, symbols block+border, work 0.5"),
    # BASELINE.json configs[3]: the animation (frames sharded over the ranks)
    "c4_anim_300f_1080p": dict(
        src=(1920, 1080), cells=(240, 67), frames=300,
        cfg=dict(mode=capi.CANVAS_MODE_TRUECOLOR, symbols="block+border", work=0.5),
        desc="300 frames of 1920x1080 RGBA8 (shapes translated by one pixel per frame) -> 240x67 cells, truecolor, "
             "symbols block+border, work 0.5; one step = one frame"),
    # BASELINE.json configs[4] on one GPU
    "c5_8k_din99d_work9": dict(
        src=(7680, 4320), cells=(960, 540),
        cfg=dict(mode=capi.CANVAS_MODE_INDEXED_256, symbols="all", work=1.0, color_space=capi.COLOR_SPACE_DIN99D),
        desc="7680x4320 RGBA8 -> 960x540 cells, DIN99d, INDEXED_256, symbols all, work 1.0 (exhaustive)"),
    # BASELINE.json configs[2]: one diffusion chain per image (serial by construction): replicas
    "c3_1080p_sixel_fs": dict(
        src=(1920, 1080), cells=(240, 135),
        cfg=dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1)),
        desc="1920x1080 RGBA8 -> 1920x1080 sixels (240x135 cells of 8x8), 256-colour generated palette, "
             "Floyd-Steinberg"),
    # a large sixel still for the band-sharded leg (the palette's colour bins are all-reduced)
    "c3n_4k_sixel_noise": dict(
        src=(3840, 2160), cells=(480, 270),
        cfg=dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_NOISE, grain=(1, 1)),
        desc="3840x2160 RGBA8 -> 3840x2160 sixels, 256-colour generated palette, blue-noise dither"),
    "c3n_1080p_sixel_noise": dict(
        src=(1920, 1080), cells=(240, 135),
        cfg=dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_NOISE, grain=(1, 1)),
        desc="1920x1080 RGBA8 -> 1920x1080 sixels, 256-colour generated palette, blue-noise dither"),
}
STAGES = ["h2d", "scale", "prep", "match", "match_wide", "resolve", "print_measure", "print_emit", "d2h"]
E2E_THREADS = 8   # host threads of the e2e leg on one rank; fewer per rank when ranks share the host
N_DEV_CANVASES = int(os.environ.get("CHAFA_B200_BENCH_CANVASES", "0"))   # override of dev_canvases () for measurements
N_ROTATE = 8   # distinct source images cycled through so every step reads a cold (non-L2-resident) frame
N_REPLICAS = 444  # sixel diffusion: stills in flight.  One serial chain each; their chains are queued as one launch (a
                  # process has 32 hardware queues, so chains on separate streams stop overlapping at 32), a block per
                  # image and three blocks per SM (a chain waits on its own loads most of the time: 128 / 148 / 296 / 444
                  # stills per launch: 190 / 201 / 247 / 267 frames/s with a 3.07 ms pen table per image, 426 at 444 with
                  # the 1.42 ms one; the pen table of all 2^24 colours, built on the whole GPU per image, bounds it)


def host_facts(threads_per_rank, world):
    """The host side of the e2e leg: cores, feeding threads, and how the GPUs hang off the host."""
    facts = {"cpus": host_cpus(), "feeder_threads_per_rank": threads_per_rank, "ranks_on_this_host": world}
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                             timeout=20).stdout.splitlines()
        rows = [ln.split() for ln in out if ln.startswith("GPU")]
        if rows:
            # link of GPU0 to every other GPU, and the CPU / NUMA affinity columns of every GPU
            facts["gpu0_links"] = " ".join(rows[0][1:1 + len(rows)])
            facts["numa_affinity"] = sorted({r[-2] for r in rows if len(r) > len(rows) + 2})
    except Exception:  # noqa: BLE001
        pass
    return facts


def dev_canvases(n_cells, world=1):
    """Canvases (CUDA streams, host threads) the device-resident leg spreads its frames over: four for
    frames of a few ten thousand cells (their kernels are short and leave SMs free: C4 52 700 frames/s
    with four, 48 300 with three), three up to C2's size, one beyond (a frame of more than ~200 k cells
    fills the GPU on its own: measured no gain, C5 loses 5 %)."""
    if N_DEV_CANVASES > 0:
        return N_DEV_CANVASES
    if n_cells <= 50000:
        # every canvas has a host thread: not more of them (plus the main thread) than this rank's share of the cores
        return 4 if world * 5 <= host_cpus() else 3
    return 3 if n_cells <= 200000 else 1


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                mhz.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mhz.sort()
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


def median(xs):
    return sorted(xs)[len(xs) // 2]


def is_sixel(wl):
    return wl["cfg"].get("pixel_mode") == capi.PIXEL_MODE_SIXELS


def is_serial_chain(wl):
    """Sixel Floyd-Steinberg: one dependency chain per image; throughput comes from replicas."""
    return is_sixel(wl) and wl["cfg"].get("dither") == capi.DITHER_DIFFUSION


def make_sources(wl, kind, n):
    """@n distinct source frames of the workload (the animation: its first @n frames)."""
    w, h = wl["src"]
    if "frames" in wl:
        frame = synth.animation(w, h, wl["frames"], seed=100)
        return [frame(i) for i in range(min(n, wl["frames"]))]
    return [synth.KINDS[kind](w, h, seed=100 + i) for i in range(n)]


def bench_config(args, wl):
    """The `config` object: identical on the product arm and the reference arm."""
    return {"workload": args.workload, "desc": wl["desc"], "image": "shapes" if "frames" in wl else args.image}


def host_cpus():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 16


def parity_check(lib, wl, src):
    """What was timed is what is checked: one source frame of the timed workload through the
    product's host API and through the compiled reference (oracle/_ref), outputs compared."""
    try:
        ref = capi.load_reference()
    except Exception as e:  # noqa: BLE001
        return {"bytes_equal": None, "note": f"reference unavailable: {e}"}
    ref.chafa_set_n_threads(-1)
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    cr, cp = capi.Canvas(ref, cw, ch, **wl["cfg"]), capi.Canvas(lib, cw, ch, **wl["cfg"])
    cr.draw(src, w, h)
    cp.draw(src, w, h)
    tr, tp = cr.print(), cp.print()
    out = {"bytes_equal": tr == tp, "printed_bytes": len(tp), "reference_printed_bytes": len(tr),
           "against": "oracle/_ref (compiled reference), same source frame, host API on both sides"}
    if not is_sixel(wl):
        a, b = cr.cells(), cp.cells()
        out["n_cells"] = int(cw * ch)
        out["cells_differing"] = int((a != b).any(axis=2).sum())
        out["symbols_differing"] = int((a[..., 0] != b[..., 0]).sum())
        out["symbols_differing_fraction"] = out["symbols_differing"] / float(cw * ch)
        if wl["cfg"].get("color_space") == capi.COLOR_SPACE_DIN99D:
            out["tolerance"] = "DIN99d (double transcendental maths): <= 0.5 % of cells may differ"
    cr.close()
    cp.close()
    return out


# ---- the reference's own CPU implementation -------------------------------------------------------

def reference_replicas(ref, wl, srcs, n_canvases, frames_each, warm=1):
    """@n_canvases canvases rendered concurrently, one host thread each with the library's own
    threading off: how a caller gets throughput out of the reference when one frame is a serial
    chain.  Returns seconds for n_canvases * frames_each frames."""
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    canvases = [capi.Canvas(ref, cw, ch, **wl["cfg"]) for _ in range(n_canvases)]
    start, done = threading.Barrier(n_canvases + 1), threading.Barrier(n_canvases + 1)

    def worker(t):
        for i in range(warm):
            canvases[t].draw(srcs[(t + i) % len(srcs)], w, h)
            canvases[t].print_len()
        start.wait()
        for i in range(frames_each):
            canvases[t].draw(srcs[(t + i) % len(srcs)], w, h)
            canvases[t].print_len()
        done.wait()

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_canvases)]
    for th in threads:
        th.start()
    start.wait()
    t0 = time.perf_counter()
    done.wait()
    dt = time.perf_counter() - t0
    for th in threads:
        th.join()
    for c in canvases:
        c.close()
    return dt


def bench_reference(args, wl):
    """The reference's own CPU implementation (oracle/_ref) on all host threads."""
    ref = capi.load_reference()
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    n_cells = cw * ch
    srcs = make_sources(wl, args.image, wl["frames"] if "frames" in wl else 2)
    windows = []

    if is_serial_chain(wl):
        # one frame is a serial chain: the reference's throughput configuration is one canvas per core
        cores = host_cpus()
        ref.chafa_set_n_threads(1)
        threads = cores
        frames_each = max(1, args.steps // cores)
        n_frames = frames_each * cores
        for r in range(args.repeats):
            windows.append(reference_replicas(ref, wl, srcs, cores, frames_each, warm=1 if r == 0 else 0))
        sample = (f"{args.repeats} windows of {n_frames} full frames, {cores} canvases in flight (one host thread "
                  f"each, library threading off), median window")
    else:
        ref.chafa_set_n_threads(-1)
        threads = ref.chafa_get_n_actual_threads()
        canvas = capi.Canvas(ref, cw, ch, **wl["cfg"])
        n_frames = wl["frames"] if "frames" in wl else args.steps

        def step(i):
            canvas.draw(srcs[i % len(srcs)], w, h)
            return canvas.print_len()

        for i in range(args.warmup):
            step(i)
        for r in range(args.repeats):
            t0 = time.perf_counter()
            for i in range(n_frames):
                step(r * n_frames + i)
            windows.append(time.perf_counter() - t0)
        canvas.close()
        sample = (f"{args.repeats} windows of {n_frames} full frames ({args.warmup} warm-up), median window, "
                  f"one canvas reused, {threads} threads")

    dt = median(windows)
    value = n_cells * n_frames / dt
    line = {
        "impl": "reference", "metric": "cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": n_frames, "warmup": args.warmup, "ms_per_step": dt / n_frames * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args, wl),
        "repeats": {"windows": args.repeats, "ms_per_step_min": min(windows) / n_frames * 1e3,
                    "ms_per_step_median": dt / n_frames * 1e3, "ms_per_step_max": max(windows) / n_frames * 1e3},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_sec": n_frames / dt, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(args, wl, budget_s=12.0):
    """Reference on host cores, bounded: frames until ~budget_s of CPU wall time."""
    try:
        ref = capi.load_reference()
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "cells/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    srcs = make_sources(wl, args.image, 2)
    if is_serial_chain(wl):
        cores = host_cpus()
        ref.chafa_set_n_threads(1)
        dt = reference_replicas(ref, wl, srcs, cores, 2)
        n = 2 * cores
        ref.chafa_set_n_threads(-1)
        c = capi.Canvas(ref, cw, ch, **wl["cfg"])
        c.draw(srcs[0], w, h)
        c.print_len()
        t0 = time.perf_counter()
        c.draw(srcs[0], w, h)
        c.print_len()
        one = time.perf_counter() - t0
        c.close()
        return {"value": cw * ch * n / dt, "unit": "cells/s", "cores": cores, "kind": "reference",
                "frames_per_sec": n / dt, "single_frame_ms": one * 1e3,
                "sample": f"{n} full frames, {cores} canvases in flight on {cores} host threads (library threading "
                          f"off: one frame is a serial chain), {dt:.1f} s; single_frame_ms = one frame alone with "
                          f"the library's own threading"}
    ref.chafa_set_n_threads(-1)
    threads = ref.chafa_get_n_actual_threads()
    canvas = capi.Canvas(ref, cw, ch, **wl["cfg"])
    canvas.draw(srcs[0], w, h)
    canvas.print_len()
    n, t0 = 0, time.perf_counter()
    while True:
        canvas.draw(srcs[n % len(srcs)], w, h)
        canvas.print_len()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 50:
            break
    canvas.close()
    return {"value": cw * ch * n / dt, "unit": "cells/s", "cores": threads, "kind": "reference",
            "sample": f"{n} full frames of the same workload after 1 warm-up, {threads} threads "
                      f"({os.cpu_count()} host cpus), {dt:.1f} s"}


# ---- the B200 library ----------------------------------------------------------------------------

class Ctx:
    """Per-process state shared by the legs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.distributed = self.world > 1
        torch.cuda.set_device(self.local_rank)
        if self.distributed:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        self.lib = capi.load_product()
        if not self.lib.chafa_b200_available():
            raise RuntimeError("libchafa_b200: no CUDA device (there is no CPU fallback)")
        self.lib.chafa_b200_set_device(self.local_rank)
        lib = self.lib.lib
        lib.chafa_b200_canvas_set_kernel_timing.argtypes = [C.c_void_p, C.c_int]
        lib.chafa_b200_canvas_get_kernel_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.chafa_b200_canvas_copy_print_length.argtypes = [C.c_void_p, C.c_void_p]
        lib.chafa_b200_canvas_copy_print_length.restype = C.c_int
        lib.chafa_b200_canvas_band_scatter.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                       C.c_size_t, C.c_void_p]
        lib.chafa_b200_canvas_band_scatter.restype = C.c_int
        lib.chafa_b200_ipc_export.argtypes = [C.c_void_p, C.c_void_p]
        lib.chafa_b200_ipc_export.restype = C.c_int
        lib.chafa_b200_ipc_open.argtypes = [C.c_void_p]
        lib.chafa_b200_ipc_open.restype = C.c_void_p
        lib.chafa_b200_ipc_close.argtypes = [C.c_void_p]

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        values = list(values)
        if not self.distributed or not values:
            return values
        t = self.torch.tensor(values, device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def read_u64(self, d_ptr):
        buf = C.c_uint64(0)
        if not self.lib.chafa_b200_copy_from_device(C.addressof(buf), d_ptr, 8):
            raise RuntimeError("could not read a device word back")
        return int(buf.value)


class QueuedFrames:
    """Device-resident leg: frames queued back to back, round-robin over a few canvases on their
    own CUDA streams (frames are independent -- the reference's CLI makes a canvas per frame -- so
    one frame's short kernels run in the gaps of another frame's matching); text and its length
    stay in HBM; nothing is waited for inside the window.  Every canvas is driven by a host thread
    of its own, as a caller with several frames in flight would: queueing one short frame costs
    about 18 us of host time (four kernel launches), which one thread alone cannot hide behind a
    27 us frame."""

    def __init__(self, ctx, wl, n_canvases):
        torch = ctx.torch
        self.ctx, self.wl = ctx, wl
        (self.w, self.h), (cw, ch) = wl["src"], wl["cells"]
        self.streams = [torch.cuda.Stream() for _ in range(n_canvases)]
        self.canvases = [capi.Canvas(ctx.lib, cw, ch, **wl["cfg"]) for _ in range(n_canvases)]
        for cv, st in zip(self.canvases, self.streams):
            ctx.lib.chafa_b200_canvas_set_stream(cv.handle, st.cuda_stream)
        self.last = None
        self.pool = ThreadPoolExecutor(max_workers=n_canvases) if n_canvases > 1 else None

    def frame(self, i, d_src, n_active=None):
        cv = self.canvases[i % (n_active or len(self.canvases))]
        cv.draw_device(d_src.data_ptr(), self.w, self.h)
        self.last = cv.print_device_enqueue()

    def _queue_share(self, c, ptrs):
        """Worker thread: the frames of canvas @c, in order."""
        self.ctx.lib.chafa_b200_set_device(self.ctx.local_rank)
        cv, last = self.canvases[c], None
        for ptr in ptrs:
            cv.draw_device(ptr, self.w, self.h)
            last = cv.print_device_enqueue()
        return last

    def window(self, d_srcs, first, n, n_active=None, collective=True):
        """Exactly @n frames between two events on the first stream (the other streams are fenced
        to it on both sides); barrier + synchronize before and after.  Returns milliseconds."""
        torch, ctx = self.ctx.torch, self.ctx
        sync = ctx.barrier if collective else torch.cuda.synchronize
        n_act = n_active or len(self.canvases)
        ptrs = [d_srcs[(first + i) % len(d_srcs)].data_ptr() for i in range(n)]
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main, others = self.streams[0], self.streams[1:n_act]
        with torch.cuda.stream(main):
            e0.record()
            for st in others:
                st.wait_event(e0)                # nothing of the timed frames starts before e0
            if n_act == 1 or self.pool is None:
                for i in range(n):
                    self.frame(first + i, d_srcs[(first + i) % len(d_srcs)], n_act)
            else:
                shares = [self.pool.submit(self._queue_share, c, [ptrs[i] for i in range(n) if (first + i) % n_act == c])
                          for c in range(n_act)]
                lasts = [f.result() for f in shares]
                if n:
                    self.last = lasts[(first + n - 1) % n_act]
            for st in others:
                done = torch.cuda.Event()
                done.record(st)
                main.wait_event(done)            # e1 follows the last kernel of every stream
            e1.record()
        sync()
        return e0.elapsed_time(e1)

    def last_length(self):
        self.ctx.torch.cuda.synchronize()
        return self.ctx.read_u64(self.last[1]) if self.last else 0

    def close(self):
        if self.pool is not None:
            self.pool.shutdown()
        for cv in self.canvases:
            cv.close()


class BandedStill:
    """Strong scaling of ONE still: every rank renders its band of cell rows of the same frame;
    the band lengths are all-gathered on the device and every rank's text goes straight to its
    place in rank 0's buffer by peer stores (NVLink).  No host round trip per frame."""

    def __init__(self, ctx, wl):
        torch, dist, lib = ctx.torch, ctx.dist, ctx.lib
        self.ctx, self.wl = ctx, wl
        (self.w, self.h), (cw, ch) = wl["src"], wl["cells"]
        self.stream = torch.cuda.Stream()
        self.canvas = capi.Canvas(lib, cw, ch, **wl["cfg"])
        lib.chafa_b200_canvas_set_stream(self.canvas.handle, self.stream.cuda_stream)
        self.first, self.n = shard.split_rows(ch, ctx.world)[ctx.rank]
        lib.chafa_b200_canvas_set_band(self.canvas.handle, self.first, self.n)
        self.capacity = cw * ch * 72 + 4096
        self.my_len = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.lens = torch.zeros(ctx.world, dtype=torch.int64, device="cuda")
        self.total = torch.zeros(1, dtype=torch.int64, device="cuda")
        # modes that normalise against a whole-image histogram: all-reduce it between the two draw phases
        self.hist = torch.zeros(2051, dtype=torch.int32, device="cuda")
        lib.chafa_b200_canvas_set_histogram_buffer(self.canvas.handle, self.hist.data_ptr())
        # rank 0 owns the assembled text; the others map it
        handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
        self.full = None
        if ctx.rank == 0:
            self.full = torch.zeros(self.capacity, dtype=torch.uint8, device="cuda")
            h = (C.c_ubyte * 64)()
            if not lib.lib.chafa_b200_ipc_export(self.full.data_ptr(), h):
                raise RuntimeError("cudaIpcGetMemHandle failed")
            handle.copy_(torch.tensor(list(h), dtype=torch.uint8))
        dist.broadcast(handle, src=0)
        torch.cuda.synchronize()
        if ctx.rank == 0:
            self.dest = self.full.data_ptr()
        else:
            h = (C.c_ubyte * 64)(*handle.cpu().tolist())
            self.dest = lib.lib.chafa_b200_ipc_open(h)
            if not self.dest:
                raise RuntimeError("cudaIpcOpenMemHandle failed: " + lib.chafa_b200_last_error().decode())

    def frame(self, d_src):
        ctx, lib, torch = self.ctx, self.ctx.lib, self.ctx.torch
        with torch.cuda.stream(self.stream):
            need = lib.chafa_b200_canvas_draw_begin_device(self.canvas.handle, capi.PIXEL_RGBA8_UNASSOCIATED,
                                                           d_src.data_ptr(), self.w, self.h, self.w * 4, None, None)
            if need:
                ctx.dist.all_reduce(self.hist[:2049])
            lib.chafa_b200_canvas_draw_finish(self.canvas.handle)
            self.canvas.print_device_enqueue()
            lib.lib.chafa_b200_canvas_copy_print_length(self.canvas.handle, self.my_len.data_ptr())
            ctx.dist.all_gather_into_tensor(self.lens, self.my_len)
            lib.lib.chafa_b200_canvas_band_scatter(self.canvas.handle, self.lens.data_ptr(), ctx.world, ctx.rank,
                                                   self.dest, self.capacity, self.total.data_ptr())

    def window(self, d_srcs, first, n):
        torch, ctx = self.ctx.torch, self.ctx
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            e0.record()
            for i in range(n):
                self.frame(d_srcs[(first + i) % len(d_srcs)])
            e1.record()
        ctx.barrier()
        return e0.elapsed_time(e1)

    def text(self):
        """Rank 0: the assembled text of the last frame."""
        self.ctx.barrier()
        if self.ctx.rank != 0:
            return None
        n = int(self.total.item())
        return self.full[:n].cpu().numpy().tobytes()

    def close(self):
        self.ctx.barrier()
        if self.ctx.rank != 0:
            self.ctx.lib.lib.chafa_b200_ipc_close(self.dest)
        self.canvas.close()


PRINT_THREADS = 4   # printing a sixel canvas waits for its encoder twice (length, text): a few callers overlap those waits


def print_all(ctx, pool, canvases):
    """chafa_canvas_print of every canvas (host GString each), from PRINT_THREADS threads; total bytes."""
    def some(k):
        ctx.lib.chafa_b200_set_device(ctx.local_rank)
        return sum(cv.print_len() for cv in canvases[k::PRINT_THREADS])
    return sum(pool.map(some, range(PRINT_THREADS)))


def replicas_leg(ctx, wl, d_srcs, n_canvases, frames_each, repeats):
    """Sixel diffusion: @n_canvases stills in flight.  Every frame is one serial chain on one SM;
    the chains of all the canvases are queued as ONE launch, a block per image
    (chafa_b200_canvases_draw_batch_device), then every canvas is printed (text to the host: the
    sixel header is assembled there).  Device-resident sources.  Returns ([window ms], bytes per frame)."""
    torch, lib = ctx.torch, ctx.lib
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    streams = [torch.cuda.Stream() for _ in range(n_canvases)]
    canvases = [capi.Canvas(lib, cw, ch, **wl["cfg"]) for _ in range(n_canvases)]
    for cv, st in zip(canvases, streams):
        lib.chafa_b200_canvas_set_stream(cv.handle, st.cuda_stream)
    out_bytes = 0

    pool = ThreadPoolExecutor(max_workers=PRINT_THREADS)

    def one_round(r):
        nonlocal out_bytes
        ptrs = [d_srcs[(t + r) % len(d_srcs)].data_ptr() for t in range(n_canvases)]
        if not capi.draw_batch_device(lib, canvases, ptrs, w, h):
            raise RuntimeError("batched draw failed: " + lib.chafa_b200_last_error().decode())
        out_bytes = print_all(ctx, pool, canvases) // n_canvases

    one_round(0)        # warm-up: allocations, tables
    windows = []
    for r in range(repeats):
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(frames_each):
            one_round(r * frames_each + i + 1)
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            torch.cuda.current_stream().wait_event(ev)
        e1.record()
        torch.cuda.synchronize()
        windows.append(e0.elapsed_time(e1))
    pool.shutdown()
    for cv in canvases:
        cv.close()
    return windows, out_bytes


def e2e_leg(ctx, wl, h_srcs, n_threads, steps, repeats, warm, band=None):
    """Host buffers through the reference-facing API: @n_threads host threads, one canvas each, as a
    caller rendering a stream of frames would run it (the reference is multi-threaded inside one
    call; here the parallelism a caller can add is across frames).  Every frame goes through
    chafa_canvas_draw_all_pixels with a host pointer (H2D inside) and chafa_canvas_print returning
    a host GString (D2H inside); thread t handles frames t, t + n_threads, ...  ctypes releases the
    GIL during the calls, so one thread's H2D copy overlaps the others' kernels and D2H.
    @h_srcs: host addresses.  Returns ([window ms], printed bytes per window)."""
    lib = ctx.lib
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    canvases = [capi.Canvas(lib, cw, ch, **wl["cfg"]) for _ in range(n_threads)]   # each on its own stream
    if band:
        for c in canvases:
            lib.chafa_b200_canvas_set_band(c.handle, band[0], band[1])
    d2h = [0] * n_threads
    n_warm = max(warm, n_threads)
    gate_start, gate_done = threading.Barrier(n_threads + 1), threading.Barrier(n_threads + 1)

    def frames(t, first, n):
        for i in range(first + t, first + n, n_threads):
            canvases[t].draw(h_srcs[i % len(h_srcs)], w, h)
            d2h[t] += canvases[t].print_len()      # host GString produced and freed; no Python copy

    def worker(t):
        try:
            lib.chafa_b200_set_device(ctx.local_rank)
            frames(t, 0, n_warm)
            for r in range(repeats):
                d2h[t] = 0
                gate_start.wait()
                frames(t, n_warm + r * steps, steps)
                gate_done.wait()
        except BaseException:
            gate_start.abort()      # a failed worker must not leave the others waiting
            gate_done.abort()
            raise

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    windows = []
    for r in range(repeats):
        ctx.barrier()
        gate_start.wait()
        t0 = time.perf_counter()
        gate_done.wait()
        ctx.torch.cuda.synchronize()
        windows.append((time.perf_counter() - t0) * 1e3)
    for th in threads:
        th.join()
    total = sum(d2h)
    for c in canvases:
        c.close()
    return windows, total


def e2e_batch_leg(ctx, wl, h_srcs, n_canvases, rounds, repeats):
    """Host buffers, many stills at once: chafa_b200_canvases_draw_batch (host pointers, H2D inside,
    the diffusion chains of all canvases in one launch), then chafa_canvas_print of every canvas (host
    GString, D2H inside).  One calling thread.  Returns ([window ms], printed bytes per window)."""
    lib = ctx.lib
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    canvases = [capi.Canvas(lib, cw, ch, **wl["cfg"]) for _ in range(n_canvases)]
    pool = ThreadPoolExecutor(max_workers=PRINT_THREADS)

    def one_round(r):
        ptrs = [h_srcs[(t + r) % len(h_srcs)] for t in range(n_canvases)]
        if not capi.draw_batch_device(lib, canvases, ptrs, w, h, on_device=False):
            raise RuntimeError("batched draw failed: " + lib.chafa_b200_last_error().decode())
        return print_all(ctx, pool, canvases)

    one_round(0)
    windows, d2h = [], 0
    for r in range(repeats):
        ctx.barrier()
        t0 = time.perf_counter()
        d2h = 0
        for i in range(rounds):
            d2h += one_round(1 + r * rounds + i)
        ctx.torch.cuda.synchronize()
        windows.append((time.perf_counter() - t0) * 1e3)
    pool.shutdown()
    for cv in canvases:
        cv.close()
    return windows, d2h


def kernel_times(ctx, canvas, run_frame, n):
    """Per-stage CUDA-event times of @n frames on @canvas (synchronises every step: not a timed leg)."""
    lib = ctx.lib.lib
    lib.chafa_b200_canvas_set_kernel_timing(canvas.handle, 1)
    for i in range(n):
        run_frame(i)
    ms = (C.c_double * 16)()
    cnt = (C.c_uint64 * 16)()
    lib.chafa_b200_canvas_get_kernel_times(canvas.handle, ms, cnt, 16)
    lib.chafa_b200_canvas_set_kernel_timing(canvas.handle, 0)
    return {STAGES[i]: (ms[i] / cnt[i] if cnt[i] else 0.0) for i in range(len(STAGES))}


def extra_c4_frames(ctx, repeats=3):
    """BASELINE configs[3]: the 300-frame 1080p animation, frames round-robin over the ranks, each
    rank's frames resident in its HBM and queued back to back.  frames/s = 300 / slowest rank."""
    torch = ctx.torch
    wl = WORKLOADS["c4_anim_300f_1080p"]
    w, h = wl["src"]
    n_frames = wl["frames"]
    frame = synth.animation(w, h, n_frames, seed=100)
    mine = list(shard.frames_of_rank(n_frames, ctx.world, ctx.rank))
    d_srcs = [torch.from_numpy(frame(f)).cuda() for f in mine]
    q = QueuedFrames(ctx, wl, dev_canvases(wl["cells"][0] * wl["cells"][1], ctx.world))
    q.window(d_srcs, 0, min(len(d_srcs), 6))        # warm-up
    windows = ctx.max_over_ranks([q.window(d_srcs, 0, len(d_srcs)) for _ in range(repeats)])
    q.close()
    ms = median(windows)
    return {"workload": "c4_anim_300f_1080p", "frames": n_frames, "frames_per_rank": len(mine),
            "frames_per_sec": n_frames / (ms * 1e-3), "ms_total": ms,
            "cells_per_sec": n_frames * wl["cells"][0] * wl["cells"][1] / (ms * 1e-3),
            "sharding": "frames round-robin over the ranks, no data-path collective", "windows": len(windows)}


def extra_c5_bands(ctx, steps=5, repeats=3):
    """BASELINE configs[4]: one 8K DIN99d still (every symbol scored for every cell).  At N > 1 the
    still is rendered as cell-row bands, one per rank, and assembled on rank 0 over NVLink; rank 0
    also renders it alone in the same run, so the speed-up is measured, not derived."""
    torch = ctx.torch
    wl = WORKLOADS["c5_8k_din99d_work9"]
    w, h = wl["src"]
    src = synth.noise(w, h, seed=100)
    d_srcs = [torch.from_numpy(src).cuda()]
    out = {"workload": "c5_8k_din99d_work9", "image": "noise", "steps": steps}
    full_text = None

    # one GPU alone (rank 0; the others wait)
    ctx.barrier()
    if ctx.rank == 0:
        q = QueuedFrames(ctx, wl, 1)
        q.window(d_srcs, 0, 2, collective=False)
        single = [q.window(d_srcs, 0, steps, collective=False) / steps for _ in range(repeats)]
        out["ms_per_frame_one_gpu"] = median(single)
        ptr, n = q.canvases[0].print_device()
        buf = (C.c_char * n)()
        ctx.lib.chafa_b200_copy_from_device(C.addressof(buf), ptr, n)
        full_text = bytes(buf)
        q.close()
    ctx.barrier()
    if ctx.world == 1:
        out["ms_per_frame"] = out["ms_per_frame_one_gpu"]
        out["n_bands"] = 1
        return out

    b = BandedStill(ctx, wl)
    b.window(d_srcs, 0, 2)        # warm-up
    windows = ctx.max_over_ranks([b.window(d_srcs, 0, steps) for _ in range(repeats)])
    text = b.text()
    b.close()
    out["n_bands"] = ctx.world
    out["ms_per_frame"] = median(windows) / steps
    if ctx.rank == 0:
        out["speedup_vs_one_gpu"] = out["ms_per_frame_one_gpu"] / out["ms_per_frame"]
        out["bands_equal_full"] = text == full_text
        out["assembly"] = "band lengths all-gathered on the device (NCCL), text stored to rank 0 over NVLink (CUDA IPC)"
    return out


class BandedSixelStill:
    """One sixel still as row bands, one per rank (bands lie on sixel rows: shard.split_sixel_rows).
    The generated palette needs the colour-cube bins of the WHOLE image: every rank samples its
    rows, the bins are summed with one NCCL all-reduce (exact 32-bit integer sums,
    chafa-palette.c:513-536), the float sums of bins beyond float's exact range travel from rank to
    rank in band order (one small broadcast per rank), and every rank then derives the same palette
    and quantises and prints its band."""

    def __init__(self, ctx, wl):
        torch, lib = ctx.torch, ctx.lib
        self.ctx, self.wl = ctx, wl
        (self.w, self.h), (cw, ch) = wl["src"], wl["cells"]
        self.stream = torch.cuda.Stream()
        self.canvas = capi.Canvas(lib, cw, ch, **wl["cfg"])
        lib.chafa_b200_canvas_set_stream(self.canvas.handle, self.stream.cuda_stream)
        cell_h = wl["cfg"].get("cell", (8, 8))[1]
        self.first, self.n = shard.split_sixel_rows(ch, cell_h, ctx.world)[ctx.rank]
        if self.n == 0:
            raise RuntimeError("more ranks than sixel bands")
        lib.chafa_b200_canvas_set_band(self.canvas.handle, self.first, self.n)
        n_bytes = lib.chafa_b200_canvas_exchange_buffer_bytes(self.canvas.handle)
        self.head = torch.zeros(n_bytes // 4, dtype=torch.int32, device="cuda")
        lib.chafa_b200_canvas_set_histogram_buffer(self.canvas.handle, self.head.data_ptr())
        self.reduced_words = 0
        self.text = b""

    def frame(self, d_src):
        ctx, lib, torch = self.ctx, self.ctx.lib, self.ctx.torch
        with torch.cuda.stream(self.stream):
            buf, n_words = C.c_void_p(), C.c_size_t()
            need = lib.chafa_b200_canvas_draw_begin_device(self.canvas.handle, capi.PIXEL_RGBA8_UNASSOCIATED,
                                                           d_src.data_ptr(), self.w, self.h, self.w * 4,
                                                           C.byref(buf), C.byref(n_words))
            if need:
                assert buf.value == self.head.data_ptr()
                self.reduced_words = n_words.value
                ctx.dist.all_reduce(self.head[:n_words.value])
                sums = self.head[n_words.value:].view(torch.float32)
                for owner in range(ctx.world):
                    if owner == ctx.rank:
                        lib.chafa_b200_canvas_draw_float_sums_device(self.canvas.handle, None, None)
                    ctx.dist.broadcast(sums, src=owner)
            lib.chafa_b200_canvas_draw_finish(self.canvas.handle)
            self.text = self.canvas.print()

    def window(self, d_srcs, first, n):
        torch, ctx = self.ctx.torch, self.ctx
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            e0.record()
            for i in range(n):
                self.frame(d_srcs[(first + i) % len(d_srcs)])
            e1.record()
        ctx.barrier()
        return e0.elapsed_time(e1)

    def gathered_text(self):
        """Rank 0: the bands' texts concatenated (after the timed region; for the check)."""
        ctx = self.ctx
        parts = [None] * ctx.world
        ctx.dist.all_gather_object(parts, self.text)
        return b"".join(parts) if ctx.rank == 0 else None

    def close(self):
        self.ctx.barrier()
        self.canvas.close()


def extra_sixel_bands(ctx, steps=5, repeats=3):
    """A 4K sixel still with a generated palette as row bands over the ranks: the one place the
    path has a collective on its data path (the all-reduce of the palette's colour bins).  Rank 0
    also renders the still alone in the same run."""
    torch = ctx.torch
    wl = WORKLOADS["c3n_4k_sixel_noise"]
    w, h = wl["src"]
    src = synth.shapes(w, h, seed=100)
    d_srcs = [torch.from_numpy(src).cuda()]
    out = {"workload": "c3n_4k_sixel_noise", "image": "shapes", "steps": steps}
    full_text = None

    def alone():
        cv = capi.Canvas(ctx.lib, *wl["cells"], **wl["cfg"])
        st = torch.cuda.Stream()
        ctx.lib.chafa_b200_canvas_set_stream(cv.handle, st.cuda_stream)
        text, times = b"", []
        for r in range(repeats + 1):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record()
                for i in range(steps):
                    cv.draw_device(d_srcs[0].data_ptr(), w, h)
                    text = cv.print()
                e1.record()
            torch.cuda.synchronize()
            if r:
                times.append(e0.elapsed_time(e1) / steps)
        cv.close()
        return text, median(times)

    ctx.barrier()
    if ctx.rank == 0:
        full_text, out["ms_per_frame_one_gpu"] = alone()
    ctx.barrier()
    if ctx.world == 1:
        out["ms_per_frame"] = out["ms_per_frame_one_gpu"]
        out["n_bands"] = 1
        return out

    b = BandedSixelStill(ctx, wl)
    b.window(d_srcs, 0, 2)        # warm-up
    windows = ctx.max_over_ranks([b.window(d_srcs, 0, steps) for _ in range(repeats)])
    text = b.gathered_text()
    out["n_bands"] = ctx.world
    out["ms_per_frame"] = median(windows) / steps
    out["all_reduce_bytes"] = b.reduced_words * 4
    b.close()
    if ctx.rank == 0:
        out["speedup_vs_one_gpu"] = out["ms_per_frame_one_gpu"] / out["ms_per_frame"]
        out["bands_equal_full"] = text == full_text
        out["exchange"] = ("colour-cube bins of the palette all-reduced (NCCL, exact int32 sums), float sums relayed "
                           "in band order (one broadcast per rank); every rank prints its band to host memory, "
                           "texts gathered after the timed region for the check")
    return out


def bench_b200(args, wl):
    ctx = Ctx(args)
    torch, lib, rank, world = ctx.torch, ctx.lib, ctx.rank, ctx.world
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    n_cells = cw * ch
    chain = is_serial_chain(wl)
    bands = args.sharding == "bands" and ctx.distributed and not is_sixel(wl)
    animation = "frames" in wl
    steps = wl["frames"] // world if animation else args.steps

    if animation:
        frame = synth.animation(w, h, wl["frames"], seed=100)
        srcs = [frame(f) for f in shard.frames_of_rank(wl["frames"], world, rank)]
    else:
        srcs = make_sources(wl, args.image, N_ROTATE)
    d_srcs = [torch.from_numpy(s).cuda() for s in srcs]           # resident in HBM before timing
    n_host = min(len(srcs), N_ROTATE)
    h_srcs = [torch.from_numpy(s).pin_memory() for s in srcs[:n_host]]     # pinned host frames for the e2e leg
    torch.cuda.synchronize()

    sampler = ClockSampler(ctx.local_rank)
    lat_windows, dev_windows, out_len, launches = [], [], 0, 0
    main = still = None

    def timed_windows(window):
        nonlocal launches
        for r in range(args.repeats):
            l0 = lib.chafa_b200_launch_count()
            dev_windows.append(window(args.warmup + r * steps, steps))
            launches = lib.chafa_b200_launch_count() - l0

    # ---- device-resident leg ----
    if chain:
        frames_each = 1         # one round of N_REPLICAS stills per window
        if rank == 0:
            sampler.start()
        l0 = lib.chafa_b200_launch_count()
        windows, out_len = replicas_leg(ctx, wl, d_srcs, N_REPLICAS, frames_each, args.repeats)
        dev_windows.extend(windows)
        frames_per_window = N_REPLICAS * frames_each
        launches = (lib.chafa_b200_launch_count() - l0) * frames_each // (frames_each * args.repeats + 1)
        device_leg = (f"{N_REPLICAS} stills in flight, their diffusion chains in one launch (a block = an SM per "
                      f"image), {frames_each} round(s) per window")
        main = QueuedFrames(ctx, wl, 1)
    elif bands:
        still = BandedStill(ctx, wl)
        still.window(d_srcs, 0, max(args.warmup, 3))
        if rank == 0:
            sampler.start()
        timed_windows(lambda first, n: still.window(d_srcs, first, n))
        frames_per_window = steps
        device_leg = "one canvas per rank rendering its band; lengths all-gathered, text stored to rank 0 over NVLink"
        text = still.text()
        out_len = len(text) if text is not None else 0
    elif is_sixel(wl):
        # sixel, parallel dithers: the text is assembled with a host-side header: host print per frame
        main = QueuedFrames(ctx, wl, 1)
        canvas, st = main.canvases[0], main.streams[0]

        def sixel_window(first, n):
            nonlocal out_len
            ctx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(st):
                e0.record()
                for i in range(n):
                    canvas.draw_device(d_srcs[(first + i) % len(d_srcs)].data_ptr(), w, h)
                    out_len = canvas.print_len()
                e1.record()
            ctx.barrier()
            return e0.elapsed_time(e1)

        sixel_window(0, max(args.warmup, 3))
        if rank == 0:
            sampler.start()
        timed_windows(sixel_window)
        frames_per_window = steps
        device_leg = "one canvas, frames back to back, text returned to the host (header assembled there)"
    else:
        # (frames of more than ~200 k cells fill the GPU on their own: measured no gain, C5 loses 5 %)
        n_canvases = dev_canvases(n_cells, world)
        main = QueuedFrames(ctx, wl, n_canvases)
        main.window(d_srcs, 0, max(args.warmup, n_canvases))
        if rank == 0:
            sampler.start()
        timed_windows(lambda first, n: main.window(d_srcs, first, n))
        out_len = main.last_length()
        if n_canvases > 1:      # single-canvas latency: the same frames on ONE canvas / stream
            lat_windows = [main.window(d_srcs, args.warmup + r * steps, steps, n_active=1)
                           for r in range(min(args.repeats, 3))]
        frames_per_window = steps
        device_leg = (f"{n_canvases} canvas(es) / CUDA stream(s), one host thread each, frames round-robin, "
                      "queued back to back")

    # ---- end-to-end leg: host buffers through the reference-facing API ----
    # all ranks share one host: keep the total number of feeding threads near twice its cores
    n_threads = max(2, min(E2E_THREADS, (2 * host_cpus()) // world))
    band = (still.first, still.n) if bands else None
    if chain:
        # Diffusion chains: a caller with many stills hands them over at once (the chains share a launch)
        e2e_steps = N_REPLICAS
        e2e_repeats = min(args.repeats, 3)
        e2e_windows, d2h = e2e_batch_leg(ctx, wl, [t.data_ptr() for t in h_srcs], N_REPLICAS, 1, e2e_repeats)
        pg_windows, _ = e2e_batch_leg(ctx, wl, [s.ctypes.data for s in srcs[:n_host]], N_REPLICAS, 1, e2e_repeats)
        e2e_pipeline = (f"one host thread: chafa_b200_canvases_draw_batch of {N_REPLICAS} stills (host images), then "
                        f"chafa_canvas_print of each from {PRINT_THREADS} threads")
        n_threads = 1
    else:
        e2e_steps = steps
        e2e_repeats = args.repeats
        e2e_windows, d2h = e2e_leg(ctx, wl, [t.data_ptr() for t in h_srcs], n_threads, e2e_steps, e2e_repeats,
                                   args.warmup, band)
        # the same with the frames in ordinary pageable memory (what a drop-in caller hands over)
        pg_windows, _ = e2e_leg(ctx, wl, [s.ctypes.data for s in srcs[:n_host]], n_threads, e2e_steps,
                                min(e2e_repeats, 3), args.warmup, band)
        e2e_pipeline = f"{n_threads} host threads, one canvas each, frames interleaved, pinned source frames"
    clocks = sampler.stop() if rank == 0 else None      # sampled across the timed legs
    ctx.barrier()

    # ---- per-kernel pass (separate from the timed legs: it synchronises every step) ----
    stage_ms = {k: 0.0 for k in STAGES}
    if main is not None:
        cv = main.canvases[0]

        def one(i):
            cv.draw_device(d_srcs[i % len(d_srcs)].data_ptr(), w, h)
            if is_sixel(wl):
                cv.print_len()
            else:
                cv.print_device_enqueue()
        stage_ms = kernel_times(ctx, cv, one, 2 if chain else min(max(steps, 5), 20))

    # every window: max over ranks
    allw = ctx.max_over_ranks(dev_windows + e2e_windows + pg_windows + lat_windows)
    a, b, c = len(dev_windows), len(e2e_windows), len(pg_windows)
    dev_windows, e2e_windows, pg_windows, lat_windows = allw[:a], allw[a:a + b], allw[a + b:a + b + c], allw[a + b + c:]
    dev_ms, e2e_ms, pg_ms = median(dev_windows), median(e2e_windows), median(pg_windows)

    # ---- the north-star multi-GPU splits, on the default line (every N, so the 1 -> 8 curve lands in SCALE) ----
    extras = {}
    first_src = srcs[0]
    if args.workload == "c2_4k_256c_ordered_allwide" and not bands and not args.no_extras:
        if main is not None:
            main.close()
            main = None
        del d_srcs, h_srcs
        torch.cuda.empty_cache()
        extras["c4_frames"] = extra_c4_frames(ctx)
        torch.cuda.empty_cache()
        extras["strong_c5_bands"] = extra_c5_bands(ctx)
        extras["sixel_bands"] = extra_sixel_bands(ctx)

    if rank == 0:
        peak, peak_src = read_peaks()
        ranks_share = 1 if bands else world
        value = n_cells * frames_per_window * ranks_share / (dev_ms * 1e-3)
        e2e_cells = n_cells * e2e_steps * ranks_share
        # Dominant kernel and its algorithmic bytes (DESIGN.md "Kernels"): the matcher reads the
        # prepared pixmap once (Wpx*Hpx*4) and writes one 20-byte evaluation record per cell.
        kernel_stages = {k: v for k, v in stage_ms.items() if k not in ("h2d", "d2h")}
        top = max(kernel_stages, key=kernel_stages.get)
        wpx, hpx = cw * 8, ch * 8
        # stage "match" = narrow (+ wide, when the tensor-core kernel or the exhaustive kernel handles
        # both in one launch: then "match_wide" stays 0); "print_emit" = the single-launch printer
        fused_wide = stage_ms.get("match_wide", 0.0) == 0.0 and "wide" in str(wl["cfg"].get("symbols", ""))
        alg = {
            "scale": w * h * 4 + wpx * hpx * 4,
            "prep": 2 * wpx * hpx * 4,
            "match": wpx * hpx * 4 + n_cells * (40 if fused_wide else 20),
            "match_wide": wpx * hpx * 4 + n_cells * 20,
            "resolve": n_cells * (20 + 20 + 12),
            "print_measure": n_cells * (12 + 4),
            "print_emit": n_cells * 12 + out_len,
        }[top]
        achieved = alg / (stage_ms[top] * 1e-3) / 1e9 if stage_ms[top] > 0 else 0.0
        traffic, issue = None, None
        try:   # DRAM bytes / warp instructions per launch of that kernel from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                prof = json.load(f)
            traffic = prof.get(args.workload, {}).get(top)
            n_inst = prof.get(args.workload + "_warp_instructions", {}).get(top)
            if n_inst and stage_ms[top] > 0:
                # the axis this kernel is actually bound on: warp-instruction issue slots
                # (4 schedulers per SM x 148 SMs x SM clock); instruction count per launch from ncu
                sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
                peak_issue = 148 * 4 * sm_mhz * 1e6
                ach = n_inst / (stage_ms[top] * 1e-3)
                issue = {"bound": "warp-instruction issue", "kernel": top, "warp_instructions_per_launch": n_inst,
                         "warp_instructions_per_cell": round(n_inst / n_cells, 1), "achieved": ach,
                         "peak": peak_issue, "unit": "warp-inst/s", "frac": ach / peak_issue}
        except Exception:  # noqa: BLE001
            pass
        line = {
            "metric": "cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": world, "steps": frames_per_window,
            "warmup": args.warmup, "ms_per_step": dev_ms / frames_per_window, "higher_is_better": True,
            "scaling": "strong" if bands else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": bench_config(args, wl),
            "run": {"sharding": ("cell-row bands of one still per rank, text assembled on rank 0 over NVLink"
                                 if bands else "frames per rank, no data-path collective"),
                    "device_leg": device_leg,
                    "l2": f"{len(srcs)} distinct source frames rotated ({len(srcs) * w * h * 4 / 1e6:.0f} MB > 126 MB L2)",
                    "printed_bytes_per_frame": out_len},
            "repeats": {"windows": len(dev_windows), "statistic": "median window; every window is the max over ranks",
                        "ms_per_step_min": min(dev_windows) / frames_per_window,
                        "ms_per_step_median": dev_ms / frames_per_window,
                        "ms_per_step_max": max(dev_windows) / frames_per_window,
                        "e2e_ms_per_step_min": min(e2e_windows) / e2e_steps,
                        "e2e_ms_per_step_max": max(e2e_windows) / e2e_steps},
            "frames_per_sec": frames_per_window * ranks_share / (dev_ms * 1e-3),
            "e2e": {"value": e2e_cells / (e2e_ms * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": w * h * 4,
                    "d2h_bytes_per_step": d2h // max(e2e_steps, 1), "ms_per_step": e2e_ms / e2e_steps,
                    "pipeline": e2e_pipeline,
                    "pageable": {"value": e2e_cells / (pg_ms * 1e-3), "unit": "cells/s", "ms_per_step": pg_ms / e2e_steps,
                                 "note": "the same with the source frames in ordinary pageable memory, as a drop-in "
                                         "caller has them"},
                    # what the leg asks of the host: every rank's frames cross the host's memory and its PCIe root
                    "h2d_gb_per_s_all_ranks": (w * h * 4) * e2e_steps * ranks_share / (e2e_ms * 1e-3) / 1e9,
                    "host": host_facts(n_threads, world)},
            "gpu_launches": int(launches),
            "kernel_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg,
                         "note": "matching is bound by warp-instruction issue, not by HBM (SURVEY.md finding 10): "
                                 "see roofline_issue; the HBM fraction is reported because the contract asks for it"},
            "clocks": clocks,
        }
        if issue:
            line["roofline_issue"] = issue
        if lat_windows:
            line["single_canvas"] = {"ms_per_step": median(lat_windows) / frames_per_window,
                                     "note": "the same frames on ONE canvas / CUDA stream, queued back to back"}
        if chain:
            line["single_canvas"] = {"ms_per_step": sum(stage_ms.values()),
                                     "note": "one frame alone on one canvas (the serial chain's latency): sum of the "
                                             "per-stage CUDA events"}
        line.update(extras)
        if not args.no_cpu_baseline and not bands:
            line["parity_check"] = parity_check(lib, wl, first_src)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(args, wl)
        print(json.dumps(line), flush=True)

    if still is not None:
        still.close()
    if main is not None:
        main.close()
    if ctx.distributed:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2_4k_256c_ordered_allwide", choices=sorted(WORKLOADS))
    ap.add_argument("--image", default="noise", choices=sorted(synth.KINDS))
    ap.add_argument("--repeats", type=int, default=5,
                    help="timed windows of --steps frames each; the line reports the median window")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="default workload: skip the c4_frames / strong_c5_bands measurements")
    ap.add_argument("--sharding", default="frames", choices=["frames", "bands"],
                    help="N > 1: frames = every rank renders its own frames (weak scaling, default); "
                         "bands = all ranks render row bands of the same still (strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    wl = WORKLOADS[args.workload]

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        bench_reference(args, wl)
    else:
        bench_b200(args, wl)


if __name__ == "__main__":
    main()
