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
             back to back on the canvas stream, CUDA events around them
  e2e        cells/s through the reference-facing C API with HOST buffers: H2D of the frame from
             pinned memory and D2H of the printed text inside the timed region
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event duration vs measured HBM peak
  cpu_baseline  the compiled reference (oracle/_ref) on this box's host cores, bounded sample

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
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

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
    # BASELINE.json configs[4] on one GPU
    "c5_8k_din99d_work9": dict(
        src=(7680, 4320), cells=(960, 540),
        cfg=dict(mode=capi.CANVAS_MODE_INDEXED_256, symbols="all", work=1.0, color_space=capi.COLOR_SPACE_DIN99D),
        desc="7680x4320 RGBA8 -> 960x540 cells, DIN99d, INDEXED_256, symbols all, work 1.0 (exhaustive)"),
    # BASELINE.json configs[2]: one diffusion chain per image (serial by construction)
    "c3_1080p_sixel_fs": dict(
        src=(1920, 1080), cells=(240, 135),
        cfg=dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_DIFFUSION, grain=(1, 1)),
        desc="1920x1080 RGBA8 -> 1920x1080 sixels (240x135 cells of 8x8), 256-colour generated palette, "
             "Floyd-Steinberg"),
    "c3n_1080p_sixel_noise": dict(
        src=(1920, 1080), cells=(240, 135),
        cfg=dict(pixel_mode=capi.PIXEL_MODE_SIXELS, symbols=None, dither=capi.DITHER_NOISE, grain=(1, 1)),
        desc="1920x1080 RGBA8 -> 1920x1080 sixels, 256-colour generated palette, blue-noise dither"),
}
STAGES = ["h2d", "scale", "prep", "match", "match_wide", "resolve", "print_measure", "print_emit", "d2h"]
E2E_THREADS = 8   # host threads of the e2e leg on one rank; fewer per rank when ranks share the host (see bench_b200)
N_DEV_CANVASES = 3   # canvases (CUDA streams) the device-resident leg spreads its frames over
N_ROTATE = 8   # distinct source images cycled through so every step reads a cold (non-L2-resident) frame


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


def make_sources(kind, w, h, n):
    return [synth.KINDS[kind](w, h, seed=100 + i) for i in range(n)]


def bench_config(args, wl):
    """The `config` object: identical on the product arm and the reference arm."""
    return {"workload": args.workload, "desc": wl["desc"], "image": args.image}


def parity_check(lib, args, wl, src):
    """What was timed is what is checked: one source frame of the timed workload through the
    product's host API and through the compiled reference (oracle/_ref), outputs compared."""
    import numpy as np
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
    if wl["cfg"].get("pixel_mode", capi.PIXEL_MODE_SYMBOLS) == capi.PIXEL_MODE_SYMBOLS:
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


def bench_reference(args, wl):
    """The reference's own CPU implementation (oracle/_ref) on all host threads."""
    ref = capi.load_reference()
    ref.chafa_set_n_threads(-1)
    threads = ref.chafa_get_n_actual_threads()
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    srcs = make_sources(args.image, w, h, 2)
    canvas = capi.Canvas(ref, cw, ch, **wl["cfg"])
    n_cells = cw * ch

    def step(i):
        canvas.draw(srcs[i % len(srcs)], w, h)
        return canvas.print_len()

    for i in range(args.warmup):
        step(i)
    windows = []
    for r in range(args.repeats):
        t0 = time.perf_counter()
        for i in range(args.steps):
            step(i)
        windows.append(time.perf_counter() - t0)
    canvas.close()
    dt = sorted(windows)[len(windows) // 2]
    value = n_cells * args.steps / dt
    sample = (f"{args.repeats} windows of {args.steps} full frames ({args.warmup} warm-up), median window, "
              f"one canvas reused, {threads} threads")
    line = {
        "impl": "reference", "metric": "cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(args, wl),
        "repeats": {"windows": args.repeats, "ms_per_step_min": min(windows) / args.steps * 1e3,
                    "ms_per_step_median": dt / args.steps * 1e3, "ms_per_step_max": max(windows) / args.steps * 1e3},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_sec": args.steps / dt, "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(args, wl, budget_s=12.0):
    """Reference on host cores, bounded: frames until ~budget_s of CPU wall time."""
    try:
        ref = capi.load_reference()
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "cells/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    ref.chafa_set_n_threads(-1)
    threads = ref.chafa_get_n_actual_threads()
    (w, h), (cw, ch) = wl["src"], wl["cells"]
    src = make_sources(args.image, w, h, 1)[0]
    canvas = capi.Canvas(ref, cw, ch, **wl["cfg"])
    canvas.draw(src, w, h)
    canvas.print_len()
    n, t0 = 0, time.perf_counter()
    while True:
        canvas.draw(src, w, h)
        canvas.print_len()
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 50:
            break
    canvas.close()
    return {"value": cw * ch * n / dt, "unit": "cells/s", "cores": threads, "kind": "reference",
            "sample": f"{n} full frames of the same workload after 1 warm-up, {threads} threads "
                      f"({os.cpu_count()} host cpus), {dt:.1f} s"}


def bench_b200(args, wl):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lib = capi.load_product()
    if not lib.chafa_b200_available():
        raise RuntimeError("libchafa_b200: no CUDA device (there is no CPU fallback)")
    lib.chafa_b200_set_device(local_rank)

    (w, h), (cw, ch) = wl["src"], wl["cells"]
    n_cells = cw * ch
    srcs = make_sources(args.image, w, h, N_ROTATE)
    stream = torch.cuda.Stream()
    canvas = capi.Canvas(lib, cw, ch, **wl["cfg"])
    lib.chafa_b200_canvas_set_stream(canvas.handle, stream.cuda_stream)
    # Frames are independent (the reference's CLI makes a canvas per frame): the device-resident leg
    # renders the stream of frames round-robin over N_DEV_CANVASES canvases, each on its own CUDA
    # stream, so that one frame's short kernels (scale, resolve, print) run in the gaps of another
    # frame's matching.  Band sharding (one still) and sixels keep the single canvas.
    extra_streams, extra_canvases = [], []

    d_srcs = [torch.from_numpy(s).cuda() for s in srcs]           # resident in HBM before timing
    h_srcs = [torch.from_numpy(s).pin_memory() for s in srcs]     # pinned host frames for the e2e leg
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    is_sixel = wl["cfg"].get("pixel_mode") == capi.PIXEL_MODE_SIXELS
    bands = args.sharding == "bands" and distributed and not is_sixel
    if bands:
        # Strong scaling of ONE still: every rank renders its band of cell rows of the same
        # frame and rank 0 gathers the text over NCCL (the only exchange step of the path).
        band_first, band_n = shard.split_rows(ch, world)[rank]
        lib.chafa_b200_canvas_set_band(canvas.handle, band_first, band_n)
        slot = max(shard.split_rows(ch, world)[0][1] * cw * 72 + 4096, 4096)
        band_buf = torch.zeros(slot, dtype=torch.uint8, device="cuda")
        gather_bufs = [torch.zeros(slot, dtype=torch.uint8, device="cuda") for _ in range(world)] if rank == 0 else None
        band_len = torch.zeros(1, dtype=torch.int64, device="cuda")
        all_lens = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)] if rank == 0 else None
        # modes that normalise against a whole-image histogram: all-reduce it between the two draw phases
        hist = torch.zeros(2051, dtype=torch.int32, device="cuda")
        lib.chafa_b200_canvas_set_histogram_buffer(canvas.handle, hist.data_ptr())

    # (frames of more than ~200 k cells fill the GPU on their own: measured no gain, C5 loses 5 %)
    if not bands and not is_sixel and n_cells <= 200000:
        for _ in range(N_DEV_CANVASES - 1):
            st = torch.cuda.Stream()
            cv = capi.Canvas(lib, cw, ch, **wl["cfg"])
            lib.chafa_b200_canvas_set_stream(cv.handle, st.cuda_stream)
            extra_streams.append(st)
            extra_canvases.append(cv)
    dev_canvases = [canvas] + extra_canvases
    spread = [False]        # the per-kernel pass at the end runs on the first canvas only

    def step_device(i):
        if bands:
            with torch.cuda.stream(stream):
                need = lib.chafa_b200_canvas_draw_begin_device(canvas.handle, capi.PIXEL_RGBA8_UNASSOCIATED,
                                                               d_srcs[i % N_ROTATE].data_ptr(), w, h, w * 4, None, None)
                if need:
                    dist.all_reduce(hist[:2049])
                lib.chafa_b200_canvas_draw_finish(canvas.handle)
        else:
            cvs = dev_canvases[i % len(dev_canvases)] if spread[0] else canvas
            cvs.draw_device(d_srcs[i % N_ROTATE].data_ptr(), w, h)
        if bands:
            n = lib.chafa_b200_canvas_print_into(canvas.handle, None, band_buf.data_ptr(), slot)
            band_len[0] = n
            dist.gather(band_len, all_lens, dst=0)
            dist.gather(band_buf, gather_bufs, dst=0)
            return int(n)
        if is_sixel:
            return len(canvas.print())      # the sixel text is assembled with a host-side header
        # text and its length stay in HBM; nothing is waited for, frames queue back to back
        last_print[0] = cvs.print_device_enqueue()
        return None

    last_print = [None]

    def read_out_len(n):
        """Length of the last frame's text, read back after the timed region."""
        if n is not None or last_print[0] is None:
            return n
        torch.cuda.synchronize()
        buf = C.c_uint64(0)
        if not lib.chafa_b200_copy_from_device(C.addressof(buf), last_print[0][1], 8):
            raise RuntimeError("could not read the printed length back")
        return int(buf.value)

    # ---- device-resident leg ----
    spread[0] = True
    for i in range(max(args.warmup, len(dev_canvases))):
        out_len = step_device(i)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()

    def timed_window(first):
        """Exactly args.steps frames between two events on the canvas stream (the other streams
        are fenced to it on both sides); barrier + synchronize before and after."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = None
        with torch.cuda.stream(stream):
            e0.record()
            for st in extra_streams:
                st.wait_event(e0)                # nothing of the timed frames starts before e0
            for i in range(args.steps):
                n = step_device(first + i)
            for st in extra_streams:
                done = torch.cuda.Event()
                done.record(st)
                stream.wait_event(done)          # e1 follows the last kernel of every stream
            e1.record()
        barrier()
        return e0.elapsed_time(e1), n

    dev_windows = []
    for r in range(args.repeats):
        launches0 = lib.chafa_b200_launch_count()
        ms_r, out_len = timed_window(args.warmup + r * args.steps)
        dev_windows.append(ms_r)
        launches = lib.chafa_b200_launch_count() - launches0
    out_len = read_out_len(out_len)

    # single-canvas latency: the same frames on ONE canvas / stream, queued back to back
    spread[0] = False
    saved_extra, extra_streams[:] = list(extra_streams), []
    lat_windows = [timed_window(args.warmup + r * args.steps)[0] for r in range(min(args.repeats, 3))] \
        if saved_extra else []
    extra_streams[:] = saved_extra

    # ---- end-to-end leg: host buffers through the reference-facing API ----
    # E2E_THREADS host threads, one canvas each, as a caller rendering a stream of frames would
    # run it (the reference is multi-threaded inside one call; here the parallelism a caller
    # can add is across frames).  Every frame goes through chafa_canvas_draw_all_pixels with a
    # host pointer (H2D inside) and chafa_canvas_print returning a host GString (D2H inside);
    # thread t handles frames t, t + E2E_THREADS, ...  ctypes releases the GIL during the calls,
    # so one thread's H2D copy overlaps the other's kernels and D2H.
    # all ranks share one host: keep the total number of feeding threads near twice its cores
    try:
        host_cpus = len(os.sched_getaffinity(0))
    except AttributeError:
        host_cpus = os.cpu_count() or 16
    E2E_THREADS = max(2, min(globals()["E2E_THREADS"], (2 * host_cpus) // world))
    canvases = [capi.Canvas(lib, cw, ch, **wl["cfg"]) for _ in range(E2E_THREADS)]   # each on its own stream
    if bands:
        for c in canvases:
            lib.chafa_b200_canvas_set_band(c.handle, band_first, band_n)
    d2h_total = [0] * E2E_THREADS

    # The worker threads live for the whole leg (a caller's pool): they run the warm-up frames,
    # meet at a barrier, and the clock starts when the barrier opens.
    n_warm = max(args.warmup, E2E_THREADS)
    gate_start = threading.Barrier(E2E_THREADS + 1)
    gate_done = threading.Barrier(E2E_THREADS + 1)

    def e2e_frames(t, first, n):
        for i in range(first + t, first + n, E2E_THREADS):
            canvases[t].draw(h_srcs[i % N_ROTATE].data_ptr(), w, h)
            d2h_total[t] += canvases[t].print_len()      # host GString produced and freed; no Python copy

    def e2e_worker(t):
        try:
            lib.chafa_b200_set_device(local_rank)
            e2e_frames(t, 0, n_warm)
            for r in range(args.repeats):
                d2h_total[t] = 0
                gate_start.wait()
                e2e_frames(t, n_warm + r * args.steps, args.steps)
                gate_done.wait()
        except BaseException:
            gate_start.abort()      # a failed worker must not leave the others waiting
            gate_done.abort()
            raise

    threads = [threading.Thread(target=e2e_worker, args=(t,)) for t in range(E2E_THREADS)]
    for th in threads:
        th.start()
    e2e_windows = []
    for r in range(args.repeats):
        barrier()
        gate_start.wait()
        t0 = time.perf_counter()
        gate_done.wait()
        torch.cuda.synchronize()
        e2e_windows.append((time.perf_counter() - t0) * 1e3)
    for th in threads:
        th.join()
    d2h = sum(d2h_total)
    clocks = sampler.stop() if rank == 0 else None      # sampled across both timed legs
    barrier()
    for c in canvases:
        c.close()

    # ---- per-kernel pass (separate from the timed legs: it synchronises every step) ----
    lib.lib.chafa_b200_canvas_set_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    lib.lib.chafa_b200_canvas_get_kernel_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.lib.chafa_b200_canvas_set_kernel_timing(canvas.handle, 1)
    for i in range(max(args.steps, 5)):
        step_device(i)
    ms = (C.c_double * 16)()
    cnt = (C.c_uint64 * 16)()
    lib.lib.chafa_b200_canvas_get_kernel_times(canvas.handle, ms, cnt, 16)
    lib.lib.chafa_b200_canvas_set_kernel_timing(canvas.handle, 0)
    stage_ms = {STAGES[i]: (ms[i] / cnt[i] if cnt[i] else 0.0) for i in range(len(STAGES))}

    if distributed:     # every window: max over ranks
        t = torch.tensor(dev_windows + e2e_windows + lat_windows, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t.tolist()
        dev_windows, e2e_windows = t[:len(dev_windows)], t[len(dev_windows):len(dev_windows) + len(e2e_windows)]
        lat_windows = t[len(dev_windows) + len(e2e_windows):]
    median = lambda xs: sorted(xs)[len(xs) // 2]      # noqa: E731
    dev_ms, e2e_ms = median(dev_windows), median(e2e_windows)

    if rank == 0:
        peak, peak_src = read_peaks()
        total_cells = n_cells * args.steps * (1 if bands else world)
        value = total_cells / (dev_ms * 1e-3)
        e2e = total_cells / (e2e_ms * 1e-3)
        # Dominant kernel and its algorithmic bytes (DESIGN.md "Kernels"): the matcher reads the
        # prepared pixmap once (Wpx*Hpx*4) and writes one 20-byte evaluation record per cell.
        kernel_stages = {k: v for k, v in stage_ms.items() if k not in ("h2d", "d2h")}
        top = max(kernel_stages, key=kernel_stages.get)
        wpx, hpx = cw * 8, ch * 8
        # stage "match" = narrow (+ wide, when the tensor-core kernel or the exhaustive kernel handles
        # both in one launch: then "match_wide" stays 0); "print_emit" = the single-launch printer
        # ("print_measure" stays 0)
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
            "metric": "cells_per_sec", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if bands else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": bench_config(args, wl),
            "run": {"sharding": ("cell-row bands of one still per rank, text gathered to rank 0 over NCCL"
                                 if bands else "frames per rank, no data-path collective"),
                    "device_leg": f"{len(dev_canvases)} canvas(es) / CUDA stream(s), frames round-robin, queued back to back",
                    "l2": f"{N_ROTATE} distinct source frames rotated ({N_ROTATE * w * h * 4 / 1e6:.0f} MB > 126 MB L2)",
                    "printed_bytes_per_frame": out_len},
            "repeats": {"windows": args.repeats, "statistic": "median window; every window is the max over ranks",
                        "ms_per_step_min": min(dev_windows) / args.steps, "ms_per_step_median": dev_ms / args.steps,
                        "ms_per_step_max": max(dev_windows) / args.steps,
                        "e2e_ms_per_step_min": min(e2e_windows) / args.steps,
                        "e2e_ms_per_step_max": max(e2e_windows) / args.steps},
            "frames_per_sec": args.steps * (1 if bands else world) / (dev_ms * 1e-3),
            "e2e": {"value": e2e, "unit": "cells/s", "h2d_bytes_per_step": w * h * 4,
                    "d2h_bytes_per_step": d2h // max(args.steps, 1), "ms_per_step": e2e_ms / args.steps,
                    "pipeline": f"{E2E_THREADS} host threads, one canvas each, frames interleaved"},
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
            line["single_canvas"] = {"ms_per_step": median(lat_windows) / args.steps,
                                     "note": "the same frames on ONE canvas / CUDA stream, queued back to back"}
        if not args.no_cpu_baseline and not bands:
            line["parity_check"] = parity_check(lib, args, wl, srcs[0])
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(args, wl)
        print(json.dumps(line), flush=True)

    canvas.close()
    for cv in extra_canvases:
        cv.close()
    if distributed:
        dist.destroy_process_group()


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
