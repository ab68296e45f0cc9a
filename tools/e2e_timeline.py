"""Timeline of the threaded end-to-end path: per-call wall times per thread."""
import sys, time, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chafa_b200 import capi, synth
import bench
kind = sys.argv[1] if len(sys.argv) > 1 else "noise"
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.WORKLOADS["c2_4k_256c_ordered_allwide"]
lib = capi.load_product()
(w, h), (cw, ch) = wl["src"], wl["cells"]
srcs = [torch.from_numpy(synth.KINDS[kind](w, h, 100 + i)).pin_memory() for i in range(4)]
cs = [capi.Canvas(lib, cw, ch, **wl["cfg"]) for _ in range(nthreads)]
log = []
def worker(t, n):
    for i in range(n):
        t0 = time.perf_counter(); cs[t].draw(srcs[(i + t) % 4].data_ptr(), w, h)
        t1 = time.perf_counter(); s = cs[t].print(); t2 = time.perf_counter()
        log.append((t, i, t0, t1, t2, len(s)))
for n in (4, 10):
    log.clear()
    ths = [threading.Thread(target=worker, args=(t, n)) for t in range(nthreads)]
    T0 = time.perf_counter()
    [t.start() for t in ths]; [t.join() for t in ths]
    T1 = time.perf_counter()
print("frames %d in %.3f ms -> %.3f ms/frame" % (n * nthreads, (T1 - T0) * 1e3, (T1 - T0) * 1e3 / (n * nthreads)))
for t, i, t0, t1, t2, L in sorted(log, key=lambda r: r[2])[:12]:
    print("thr %d frame %d: draw %.3f..%.3f (%.3f) print ..%.3f (%.3f) len %d" % (t, i, (t0 - T0) * 1e3, (t1 - T0) * 1e3, (t1 - t0) * 1e3, (t2 - T0) * 1e3, (t2 - t1) * 1e3, L))
