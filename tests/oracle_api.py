"""ctypes access to the plain-C oracle restatement (oracle/chafa_oracle.c -> oracle/_ref/libchafa_oracle.so).
TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it."""

import ctypes as C
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "_ref", "libchafa_oracle.so")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class OracleCanvas(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("mode", C.c_int32), ("work_factor_int", C.c_int32),
                ("alpha_threshold", C.c_int32), ("optimizations", C.c_int32), ("blank_char", C.c_uint32),
                ("solid_char", C.c_uint32), ("n_syms", C.c_int32), ("bitmaps", C.c_void_p), ("chars", C.c_void_p),
                ("n_syms2", C.c_int32), ("bitmaps2", C.c_void_p), ("chars2", C.c_void_p)]


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(ORACLE_LIB)
        _lib.oracle_update_cells.argtypes = [C.POINTER(OracleCanvas), C.c_void_p, C.c_void_p]
        _lib.oracle_update_cells.restype = None
        _lib.oracle_print.argtypes = [C.POINTER(OracleCanvas), C.c_void_p, C.c_void_p]
        _lib.oracle_print.restype = C.c_size_t
    return _lib


def golden_names(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(prefix))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: g[k] for k in g.files}
    d["config"] = json.loads(str(d["config"]))
    return d


def run_oracle(g):
    """Cells and printed bytes of the restatement for one golden case (from its prepared pixmap)."""
    lib = load()
    cw, ch = g["config"]["cells"]
    cfg = g["config"]["cfg"]
    info = g["info"]
    keep = [np.ascontiguousarray(g[k]) for k in ("bitmaps", "chars", "bitmaps2", "chars2", "pixmap")]
    cv = OracleCanvas(cw, ch, cfg.get("mode", 0), int(info[2]), 127, 0x7FFFFFFF, int(info[3]), int(info[4]),
                      len(keep[1]), keep[0].ctypes.data, keep[1].ctypes.data,
                      len(keep[3]), keep[2].ctypes.data, keep[3].ctypes.data)
    cells = np.zeros((ch, cw, 3), np.uint32)
    lib.oracle_update_cells(C.byref(cv), keep[4].ctypes.data, cells.ctypes.data)
    out = C.create_string_buffer((cw * ch + ch) * 64 + 64)
    n = lib.oracle_print(C.byref(cv), cells.ctypes.data, out)
    return cells, out.raw[:n]
