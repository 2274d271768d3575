"""Development aid: per-warp phase timeline of one tile pass (needs the -DXYK_TRACE build of the library,
tools/_build/libxyk_trace.so; see tools/README.md).  Writes gpurun_out/trace_n{n}_p{pass}.npy:
[warp slot][256] words, word 0 = SM id, then (clock << 8 | code) events:
1/2 = tile wait start/end, 3 = round start (barrier released), 4 = loads returned, 5 = gates done,
6 = stores issued."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scpn_quantum_control_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "tools", "_build", "libxyk_trace.so")  # make -C scpn_quantum_control_b200/csrc trace
from oracle import xy_oracle as oc  # noqa: E402  (problem generator only)
from scpn_quantum_control_b200.engine import XYKEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
passes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 3]
k0 = int(sys.argv[3]) if len(sys.argv) > 3 else 100
nk = int(sys.argv[4]) if len(sys.argv) > 4 else 4
lib = _lib.load()
lib.xyk_debug_trace_begin.restype = ctypes.c_int
lib.xyk_debug_trace_begin.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
lib.xyk_debug_trace_read.restype = ctypes.c_int64
lib.xyk_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int64]
K, w = oc.random_problem(n, 20260700 + n)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with XYKEngine(n) as eng:
    eng.set_problem(K, w)
    eng.set_engine("tiled")
    eng.init_product_state()
    eng.enable_timing(True)
    eng.apply_layers(0.02, 1, 1)
    eng.apply_layers(0.02, 1, 1)
    for p in passes:
        assert lib.xyk_debug_trace_begin(p, k0, nk) == 0
        eng.apply_layers(0.02, 1, 1)
        buf = np.zeros(148 * 3 * 4 * 256, dtype=np.uint64)
        got = lib.xyk_debug_trace_read(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
        assert got == buf.size, got
        np.save(os.path.join(ROOT, "gpurun_out", f"trace_n{n}_p{p}.npy"), buf.reshape(-1, 256))
        print(f"pass {p}: traced, layer {eng.stats()['last_pass_ms']:.2f} ms")
