"""Timing ablations of the fused tile pass (development aid; XYK_ABLATE makes results wrong on purpose)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import XYKEngine
n = int(sys.argv[1])
K, w = oc.random_problem(n, 20260700 + n)
with XYKEngine(n) as eng:
    eng.set_problem(K, w); eng.set_engine("tiled"); eng.init_product_state(); eng.enable_timing(True)
    eng.apply_layers(0.02, 1, 1); eng.apply_layers(0.02, 1, 1)
    ts = []
    for _ in range(3):
        eng.apply_layers(0.02, 1, 1); ts.append(eng.stats()["last_pass_ms"])
    st = eng.stats()
    print("n=%%d ablate=%%s layer %%.2f ms passes=%%d rounds=%%d  per-pass %%.3f ms  sweep-BW %%.0f GB/s" %% (n, os.environ.get("XYK_ABLATE", "0"), min(ts), st["last_passes"], st["last_rounds"], min(ts) / st["last_passes"], st["last_passes"] * 32 * 2**n / (min(ts) * 1e-3) / 1e9))
''' % ROOT
n = sys.argv[1] if len(sys.argv) > 1 else "30"
for ab in (sys.argv[2:] or ["0", "1", "2", "4", "5", "8", "9"]):
    env = dict(os.environ, XYK_ABLATE=ab)
    subprocess.run([sys.executable, "-c", code, n], env=env)
