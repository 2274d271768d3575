"""Timing experiment (results intentionally wrong in modes 1/2): where does a pass spend its time?
mode 0: normal; mode 1: all exchanges, no rotations; mode 2: one empty round per pass (pure permuting copy)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import XYKEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
K, w = oc.random_problem(n, 20260700 + n)
with XYKEngine(n) as eng:
    eng.set_problem(K, w)
    eng.set_engine("tiled")
    for mode in (0, 1, 3, 0):
        eng.set_option(100, mode)
        eng.init_product_state()
        eng.enable_timing(True)
        eng.apply_layers(0.02, 1, 1)
        ts = []
        for _ in range(3):
            eng.apply_layers(0.02, 1, 1)
            ts.append(eng.stats()["last_pass_ms"])
        st = eng.stats()
        print(f"n={n} mode={mode} layer {min(ts):.3f} ms  per-pass {min(ts)/st['last_passes']:.3f} ms passes={st['last_passes']}", flush=True)
