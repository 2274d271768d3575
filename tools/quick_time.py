"""Quick timing of first-order layers through the tiled engine at several n (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import XYKEngine

for n in (20, 24, 26, 28, 30):
    K, w = oc.random_problem(n, 20260700 + n)
    with XYKEngine(n) as eng:
        eng.set_problem(K, w)
        eng.set_engine("tiled")
        eng.init_product_state()
        eng.enable_timing(True)
        eng.apply_layers(0.02, 1, 1)  # warm-up + plan
        eng.apply_layers(0.02, 1, 1)
        ts = []
        for _ in range(3):
            eng.apply_layers(0.02, 1, 1)
            ts.append(eng.stats()["last_pass_ms"])
        st = eng.stats()
        t0 = time.time(); R, _ = eng.order_parameter(); t1 = time.time()
        nrm = eng.norm2()
        ms = min(ts)
        sweeps = st["last_passes"] + 1
        gbs = sweeps * 2 * 16 * 2**n / (ms * 1e-3) / 1e9
        print(f"n={n} layer {ms:.3f} ms ({1e3/ms:.2f} layers/s) passes={st['last_passes']} rounds={st['last_rounds']} "
              f"gates={st['last_gates']} eff-sweep-BW={gbs:.0f} GB/s  R-eval {1e3*(t1-t0):.1f} ms norm2={nrm:.12f}", flush=True)
