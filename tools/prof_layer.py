"""Run a few first-order layers through the tiled engine (profiling target: short, one GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _devlib import use_dev_library  # noqa: E402
use_dev_library()
from oracle import xy_oracle as oc  # noqa: E402  (problem generator only)
from scpn_quantum_control_b200.engine import XYKEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K, w = oc.random_problem(n, 20260700 + n)
with XYKEngine(n) as eng:
    eng.set_problem(K, w)
    eng.set_engine("tiled")
    eng.init_product_state()
    eng.enable_timing(True)
    for _ in range(layers):
        eng.apply_layers(0.02, 1, 1)
    st = eng.stats()
    R, _ = eng.order_parameter()
    print(f"n={n} last layer {st['last_pass_ms']:.3f} ms passes={st['last_passes']} R={R:.6f}")
