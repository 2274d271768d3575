"""Development aid: wall time (synchronised) of the product-state init and of one R(t) evaluation at n."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _devlib import use_dev_library  # noqa: E402
use_dev_library()
import numpy as np
from oracle import xy_oracle as oc  # problem generator only
from scpn_quantum_control_b200.engine import XYKEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
K, w = oc.random_problem(n, 20260700 + n)
with XYKEngine(n) as eng:
    eng.set_problem(K, w)
    eng.set_engine("tiled")
    for name, fn in (("init", eng.init_product_state), ("order_parameter", eng.order_parameter), ("z_expectations", eng.z_expectations)):
        fn(); eng.synchronize()
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); fn(); eng.synchronize(); ts.append(time.perf_counter() - t0)
        print(f"n={n} {name}: {1e3 * min(ts):.2f} ms")
    eng.apply_layers(0.02, 1, 1)
    ex, ey = eng.xy_expectations()
    eng.set_option(3, 0)  # OPT_TILED_OBS off: the five-qubits-per-sweep kernel as cross-check
    ex2, ey2 = eng.xy_expectations()
    print("tile vs multi kernel max diff", float(max(np.abs(ex - ex2).max(), np.abs(ey - ey2).max())))
