"""A/B timing of two builds of libxyk.so on the same box: XYK_LIB_AB=<path> selects the library (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _devlib import use_dev_library  # noqa: E402
use_dev_library()
from oracle import xy_oracle as oc  # noqa: E402
from scpn_quantum_control_b200.engine import XYKEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
K, w = oc.random_problem(n, 20260700 + n)
with XYKEngine(n) as eng:
    eng.set_problem(K, w)
    eng.set_engine("tiled")
    eng.init_product_state()
    eng.enable_timing(True)
    ts = []
    for i in range(6):
        eng.apply_layers(0.02, 1, 1)
        if i >= 2:
            ts.append(eng.stats()["last_pass_ms"])
    print(f"{os.environ.get('XYK_LIB_AB', 'default')}: n={n} layer ms {' '.join('%.2f' % t for t in ts)}  min {min(ts):.2f} norm2 {eng.norm2():.12f}")
