import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import XYKEngine
n, order, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
K, w = oc.random_problem(n, 70 + n)
Kc, wc = oc.canonicalise(n, K, w)
z, p = oc.term_list(Kc, wc)
ref = oc.evolve_product_formula(oc.initial_state(wc), z, p, 0.1, reps, order)
with XYKEngine(n) as eng:
    eng.set_problem(Kc, wc); eng.set_engine("tiled"); eng.init_product_state(); eng.apply_layers(0.1, order, reps)
    psi = eng.get_state()
print("n", n, "order", order, "reps", reps, "infid %%.3e" %% (1 - oc.fidelity(psi, ref)))
''' % ROOT
for env in ({"XYK_SPLIT_MIN_LEFT": "3", "XYK_BEAM": "3"}, {"XYK_SPLIT_MIN_LEFT": "2", "XYK_BEAM": "1"}, {"XYK_SPLIT_MIN_LEFT": "2", "XYK_BEAM": "3", "XYK_FUSE_IO": "0"}, {"XYK_SPLIT_MIN_LEFT": "2", "XYK_BEAM": "3", "XYK_ABLATE": "8"}):
    for args in (("20", "2", "1"),):
        e = dict(os.environ, **env)
        print(env, end=" ", flush=True)
        subprocess.run([sys.executable, "-c", code, *args], env=e)
