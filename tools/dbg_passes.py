"""Debugging aid: GPU state after the first k passes of a program vs the NumPy emulation of the same passes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import XYKEngine, plan_describe
from tests.plan_sim import permute_bits

n, order, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
K, w = oc.random_problem(n, 70 + n)
Kc, wc = oc.canonicalise(n, K, w)
plan = plan_describe(Kc, wc, 0.1, order, reps)
TB = plan["TB"]
perm = list(plan["start_perm"])
psi = permute_bits(oc.initial_state(wc), n, perm)
refs = []
for p in plan["passes"]:
    if p["has_phase"]:
        oc.apply_z_phase(psi, [(perm[q], float(wc[q])) for q in range(n) if abs(wc[q]) > 1e-15], p["ztau"])
    for r in p["rounds"]:
        for i, j, phi in r["gates"]:
            oc.apply_pair(psi, perm[i], perm[j], phi)
    psi = permute_bits(psi, n, p["out_of_in"])
    perm = [p["out_of_in"][perm[q]] for q in range(n)]
    inv = [0] * n
    for q in range(n):
        inv[perm[q]] = q
    refs.append(permute_bits(psi, n, inv))
for k in range(1, len(plan["passes"]) + 1):
    os.environ["XYK_STOP_AFTER_PASS"] = str(k)
    with XYKEngine(n) as eng:
        eng.set_problem(Kc, wc); eng.set_engine("tiled"); eng.init_product_state(); eng.apply_layers(0.1, order, reps)
        got = eng.get_state()
    p = plan["passes"][k - 1]
    desc = [("S%d|%d" % (r["nleft"], len(r["regq"]) - r["nleft"]) if r["split"] else "C") + ":" + str(len(r["gates"])) for r in p["rounds"]]
    print(k, "infid %.3e maxdiff %.3e" % (1 - oc.fidelity(got, refs[k - 1]), np.abs(got - refs[k - 1]).max()), "fl", p["fuse_load"], "fs", p["fuse_store"], desc, flush=True)
