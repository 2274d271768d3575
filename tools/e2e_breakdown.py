"""Where the end-to-end time of solver.update_problem + solver.run goes at n (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import xy_oracle as oc  # problem generator only
from scpn_quantum_control_b200 import QuantumKuramotoSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
K, w = oc.random_problem(n, 20260700 + n)
s = QuantumKuramotoSolver(n, K, w, evolution="trotter")
s.run(0.1, 0.1, 5)
eng = s.engine()
eng.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); s.update_problem(K, w); t1 = time.perf_counter()
    st0 = eng.stats()
    res = s.run(0.2, 0.1, 5); t2 = time.perf_counter()
    st1 = eng.stats()
    print(f"update_problem {1e3*(t1-t0):.2f} ms   run {1e3*(t2-t1):.2f} ms   launches {st1['kernel_launches']-st0['kernel_launches']} sweeps {st1['state_sweeps']-st0['state_sweeps']} passes {st1['pass_launches']-st0['pass_launches']}")
# pieces
eng.synchronize(); t0 = time.perf_counter(); eng.init_product_state(); eng.synchronize(); t1 = time.perf_counter()
print(f"init_product_state {1e3*(t1-t0):.2f} ms")
t0 = time.perf_counter(); eng.order_parameter(); t1 = time.perf_counter(); print(f"order_parameter {1e3*(t1-t0):.2f} ms")
t0 = time.perf_counter(); eng.apply_layers(0.02, 1, 1); eng.synchronize(); t1 = time.perf_counter(); print(f"first layer after init (incl. re-layout) {1e3*(t1-t0):.2f} ms")
t0 = time.perf_counter(); eng.apply_layers(0.02, 1, 5); eng.synchronize(); t1 = time.perf_counter(); print(f"5 layers {1e3*(t1-t0):.2f} ms")
t0 = time.perf_counter(); eng.order_parameter(); t1 = time.perf_counter(); print(f"order_parameter {1e3*(t1-t0):.2f} ms")
s.close()
