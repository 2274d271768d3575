"""The other BASELINE.json configurations (parity-test cases, not the headline bench line), timed end to end through
the public API on one B200 with the oracle's CPU path timed beside them:
  configs[0]  n=4 Kuramoto-XY ring, run(t_max=1, dt=0.1, 5 layers per step), R(t) trace;
  configs[1]  n=16 all-to-all random K, 200 first-order layers (run(4.0, 0.1, 5)), complex128;
  configs[2]  4096 independent n=12 problems, 100 layers each (batch kernel).
Prints one JSON object; `python tools/bench_configs.py > profiles/...json`."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import xy_oracle as oc
from scpn_quantum_control_b200 import QuantumKuramotoSolver
from scpn_quantum_control_b200.engine import batch_run

out = {}

def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0)
    return min(ts), r

# configs[0]
K, w = oc.ring_K(4), oc.OMEGA_N_16[:4]
def gpu0():
    s = QuantumKuramotoSolver(4, K, w, evolution="trotter"); r = s.run(1.0, 0.1, 5); s.close(); return np.asarray(r["R"])
t_gpu, R = timed(gpu0)
t0 = time.perf_counter(); _, Rref = oc.run(4, K, w, 1.0, 0.1, 5, order=1); t_cpu = time.perf_counter() - t0
out["config0_n4_ring"] = {"gpu_s": t_gpu, "numpy_oracle_s": t_cpu, "max_dR": float(np.abs(R - Rref).max()), "layers": 50}

# configs[1]
K, w = oc.random_problem(16, 20260716)
def gpu1():
    s = QuantumKuramotoSolver(16, K, w, evolution="trotter"); r = s.run(4.0, 0.1, 5); s.close(); return np.asarray(r["R"])
t_gpu, R = timed(gpu1)
t0 = time.perf_counter(); _, Rref = oc.run(16, K, w, 4.0, 0.1, 5, order=1); t_cpu = time.perf_counter() - t0
out["config1_n16_200layers"] = {"gpu_s": t_gpu, "gpu_layers_per_s": 200 / t_gpu, "numpy_oracle_s": t_cpu, "numpy_layers_per_s": 200 / t_cpu,
                                 "max_dR": float(np.abs(R - Rref).max())}

# configs[2]
B, n = 4096, 12
Ks = np.empty((B, n, n)); ws = np.empty((B, n))
for b in range(B):
    Ks[b], ws[b] = oc.random_problem(n, 20260716 + b)
steps = np.full(20, 0.1)
t_gpu, R = timed(lambda: batch_run(Ks, ws, steps, 5, 1))
sub = np.random.default_rng(1).choice(B, 16, replace=False)
t0 = time.perf_counter(); err = 0.0
for b in sub:
    _, Rref = oc.run(n, Ks[b], ws[b], 2.0, 0.1, 5, order=1); err = max(err, float(np.abs(R[b] - Rref).max()))
t_cpu = (time.perf_counter() - t0) / len(sub)
out["config2_batch_4096_n12"] = {"gpu_s": t_gpu, "gpu_problem_layers_per_s": B * 100 / t_gpu, "numpy_oracle_s_per_problem": t_cpu,
                                  "numpy_problem_layers_per_s": 100 / t_cpu, "max_dR_subset16": err}
print(json.dumps(out, indent=1))
