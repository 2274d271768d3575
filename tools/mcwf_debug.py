"""Debugging aid: one MCWF trajectory on the device next to the NumPy device twin, operation by operation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import xy_oracle as oc
from scpn_quantum_control_b200 import tensor_jump as tj
from tests.test_next_rows_cpu import _NumpyDevice

n, ga, gd, seed = int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
K = oc.paper27_K(16)[:n, :n] * 1.5
w = oc.OMEGA_N_16[:n]
dev = tj._Device(n, K, w, 0, None, 2e-10)
twin = _NumpyDevice(n, K, w)
theta = np.mod(w, 2 * np.pi)
for d in (dev, twin):
    d.prepare(theta[::-1].copy())
def cmp(tag):
    a, b = dev.state(), twin.state()
    print(f"{tag:28s} max|dpsi| {np.abs(a-b).max():.3e}  norm dev {dev.norm2():.12f} twin {twin.norm2():.12f}  R dev {abs(dev.xy_sum()/n):.9f} twin {abs(twin.xy_sum()/n):.9f}")
cmp("prepare")
print("z dev", np.round(dev.z_expectations(), 6), "twin", np.round(twin.z_expectations(), 6))
snap = dev.snapshot(); tsnap = twin.snapshot()
cmp("after snapshot")
for d in (dev, twin):
    d.exact_step(0.1, 10)
cmp("exact_step")
for d in (dev, twin):
    d.decay(0.5 * ga * 0.1, float(np.exp(-0.25 * gd * n * 0.1)))
cmp("decay")
ns = dev.norm2()
for d in (dev, twin):
    d.decay(0.0, 1.0 / np.sqrt(ns))
cmp("renormalise")
dev.restore(snap); twin.restore(tsnap)
cmp("restore")
for d in (dev, twin):
    d.jump(n - 1, 0)
cmp("jump top bit kind 0")
for d in (dev, twin):
    d.jump(0, 1)
cmp("jump bit 0 kind 1")
dev.close()
