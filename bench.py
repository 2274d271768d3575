#!/usr/bin/env python
"""bench.py -- XY Trotter steps/s on the BASELINE.json headline workload.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--n QUBITS] [--weak]

A "step" is ONE first-order Trotter layer (one omega.Z diagonal phase + every pair rotation
exp(-i theta_ij (XX+YY)) in the reference's order) applied to the resident statevector:
n = 30 all-to-all random K_nm, complex128 (16 GiB state; BASELINE.json configs[3], the
configuration the metric "XY Trotter steps/sec at n=30" is quoted on).  With N > 1 ranks the SAME
n = 30 state is sharded by its top log2 N qubits (strong scaling, the default: the metric is quoted
"at n=30 (1/2/4/8 B200)"); ``--weak`` keeps 2^30 amplitudes per GPU instead (n = 33 on 8 GPUs).

The K timed layers are applied the way the reference's ``run()`` applies them: one call per dt step of
``trotter_per_step`` = 5 layers (its default, phase/xy_kuramoto.py:54-70) -- K // 5 calls of five layers plus K % 5
single-layer calls; ``steps`` counts layers.  The call size changes nothing on one GPU (cyclic pass program); a sharded
state restores its sharding once per call.

Printed JSON (one line, rank 0):
  value / ms_per_step  device-timed (CUDA events on the engine's stream) steps with the state
                       resident in HBM; max over ranks.
  e2e                  the same metric through the public API: ``solver.update_problem(K, omega)``
                       (host K, omega -> device) + ``solver.run(t_max, dt, 5)`` (device state
                       preparation, layers, R(t) every 5 layers, read-back of R(t) to the host).
  roofline             dominant kernel = tile_pass_kernel; achieved = ALGORITHMIC bytes per launch
                       (one read + one write of the state, 2 * 16 B * 2^n) / mean launch duration.
  check                parity evidence of THIS run: sharded engine vs the C oracle at small n
                       (N > 1), and at the benchmark size an order-2 echo against the analytic
                       product state, <sum Z> and the norm.  A failed check exits non-zero.
  cpu_baseline         (N = 1) the oracle's C/OpenMP restatement of the reference path timed on
                       this box's host cores on a bounded sample of the real n = 30 layer.

``--impl reference`` times that C/OpenMP restatement on ONE REAL n = 30 layer (16 GiB of host
memory): the K timed steps are consecutive chunks of ceil(435 / K) pair rotations, so the timed
region applies one complete layer (or more) and ``value`` is measured, not extrapolated.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "XY Trotter steps/sec at n=30"
UNIT = "steps/s"
BASE_N = 30
SEED = 20260730


def synthetic_problem(n: int, seed: int):
    """K = sym(U(0.05, 1)), zero diagonal, omega ~ U(0.5, 4); all pairs non-zero."""
    rng = np.random.default_rng(seed)
    A = rng.uniform(0.05, 1.0, (n, n))
    K = (A + A.T) / 2.0
    np.fill_diagonal(K, 0.0)
    return K, rng.uniform(0.5, 4.0, n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.samples: list[list[str]] = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._loop, daemon=True)

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.device)],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=5)

    def summary(self) -> dict:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for name, val in zip(names, s[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": float(self.samples[0][2]) if self.samples[0][2].replace(".", "").isdigit() else None,
            "power_w_max": max((float(s[3]) for s in self.samples if s[3].replace(".", "").isdigit()), default=None),
            "reasons": sorted(reasons),
            "samples": len(self.samples),
        }


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(n: int):
    """dram bytes (read + write) per tile-pass launch from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "tile_pass_traffic.json")
    try:
        d = json.load(open(path))
        if int(d["n"]) == n:
            return float(d["dram_bytes_per_launch"])
        return float(d["dram_bytes_per_launch"]) * 2.0 ** (n - int(d["n"]))
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle's C/OpenMP restatement (test infrastructure, executed here ONLY as the
# timed CPU baseline / --impl reference arm)
# ----------------------------------------------------------------------------------------------
class CpuLayerStepper:
    """The oracle's C/OpenMP gate-by-gate path on a real n-qubit state, advanced a few gates at a time.

    The gate sequence is the reference's first-order layer repeated: diagonal phase, then the pairs in
    lexicographic order (oracle/xy_oracle.c, bridge/knm_hamiltonian.py:176-195)."""

    def __init__(self, n_target: int):
        import ctypes

        so = os.path.join(ROOT, "oracle", "_build", "libxyoracle.so")
        if not os.path.exists(so):
            env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], env=env, stdout=subprocess.DEVNULL)
        lib = ctypes.CDLL(so)
        self.dp = dp = ctypes.POINTER(ctypes.c_double)
        lib.xyo_num_threads.restype = ctypes.c_int
        lib.xyo_set_num_threads.argtypes = [ctypes.c_int]
        lib.xyo_init_state.argtypes = [ctypes.c_int, dp, dp]
        lib.xyo_apply_z_phase.argtypes = [ctypes.c_int, dp, ctypes.c_double, dp]
        lib.xyo_apply_pair.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp]
        # every host core this process may use, whatever OMP_NUM_THREADS the launcher exported (torchrun sets 1)
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        lib.xyo_set_num_threads(cores)
        self.threads = int(lib.xyo_num_threads())
        self.lib = lib
        # the real size when host memory allows (16 B * 2^n + headroom), else the largest that fits -- stated in `sample`
        n = n_target
        avail = _host_available_bytes()
        while n > 20 and avail is not None and 16 * 2 ** n * 1.2 > avail:
            n -= 1
        self.n, self.n_target = n, n_target
        self.K, self.w = synthetic_problem(n, SEED)
        self.pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
        self.tau = 0.02
        self.psi = np.empty(2 << n)
        lib.xyo_init_state(n, self.w.ctypes.data_as(dp), self.psi.ctypes.data_as(dp))
        self.pos = 0  # next gate of the layer sequence

    def advance(self, gates: int) -> None:
        lib, dp, n = self.lib, self.dp, self.n
        for _ in range(gates):
            if self.pos == 0:
                lib.xyo_apply_z_phase(n, self.w.ctypes.data_as(dp), self.tau, self.psi.ctypes.data_as(dp))
            i, j = self.pairs[self.pos]
            lib.xyo_apply_pair(n, i, j, 2.0 * self.K[i, j] * self.tau, self.psi.ctypes.data_as(dp))
            self.pos = (self.pos + 1) % len(self.pairs)

    def scale_to_target(self) -> float:
        """cost of a layer ~ (#pairs) * 2^n: 1.0 when the real size fitted"""
        pairs = lambda q: q * (q - 1) / 2.0  # noqa: E731
        return (pairs(self.n_target) / pairs(self.n)) * 2.0 ** (self.n_target - self.n)


def _host_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except OSError:
        pass
    return None


def cpu_baseline_sample(n_target: int, gates: int = 24):
    """bounded CPU sample for the GPU arm's line: the phase + the first `gates` pair rotations of the real layer"""
    st = CpuLayerStepper(n_target)
    st.advance(1)  # warm (threads, page faults of the first sweep)
    st.pos = 0
    t0 = time.perf_counter()
    st.advance(gates)
    sec = time.perf_counter() - t0
    layer_sec = sec * (len(st.pairs) / gates) * st.scale_to_target()
    sample = (f"diagonal phase + the first {gates} of the {len(st.pairs)} pair rotations of one first-order layer on a real "
              f"n={st.n} complex128 state (C/OpenMP oracle port, {st.threads} threads), scaled by {len(st.pairs)}/{gates}"
              + ("" if st.n == n_target else f" and to n={n_target} by pairs*2^n (host memory too small for n={n_target})"))
    return layer_sec, st.threads, sample


LAYERS_PER_CALL = 5   # the reference's default trotter_per_step (TrotterEvolutionConfig.run_steps_per_step)


def common_config(n: int, tau: float = 0.02) -> dict:
    """the workload both arms run (identical in the two JSON lines)"""
    return {"workload": f"n={n} all-to-all XY Trotter layer (first order), complex128", "n_qubits": n,
            "pairs_per_layer": n * (n - 1) // 2, "seed": SEED, "tau": tau}


def metric_name(n: int, weak: bool) -> str:
    if n == BASE_N:
        return METRIC
    return f"XY Trotter steps/sec at n={n}" + (f" (weak scaling: 2^{BASE_N} amplitudes per GPU)" if weak else "")


def run_reference_arm(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    n = args.n + (max(0, world.bit_length() - 1) if args.weak else 0)
    st = CpuLayerStepper(n)
    npairs = len(st.pairs)
    gates_per_step = max(1, -(-npairs // max(1, args.steps)))  # K steps together = at least one complete layer
    for _ in range(args.warmup):
        st.advance(1)
    st.pos = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st.advance(gates_per_step)
    sec = time.perf_counter() - t0
    layers = args.steps * gates_per_step / npairs
    layer_sec = sec / layers * st.scale_to_target()
    value = 1.0 / layer_sec
    sample = (f"each step = {gates_per_step} consecutive pair rotations (+ the diagonal phase at the layer start) of the real "
              f"n={st.n} first-order layer, C/OpenMP oracle port on {st.threads} threads; the {args.steps} timed steps apply "
              f"{layers:.3f} complete layers in {sec:.1f} s; value = layers / seconds"
              + ("" if st.n == n else f", scaled to n={n} by pairs*2^n (host memory too small)"))
    line = {
        "impl": "reference",
        "metric": metric_name(n, args.weak), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / args.steps, "layer_fraction_per_step": gates_per_step / npairs,
        "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": common_config(n),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": st.threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = Python+Qiskit+Rust; neither Qiskit nor cargo exists in this image, so the timed CPU arm "
                "is the oracle's gate-by-gate C/OpenMP restatement of the same path; a step is a bounded chunk of the layer "
                "(ms_per_step is the chunk time), value is whole layers per second",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def _c_oracle_evolve(n, K, omega, delta, reps, order):
    """checker only: the oracle's C/OpenMP restatement (small n)"""
    import ctypes

    from oracle import xy_oracle as oc

    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libxyoracle.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    Kc, wc = oc.canonicalise(n, K, omega)
    Ks = np.ascontiguousarray((Kc + Kc.T) / 2.0)
    psi = np.zeros(2 << n)
    lib.xyo_init_state.argtypes = [ctypes.c_int, dp, dp]
    lib.xyo_init_state(n, wc.ctypes.data_as(dp), psi.ctypes.data_as(dp))
    lib.xyo_evolve.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, dp]
    lib.xyo_evolve(n, Ks.ctypes.data_as(dp), wc.ctypes.data_as(dp), float(delta), reps, order, psi.ctypes.data_as(dp))
    return psi.view(np.complex128)


def matching_check(eng, n: int, w: np.ndarray, tau: float) -> float:
    """with a perfect-matching coupling the state stays a product of two-qubit states, so two
    second-order layers of the whole n-qubit state have closed-form single-qubit observables (4 x 4 arithmetic per pair).
    The matching pairs qubit q with n-1-q: on a sharded state that covers local-global pairs and, from 4 GPUs on,
    the partners of the other global qubits.  Returns the largest |<X>, <Y>, <Z>| error over all qubits."""
    rng = np.random.default_rng(SEED + 7)
    Km = np.zeros((n, n))
    pairs = [(q, n - 1 - q) for q in range(n // 2)]
    for a, b in pairs:
        Km[a, b] = Km[b, a] = rng.uniform(0.3, 1.0)
    eng.set_problem(Km, w)
    eng.init_product_state()
    eng.apply_layers(2 * tau, 2, 2)
    ex, ey = eng.xy_expectations()
    ez = eng.z_expectations()
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    Y = np.array([[0, -1j], [1j, 0]])
    Z = np.diag([1.0, -1.0]).astype(complex)
    I2 = np.eye(2, dtype=complex)
    worst = 0.0
    theta = np.mod(w, 2.0 * np.pi)
    for a, b in pairs:   # two-qubit register: qubit a = bit 0, qubit b = bit 1 (little endian)
        ka = Km[a, b]

        def zph(t):
            return np.diag(np.exp(1j * t * np.array([w[a] + w[b], -w[a] + w[b], w[a] - w[b], -w[a] - w[b]])))

        def rot(phi):   # exp(+i (phi/2)(XX + YY)): mixes |01> and |10>
            u = np.eye(4, dtype=complex)
            u[1, 1] = u[2, 2] = np.cos(phi)
            u[1, 2] = u[2, 1] = 1j * np.sin(phi)
            return u

        psi2 = np.kron([np.cos(theta[b] / 2), np.sin(theta[b] / 2)], [np.cos(theta[a] / 2), np.sin(theta[a] / 2)]).astype(complex)
        for _ in range(2):   # one pair term: the second-order layer is Z(tau/2) pair(tau) Z(tau/2) with step tau
            psi2 = zph(tau / 2) @ rot(2 * ka * tau) @ zph(tau / 2) @ psi2
        for q, ops in ((a, [np.kron(I2, P) for P in (X, Y, Z)]), (b, [np.kron(P, I2) for P in (X, Y, Z)])):
            got = (ex[q], ey[q], ez[q])
            for g, op in zip(got, ops):
                worst = max(worst, abs(g - float(np.real(np.vdot(psi2, op @ psi2)))))
    if n % 2:   # the middle qubit is uncoupled: it only precesses
        q = n // 2
        ph = 2 * tau * 2 * w[q]   # exp(+i t w Z) for t = 2 tau: <X + iY> rotates by -2 t w
        worst = max(worst, abs(ex[q] - np.sin(theta[q]) * np.cos(ph)), abs(ey[q] + np.sin(theta[q]) * np.sin(ph)),
                    abs(ez[q] - np.cos(theta[q])))
    return float(worst)


def parity_check(eng, n: int, w: np.ndarray, world: int, rank: int, device: int, tau: float, c64: bool = False) -> dict:
    """Parity evidence of this very run: (i) N > 1: the sharded engine against the C oracle at
    n = 17 + log2 N (fidelity, |dR|, orders 1 and 2); (ii) at the benchmark size: an order-2 echo (+tau, -tau)
    against the analytic product state (<X_q> = sin theta_q, <Y_q> = 0, <Z_q> = cos theta_q), <sum Z> and the
    norm after one first-order layer.  Tolerances: infidelity 1e-10, |dR| 1e-9, observables 1e-9, norm 1e-10."""
    out: dict = {"ok": True}
    if world > 1:
        from oracle import xy_oracle as oc
        from scpn_quantum_control_b200.sharded import ShardedXYKEngine

        ns = 17 + world.bit_length() - 1
        Ks, ws = synthetic_problem(ns, SEED + 1)
        small = ShardedXYKEngine(ns, device=device)
        small.set_problem(Ks, ws)
        rows = []
        for order, reps in ((1, 2), (2, 1), (1, LAYERS_PER_CALL)):   # the last one: the call shape of the timed loop
            small.init_product_state()
            small.apply_layers(0.1, order, reps)
            R, _ = small.order_parameter()
            psi = small.gather_state()
            if rank == 0:
                ref = _c_oracle_evolve(ns, Ks, ws, 0.1, reps, order)
                infid = float(1.0 - oc.fidelity(psi, ref))
                dR = float(abs(R - oc.order_parameter(ref, ns)[0]))
                rows.append({"order": order, "reps": reps, "infidelity": infid, "dR": dR})
                out["ok"] &= infid < 1e-10 and dR < 1e-9
        small.close()
        out["sharded_vs_oracle"] = {"n": ns, "gpus": world, "cases": rows}
    theta = np.mod(w, 2.0 * np.pi)
    eng.init_product_state()
    eng.apply_layers(tau, 2, 1)
    eng.apply_layers(-tau, 2, 1)
    ex, ey = eng.xy_expectations()
    ez = eng.z_expectations()
    echo = float(max(np.abs(ex - np.sin(theta)).max(), np.abs(ey).max(), np.abs(ez - np.cos(theta)).max()))
    eng.init_product_state()
    eng.apply_layers(tau, 1, 1)
    zsum = float(abs(eng.z_expectations().sum() - np.cos(theta).sum()))
    nrm = float(abs(eng.norm2() - 1.0))
    out["at_bench_size"] = {"n": n, "echo_order2_max_observable_error": echo, "sumZ_error_after_one_layer": zsum,
                            "norm2_error": nrm}
    out["at_bench_size"]["disjoint_pairs_max_observable_error"] = match = matching_check(eng, n, w, tau)
    # complex64 storage: the stated bound of that variant instead of the complex128 tolerances
    tol_obs, tol_nrm = (1e-4, 1e-4) if c64 else (1e-9, 1e-10)
    out["ok"] = bool(out["ok"] and echo < tol_obs and zsum < tol_obs * (n if c64 else 1) and nrm < tol_nrm and match < tol_obs)
    return out


def other_configs(device: int) -> dict:
    """BASELINE configs[0..2] end to end through the public API (host K, omega -> host R(t)), checked against the oracle"""
    import ctypes

    from oracle import xy_oracle as oc
    from scpn_quantum_control_b200 import QuantumKuramotoSolver
    from scpn_quantum_control_b200.engine import batch_run

    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libxyoracle.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.xyo_run.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, dp]

    def c_run(n, K, w, steps, reps, order):
        Kc, wc = oc.canonicalise(n, K, w)
        Ks = np.ascontiguousarray((Kc + Kc.T) / 2.0)
        psi, R = np.zeros(2 << n), np.zeros(len(steps) + 1)
        st = np.ascontiguousarray(steps, dtype=np.float64)
        lib.xyo_run(n, Ks.ctypes.data_as(dp), wc.ctypes.data_as(dp), st.ctypes.data_as(dp), len(st), reps, order,
                    psi.ctypes.data_as(dp), R.ctypes.data_as(dp))
        return R

    def best(fn, reps=3):
        fn()
        ts, r = [], None
        for _ in range(reps):
            t0 = time.perf_counter()
            r = fn()
            ts.append(time.perf_counter() - t0)
        return min(ts), r

    out = {}
    K, w = oc.ring_K(4), oc.OMEGA_N_16[:4]
    s0 = QuantumKuramotoSolver(4, K, w, evolution="trotter", device=device)
    t, R = best(lambda: np.asarray(s0.run(1.0, 0.1, 5)["R"]))
    s0.close()
    out["config0_n4_ring_50_layers"] = {"ms": 1e3 * t, "max_dR_vs_oracle": float(np.abs(R - c_run(4, K, w, np.full(10, 0.1), 5, 1)).max())}
    K, w = oc.random_problem(16, 20260716)
    s1 = QuantumKuramotoSolver(16, K, w, evolution="trotter", device=device)
    t, R = best(lambda: np.asarray(s1.run(4.0, 0.1, 5)["R"]))
    s1.close()
    out["config1_n16_200_layers"] = {"ms": 1e3 * t, "layers_per_s": 200 / t,
                                     "max_dR_vs_oracle": float(np.abs(R - c_run(16, K, w, np.full(40, 0.1), 5, 1)).max())}
    B, n = 4096, 12
    Ks, ws = np.empty((B, n, n)), np.empty((B, n))
    for b in range(B):
        Ks[b], ws[b] = oc.random_problem(n, 20260716 + b)
    steps = np.full(20, 0.1)
    t, R = best(lambda: batch_run(Ks, ws, steps, 5, 1, device=device))
    sub = np.random.default_rng(1).choice(B, 64, replace=False)
    err = max(float(np.abs(R[b] - c_run(n, Ks[b], ws[b], steps, 5, 1)).max()) for b in sub)
    out["config2_batch_4096_n12_100_layers"] = {"ms": 1e3 * t, "problem_layers_per_s": B * 100 / t, "max_dR_vs_oracle_64_subset": err}
    return out


def run_gpu_arm(args, rank: int, world: int, local_rank: int) -> None:
    import torch  # device memory / streams / torch.distributed plumbing only

    lpc = max(1, int(args.layers_per_call))

    from scpn_quantum_control_b200 import QuantumKuramotoSolver
    from scpn_quantum_control_b200.engine import XYKEngine, device_count

    if device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; the product path has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    gbits = world.bit_length() - 1
    n = args.n + (gbits if args.weak else 0)   # total qubits of the (sharded) state
    K, w = synthetic_problem(n, SEED)
    tau = 0.02

    c64 = args.dtype == "complex64"
    if c64 and world > 1:
        raise SystemExit("bench.py: --dtype complex64 runs on one GPU (sharded states are complex128)")
    amp_bytes = 8.0 if c64 else 16.0
    if world == 1:
        eng = XYKEngine(n, dtype=args.dtype, device=device)
        eng.set_problem(K, w)
        eng.set_engine("tiled" if n >= 12 else "simple")
        raw = eng
    else:
        from scpn_quantum_control_b200.sharded import ShardedXYKEngine

        eng = ShardedXYKEngine(n, device=device)
        eng.set_problem(K, w)
        raw = eng.eng
    check = parity_check(eng, n, w, world, rank, device, tau, c64) if not args.no_check else None
    if check is not None:
        eng.set_problem(K, w)   # the matching check re-parameterised the engine
    # One timed call = one dt step of the reference's run(): trotter_per_step = 5 first-order layers (its default,
    # phase/xy_kuramoto.py:54-70, run_steps_per_step) -- a step of the metric stays ONE layer, `steps` layers are applied in steps // 5 calls of five
    # layers plus steps % 5 single-layer calls.  On one GPU the call size changes nothing (the pass program is cyclic); a
    # sharded state pays the exchange that restores its sharding once per call.
    def run_layers(count: int) -> None:
        full, rem = divmod(count, lpc)
        for _ in range(full):
            eng.apply_layers(lpc * tau, 1, lpc)
        for _ in range(rem):
            eng.apply_layers(tau, 1, 1)

    eng.init_product_state()
    warm_layers = max(args.warmup, 3)
    if args.steps >= lpc:
        warm_layers = max(warm_layers, lpc)          # the five-layer program is compiled and warm
    if args.steps % lpc and warm_layers % lpc == 0:
        warm_layers += 1                                          # ... and so is the single-layer program
    run_layers(warm_layers)
    raw.synchronize()
    raw.enable_timing(True)
    st0 = eng.stats()
    stream = torch.cuda.ExternalStream(raw.stream(), device=device)
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(device)
    with ClockSampler(device) as clocks:
        ev0.record(stream)
        run_layers(args.steps)
        ev1.record(stream)
        ev1.synchronize()
        torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    st1 = eng.stats()
    norm2 = eng.norm2()
    if dist is not None:
        t = torch.tensor([ms_total], device=f"cuda:{device}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    passes = st1["pass_launches"] - st0["pass_launches"]
    sweeps = st1["state_sweeps"] - st0["state_sweeps"]
    pass_ms = (st1["pass_kernel_ms"] - st0["pass_kernel_ms"]) / max(1, st1["pass_kernel_timed"] - st0["pass_kernel_timed"])
    peer_n = st1["peer_passes"] - st0["peer_passes"]
    peer_ms = (st1["peer_pass_ms"] - st0["peer_pass_ms"]) / max(1, peer_n)
    rounds = st1["last_rounds"]                                     # rounds of the last call ...
    last_call_layers = 1 if args.steps % lpc else lpc   # ... which applied this many layers
    exchange_log = list(getattr(eng, "exchange_log", []))
    eng.close()

    # ---- e2e through the public API, host buffers in, host R(t) out ----
    e2e_steps = 2
    layers_per_step = 5
    solver = QuantumKuramotoSolver(n, K, w, evolution="trotter", device=device, distributed=world > 1, dtype=args.dtype)
    solver.run(0.1 * 1, 0.1, layers_per_step)  # creates the engine (buffers, plan) and warms it
    torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    solver.update_problem(K, w)  # host K, omega -> device (re-validated, gate table and plan rebuilt)
    res = solver.run(0.1 * e2e_steps, 0.1, layers_per_step)
    torch.cuda.synchronize(device)
    e2e_sec = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_sec], device=f"cuda:{device}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    solver.close()
    e2e_layers = e2e_steps * layers_per_step
    h2d = (K.nbytes + w.nbytes) // e2e_layers               # the step's inputs: K, omega (host -> device every timed call)
    d2h = (2 * n * 8 * (e2e_steps + 1)) // e2e_layers       # <X_q>, <Y_q> per R(t) point (device -> host)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    n_shard = n - gbits
    alg_bytes = 2.0 * amp_bytes * 2.0 ** n_shard   # per launch: one read + one write of this GPU's shard
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else 0.0
    # one step = one first-order Trotter layer of the WHOLE n-qubit state (all ranks together)
    value = 1e3 / ms_per_step
    engine_info = {
        "state": f"{amp_bytes * 2 ** n_shard / 2 ** 30:.2f} GiB shard per GPU"
                 + ("" if world == 1 else f", sharded by the top {gbits} physical qubits over {world} GPUs; qubit exchanges are fused into "
                    "the preceding tile pass as peer stores over NVLink (CUDA IPC mapped buffers), NCCL all-to-all as fallback"),
        "l2_policy": f"state shard ({amp_bytes * 2 ** n_shard / 2 ** 30:.0f} GiB per GPU) >> L2 (126 MB): every pass streams from HBM, no flush needed",
        "passes_per_layer": passes / args.steps, "rounds_per_layer": rounds / last_call_layers, "sweeps_per_layer": sweeps / args.steps,
        "layers_per_call": lpc,
        "calls": f"{args.steps // lpc} x apply_layers({lpc} tau, order 1, reps {lpc})"
                 + (f" + {args.steps % lpc} x apply_layers(tau, 1, 1)" if args.steps % lpc else "")
                 + ": one call = one dt step of the reference's run() (trotter_per_step = 5, its default); a step of the metric is one layer",
    }
    if world > 1:
        peer_bytes = 16.0 * 2.0 ** n_shard * (world - 1) / world   # leaves this GPU per fused exchange
        engine_info.update({
            "nccl_exchanges_per_layer": (st1["exchanges"] - st0.get("exchanges", 0)) / args.steps,
            "nccl_exchange_ms_mean": float(np.mean(exchange_log[-8:])) if exchange_log else None,
            "peer_passes_per_layer": peer_n / args.steps,
            "peer_pass_ms_mean": peer_ms if peer_n else None,
            "nvlink_bytes_out_per_peer_pass_per_gpu": int(peer_bytes),
            "nvlink_gbs_out_per_gpu_during_peer_pass": peer_bytes / (peer_ms * 1e-3) / 1e9 if peer_n and peer_ms > 0 else None,
        })
    cpu = None
    if world == 1 and not args.no_cpu and not c64:
        layer_sec, threads, sample = cpu_baseline_sample(n)
        cpu = {"value": 1.0 / layer_sec, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    line = {
        "metric": metric_name(n, args.weak), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm_layers, "warmup_requested": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64" if not c64 else "f64 arithmetic on complex64 storage", "data": "synthetic",
        "config": common_config(n, tau) if not c64 else dict(common_config(n, tau), workload=f"n={n} all-to-all XY Trotter layer (first order), "
                                                             "complex64 state storage, FP64 arithmetic"),
        "engine": engine_info,
        "roofline": {
            "bound": "hbm", "kernel": "tile_pass_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src, "traffic": None if c64 else ncu_traffic(n_shard),
            "algorithmic_bytes_per_launch": alg_bytes, "mean_launch_ms": pass_ms,
            "layer_algorithmic_frac": (1e3 / ms_per_step) * alg_bytes / 1e9 / peak,
            "layer_sweeps_frac": (1e3 / ms_per_step) * (sweeps / args.steps) * alg_bytes / 1e9 / peak,
        },
        "cpu_baseline": cpu,
        "e2e": {
            "value": e2e_layers / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "call": f"solver.update_problem(K, omega); solver.run({0.1 * e2e_steps:.1f}, 0.1, {layers_per_step})",
            "includes": "timed: K/omega upload (validation, gate table, plan), device state preparation, "
                        f"{e2e_layers} layers, R(t) every {layers_per_step} layers, read-back; NOT timed: creation of the solver's "
                        "engine (two state buffers) in a warm call before",
        },
        "check": check,
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "norm2_after": norm2,
    }
    if world == 1 and not args.no_other and not c64:
        try:
            line["other_configs"] = other_configs(device)
        except Exception as exc:  # the headline line must survive a failure of the side measurements
            line["other_configs"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if check is not None and not check["ok"]:
        raise SystemExit("bench.py: parity check FAILED: " + json.dumps(check))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--qubits", dest="n", type=int, default=BASE_N,
                    help="total qubits (with --weak: qubits per GPU); default: the headline n=30.  Under torchrun use --qubits "
                         "(torchrun's own parser claims the prefix --n)")
    ap.add_argument("--dtype", default="complex128", choices=["complex128", "complex64"],
                    help="state storage (complex64: 8-byte amplitudes in HBM, FP64 arithmetic on chip; one GPU only)")
    ap.add_argument("--layers-per-call", type=int, default=LAYERS_PER_CALL,
                    help="layers applied per engine call in the timed loop (default: the reference's trotter_per_step = 5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-check", action="store_true", help="skip the parity check block")
    ap.add_argument("--no-other", action="store_true", help="skip the other BASELINE configs (N=1)")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling: --n qubits PER GPU (n=33 on 8 GPUs); default is strong scaling at n total")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
