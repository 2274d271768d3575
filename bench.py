#!/usr/bin/env python
"""bench.py -- XY Trotter steps/s on the BASELINE.json headline workload.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--n QUBITS]

A "step" is ONE first-order Trotter layer (one omega.Z diagonal phase + every pair rotation
exp(-i theta_ij (XX+YY)) in the reference's order) applied to the resident statevector:
n = 30 all-to-all random K_nm, complex128 (16 GiB state) on one B200 (BASELINE.json configs[3],
the configuration the metric "XY Trotter steps/sec at n=30" is quoted on).  With N > 1 ranks the
per-GPU shard stays 2^30 amplitudes (weak scaling, n = 30 + log2 N; n = 33 on 8 GPUs).

Printed JSON (one line, rank 0):
  value / ms_per_step  device-timed (CUDA events on the engine's stream) steps with the state
                       resident in HBM; max over ranks.
  e2e                  the same metric through the public API
                       ``QuantumKuramotoSolver(n, K, omega, evolution="trotter").run(t_max, dt, 5)``
                       from HOST inputs (K, omega) to HOST outputs (R(t)): includes problem upload,
                       state preparation, an R(t) evaluation every 5 layers and the read-back.
  roofline             dominant kernel = tile_pass_kernel; achieved = ALGORITHMIC bytes per launch
                       (one read + one write of the state, 2 * 16 B * 2^n) / mean launch duration.
  cpu_baseline         the oracle's C/OpenMP restatement of the reference path timed on this box's
                       host cores on a bounded sample (one layer at n = 26, scaled to n = 30).
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
CPU_SAMPLE_N = 26
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
def cpu_layers_per_sec(steps: int, warmup: int, sample_n: int = CPU_SAMPLE_N):
    import ctypes

    so = os.path.join(ROOT, "oracle", "_build", "libxyoracle.so")
    if not os.path.exists(so):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], env=env, stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.xyo_layer_order1.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, dp]
    lib.xyo_num_threads.restype = ctypes.c_int
    threads = int(lib.xyo_num_threads())
    K, w = synthetic_problem(sample_n, SEED)
    Ks = np.ascontiguousarray(K)
    psi = np.zeros(2 << sample_n)
    lib.xyo_init_state(sample_n, w.ctypes.data_as(dp), psi.ctypes.data_as(dp))
    for _ in range(warmup):
        lib.xyo_layer_order1(sample_n, Ks.ctypes.data_as(dp), w.ctypes.data_as(dp), 0.02, psi.ctypes.data_as(dp))
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.xyo_layer_order1(sample_n, Ks.ctypes.data_as(dp), w.ctypes.data_as(dp), 0.02, psi.ctypes.data_as(dp))
    dt = (time.perf_counter() - t0) / max(1, steps)
    return dt, threads


def scale_to_n30(sec_per_layer_sample: float, sample_n: int, target_n: int) -> float:
    """cost of a layer ~ (#pairs) * 2^n: extrapolate the sample to the target size."""
    pairs = lambda n: n * (n - 1) / 2.0  # noqa: E731
    return sec_per_layer_sample * (pairs(target_n) / pairs(sample_n)) * 2.0 ** (target_n - sample_n)


def run_reference_arm(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    n = args.n + (max(0, world.bit_length() - 1) if args.weak else 0)
    sec, threads = cpu_layers_per_sec(args.steps, min(args.warmup, 1))
    sec_full = scale_to_n30(sec, CPU_SAMPLE_N, n)
    value = 1.0 / sec_full
    sample = (f"each step = one first-order layer of the C/OpenMP oracle port at n={CPU_SAMPLE_N} "
              f"(seed {SEED}), scaled by (pairs({n})/pairs({CPU_SAMPLE_N})) * 2^{n - CPU_SAMPLE_N} to n={n}")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec_full, "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"n={n} all-to-all XY Trotter layer (first order), complex128", "n_qubits": n,
                   "pairs_per_layer": n * (n - 1) // 2, "seed": SEED},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = Python+Qiskit+Rust; neither Qiskit nor cargo exists in this image, so the timed CPU arm "
                "is the oracle's gate-by-gate C/OpenMP restatement of the same path",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_gpu_arm(args, rank: int, world: int, local_rank: int) -> None:
    import torch  # device memory / streams / torch.distributed plumbing only

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

    if world == 1:
        eng = XYKEngine(n, device=device)
        eng.set_problem(K, w)
        eng.set_engine("tiled" if n >= 12 else "simple")
        raw = eng
    else:
        from scpn_quantum_control_b200.sharded import ShardedXYKEngine

        eng = ShardedXYKEngine(n, device=device)
        eng.set_problem(K, w)
        raw = eng.eng
    eng.init_product_state()
    for _ in range(max(args.warmup, 3)):
        eng.apply_layers(tau, 1, 1)
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
        for _ in range(args.steps):
            eng.apply_layers(tau, 1, 1)
        ev1.record(stream)
        ev1.synchronize()
        torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    st1 = eng.stats()
    norm2 = eng.norm2()
    if world > 1:
        st1 = dict(st1)
    if dist is not None:
        t = torch.tensor([ms_total], device=f"cuda:{device}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    passes = st1["pass_launches"] - st0["pass_launches"]
    sweeps = st1["state_sweeps"] - st0["state_sweeps"]
    pass_ms = (st1["pass_kernel_ms"] - st0["pass_kernel_ms"]) / max(1, st1["pass_kernel_timed"] - st0["pass_kernel_timed"])
    rounds = st1["last_rounds"]
    eng.close()

    # ---- e2e through the public API, host buffers in, host R(t) out ----
    e2e_steps = 2
    layers_per_step = 5
    solver = QuantumKuramotoSolver(n, K, w, evolution="trotter", device=device, distributed=world > 1)
    solver.run(0.1 * 1, 0.1, layers_per_step)  # warm (plans, allocations)
    torch.cuda.synchronize(device)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    res = solver.run(0.1 * e2e_steps, 0.1, layers_per_step)
    torch.cuda.synchronize(device)
    e2e_sec = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_sec], device=f"cuda:{device}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    solver.close()
    e2e_layers = e2e_steps * layers_per_step
    h2d = (K.nbytes + w.nbytes) // e2e_layers + 4000  # problem upload amortised + one pass descriptor set per layer
    d2h = (res["R"].nbytes + 2 * n * 8 * (e2e_steps + 1)) // e2e_layers

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    n_shard = n - gbits
    alg_bytes = 2.0 * 16.0 * 2.0 ** n_shard   # per launch: one read + one write of this GPU's shard
    achieved = alg_bytes / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else 0.0
    # one step = one first-order Trotter layer of the WHOLE n-qubit state (all ranks together)
    value = 1e3 / ms_per_step
    exch = {}
    if world > 1:
        exch = {"exchanges_per_layer": (st1["exchanges"] - st0.get("exchanges", 0)) / args.steps,
                "exchange_ms_mean": float(np.mean(eng.exchange_log[-8:])) if eng.exchange_log else None,
                "nvlink_bytes_per_exchange_per_gpu": int(16 * 2 ** n_shard * (world - 1) // world)}
    cpu_sec, threads = cpu_layers_per_sec(1, 1) if not args.no_cpu else (float("nan"), 0)
    cpu_sec_full = scale_to_n30(cpu_sec, CPU_SAMPLE_N, n)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"n={n} all-to-all XY Trotter layer (first order), complex128, {16 * 2 ** n_shard / 2 ** 30:.2f} GiB shard per GPU"
                        + ("" if world == 1 else f", state sharded by its top {gbits} qubits over {world} GPUs, NCCL all-to-all qubit exchanges"),
            "n_qubits": n, "pairs_per_layer": n * (n - 1) // 2, "seed": SEED, "tau": tau,
            "l2_policy": "state (16 GiB) >> L2 (126 MB): every pass streams from HBM, no flush needed",
            "passes_per_layer": passes / args.steps, "rounds_per_layer": rounds, "sweeps_per_layer": sweeps / args.steps,
            **exch,
        },
        "roofline": {
            "bound": "hbm", "kernel": "tile_pass_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(n_shard),
            "algorithmic_bytes_per_launch": alg_bytes, "mean_launch_ms": pass_ms,
            "layer_algorithmic_frac": (1e3 / ms_per_step) * alg_bytes / 1e9 / peak,
            "layer_sweeps_frac": (1e3 / ms_per_step) * (sweeps / args.steps) * alg_bytes / 1e9 / peak,
        },
        "cpu_baseline": {
            "value": 1.0 / cpu_sec_full if cpu_sec_full == cpu_sec_full else None, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one first-order layer of the C/OpenMP oracle port at n={CPU_SAMPLE_N}, scaled by "
                      f"(pairs({n})/pairs({CPU_SAMPLE_N})) * 2^{n - CPU_SAMPLE_N}",
        },
        "e2e": {
            "value": e2e_layers / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "call": f"QuantumKuramotoSolver(n, K, omega, evolution='trotter').run({0.1 * e2e_steps:.1f}, 0.1, {layers_per_step})",
            "includes": "problem upload, device state preparation, R(t) every 5 layers, read-back",
        },
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "norm2_after": norm2,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=BASE_N, help="qubits per GPU (default: the headline n=30)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
