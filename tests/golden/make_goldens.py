#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Run in the BUILD container (where /root/reference is mounted); the GPU box never reads
/root/reference, only the JSON files this script writes.

Two files are produced:

* ``reference_goldens.json`` -- numbers copied verbatim from the reference's own tests,
  result files and docs (data, not code).  These pin the oracle (Mode E to ~1e-15, Mode T
  to the 3 printed digits the reference docs hold).
* ``oracle_vectors.json`` -- outputs of ``oracle/xy_oracle.py`` on seeded inputs (Mode T
  R(t) traces and final-state checksums) used by the ``-m gpu`` parity tests so that they do
  not need to re-run the slow exact path.  They are oracle-derived, not reference-derived:
  Qiskit is not importable in the build container, so the reference itself cannot be run.
"""

from __future__ import annotations

import json
import re
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

from oracle import xy_oracle as oc  # noqa: E402


def reference_goldens() -> dict:
    out: dict = {"_source": "anulum/scpn-quantum-control (paths relative to the reference root)"}

    c16 = json.loads((REF / "results/classical_16q_reference.json").read_text())
    out["classical_16q_reference"] = {
        "_file": "results/classical_16q_reference.json",
        "evo_R_dt0.1": {str(n): c16[f"evo_{n}q_dt0.1"]["R"] for n in (2, 4, 6, 8, 10, 12, 14, 16)},
        "evo_16q_single_step": {k: c16[f"evo_16q_dt{k}"]["R"] for k in ("0.01", "0.02", "0.05", "0.1", "0.2")},
        "evo_16q_8step_dt0.05": c16["evo_16q_8step_dt0.05"]["R_trajectory"],
        "evo_16q_8step_times": c16["evo_16q_8step_dt0.05"]["times"],
    }

    # the same R(0.1) values as the reference test pins them (tests/...regression.py:35-46)
    txt = (REF / "tests/test_classical_reference_regression.py").read_text()
    block = txt.split("def test_evolution_R_dt01(")[0]
    pairs = re.findall(r"\((\d+),\s*([0-9.]+)\)", block)
    out["test_classical_reference_regression"] = {
        "_file": "tests/test_classical_reference_regression.py:35-46",
        "R_dt01": {n: float(v) for n, v in pairs[:6]},
        "abs_tol": 1e-10,
    }

    qt = json.loads((REF / "results/qutip_comparison_2026-03-30.json").read_text())
    out["qutip_comparison_qiskit_unitary"] = {
        "_file": "results/qutip_comparison_2026-03-30.json (qiskit_unitary.R); problem from "
        "notebooks/48_qutip_comparison.py:32-36,148-149",
        "n": 4,
        "K": "0.45*exp(-0.3*|i-j|)",
        "omega": "linspace(0.8,1.2,4)",
        "t_max": 2.0,
        "dt": 0.05,
        "R": qt["qiskit_unitary"]["R"],
    }

    n8 = json.loads((REF / "data/classical_quantum_comparison/reproducible_comparison_n8.json").read_text())
    rows = {r["method"]: r for r in n8["rows"]}
    out["reproducible_comparison_n8"] = {
        "_file": "data/classical_quantum_comparison/reproducible_comparison_n8.json",
        "n": 8,
        "t_max": n8["t_max"],
        "dt": n8["dt"],
        "trotter_per_step": n8["trotter_per_step"],
        "quantum_trotter_r_final": rows["quantum_trotter"]["r_final"],
        "classical_exact_r_final": rows["classical_exact"]["r_final"],
    }

    out["doc_traces_3_digits"] = {
        "_file": "docs/pipeline_performance.md:376,410 (older Qiskit => product formula, Mode T order 1 x 5 reps)",
        "run_t0.3_dt0.1_n4": [0.806, 0.796, 0.766],
        "upde_run_5steps_dt0.05_n4": [0.806, 0.804, 0.796, 0.783, 0.765, 0.743],
    }
    out["time_grid_example"] = {
        "_file": "tests/test_xy_kuramoto.py:393-411",
        "t_max": 0.35,
        "dt": 0.1,
        "times": [0.0, 0.1, 0.2, 0.3, 0.35],
        "step_sizes": [0.1, 0.1, 0.1, 0.05],
    }
    return out


def state_digest(psi: np.ndarray) -> dict:
    """Size-independent checksums of a state: a few projections onto seeded random vectors."""
    rng = np.random.default_rng(12345)
    dim = psi.size
    outs = []
    for _ in range(4):
        v = rng.standard_normal(dim) + 1j * rng.standard_normal(dim)
        s = np.vdot(v, psi)
        outs.append([float(s.real), float(s.imag)])
    return {"proj": outs, "norm": float(np.vdot(psi, psi).real)}


def oracle_vectors() -> dict:
    out: dict = {"_source": "oracle/xy_oracle.py (oracle-derived; see make_goldens.py docstring)"}
    cases = {}

    def add(name, n, K, omega, t_max, dt, reps, order, mode="T"):
        times, R, psi = oc.run(n, K, omega, t_max, dt, reps, order=order, mode=mode, return_state=True)
        cases[name] = {
            "n": n,
            "K": np.asarray(K).tolist(),
            "omega": np.asarray(omega).tolist(),
            "t_max": t_max,
            "dt": dt,
            "trotter_per_step": reps,
            "order": order,
            "mode": mode,
            "times": times.tolist(),
            "R": R.tolist(),
            "digest": state_digest(psi),
        }
        if n <= 6:
            cases[name]["state_re"] = psi.real.tolist()
            cases[name]["state_im"] = psi.imag.tolist()

    w4 = oc.OMEGA_N_16[:4]
    for order in (1, 2):
        add(f"n4_paper27_o{order}", 4, oc.paper27_K(4), w4, 1.0, 0.1, 5, order)
        add(f"n4_ring_o{order}", 4, oc.ring_K(4), w4, 1.0, 0.1, 5, order)
    add("n4_paper27_exact", 4, oc.paper27_K(4), w4, 1.0, 0.1, 5, 1, mode="E")
    add("n4_ring_exact", 4, oc.ring_K(4), w4, 1.0, 0.1, 5, 1, mode="E")
    add("n4_ring_partial_grid_o1", 4, oc.ring_K(4), w4, 0.35, 0.1, 3, 1)
    add("n8_paper27_o1", 8, oc.paper27_K(8), oc.OMEGA_N_16[:8], 1.0, 0.1, 5, 1)
    add("n8_paper27_o2", 8, oc.paper27_K(8), oc.OMEGA_N_16[:8], 1.0, 0.1, 5, 2)
    add("n8_paper27_exact", 8, oc.paper27_K(8), oc.OMEGA_N_16[:8], 1.0, 0.1, 5, 1, mode="E")
    K, w = oc.random_problem(10, 20260710)
    add("n10_random_o1", 10, K, w, 0.5, 0.1, 5, 1)
    add("n10_random_o2", 10, K, w, 0.5, 0.1, 5, 2)
    K, w = oc.random_problem(12, 20260716)
    add("n12_random_o1", 12, K, w, 0.4, 0.1, 5, 1)
    add("n12_random_exact", 12, K, w, 0.2, 0.1, 5, 1, mode="E")
    # config 2 of BASELINE.json (n=16 all-to-all random K, 200 layers): R trace only
    K, w = oc.random_problem(16, 20260716)
    add("n16_random_o1_200layers", 16, K, w, 4.0, 0.1, 5, 1)
    add("n16_paper27_o1_8step", 16, oc.paper27_K(16), oc.OMEGA_N_16, 0.4, 0.05, 5, 1)
    out["cases"] = cases
    return out


def main() -> None:
    if REF.exists():
        (HERE / "reference_goldens.json").write_text(json.dumps(reference_goldens(), indent=1) + "\n")
        print("wrote reference_goldens.json")
    else:
        print("no /root/reference: keeping the committed reference_goldens.json")
    (HERE / "oracle_vectors.json").write_text(json.dumps(oracle_vectors()) + "\n")
    print("wrote oracle_vectors.json")


if __name__ == "__main__":
    main()
