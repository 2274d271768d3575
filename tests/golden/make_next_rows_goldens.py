#!/usr/bin/env python
"""Generate tests/golden/next_rows_vectors.json: outputs of ``oracle/xy_oracle_next.py`` on seeded inputs for the SURVEY
8(f) rows (feedback controller history, XXZ R(t) traces in Mode T / Mode E, Floquet drive).  Oracle-derived, not
reference-derived: the reference's modules for these rows import Qiskit, which does not exist in the build container, so
they cannot be run here.  The fixtures pin the oracle against regressions (``-m "not gpu"``) and give the ``-m gpu`` tests
committed numbers to compare with."""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import xy_oracle as oc  # noqa: E402
from oracle import xy_oracle_next as on  # noqa: E402


def main() -> None:
    out: dict = {"_source": "oracle/xy_oracle_next.py (seeded inputs; see tests/golden/make_next_rows_goldens.py)"}
    n = 4
    K, w = oc.paper27_K(16)[:n, :n] * 2.0, oc.OMEGA_N_16[:n]
    kw = dict(target_r=0.6, deadband=0.02, measurement_shots=64, correction_angle=0.2)
    rows, _ = on.feedback_run(K, w, 6, 123, order=1, mode="T", **kw)
    out["feedback_n4_seed123_trotter_order1"] = {
        "K": "2 * paper27_K(16)[:4, :4]", "omega": "OMEGA_N_16[:4]", "config": kw, "n_steps": 6, "seed": 123,
        "history": [{k: v for k, v in r.items() if k != "readout_counts"} | {"readout_counts": r["readout_counts"]} for r in rows],
    }
    n = 6
    K, w = oc.random_problem(n, 11)
    xxz = {"n": n, "problem": "oracle.random_problem(6, 11)", "delta": 0.7, "t_max": 0.3, "dt": 0.1}
    for order in (1, 2):
        _, R = on.run_xxz(n, K, w, 0.7, 0.3, 0.1, 4, order=order, mode="T")
        xxz[f"R_trotter_order{order}_4_per_step"] = R.tolist()
    _, RE = on.run_xxz(n, K, w, 0.7, 0.3, 0.1, mode="E")
    xxz["R_exact"] = RE.tolist()
    out["xxz_n6"] = xxz
    n = 4
    K = oc.paper27_K(16)[:n, :n]
    fl = on.floquet_evolve(K / K.max(), oc.OMEGA_N_16[:n], 0.5, 0.6, 2.0, 2, 8)
    out["floquet_n4"] = {"K_topology": "paper27_K(16)[:4, :4] / max", "omega": "OMEGA_N_16[:4]", "K_base": 0.5, "drive_amplitude": 0.6,
                         "drive_frequency": 2.0, "n_periods": 2, "steps_per_period": 8, "times": fl["times"].tolist(),
                         "R_values": fl["R_values"].tolist(), "drive_signal": fl["drive_signal"].tolist(),
                         "subharmonic_ratio": fl["subharmonic_ratio"]}
    n = 4
    K = oc.paper27_K(16)[:n, :n] * 1.5
    tr = on.mcwf_trajectory(K, oc.OMEGA_N_16[:n], 0.6, 0.4, 1.0, 0.1, 7)
    out["mcwf_n4_seed7"] = {"K": "1.5 * paper27_K(16)[:4, :4]", "omega": "OMEGA_N_16[:4]", "gamma_amp": 0.6, "gamma_deph": 0.4,
                            "t_max": 1.0, "dt": 0.1, "seed": 7, "R": tr["R"].tolist(), "n_jumps": tr["n_jumps"],
                            "psi_final_re": tr["psi_final"].real.tolist(), "psi_final_im": tr["psi_final"].imag.tolist()}
    (HERE / "next_rows_vectors.json").write_text(json.dumps(out, indent=1))
    print("wrote", HERE / "next_rows_vectors.json")


if __name__ == "__main__":
    main()
