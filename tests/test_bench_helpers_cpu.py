"""CPU tests of bench.py's checker logic (no GPU): the closed-form disjoint-pair check of the parity block, run on a NumPy
engine twin, must agree with the oracle's product formula -- so that a red `check` in a bench line means the engine is wrong,
not the checker."""
import numpy as np
import pytest

import bench
from oracle import xy_oracle as oc


class _OracleEngine:
    """the engine methods matching_check uses, backed by the oracle"""

    def __init__(self, n):
        self.n = n

    def set_problem(self, K, w):
        self.K, self.w = oc.canonicalise(self.n, K, w)

    def init_product_state(self):
        self.psi = oc.initial_state(self.w)

    def apply_layers(self, delta, order, reps):
        z, p = oc.term_list(self.K, self.w)
        self.psi = oc.evolve_product_formula(self.psi, z, p, delta, reps, order)

    def xy_expectations(self):
        return oc.xy_expectations(self.psi, self.n)

    def z_expectations(self):
        return oc.z_expectations(self.psi, self.n)


@pytest.mark.parametrize("n", [2, 5, 8, 11])
def test_disjoint_pair_closed_form_matches_the_oracle(n):
    _K, w = bench.synthetic_problem(n, 5)
    assert bench.matching_check(_OracleEngine(n), n, w, 0.02) < 1e-13


def test_disjoint_pair_check_detects_a_wrong_engine():
    class Broken(_OracleEngine):
        def apply_layers(self, delta, order, reps):
            super().apply_layers(delta * 1.01, order, reps)   # a 1 % error in the step

    n = 6
    _K, w = bench.synthetic_problem(n, 5)
    assert bench.matching_check(Broken(n), n, w, 0.02) > 1e-6


def test_bench_metric_and_config_labels():
    assert bench.metric_name(30, False) == bench.METRIC
    assert "n=33" in bench.metric_name(33, True) and "weak" in bench.metric_name(33, True)
    cfg = bench.common_config(30)
    assert cfg["n_qubits"] == 30 and cfg["pairs_per_layer"] == 435 and "complex128" in cfg["workload"]
    assert bench.LAYERS_PER_CALL == 5
