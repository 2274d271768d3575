"""Property test of the schedule compiler on random coupling PATTERNS (host only): whatever the sparsity -- empty rows,
isolated qubits, single pairs, dense blocks -- the compiled pass program, executed in NumPy with its own layouts, is the
reference's gate order (oracle Mode T).  Strategy modelled on the reference's ``st_coupling_matrix`` / ``st_frequencies``
(tests/conftest.py:239-332: symmetric, 0..2, zero diagonal; 0.1..5)."""
import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import xy_oracle as oc
from tests.test_schedule_cpu import _check


@st.composite
def problems(draw):
    n = draw(st.integers(min_value=12, max_value=14))
    seed = draw(st.integers(min_value=0, max_value=2**31 - 1))
    density = draw(st.sampled_from([0.0, 0.05, 0.3, 0.7, 1.0]))
    rng = np.random.default_rng(seed)
    A = rng.uniform(0.0, 2.0, (n, n)) * (rng.uniform(size=(n, n)) < density)
    K = np.triu(A, 1)
    K = K + K.T
    if draw(st.booleans()) and n > 2:   # an isolated qubit
        q = draw(st.integers(min_value=0, max_value=n - 1))
        K[q, :] = 0.0
        K[:, q] = 0.0
    w = rng.uniform(0.1, 5.0, n)
    if draw(st.booleans()):
        w[draw(st.integers(min_value=0, max_value=n - 1))] = 0.0   # a qubit without a Z term
    order = draw(st.sampled_from([1, 2]))
    reps = draw(st.integers(min_value=1, max_value=2))
    delta = draw(st.sampled_from([0.01, 0.05, 0.1, -0.07]))
    return n, K, w, delta, order, reps


@settings(max_examples=60, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(problems())
def test_compiled_program_is_the_reference_gate_order(problem):
    n, K, w, delta, order, reps = problem
    plan = _check(n, K, w, delta, order, reps)
    npairs = int(np.count_nonzero(np.abs(np.triu(K, 1)) >= 1e-15))
    per_rep = npairs if order == 1 else max(0, 2 * npairs - 1)
    assert plan["total_gates"] == per_rep * reps
