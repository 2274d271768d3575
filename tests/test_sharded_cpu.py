"""Sharded schedule on the CPU: (1) the compiled sharded program (classic all-to-all exchanges and
exchanges fused into peer-storing passes) emulated in NumPy against the oracle; (2) the all-to-all the
host performs (torch.distributed, gloo, world size 2) is exactly the qubit swap the engine assumes."""

import os
import socket
import sys

import numpy as np
import pytest

from oracle import xy_oracle as oc
from scpn_quantum_control_b200.engine import plan_describe
from tests.plan_sim import permute_bits
from tests.shard_sim import simulate_sharded


@pytest.mark.parametrize("peer", [False, True])
@pytest.mark.parametrize("n,g,order", [(14, 1, 1), (15, 2, 1), (16, 3, 1), (14, 2, 2)])
def test_sharded_program_matches_reference_order(n, g, order, peer):
    K, w = oc.random_problem(n, 500 + n + g)
    Kc, wc = oc.canonicalise(n, K, w)
    prog = plan_describe(Kc, wc, 0.1, order, 2, n_local=n - g, peer_exchange=peer)
    z, p = oc.term_list(Kc, wc)
    psi0 = oc.initial_state(wc)
    ref = oc.evolve_product_formula(psi0.copy(), z, p, 0.1, 2, order)
    got, perm, n_ex, n_fused = simulate_sharded(prog, n, n - g, psi0, wc)
    inv = [0] * n
    for q in range(n):
        inv[perm[q]] = q
    back = permute_bits(got, n, inv)
    assert 1.0 - oc.fidelity(back, ref) < 1e-12
    assert np.abs(back - ref).max() < 1e-12
    assert n_ex + n_fused >= 2
    if peer:
        assert n_fused > 0, "no exchange was fused into a pass"


@pytest.mark.parametrize("peer", [False, True])
@pytest.mark.parametrize("n,g", [(14, 1), (16, 3)])
def test_sharded_program_restores_its_sharding(n, g, peer):
    """With restore_sharding the program ends with the qubits that were sharded on entry back on their rank bits
    (one more exchange), so that the engine can replay the compiled program layer after layer."""
    K, w = oc.random_problem(n, 900 + n + g)
    Kc, wc = oc.canonicalise(n, K, w)
    prog = plan_describe(Kc, wc, 0.1, 1, 1, n_local=n - g, peer_exchange=peer, restore_sharding=True)
    z, p = oc.term_list(Kc, wc)
    psi0 = oc.initial_state(wc)
    ref = oc.evolve_product_formula(psi0.copy(), z, p, 0.1, 1, 1)
    got, perm, n_ex, n_fused = simulate_sharded(prog, n, n - g, psi0, wc)
    assert [perm[q] for q in range(n - g, n)] == list(range(n - g, n)), "sharded qubits did not return to their rank bits"
    inv = [0] * n
    for q in range(n):
        inv[perm[q]] = q
    back = permute_bits(got, n, inv)
    assert 1.0 - oc.fidelity(back, ref) < 1e-12
    assert n_ex + n_fused >= 3
    if peer:
        # the restoring exchange is fused into the last pass too, and that pass leaves the local qubits in the START
        # layout of the first program: the next call needs neither an exchange nor a re-layout
        assert n_ex == 0 and all(op["kind"] == "local" and op["peer_last"] for op in prog["ops"])
        first = prog["ops"][0]
        assert [perm[q] for q in first["loc2log"]] == first["plan"]["start_perm"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _a2a_worker(rank, world, port, n, nl, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    full = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    shard = full[rank << nl:(rank + 1) << nl].copy()
    tin = torch.from_numpy(shard.view(np.float64))
    tout = torch.empty_like(tin)
    dist.all_to_all_single(tout, tin)  # what ShardedXYKEngine._exchange does with NCCL
    np.save(os.path.join(out_dir, f"shard{rank}.npy"), tout.numpy().view(np.complex128))
    dist.destroy_process_group()


def test_gloo_all_to_all_is_the_qubit_swap(tmp_path):
    """world_size 2 (gloo, CPU): all_to_all_single on the shards == swapping the top local bit with the rank bit."""
    import torch.multiprocessing as mp

    n, g = 8, 1
    nl = n - g
    port = _free_port()
    mp.spawn(_a2a_worker, args=(2, port, n, nl, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(7)
    full = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    got = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(2)])
    swap = list(range(n))
    swap[nl - 1], swap[nl] = nl, nl - 1
    want = permute_bits(full, n, swap)
    assert np.array_equal(got, want)


def test_random_sharded_programs():
    """Random coupling graphs, 1-3 sharded qubits, classic and fused exchanges, with and without restoring the sharding:
    emulated result == oracle, structural invariants of every local program hold."""
    from tests.plan_check import check_plan

    rng = np.random.default_rng(777)
    for _ in range(10):
        g = int(rng.integers(1, 4))
        n = min(int(rng.integers(13 + g, 16 + g)), 16)
        order, reps = int(rng.choice([1, 2])), int(rng.choice([1, 2]))
        dens = float(rng.choice([0.3, 0.6, 1.0]))
        peer, rest = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        A = rng.uniform(0.05, 1.0, (n, n)) * (rng.random((n, n)) < dens)
        K = (A + A.T) / 2.0
        np.fill_diagonal(K, 0.0)
        w = rng.uniform(0.5, 4.0, n)
        Kc, wc = oc.canonicalise(n, K, w)
        prog = plan_describe(Kc, wc, 0.1, order, reps, n_local=n - g, peer_exchange=peer, restore_sharding=rest)
        z, p = oc.term_list(Kc, wc)
        psi0 = oc.initial_state(wc)
        ref = oc.evolve_product_formula(psi0.copy(), z, p, 0.1, reps, order)
        got, perm, _, _ = simulate_sharded(prog, n, n - g, psi0, wc)
        inv = [0] * n
        for q in range(n):
            inv[perm[q]] = q
        assert 1.0 - oc.fidelity(permute_bits(got, n, inv), ref) < 1e-12
        for op in prog["ops"]:
            if op["kind"] == "local":
                check_plan(op["plan"])


@pytest.mark.parametrize("n,g", [(14, 1), (15, 2), (16, 3)])
def test_five_layer_sharded_call_matches_reference_order(n, g):
    """one call of five first-order layers (a dt step of the reference's run(), trotter_per_step = 5) with fused
    exchanges and the restoring exchange at its end: the program shape bench.py times on sharded states"""
    K, w = oc.random_problem(n, 1500 + n + g)
    Kc, wc = oc.canonicalise(n, K, w)
    prog = plan_describe(Kc, wc, 0.1, 1, 5, n_local=n - g, peer_exchange=True, restore_sharding=True)
    z, p = oc.term_list(Kc, wc)
    psi0 = oc.initial_state(wc)
    ref = oc.evolve_product_formula(psi0.copy(), z, p, 0.1, 5, 1)
    got, perm, n_ex, n_fused = simulate_sharded(prog, n, n - g, psi0, wc)
    assert [perm[q] for q in range(n - g, n)] == list(range(n - g, n)), "sharded qubits did not return to their rank bits"
    inv = [0] * n
    for q in range(n):
        inv[perm[q]] = q
    back = permute_bits(got, n, inv)
    assert 1.0 - oc.fidelity(back, ref) < 1e-12
    assert np.abs(back - ref).max() < 1e-12
    assert n_ex + n_fused < 15, "five layers in one call need fewer exchanges than five single-layer calls (3 each)"
