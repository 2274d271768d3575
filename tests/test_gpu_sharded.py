"""Sharded (multi-GPU) path: P-GPU result == oracle at small n, echo, conservation laws.
Needs >= 2 GPUs on the box (skipped otherwise); runs tests/dist_sharded_check.py under torchrun."""

import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    from scpn_quantum_control_b200.engine import device_count

    return device_count()


@pytest.mark.parametrize("world,n", [(2, 15), (4, 16), (8, 18)])
def test_sharded_matches_oracle(world, n):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world),
           os.path.join(ROOT, "tests", "dist_sharded_check.py"), str(n)]
    proc = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    line = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1]
    summary = json.loads(line)
    assert summary["ok"] and summary["exchanges"] > 0


@pytest.mark.parametrize("world,n", [(2, 28), (4, 28)])
def test_sharded_equals_single_gpu_at_scale(world, n):
    """1-GPU == P-GPU observables at n = 28 (no CPU oracle at this size): tests/dist_equiv_check.py under torchrun"""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29560 + world),
           os.path.join(ROOT, "tests", "dist_equiv_check.py"), str(n)]
    proc = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    summary = json.loads([ln for ln in proc.stdout.splitlines() if ln.startswith("{")][-1])
    assert summary["ok"] and summary["max_observable_difference_1gpu_vs_sharded"] < 1e-10
