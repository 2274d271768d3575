"""torchrun helper (one process per GPU): the P-GPU sharded engine and the single-GPU engine give the same observables on
identical inputs at a size no CPU oracle reaches in seconds (default n = 28): <X_q>, <Y_q>, <Z_q>, R after
first-order and second-order layers.  Rank 0 prints one JSON line and every rank exits 0 iff the check passed."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import xy_oracle as oc  # noqa: E402  (problem generator only)
from scpn_quantum_control_b200.engine import XYKEngine  # noqa: E402
from scpn_quantum_control_b200.sharded import ShardedXYKEngine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = dist.get_world_size()
    K, w = oc.random_problem(n, 8000 + n)
    sequence = ((0.05, 1, 2), (0.04, 2, 1))   # (delta, order, reps), applied one after the other
    eng = ShardedXYKEngine(n, device=local_rank)
    eng.set_problem(K, w)
    eng.init_product_state()
    got = []
    for delta, order, reps in sequence:
        eng.apply_layers(delta, order, reps)
        ex, ey = eng.xy_expectations()
        got.append(np.concatenate([ex, ey, eng.z_expectations()]))
    exchanges = eng.stats()["exchanges"] + eng.stats()["fused_exchanges"]
    eng.close()
    ok = True
    if rank == 0:
        worst = 0.0
        with XYKEngine(n, device=local_rank) as one:
            one.set_problem(K, w)
            one.set_engine("tiled")
            one.init_product_state()
            for k, (delta, order, reps) in enumerate(sequence):
                one.apply_layers(delta, order, reps)
                ex, ey = one.xy_expectations()
                ref = np.concatenate([ex, ey, one.z_expectations()])
                worst = max(worst, float(np.abs(got[k] - ref).max()))
        ok = worst < 1e-10 and exchanges > 0
        print(json.dumps({"n": n, "world": world, "max_observable_difference_1gpu_vs_sharded": worst, "exchanges": exchanges,
                          "ok": bool(ok)}), flush=True)
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local_rank}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
