"""Multi-GPU check, launched by torchrun (one rank per GPU):

    torchrun --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29533 tests/dist_sharded_check.py [n]

(1) the sharded engine reproduces the single-state result of the C oracle at small n (state fidelity
    and R), for orders 1 and 2;  (2) echo: L layers forward then the exact inverse sequence returns
    the analytic product state;  (3) norm and <sum Z> are conserved.
Exit code 0 = all good (rank 0 prints a JSON summary)."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from oracle import xy_oracle as oc  # noqa: E402
from scpn_quantum_control_b200.sharded import ShardedXYKEngine  # noqa: E402


def c_oracle_evolve(n, K, omega, delta, reps, order):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "libxyoracle.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    Kc, wc = oc.canonicalise(n, K, omega)
    Ks = np.ascontiguousarray((Kc + Kc.T) / 2.0)
    psi = np.zeros(2 << n)
    lib.xyo_init_state(n, wc.ctypes.data_as(dp), psi.ctypes.data_as(dp))
    lib.xyo_evolve.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, dp]
    lib.xyo_evolve(n, Ks.ctypes.data_as(dp), wc.ctypes.data_as(dp), float(delta), reps, order, psi.ctypes.data_as(dp))
    return psi.view(np.complex128)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = dist.get_world_size()
    summary = {"n": n, "world": world, "checks": []}
    ok = True
    K, w = oc.random_problem(n, 9000 + n)
    eng = ShardedXYKEngine(n, device=local_rank)
    eng.set_problem(K, w)
    for order, reps in ((1, 2), (2, 1)):
        eng.init_product_state()
        eng.apply_layers(0.1, order, reps)
        R, _ = eng.order_parameter()
        nrm = eng.norm2()
        zsum = eng.z_expectations().sum()
        psi = eng.gather_state()
        if rank == 0:
            ref = c_oracle_evolve(n, K, w, 0.1, reps, order)
            infid = 1.0 - oc.fidelity(psi, ref)
            dR = abs(R - oc.order_parameter(ref, n)[0])
            zref = np.cos(np.mod(w, 2 * np.pi)).sum()
            good = infid < 1e-10 and dR < 1e-9 and abs(nrm - 1) < 1e-10 and abs(zsum - zref) < 1e-9
            ok &= bool(good)
            summary["checks"].append({"order": order, "reps": reps, "infidelity": infid, "dR": dR, "norm2": nrm,
                                      "sumZ_err": abs(zsum - zref), "ok": bool(good)})
    # program reuse: every call ends with the sharding it started from, and so does an R(t) evaluation
    eng.init_product_state()
    g0 = eng.global_qubits()
    eng.apply_layers(0.1, 1, 1)
    g1 = eng.global_qubits()
    R1, _ = eng.order_parameter()
    g2 = eng.global_qubits()
    eng.apply_layers(0.1, 1, 1)
    R2, _ = eng.order_parameter()
    if rank == 0:
        ref = c_oracle_evolve(n, K, w, 0.2, 2, 1)
        dR = abs(R2 - oc.order_parameter(ref, n)[0])
        good = g0 == g1 == g2 and dR < 1e-9
        ok &= bool(good)
        summary["checks"].append({"sharding_restored": g0 == g1 == g2, "dR_two_calls": dR, "ok": bool(good)})
    # echo: forward L layers, then the inverse sequence = order-2 layers with negative time are exact inverses
    eng.init_product_state()
    eng.apply_layers(0.2, 2, 2)
    eng.apply_layers(-0.2, 2, 2)
    psi = eng.gather_state()
    st = eng.stats()
    if rank == 0:
        ref0 = oc.initial_state(oc.canonicalise(n, K, w)[1])
        infid = 1.0 - oc.fidelity(psi, ref0)
        good = infid < 1e-10
        ok &= bool(good)
        summary["checks"].append({"echo_infidelity": infid, "ok": bool(good)})
        summary["exchanges"] = st["exchanges"]
        summary["ok"] = bool(ok)
        print(json.dumps(summary), flush=True)
    eng.close()
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local_rank}")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
