"""torchrun timing of first-order layers on a sharded state (n_local per GPU given)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _devlib import use_dev_library
use_dev_library()  # tools/_build/libxyk_exp.so, or the build named by XYK_LIB_AB
from oracle import xy_oracle as oc
from scpn_quantum_control_b200.sharded import ShardedXYKEngine

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank = int(os.environ["RANK"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
world = dist.get_world_size(); g = world.bit_length() - 1
n = nl + g
K, w = oc.random_problem(n, 20260700 + n)
eng = ShardedXYKEngine(n, device=lr)
eng.set_problem(K, w)
eng.init_product_state()
eng.apply_layers(0.02, 1, 1)
eng.eng.enable_timing(True)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(layers):
    eng.apply_layers(0.02, 1, 1)
torch.cuda.synchronize(); dist.barrier()
dt = (time.perf_counter() - t0) / layers
st = eng.stats()
t1 = time.perf_counter(); R, _ = eng.order_parameter(); torch.cuda.synchronize(); t2 = time.perf_counter()
nrm = eng.norm2()
if rank == 0:
    print(json.dumps({"n": n, "world": world, "ms_per_layer": 1e3 * dt, "passes_last": st["last_passes"], "rounds_last": st["last_rounds"],
                      "exchanges_total": st["exchanges"], "exchange_s_total": st["exchange_seconds"],
                      "exchange_GBps": st["exchange_bytes"] / max(st["exchange_seconds"], 1e-9) / 1e9,
                      "sweeps_total": st["state_sweeps"], "pass_ms_mean": st["pass_kernel_ms"] / max(1, st["pass_kernel_timed"]), "peer_pass_ms_mean": st["peer_pass_ms"] / max(1, st["peer_passes"]), "peer_passes": st["peer_passes"], "passes_timed": st["pass_kernel_timed"], "exchange_ms": [round(x, 1) for x in eng.exchange_log], "R": R, "R_ms": 1e3 * (t2 - t1), "norm2": nrm}))
eng.close(); dist.destroy_process_group()
